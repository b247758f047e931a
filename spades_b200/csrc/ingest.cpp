// ingest.cpp -- host-side read ingest. Pure host code: needs no GPU, no context.
//
//   FASTA / FASTQ, plain or gzip   io::FastaFastqGzParser over kseq + zlib (src/common/io/reads/fasta_fastq_gz_parser.hpp:25-150,
//                                  ext/include/kseq/kseq.h): '>' / '@' records, name = first token, multi-line sequence, '+' starts
//                                  the quality which runs until it is as long as the sequence
//   N handling                     io::LongestValid (io/reads/longest_valid_wrapper.hpp:16-53): keep the FIRST longest run of
//                                  ACGTacgt (is_nucl, sequence/nucl.hpp:48-66); the tools apply it to every read
//                                  (io_helper.cpp:30-31, read_converter.cpp:115,121)
//   2-bit packing                  Sequence / RtSeq layout (sequence/rtseq.hpp:379-382): base i at bits 2(i%32) of word i/32,
//                                  A=0 C=1 G=2 T=3 (dignucl, nucl.hpp:132-146), every read on a word boundary
//   binary read streams            <prefix>.seq = ReadStreamStat {u64 read_count, max_len, total_len} (io/reads/read_stream.hpp:21-38)
//                                  then per read {u64 size, ceil(size/32) u64 words, u16 left_offset, u16 right_offset, u64 tag}
//                                  (Sequence::BinWrite sequence.hpp:817-830, SingleReadSeq::BinWrite single_read.hpp:325-338);
//                                  <prefix>.off = u64 file offset of every 100th read (BinaryWriter::CHUNK, binary_converter.cpp:96-110)
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/spades_b200.h"

struct sgpu_read_batch {
    std::vector<uint64_t> words;
    std::vector<uint64_t> offs;
    std::vector<uint32_t> lens;
    uint64_t records = 0;        // records in the file (before empty reads were dropped)
    uint64_t trimmed = 0;        // reads shortened by LongestValid
    uint64_t dropped = 0;        // reads with no valid base at all
    std::string err;
};

namespace {

inline int nucl_code(unsigned char c) {      // -1 = not a nucleotide
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

struct GzReader {
    gzFile f = nullptr;
    std::vector<unsigned char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    bool open(const char *path) {
        f = gzopen(path, "rb");       // transparently reads plain files too
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(1 << 20);
        return true;
    }
    ~GzReader() { if (f) gzclose(f); }
    int get() {
        if (pos == end) {
            if (eof) return -1;
            int n = gzread(f, buf.data(), (unsigned)buf.size());
            if (n <= 0) { eof = true; return -1; }
            pos = 0; end = (size_t)n;
        }
        return buf[pos++];
    }
    int peek() {
        int c = get();
        if (c >= 0) --pos;
        return c;
    }
};

void add_read(sgpu_read_batch *b, const std::string &seq, bool longest_valid) {
    ++b->records;
    size_t from = 0, to = seq.size();
    if (longest_valid) {
        // first longest run of nucleotides (LongestValidCoords)
        size_t best_len = 0, best_pos = 0, run = 0;
        for (size_t i = 0; i <= seq.size(); ++i) {
            if (i < seq.size() && nucl_code((unsigned char)seq[i]) >= 0) ++run;
            else {
                if (run > best_len) { best_len = run; best_pos = i - run; }
                run = 0;
            }
        }
        from = best_pos; to = best_pos + best_len;
        if (best_len < seq.size()) ++b->trimmed;
    } else {
        for (size_t i = 0; i < seq.size(); ++i)
            if (nucl_code((unsigned char)seq[i]) < 0) { from = to = 0; break; }      // without N handling an invalid read contributes nothing
    }
    const size_t len = to - from;
    if (len == 0) { ++b->dropped; return; }
    const size_t nw = (len + 31) / 32;
    const size_t w0 = b->words.size();
    b->words.resize(w0 + nw, 0);
    uint64_t *w = b->words.data() + w0;
    for (size_t i = 0; i < len; ++i) w[i >> 5] |= (uint64_t)nucl_code((unsigned char)seq[from + i]) << ((i & 31) << 1);
    b->offs.push_back((uint64_t)w0);
    b->lens.push_back((uint32_t)len);
}

// kseq_read semantics (ext/include/kseq/kseq.h): skip to the next '>' / '@'; name up to the first space; the rest of the header
// line is the comment; sequence lines are concatenated (blanks skipped) until a line starts with '>', '@' or '+'; after '+' the
// quality is read until it is at least as long as the sequence
bool parse_fastx(const char *path, bool longest_valid, sgpu_read_batch *b) {
    GzReader in;
    if (!in.open(path)) { b->err = std::string("cannot open ") + path; return false; }
    int c;
    std::string seq;
    // find the first header
    while ((c = in.get()) >= 0 && c != '>' && c != '@') {}
    while (c >= 0) {
        // header line
        while ((c = in.get()) >= 0 && c != '\n') {}
        seq.clear();
        // sequence lines
        bool at_line_start = true;
        for (;;) {
            c = in.get();
            if (c < 0) break;
            if (at_line_start && (c == '>' || c == '@' || c == '+')) break;
            if (c == '\n') { at_line_start = true; continue; }
            at_line_start = false;
            if (c == '\r' || c == ' ' || c == '\t') continue;          // kseq keeps only isgraph() characters
            seq.push_back((char)c);
        }
        add_read(b, seq, longest_valid);
        if (c == '+') {
            while ((c = in.get()) >= 0 && c != '\n') {}                  // rest of the '+' line
            size_t q = 0;
            while (q < seq.size() && (c = in.get()) >= 0) if (c != '\n' && c != '\r') ++q;
            if (q < seq.size()) { b->err = "truncated quality string"; return false; }
            while ((c = in.get()) >= 0 && c != '>' && c != '@') {}      // next header
        }
    }
    return true;
}

bool parse_seqfile(const char *prefix, sgpu_read_batch *b) {
    const std::string path = std::string(prefix) + ".seq";
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { b->err = "cannot open " + path; return false; }
    uint64_t stat[3];
    bool ok = fread(stat, 8, 3, f) == 3;
    for (uint64_t r = 0; ok && r < stat[0]; ++r) {
        uint64_t size = 0;
        ok = fread(&size, 8, 1, f) == 1 && size <= 0xffffffffull;
        if (!ok) break;
        const size_t nw = (size_t)((size + 31) / 32);
        const size_t w0 = b->words.size();
        b->words.resize(w0 + nw);
        uint8_t tail[12];
        ok = (nw == 0 || fread(b->words.data() + w0, 8, nw, f) == nw) && fread(tail, 1, 12, f) == 12;
        if (!ok) break;
        ++b->records;
        if (size == 0) { ++b->dropped; continue; }
        if (size & 31) b->words[w0 + nw - 1] &= (~0ull) >> (64 - 2 * (size & 31));      // bits above the sequence are not part of the contract
        b->offs.push_back((uint64_t)w0);
        b->lens.push_back((uint32_t)size);
    }
    fclose(f);
    if (!ok) b->err = "malformed " + path;
    return ok;
}

}  // namespace

extern "C" {

int sgpu_fastx_parse(const char *path, int longest_valid, sgpu_read_batch **out) {
    if (!path || !out) return SGPU_EINVAL;
    sgpu_read_batch *b = new sgpu_read_batch();
    *out = b;                                   // returned even on failure so that sgpu_read_batch_error() can be read
    return parse_fastx(path, longest_valid != 0, b) ? SGPU_OK : SGPU_EIO;
}
int sgpu_seqfile_parse(const char *prefix, sgpu_read_batch **out) {
    if (!prefix || !out) return SGPU_EINVAL;
    sgpu_read_batch *b = new sgpu_read_batch();
    *out = b;
    return parse_seqfile(prefix, b) ? SGPU_OK : SGPU_EIO;
}
int64_t sgpu_read_batch_num_reads(const sgpu_read_batch *b) { return b ? (int64_t)b->lens.size() : 0; }
uint64_t sgpu_read_batch_num_words(const sgpu_read_batch *b) { return b ? (uint64_t)b->words.size() : 0; }
const uint64_t *sgpu_read_batch_words(const sgpu_read_batch *b) { return b ? b->words.data() : nullptr; }
const uint64_t *sgpu_read_batch_offs(const sgpu_read_batch *b) { return b ? b->offs.data() : nullptr; }
const uint32_t *sgpu_read_batch_lens(const sgpu_read_batch *b) { return b ? b->lens.data() : nullptr; }
int sgpu_read_batch_stats(const sgpu_read_batch *b, uint64_t *out3) {
    if (!b || !out3) return SGPU_EINVAL;
    out3[0] = b->records; out3[1] = b->trimmed; out3[2] = b->dropped;
    return SGPU_OK;
}
const char *sgpu_read_batch_error(const sgpu_read_batch *b) { return b ? b->err.c_str() : "null batch"; }
void sgpu_read_batch_free(sgpu_read_batch *b) { delete b; }

int sgpu_read_batch_write_seqfile(const sgpu_read_batch *b, const char *prefix) {
    if (!b || !prefix) return SGPU_EINVAL;
    FILE *fs = fopen((std::string(prefix) + ".seq").c_str(), "wb");
    FILE *fo = fopen((std::string(prefix) + ".off").c_str(), "wb");
    if (!fs || !fo) { if (fs) fclose(fs); if (fo) fclose(fo); return SGPU_EIO; }
    uint64_t stat[3] = {(uint64_t)b->lens.size(), 0, 0};
    for (uint32_t l : b->lens) { if (l > stat[1]) stat[1] = l; stat[2] += l; }
    bool ok = fwrite(stat, 8, 3, fs) == 3;
    uint64_t pos = 24;
    const uint8_t tail[12] = {0, 0, 0, 0, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff};    // offsets 0,0 ; tag -1 (single_read.hpp:315)
    for (size_t r = 0; ok && r < b->lens.size(); ++r) {
        if (r % 100 == 0) ok = fwrite(&pos, 8, 1, fo) == 1;                               // BinaryWriter::CHUNK = 100
        const uint64_t size = b->lens[r];
        const size_t nw = (size_t)((size + 31) / 32);
        ok = ok && fwrite(&size, 8, 1, fs) == 1 && fwrite(b->words.data() + b->offs[r], 8, nw, fs) == nw && fwrite(tail, 1, 12, fs) == 12;
        pos += 8 + 8 * nw + 12;
    }
    ok = (fclose(fs) == 0) && ok;
    ok = (fclose(fo) == 0) && ok;
    return ok ? SGPU_OK : SGPU_EIO;
}

}  // extern "C"
