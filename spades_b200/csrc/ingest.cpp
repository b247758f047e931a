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
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#include <zlib.h>

#include <string>
#include <thread>
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

struct GzReader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    std::string io_error;          // set when zlib reports a read / decompression error (truncated or corrupt .gz)
    std::string carry;
    bool open(const char *path) {
        f = gzopen(path, "rb");       // transparently reads plain files too
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(4 << 20);
        return true;
    }
    ~GzReader() { if (f) gzclose(f); }
    bool refill() {
        if (eof) return false;
        const int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n <= 0) {
            // gzread() <= 0 is only a clean end of input when zlib agrees: a truncated or corrupt stream (Z_BUF_ERROR / Z_DATA_ERROR)
            // must not yield a silently smaller read set
            int zerr = Z_OK;
            const char *msg = gzerror(f, &zerr);
            if (n < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) io_error = std::string("read error: ") + (msg && *msg ? msg : "gzread failed");
            eof = true;
            return false;
        }
        pos = 0; end = (size_t)n;
        return true;
    }
    // next line without its '\n'; the view is valid until the next call. false at end of input.
    bool getline(const char *&p, size_t &n) {
        carry.clear();
        bool have_carry = false;
        for (;;) {
            if (pos == end && !refill()) {
                if (!have_carry) return false;
                p = carry.data(); n = carry.size();
                return true;
            }
            const char *nl = (const char *)memchr(buf.data() + pos, '\n', end - pos);
            if (nl) {
                const size_t len = (size_t)(nl - (buf.data() + pos));
                if (have_carry) { carry.append(buf.data() + pos, len); p = carry.data(); n = carry.size(); }
                else { p = buf.data() + pos; n = len; }
                pos += len + 1;
                return true;
            }
            carry.append(buf.data() + pos, end - pos);
            have_carry = true;
            pos = end;
        }
    }
};

static const int8_t kCode[256] = {
#define X16 -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1
    X16, X16, X16, X16,
    -1, 0, -1, 1, -1, -1, -1, 2, -1, -1, -1, -1, -1, -1, -1, -1,  -1, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,     // @ A..O, P..
    -1, 0, -1, 1, -1, -1, -1, 2, -1, -1, -1, -1, -1, -1, -1, -1,  -1, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,     // ` a..o, p..
    X16, X16, X16, X16, X16, X16, X16, X16
#undef X16
};

void add_read(sgpu_read_batch *b, const std::string &seq, bool longest_valid) {
    ++b->records;
    const unsigned char *s = (const unsigned char *)seq.data();
    const size_t sz = seq.size();
    size_t from = 0, to = sz;
    if (longest_valid) {
        // first longest run of nucleotides (LongestValidCoords)
        size_t best_len = 0, best_pos = 0, run = 0;
        for (size_t i = 0; i < sz; ++i) {
            if (kCode[s[i]] >= 0) ++run;
            else {
                if (run > best_len) { best_len = run; best_pos = i - run; }
                run = 0;
            }
        }
        if (run > best_len) { best_len = run; best_pos = sz - run; }
        from = best_pos; to = best_pos + best_len;
        if (best_len < sz) ++b->trimmed;
    } else {
        for (size_t i = 0; i < sz; ++i)
            if (kCode[s[i]] < 0) { from = to = 0; break; }      // without N handling an invalid read contributes nothing
    }
    const size_t len = to - from;
    if (len == 0) { ++b->dropped; return; }
    const size_t nw = (len + 31) / 32;
    const size_t w0 = b->words.size();
    b->words.resize(w0 + nw, 0);
    uint64_t *w = b->words.data() + w0;
    const unsigned char *q = s + from;
    for (size_t j = 0; j < nw; ++j) {
        const size_t m = len - 32 * j < 32 ? len - 32 * j : 32;
        uint64_t x = 0;
        for (size_t i = 0; i < m; ++i) x |= (uint64_t)kCode[q[32 * j + i]] << (2 * i);
        w[j] = x;
    }
    b->offs.push_back((uint64_t)w0);
    b->lens.push_back((uint32_t)len);
}

// kseq_read as vendored by the reference (ext/include/kseq/kseq.h:171-213), line by line:
//  * a record starts at the next '>' or '@' found ANYWHERE in the stream (:177) -- after a FASTQ record; after a FASTA record the
//    marker line that ended the sequence is the header (:196);
//  * sequence: every following line that does not start with '>', '@' or '+' is appended whole -- blanks included, they are
//    just invalid characters for LongestValid -- minus ONE trailing '\r' when the sequence so far is longer than 1 (:136,189-195);
//    empty lines are skipped;
//  * quality ('+'): whole lines are appended until the quality is at least as long as the sequence (at least one line is read),
//    and any other length is an error (:206-209).
// Line sources. Both give lines without their '\n', and the offset at which the line starts in the (uncompressed) input.
struct MemLines {
    const char *d; size_t pos, end;
    bool getline(const char *&p, size_t &n, size_t &at) {
        if (pos >= end) return false;
        const char *nl = (const char *)memchr(d + pos, '\n', end - pos);
        const size_t len = nl ? (size_t)(nl - (d + pos)) : end - pos;
        p = d + pos; n = len; at = pos;
        pos += len + (nl ? 1 : 0);
        return true;
    }
};
struct GzLines {
    GzReader r;
    size_t consumed = 0;
    bool getline(const char *&p, size_t &n, size_t &at) {
        if (!r.getline(p, n)) return false;
        at = consumed; consumed += n + 1;
        return true;
    }
};

// The state machine over a line source. Starts either looking for a marker (scan) or ON a header line whose first character
// is the marker (`pending` = a line already fetched by the caller is not supported: a worker that starts on a header simply
// positions the source at the start of that line and passes scan = true -- the first marker found is then that line's first
// character). Stops before starting a record whose header line begins at or after `stop_at`; *next = where that line begins
// (or the input size at end of input), *next_is_scan = whether the following record would have been looked for by scanning.
template <class Lines>
bool parse_lines(Lines &in, bool longest_valid, size_t stop_at, sgpu_read_batch *b, size_t *next, size_t input_size) {
    const char *p;
    size_t n, at = 0;
    std::string seq;
    bool scan = true;            // true: look for '>' / '@' anywhere; false: the line in (p, n) is the header
    for (;;) {
        if (scan) {
            bool have = false;
            while (in.getline(p, n, at)) {
                if (memchr(p, '>', n) || memchr(p, '@', n)) { have = true; break; }
            }
            if (!have) { *next = input_size; return true; }
        }
        if (at >= stop_at) { *next = at; return true; }
        // (p, n) is a header line: name / comment are not needed
        seq.clear();
        int stop = -1;           // the character that ended the sequence, -1 = end of input
        while (in.getline(p, n, at)) {
            if (n == 0) continue;
            if (p[0] == '>' || p[0] == '@' || p[0] == '+') { stop = p[0]; break; }
            seq.append(p, n);
            if (seq.size() > 1 && seq.back() == '\r') seq.pop_back();
        }
        add_read(b, seq, longest_valid);
        if (stop < 0) { *next = input_size; return true; }
        if (stop == '+') {
            size_t q = 0;
            do {
                if (!in.getline(p, n, at)) { b->err = "truncated quality string"; return false; }
                q += n;
                if (q > 1 && n && p[n - 1] == '\r') --q;
            } while (q < seq.size());
            if (q != seq.size()) { b->err = "quality string is of a different length than the sequence"; return false; }
            scan = true;
        } else {
            scan = false;
        }
    }
}

void append_batch(sgpu_read_batch *dst, const sgpu_read_batch &src) {
    const uint64_t w0 = dst->words.size();
    dst->words.insert(dst->words.end(), src.words.begin(), src.words.end());
    dst->lens.insert(dst->lens.end(), src.lens.begin(), src.lens.end());
    dst->offs.reserve(dst->offs.size() + src.offs.size());
    for (uint64_t o : src.offs) dst->offs.push_back(o + w0);
    dst->records += src.records; dst->trimmed += src.trimmed; dst->dropped += src.dropped;
}

// A plain (uncompressed) file in parallel, exactly: the file is cut at guessed record starts (a line that begins with '@' and
// looks like the first line of a four-line FASTQ record, or any line that begins with '>'), every worker runs the SAME state
// machine from its cut to the next one, and the pieces are accepted only if every worker stopped exactly on the next worker's
// cut -- i.e. if the sequential parser would have been in "find the next record" position there. Anything else (multi-line
// FASTQ whose cuts were guessed wrong, markers in the middle of junk lines, errors) falls back to the sequential parse.
bool parse_plain_parallel(const char *d, size_t size, bool longest_valid, int nthreads, sgpu_read_batch *b) {
    const size_t chunk = size / (size_t)nthreads;
    std::vector<size_t> cut((size_t)nthreads + 1, size);
    cut[0] = 0;
    for (int i = 1; i < nthreads; ++i) {
        // first line start at or after i*chunk ...
        size_t pos = (size_t)i * chunk;
        const char *nl = (const char *)memchr(d + pos, '\n', size - pos);
        pos = nl ? (size_t)(nl - d) + 1 : size;
        // ... that begins a record
        size_t found = size;
        for (int tries = 0; pos < size && tries < 64; ++tries) {
            MemLines ml{d, pos, size};
            const char *p[5]; size_t n[5], at[5];
            int got = 0;
            while (got < 5 && ml.getline(p[got], n[got], at[got])) ++got;
            if (got == 0) break;
            if (n[0] && p[0][0] == '>') { found = pos; break; }
            if (got >= 4 && n[0] && p[0][0] == '@' && n[2] && p[2][0] == '+' && n[1] == n[3] && n[1] && p[1][0] != '@' && p[1][0] != '>' && p[1][0] != '+' &&
                (got == 4 || (n[4] && p[4][0] == '@'))) { found = pos; break; }
            pos = got >= 2 ? at[1] : size;              // next line
        }
        cut[(size_t)i] = found;
    }
    for (int i = 1; i <= nthreads; ++i) if (cut[(size_t)i] < cut[(size_t)i - 1]) cut[(size_t)i] = cut[(size_t)i - 1];
    std::vector<sgpu_read_batch> part((size_t)nthreads);
    std::vector<size_t> next((size_t)nthreads, 0);
    std::vector<char> ok((size_t)nthreads, 0);
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i)
        th.emplace_back([&, i]() {
            if (cut[(size_t)i] >= cut[(size_t)i + 1] && i > 0) { next[(size_t)i] = cut[(size_t)i]; ok[(size_t)i] = 1; return; }     // empty piece
            MemLines ml{d, cut[(size_t)i], size};
            ok[(size_t)i] = parse_lines(ml, longest_valid, cut[(size_t)i + 1], &part[(size_t)i], &next[(size_t)i], size) ? 1 : 0;
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < nthreads; ++i) {
        if (!ok[(size_t)i]) return false;
        if (next[(size_t)i] != cut[(size_t)i + 1]) return false;          // the sequential parser would not have been at the next cut
    }
    for (int i = 0; i < nthreads; ++i) append_batch(b, part[(size_t)i]);
    return true;
}

bool parse_fastx(const char *path, bool longest_valid, int nthreads, sgpu_read_batch *b) {
    // gzip (or unreadable through mmap): one sequential stream
    bool gz = false;
    size_t size = 0;
    {
        FILE *f = fopen(path, "rb");
        if (!f) { b->err = std::string("cannot open ") + path; return false; }
        unsigned char m[2] = {0, 0};
        const size_t got = fread(m, 1, 2, f);
        gz = got == 2 && m[0] == 0x1f && m[1] == 0x8b;
        fseek(f, 0, SEEK_END);
        size = (size_t)ftell(f);
        fclose(f);
    }
    if (nthreads <= 0) {
        nthreads = (int)std::thread::hardware_concurrency();
        if (const char *e = getenv("SGPU_INGEST_THREADS")) nthreads = atoi(e);
        if (nthreads > 64) nthreads = 64;
    }
    if (!gz && nthreads > 1 && size >= ((size_t)nthreads << 16)) {
        const int fd = open(path, O_RDONLY);
        if (fd >= 0) {
            void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            close(fd);
            if (m != MAP_FAILED) {
                sgpu_read_batch tmp;
                const bool ok = parse_plain_parallel((const char *)m, size, longest_valid, nthreads, &tmp);
                munmap(m, size);
                if (ok) { append_batch(b, tmp); return true; }
                // fall through: the exact sequential parse decides (and reports errors)
            }
        }
    }
    GzLines in;
    if (!in.r.open(path)) { b->err = std::string("cannot open ") + path; return false; }
    size_t next = 0;
    const bool ok = parse_lines(in, longest_valid, (size_t)-1, b, &next, (size_t)-1);
    // a truncated / corrupt gzip stream ends the line source early: report it (SGPU_EIO) instead of returning a partial read set.
    // (The reference's FastaFastqGzParser treats kseq_read() == -2, a quality string of the wrong length, as END OF STREAM and
    // keeps what it has; this library is stricter on purpose and fails with "quality string is of a different length".)
    if (!in.r.io_error.empty()) { b->err = in.r.io_error + " (" + path + ")"; return false; }
    return ok;
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

int sgpu_fastx_parse_threads(const char *path, int longest_valid, int nthreads, sgpu_read_batch **out) {
    if (!path || !out) return SGPU_EINVAL;
    sgpu_read_batch *b = new sgpu_read_batch();
    *out = b;                                   // returned even on failure so that sgpu_read_batch_error() can be read
    return parse_fastx(path, longest_valid != 0, nthreads, b) ? SGPU_OK : SGPU_EIO;
}
int64_t sgpu_text_index_fastx(const char *text, uint64_t text_bytes, uint64_t *seq_off, uint32_t *seq_len, int64_t max_reads) {
    if (!text && text_bytes) return -1;
    if (text_bytes == 0) return 0;
    const int stride = text[0] == '>' ? 2 : (text[0] == '@' ? 4 : 0);
    if (!stride) return -1;
    int64_t n = 0;
    uint64_t pos = 0;
    int line = 0;
    uint32_t cur_len = 0;
    while (pos < text_bytes) {
        const char *nl = (const char *)memchr(text + pos, '\n', text_bytes - pos);
        uint64_t end = nl ? (uint64_t)(nl - text) : text_bytes;
        uint64_t len = end - pos;
        if (len && text[end - 1] == '\r') --len;
        const int which = line % stride;
        if (which == 0) {
            if (len == 0 && !nl) break;                                  // trailing empty line
            if (len == 0 || text[pos] != (stride == 2 ? '>' : '@')) return -1;
        } else if (which == 1) {
            if (len > 0xffffffffull) return -1;
            if (len && (text[pos] == '>' || text[pos] == '@' || text[pos] == '+')) return -1;      // would start another record / quality in kseq
            if (n >= max_reads) return -2;
            seq_off[n] = pos; seq_len[n] = (uint32_t)len; cur_len = (uint32_t)len;
            ++n;
        } else if (which == 2) {
            if (len == 0 || text[pos] != '+') return -1;
        } else {
            if (len != cur_len) return -1;
        }
        ++line;
        pos = nl ? end + 1 : text_bytes;
    }
    if (line % stride != 0) return -1;                                    // truncated record
    return n;
}
int sgpu_fastx_parse(const char *path, int longest_valid, sgpu_read_batch **out) { return sgpu_fastx_parse_threads(path, longest_valid, 0, out); }
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
