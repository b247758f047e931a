// ingest_gpu.cu -- GPU-side read packing. The host only LOCATES the sequence lines of a FASTA/FASTQ text (a newline
// scan); trimming and packing run on the device:
//   lv_k    io::LongestValid per read (io/reads/longest_valid_wrapper.hpp:16-53: the FIRST longest run of ACGT/acgt; is_nucl +
//           dignucl, sequence/nucl.hpp:33-64), or -- without N handling -- a read with any other symbol contributes nothing
//   pack_k  2 bits per nucleotide, nucleotide i at bits 2(i%32) of word i/32, every read on a word boundary: the payload of the
//           reference's binary read records (Sequence::BinWrite, sequence/sequence.hpp:817-830)
// The result replaces the context's read set (what sgpu_reads_upload does with host-packed words). Reads without a valid base keep
// their slot with length 0 (the host parser drops them; neither contributes a k-mer).
#include "sgpu_internal.h"

namespace sg {

__constant__ int8_t c_code[256];

static void init_code_table() {
    static bool done = false;
    if (done) return;
    int8_t h[256];
    for (int i = 0; i < 256; ++i) h[i] = -1;
    h['A'] = h['a'] = 0; h['C'] = h['c'] = 1; h['G'] = h['g'] = 2; h['T'] = h['t'] = 3;
    SG_CUDA(cudaMemcpyToSymbol(c_code, h, 256));
    done = true;
}

__global__ void lv_k(const uint8_t *__restrict__ text, const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ seq_len, int64_t n,
                     int longest_valid, uint32_t *__restrict__ start, uint32_t *__restrict__ len, uint32_t *__restrict__ nwords) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t *s = text + seq_off[r];
    const uint32_t sz = seq_len[r];
    uint32_t best_len = 0, best_pos = 0, run = 0;
    bool all_valid = true;
    for (uint32_t i = 0; i < sz; ++i) {
        if (c_code[s[i]] >= 0) ++run;
        else {
            all_valid = false;
            if (run > best_len) { best_len = run; best_pos = i - run; }
            run = 0;
        }
    }
    if (run > best_len) { best_len = run; best_pos = sz - run; }
    if (!longest_valid && !all_valid) { best_len = 0; best_pos = 0; }
    start[r] = best_pos; len[r] = best_len; nwords[r] = (best_len + 31u) >> 5;
}

// one thread per output word
__global__ void pack_k(const uint8_t *__restrict__ text, const uint64_t *__restrict__ seq_off, const uint32_t *__restrict__ start,
                       const uint32_t *__restrict__ len, const uint64_t *__restrict__ woff, int64_t n, uint64_t total_words, uint64_t *__restrict__ words) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total_words) return;
    int64_t lo = 0, hi = n - 1;                               // last read whose first word is <= w (reads of length 0 own no word)
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (woff[mid] <= w) lo = mid; else hi = mid - 1;
    }
    const uint32_t j0 = (uint32_t)(w - woff[lo]) * 32u;
    const uint32_t L = len[lo];
    const uint8_t *s = text + seq_off[lo] + start[lo] + j0;
    const uint32_t m = L - j0 < 32u ? L - j0 : 32u;
    uint64_t v = 0;
    for (uint32_t i = 0; i < m; ++i) v |= (uint64_t)(uint8_t)c_code[s[i]] << (2 * i);
    words[w] = v;
}

void reads_pack_text(Ctx *ctx, const char *text, uint64_t text_bytes, const uint64_t *seq_off, const uint32_t *seq_len, int64_t n, int longest_valid) {
    cudaStream_t st = ctx->stream;
    init_code_table();
    for (int64_t r = 0; r < n; ++r) SG_CHECK(seq_off[r] + seq_len[r] <= text_bytes, 2, "sequence range outside the text buffer");
    ctx->h_words.clear(); ctx->h_offs.clear(); ctx->h_lens.clear(); ctx->staged_dirty = false;
    DArr<uint8_t> d_text(ctx, text_bytes + 8);
    DArr<uint64_t> d_off(ctx, (size_t)n + 1);
    DArr<uint32_t> d_slen(ctx, (size_t)n + 1), d_start(ctx, (size_t)n + 1), d_nw(ctx, (size_t)n + 1);
    if (text_bytes) SG_CUDA(cudaMemcpyAsync(d_text.p, text, text_bytes, cudaMemcpyHostToDevice, st));
    if (n) {
        SG_CUDA(cudaMemcpyAsync(d_off.p, seq_off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        SG_CUDA(cudaMemcpyAsync(d_slen.p, seq_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
    if (ctx->r_offs.n < (size_t)n + 1) ctx->r_offs.alloc(ctx, (size_t)n + 1, true);
    if (ctx->r_lens.n < (size_t)n + 1) ctx->r_lens.alloc(ctx, (size_t)n + 1, true);
    SG_CUDA(cudaMemsetAsync(d_nw.p + n, 0, 4, st));
    if (n) {
        lv_k<<<div_up(n, 128), 128, 0, st>>>(d_text.p, d_off.p, d_slen.p, n, longest_valid, d_start.p, ctx->r_lens.p, d_nw.p);
        ctx->launches++;
    }
    exclusive_scan_u32_to_u64(ctx, d_nw.p, ctx->r_offs.p, (size_t)n + 1);
    uint64_t total_words = 0;
    SG_CUDA(cudaMemcpyAsync(&total_words, ctx->r_offs.p + n, 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (ctx->r_words.n < total_words + 4) ctx->r_words.alloc(ctx, total_words + 4, true);
    SG_CUDA(cudaMemsetAsync(ctx->r_words.p + total_words, 0, 4 * 8, st));
    if (total_words) {
        pack_k<<<div_up((int64_t)total_words, 256), 256, 0, st>>>(d_text.p, d_off.p, d_start.p, ctx->r_lens.p, ctx->r_offs.p, n, total_words, ctx->r_words.p);
        ctx->launches++;
    }
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(st));
    ctx->d_words = ctx->r_words.p; ctx->d_offs = ctx->r_offs.p; ctx->d_lens = ctx->r_lens.p; ctx->n_reads = n; ctx->n_words = total_words;
}

// the packed read set back on the host (tests; io::BinaryWriter-style consumers)
void reads_download(Ctx *ctx, uint64_t *words, uint64_t *offs, uint32_t *lens) {
    ensure_reads_on_device(ctx);
    cudaStream_t st = ctx->stream;
    if (ctx->n_words && words) SG_CUDA(cudaMemcpyAsync(words, ctx->d_words, ctx->n_words * 8, cudaMemcpyDeviceToHost, st));
    if (ctx->n_reads && offs) SG_CUDA(cudaMemcpyAsync(offs, ctx->d_offs, (size_t)ctx->n_reads * 8, cudaMemcpyDeviceToHost, st));
    if (ctx->n_reads && lens) SG_CUDA(cudaMemcpyAsync(lens, ctx->d_lens, (size_t)ctx->n_reads * 4, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
}

}  // namespace sg
