// scan.cu -- device-wide exclusive prefix sums (reduce-then-scan, recursive on the block sums) and the
// read-set upload. Small plumbing used by every phase (partition offsets, compaction offsets, MPHF ranks).
#include "sgpu_internal.h"

namespace sg {

static const int kScanThreads = 256;
static const int kScanItems = 8;
static const int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan_256(uint64_t v, uint64_t *total, uint64_t *smem /*>=8*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < (kScanThreads / 32) ? smem[lane] : 0;
        uint64_t winc = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < (kScanThreads / 32)) smem[lane] = winc - w;   // exclusive warp base
        if (lane == (kScanThreads / 32) - 1) smem[8] = winc;     // block total
    }
    __syncthreads();
    uint64_t r = smem[warp] + inc - v;
    *total = smem[8];
    __syncthreads();
    return r;
}

template <class TIn>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_k(const TIn *__restrict__ in, uint64_t *__restrict__ bsum, size_t n) {
    __shared__ uint64_t sm[9];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) s += (uint64_t)in[base + i];
    uint64_t tot;
    block_exclusive_scan_256(s, &tot, sm);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

template <class TIn>
__global__ void __launch_bounds__(kScanThreads) scan_apply_k(const TIn *__restrict__ in, uint64_t *__restrict__ out,
                                                            const uint64_t *__restrict__ bbase, size_t n) {
    __shared__ uint64_t sm[9];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint64_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? (uint64_t)in[base + i] : 0;
        s += v[i];
    }
    uint64_t tot;
    uint64_t ex = block_exclusive_scan_256(s, &tot, sm) + (bbase ? bbase[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

template <class TIn>
static void scan_impl(Ctx *ctx, const TIn *in, uint64_t *out, size_t n) {
    if (n == 0) return;
    size_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 1) {
        scan_apply_k<TIn><<<1, kScanThreads, 0, ctx->stream>>>(in, out, nullptr, n);
        ctx->launches++;
        SG_CUDA(cudaGetLastError());
        return;
    }
    DArr<uint64_t> bsum(ctx, nb);
    scan_reduce_k<TIn><<<(unsigned)nb, kScanThreads, 0, ctx->stream>>>(in, bsum.p, n);
    ctx->launches++;
    SG_CUDA(cudaGetLastError());
    scan_impl<uint64_t>(ctx, bsum.p, bsum.p, nb);
    scan_apply_k<TIn><<<(unsigned)nb, kScanThreads, 0, ctx->stream>>>(in, out, bsum.p, n);
    ctx->launches++;
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(ctx->stream));   // bsum is freed on return
}

void exclusive_scan_u64(Ctx *ctx, const uint64_t *in, uint64_t *out, size_t n) { scan_impl<uint64_t>(ctx, in, out, n); }
void exclusive_scan_u32_to_u64(Ctx *ctx, const uint32_t *in, uint64_t *out, size_t n) { scan_impl<uint32_t>(ctx, in, out, n); }

// Upload host-appended reads (once); adopted device reads are used in place.
void ensure_reads_on_device(Ctx *ctx) {
    if (!ctx->staged_dirty) return;
    size_t nw = ctx->h_words.size();
    ctx->r_words.alloc(ctx, nw + 4, true);                 // +padding: kmer_window may touch one word past a read
    ctx->r_offs.alloc(ctx, ctx->h_offs.size(), true);
    ctx->r_lens.alloc(ctx, ctx->h_lens.size(), true);
    SG_CUDA(cudaMemsetAsync(ctx->r_words.p + nw, 0, 4 * sizeof(uint64_t), ctx->stream));
    if (nw) SG_CUDA(cudaMemcpyAsync(ctx->r_words.p, ctx->h_words.data(), nw * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->h_offs.empty()) {
        SG_CUDA(cudaMemcpyAsync(ctx->r_offs.p, ctx->h_offs.data(), ctx->h_offs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        SG_CUDA(cudaMemcpyAsync(ctx->r_lens.p, ctx->h_lens.data(), ctx->h_lens.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->d_words = ctx->r_words.p; ctx->d_offs = ctx->r_offs.p; ctx->d_lens = ctx->r_lens.p;
    ctx->n_reads = (int64_t)ctx->h_lens.size();
    ctx->n_words = nw;
    ctx->staged_dirty = false;
}

}  // namespace sg
