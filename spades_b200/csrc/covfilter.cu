// Coverage pre-filter of the construction stage: the pipeline's CoverageFilter phase, stages/construction.cpp:167-198.
//
//   reference                                                         here
//   ----------------------------------------------------------------  -------------------------------------------------------
//   rolling_hash::SymmetricCyclicHash<NDNASeqHash>(k+1)               cyc_* below: fwd/rvs rolled per window, value = fwd + rvs
//     adt/cyclichash.hpp:187-259, character hashes :24-27
//   EstimateCardinalityUpperBound -> hll::hll<24>                     cov_hll_k: atomicMax into 2^24 registers; the estimate is
//     kmer_index/kmer_counting.hpp:215-249, adt/hll.hpp:34-68           finished on the host with the reference's own double sum
//   qf::cqf(maxn): key = hash & (2^(qbits+8) - 1), exact counts       an open-addressing table of (key, count) in HBM; counts
//     adt/cqf.hpp:28-37,57-63; FillCoverageHistogram + CQFProcessor     stop at the threshold like CQFProcessor (:107-120)
//     kmer_counting.hpp:96-121,251-282
//   CovFilteringWrap: median multiplicity of a read's windows >= thr  cov_filter_k: per read, windows below the threshold <= w/2
//     io/reads/coverage_filtering_read_wrapper.hpp:37-124
//
// The counting quotient filter's slot layout is not reproduced: the reference only ever asks it "count(key) >= threshold", and
// a CQF with key_bits = qbits + 8 stores every key exactly (ext/src/gqf/gqf.c:1430-1477: range = 2^key_bits), so an exact
// (key -> count) table answers identically. The hash is symmetric (a window and its reverse complement hash alike), so the
// reference's "windows of reads and their reverse complements that are IsMinimal" is "every window of the forward reads", with
// self-reverse-complementary windows counted twice (they pass the filter on both strands; SURVEY 0.6 has the same doubling).
//
// One thread walks one read (the filter's verdict is per read, and 10^8 reads are parallelism enough); the three passes re-roll
// the hash instead of materialising 8 bytes per window.
#include "sgpu_internal.h"
#include <cmath>

namespace sg {

namespace {

__device__ __forceinline__ uint64_t cyc_h(int c) {
    // NDNASeqHash(seed 0), cyclichash.hpp:24-27,51
    const uint64_t a = (c & 1) ? 0x3193c18562a02b4cULL : 0x3c8bfbb395c60474ULL;
    const uint64_t b = (c & 1) ? 0x295549f54be24456ULL : 0x20323ed082572324ULL;
    return (c & 2) ? b : a;
}
__device__ __forceinline__ uint64_t rol64d(uint64_t x, unsigned s) { s &= 63; return s ? (x << s) | (x >> (64 - s)) : x; }
__device__ __forceinline__ int base_at(const uint64_t *seq, int i) { return (int)((seq[i >> 5] >> ((i & 31) << 1)) & 3); }

struct CycHash {
    uint64_t fwd, rvs;
    __device__ __forceinline__ uint64_t value() const { return fwd + rvs; }
};
// SymmetricCyclicHash::operator() on the window at base 0 (cyclichash.hpp:231-241)
__device__ __forceinline__ CycHash cyc_init(const uint64_t *seq, int K) {
    CycHash h{0, 0};
    for (int i = 0; i < K; ++i) h.fwd = rol64d(h.fwd, 1) ^ cyc_h(base_at(seq, i));
    for (int i = 0; i < K; ++i) h.rvs = rol64d(h.rvs, 1) ^ cyc_h(3 - base_at(seq, K - 1 - i));
    return h;
}
// hash_update (cyclichash.hpp:250-256): drop `out`, append `in`
__device__ __forceinline__ void cyc_roll(CycHash &h, int out, int in, int K) {
    h.fwd = rol64d(h.fwd, 1) ^ rol64d(cyc_h(out), (unsigned)K) ^ cyc_h(in);
    h.rvs = rol64d(h.rvs, 63) ^ rol64d(cyc_h(3 - out), 63) ^ rol64d(cyc_h(3 - in), (unsigned)(K - 1));
}
// a window equal to its own reverse complement (even K only)
__device__ __forceinline__ bool window_self_rc(const uint64_t *seq, int j, int K) {
    for (int i = 0; 2 * i < K; ++i)
        if (base_at(seq, j + i) != 3 - base_at(seq, j + K - 1 - i)) return false;
    return true;
}

// ---- pass 1: HyperLogLog registers (hll.hpp:34-41) --------------------------------------------------------------------------
__global__ void cov_hll_k(const uint64_t *__restrict__ words, const uint64_t *__restrict__ offs, const uint32_t *__restrict__ lens, int64_t n, int K,
                          uint32_t *__restrict__ reg) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int L = (int)lens[r];
    if (L < K) return;
    const uint64_t *seq = words + offs[r];
    CycHash h = cyc_init(seq, K);
    for (int j = 0;; ++j) {
        const uint64_t d = h.value();
        const uint32_t id = (uint32_t)(d >> 40);
        const uint64_t low = d & ((1ull << 40) - 1);
        const uint32_t rho = (uint32_t)((low == 0 ? 64 : __clzll((long long)low)) - 24 + 1);
        if (reg[id] < rho) atomicMax(&reg[id], rho);
        if (j + K >= L) break;
        cyc_roll(h, base_at(seq, j), base_at(seq, j + K), K);
    }
}

// ---- the (key -> count) table --------------------------------------------------------------------------------------------------
// entry = (key + 1) << 16 | count; 0 = empty. key < 2^47.
struct CovTable {
    unsigned long long *e;
    uint64_t cap;
    uint64_t key_mask;
    unsigned *overflow;         // set when a probe sequence visited every slot (cannot happen while the HLL bound holds; checked on the host)
    __device__ __forceinline__ uint64_t slot_of(uint64_t key) const { return __umul64hi(key * 0x9E3779B97F4A7C15ULL, cap); }
    // CQFProcessor::ProcessKmer: nothing once the count has reached the threshold
    __device__ __forceinline__ void add(uint64_t key, unsigned thr) const {
        const unsigned long long tag = (key + 1) << 16;
        uint64_t s = slot_of(key);
        for (uint64_t probes = 0;; ++probes) {
            if (probes > cap) { *overflow = 1u; return; }
            unsigned long long cur = e[s];
            if (cur == 0) {
                const unsigned long long old = atomicCAS(&e[s], 0ull, tag | 1ull);
                if (old == 0) return;
                cur = old;
            }
            if ((cur & ~0xffffull) == tag) {
                while ((cur & 0xffffull) < thr) {
                    const unsigned long long old = atomicCAS(&e[s], cur, cur + 1);
                    if (old == cur) return;
                    cur = old;
                }
                return;
            }
            if (++s == cap) s = 0;
        }
    }
    __device__ __forceinline__ unsigned count(uint64_t key) const {
        const unsigned long long tag = (key + 1) << 16;
        uint64_t s = slot_of(key);
        for (uint64_t probes = 0; probes <= cap; ++probes) {
            const unsigned long long cur = e[s];
            if (cur == 0) return 0;
            if ((cur & ~0xffffull) == tag) return (unsigned)(cur & 0xffffull);
            if (++s == cap) s = 0;
        }
        return 0;
    }
};

// ---- pass 2: counts up to the threshold ------------------------------------------------------------------------------------------
__global__ void cov_fill_k(const uint64_t *__restrict__ words, const uint64_t *__restrict__ offs, const uint32_t *__restrict__ lens, int64_t n, int K,
                           CovTable t, unsigned thr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int L = (int)lens[r];
    if (L < K) return;
    const uint64_t *seq = words + offs[r];
    CycHash h = cyc_init(seq, K);
    for (int j = 0;; ++j) {
        const uint64_t key = h.value() & t.key_mask;
        t.add(key, thr);
        if ((K & 1) == 0 && h.fwd == h.rvs && window_self_rc(seq, j, K)) t.add(key, thr);
        if (j + K >= L) break;
        cyc_roll(h, base_at(seq, j), base_at(seq, j + K), K);
    }
}

// ---- pass 3: the verdict per read (coverage_filtering_read_wrapper.hpp:37-72) --------------------------------------------------------
__global__ void cov_filter_k(const uint64_t *__restrict__ words, const uint64_t *__restrict__ offs, const uint32_t *__restrict__ lens, int64_t n, int K,
                             CovTable t, unsigned thr, uint8_t *__restrict__ keep, uint32_t *__restrict__ keep_words) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int L = (int)lens[r];
    bool k = false;
    if (L < K) {
        k = thr == 0;                                          // CountMedianMlt returns 0 for a read shorter than K
    } else {
        const uint64_t *seq = words + offs[r];
        CycHash h = cyc_init(seq, K);
        uint32_t below = 0;
        for (int j = 0;; ++j) {
            below += t.count(h.value() & t.key_mask) < thr;
            if (j + K >= L) break;
            cyc_roll(h, base_at(seq, j), base_at(seq, j + K), K);
        }
        k = below <= (uint32_t)(L - K + 1) / 2;                // element size/2 of the sorted multiplicities >= threshold
    }
    keep[r] = k ? 1 : 0;
    keep_words[r] = k ? (uint32_t)((L + 31) >> 5) : 0u;
}

__global__ void cov_keep_count_k(const uint8_t *__restrict__ keep, int64_t n, uint32_t *__restrict__ flag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) flag[r] = keep[r];
}
// survivors keep their order: read r moves to position new_idx[r], its words to new_off[r]
__global__ void cov_compact_k(const uint64_t *__restrict__ words, const uint64_t *__restrict__ offs, const uint32_t *__restrict__ lens,
                              const uint8_t *__restrict__ keep, const uint64_t *__restrict__ new_idx, const uint64_t *__restrict__ new_off, int64_t n,
                              uint64_t *__restrict__ out_words, uint64_t *__restrict__ out_offs, uint32_t *__restrict__ out_lens) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !keep[r]) return;
    const uint32_t L = lens[r];
    const uint64_t d = new_off[r];
    out_offs[new_idx[r]] = d;
    out_lens[new_idx[r]] = L;
    const uint64_t *s = words + offs[r];
    for (uint32_t i = 0; i < ((L + 31) >> 5); ++i) out_words[d + i] = s[i];
}
__global__ void cov_distinct_k(const unsigned long long *__restrict__ e, uint64_t cap, unsigned long long *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned c = 0;
    for (; i < cap; i += (uint64_t)gridDim.x * blockDim.x) c += e[i] != 0;
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

}  // namespace

// hll<24>::cardinality / upper_bound_cardinality (adt/hll.hpp:50-68): same operations in the same order
static double hll_upper_bound(const std::vector<uint32_t> &reg) {
    const uint64_t m = 1ull << 24;
    const double alpha = 0.7213 / (1.0 + 1.079 / (double)m);
    double res = alpha * (double)m * (double)m;
    double E = 0.0;
    uint64_t zeros = 0;
    for (uint64_t i = 0; i < m; ++i) { E += exp2(-(double)reg[i]); zeros += reg[i] == 0; }
    res /= E;
    if (res <= 5.0 * (double)m / 2 && zeros > 0) res = (double)m * (std::log((double)m) - std::log((double)zeros));
    return 1.1 * res;
}

void cov_filter(Ctx *ctx, int K, unsigned thr, int apply, uint8_t *keep_out, uint64_t *stats) {
    SG_CHECK(K >= 1 && K <= 128, 2, "K must be in [1,128]");
    SG_CHECK(thr <= 60000u, 2, "coverage threshold must be at most 60000");
    ensure_reads_on_device(ctx);
    cudaStream_t st = ctx->stream;
    const int64_t n = ctx->n_reads;
    const int T = 128;
    // 1. cardinality upper bound
    std::vector<uint32_t> h_reg((size_t)1 << 24);
    {
        DArr<uint32_t> reg(ctx, (size_t)1 << 24);
        SG_CUDA(cudaMemsetAsync(reg.p, 0, reg.bytes(), st));
        if (n) { cov_hll_k<<<div_up(n, T), T, 0, st>>>(ctx->d_words, ctx->d_offs, ctx->d_lens, n, K, reg.p); ctx->launches++; }
        SG_CUDA(cudaMemcpyAsync(h_reg.data(), reg.p, reg.bytes(), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
    }
    const size_t maxn = (size_t)hll_upper_bound(h_reg);
    // 2. qf::cqf(maxn) geometry (cqf.hpp:28-37)
    const unsigned lg = maxn > 1 ? (unsigned)std::ceil(std::log2((double)maxn)) : 0u;     // (no reads: the reference's log2(0) is undefined)
    const unsigned qbits = std::max(7u, lg) + 1;
    const unsigned key_bits = qbits + 8;
    SG_CHECK(key_bits <= 47, 2, "coverage filter: more than 2^38 distinct k-mers estimated");
    CovTable t;
    t.cap = std::max<uint64_t>(1024, (uint64_t)maxn + (uint64_t)maxn / 2);
    t.key_mask = (1ull << key_bits) - 1;
    DArr<unsigned long long> table(ctx, t.cap);
    t.e = table.p;
    SG_CUDA(cudaMemsetAsync(table.p, 0, table.bytes(), st));
    DArr<uint8_t> keep(ctx, (size_t)n + 1);
    DArr<uint32_t> keep_words(ctx, (size_t)n + 1), flag(ctx, (size_t)n + 1);
    DArr<unsigned long long> d_cnt(ctx, 1);
    DArr<unsigned> d_ovf(ctx, 1);
    SG_CUDA(cudaMemsetAsync(d_cnt.p, 0, 8, st));
    SG_CUDA(cudaMemsetAsync(d_ovf.p, 0, 4, st));
    t.overflow = d_ovf.p;
    if (n) {
        cov_fill_k<<<div_up(n, T), T, 0, st>>>(ctx->d_words, ctx->d_offs, ctx->d_lens, n, K, t, thr);
        cov_filter_k<<<div_up(n, T), T, 0, st>>>(ctx->d_words, ctx->d_offs, ctx->d_lens, n, K, t, thr, keep.p, keep_words.p);
        cov_keep_count_k<<<div_up(n, 256), 256, 0, st>>>(keep.p, n, flag.p);
        ctx->launches += 3;
    }
    cov_distinct_k<<<ctx->num_sms * 4, 256, 0, st>>>(table.p, t.cap, d_cnt.p);
    ctx->launches++;
    SG_CUDA(cudaMemsetAsync(keep_words.p + n, 0, 4, st));
    SG_CUDA(cudaMemsetAsync(flag.p + n, 0, 4, st));
    DArr<uint64_t> new_off(ctx, (size_t)n + 1), new_idx(ctx, (size_t)n + 1);
    exclusive_scan_u32_to_u64(ctx, keep_words.p, new_off.p, (size_t)n + 1);
    exclusive_scan_u32_to_u64(ctx, flag.p, new_idx.p, (size_t)n + 1);
    uint64_t kept = 0, kept_words = 0;
    unsigned long long distinct = 0;
    SG_CUDA(cudaMemcpyAsync(&kept, new_idx.p + n, 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(&kept_words, new_off.p + n, 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(&distinct, d_cnt.p, 8, cudaMemcpyDeviceToHost, st));
    unsigned overflow = 0;
    SG_CUDA(cudaMemcpyAsync(&overflow, d_ovf.p, 4, cudaMemcpyDeviceToHost, st));
    if (keep_out && n) SG_CUDA(cudaMemcpyAsync(keep_out, keep.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(st));
    SG_CHECK(!overflow, 6, "coverage filter: more distinct keys than the cardinality bound allows (table full)");
    if (stats) { stats[0] = maxn; stats[1] = key_bits; stats[2] = distinct; stats[3] = kept; }
    if (!apply) return;
    // 3. the surviving reads become the context's read set (what CovFilteringWrap does to the streams)
    DArr<uint64_t> nw(ctx, kept_words + 4, true), no(ctx, kept + 1, true);
    DArr<uint32_t> nl(ctx, kept + 1, true);
    SG_CUDA(cudaMemsetAsync(nw.p + kept_words, 0, 4 * 8, st));
    if (n) { cov_compact_k<<<div_up(n, 128), 128, 0, st>>>(ctx->d_words, ctx->d_offs, ctx->d_lens, keep.p, new_idx.p, new_off.p, n, nw.p, no.p, nl.p); ctx->launches++; }
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaStreamSynchronize(st));
    ctx->h_words.clear(); ctx->h_offs.clear(); ctx->h_lens.clear(); ctx->staged_dirty = false;
    ctx->r_words = std::move(nw); ctx->r_offs = std::move(no); ctx->r_lens = std::move(nl);
    ctx->d_words = ctx->r_words.p; ctx->d_offs = ctx->r_offs.p; ctx->d_lens = ctx->r_lens.p;
    ctx->n_reads = (int64_t)kept; ctx->n_words = kept_words;
}

}  // namespace sg
