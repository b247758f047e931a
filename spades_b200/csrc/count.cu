// count.cu -- k-mer extraction, partition, sort/unique/count. The B200 replacement for
//   KMerSortingSplitter::Split / DumpBuffers   (src/common/kmer_index/kmer_mph/kmer_splitter.hpp:56-179)
//   DeBruijnReadKMerSplitter / DeBruijnKMerKMerSplitter (…/kmer_splitters.hpp:28-207)
//   ParallelSortingSplitter (projects/spades_tools/kmercount.cpp:48-122)
//   KMerDiskCounter::Count / MergeKMers (…/kmer_index_builder.hpp:306-431)
// Output per bucket == the reference's kmers.<b> file: strictly increasing W-byte records, order of
// pdqsort_pod.h:725-734; multiplicities == CoverageHashMapBuilder's second pass (coverage_hash_map_builder.hpp:18-40).
//
// Pipeline (all in HBM, see DESIGN.md "count"):
//   A  level-A partition : records -> (bucket, top rA key bits) partitions. Two kernels over the source with
//      identical static work assignment: per-CTA histograms, then scatter with CTA-private cursors held in
//      shared memory (no global atomics; 16-byte stores merge in L2).
//   B  MSD refinement    : any segment longer than the local-sort capacity is split by its next r key bits by
//      ONE CTA (histogram + scatter through shared-memory cursors), ping-ponging between two buffers; repeated
//      until every segment fits or its key bits are exhausted (then all its records are equal).
//   C  local sort        : one CTA per segment: LSD radix sort in shared memory over the remaining key bits
//      (optimistic 32-bit window + verification, full range on failure), run-length unique/count, written
//      back in place; then a compaction copy into the dense bucket-major result.
#include <time.h>

#include <algorithm>
#include <type_traits>

#include "sgpu_internal.h"
#include "pair_mailbox.cuh"

namespace sg {

// ------------------------------------------------------------------------------------------------------------
// record sources
// ------------------------------------------------------------------------------------------------------------
struct ReadsSrc {
    const uint64_t *words;
    const uint64_t *offs;
    const uint32_t *lens;
    int64_t n;          // items = reads
    uint64_t nwords;    // length of `words`
    int K;
    int both;           // 1: every window emits fwd and rc (spades-kmercount), 0: canonical form once
    __device__ __forceinline__ uint32_t nrec(int64_t item) const {
        int L = (int)lens[item];
        uint32_t w = L >= K ? (uint32_t)(L - K + 1) : 0u;
        return both ? 2u * w : w;
    }
    template <int NW, typename Ptr>
    __device__ __forceinline__ Kmer<NW> get_at(Ptr s, uint32_t j) const {
        uint32_t pos = both ? (j >> 1) : j;
        Kmer<NW> f = kmer_window<NW>(s, (int64_t)pos, K);
        Kmer<NW> r = kmer_rc<NW>(f, K);
        if (both) return (j & 1) ? r : f;
        return kmer_is_minimal<NW>(f, r) ? f : r;
    }
    template <int NW>
    __device__ __forceinline__ Kmer<NW> get(int64_t item, uint32_t j) const { return get_at<NW>(words + offs[item], j); }
    // staging of a tile's packed reads in shared memory
    static constexpr bool kStage = true;
    __device__ __forceinline__ uint64_t first_word(int64_t item) const { return offs[item]; }
    __device__ __forceinline__ uint64_t end_word(int64_t item) const { return offs[item] + (((uint64_t)lens[item] + 31) >> 5); }
    __device__ __forceinline__ uint64_t stage_word(uint64_t w) const { return words[w]; }
};

// distinct (K+1)-mers -> their two K-mers in canonical form (DeBruijnKMerKMerSplitter with add_rc + IsMinimal filter)
template <int NWS>
struct KpomerSrc {
    const uint64_t *keys;   // NWS words per record
    int64_t n;
    int K;                  // target K
    __device__ __forceinline__ uint32_t nrec(int64_t) const { return 2u; }
    static constexpr bool kStage = false;
    __device__ __forceinline__ uint64_t first_word(int64_t) const { return 0; }
    __device__ __forceinline__ uint64_t end_word(int64_t) const { return 0; }
    __device__ __forceinline__ uint64_t stage_word(uint64_t) const { return 0; }
    template <int NW, typename Ptr>
    __device__ __forceinline__ Kmer<NW> get_at(Ptr, uint32_t) const { return Kmer<NW>(); }
    template <int NW>
    __device__ __forceinline__ Kmer<NW> get(int64_t item, uint32_t j) const {
        Kmer<NWS> x;
#pragma unroll
        for (int q = 0; q < NWS; ++q) x.w[q] = keys[item * NWS + q];
        Kmer<NW> f = j ? kmer_suffix<NW, NWS>(x, K) : kmer_prefix<NW, NWS>(x, K);
        Kmer<NW> r = kmer_rc<NW>(f, K);
        return kmer_is_minimal<NW>(f, r) ? f : r;
    }
};

template <int NW>
__device__ __forceinline__ void store_rec(uint64_t *dst, const Kmer<NW> &k) {
    if (NW == 2) {
        *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(k.w[0], k.w[1]);
    } else if (NW == 4) {
        reinterpret_cast<ulonglong2 *>(dst)[0] = make_ulonglong2(k.w[0], k.w[1]);
        reinterpret_cast<ulonglong2 *>(dst)[1] = make_ulonglong2(k.w[2], k.w[3]);
    } else {
#pragma unroll
        for (int q = 0; q < NW; ++q) dst[q] = k.w[q];
    }
}
template <int NW>
__device__ __forceinline__ Kmer<NW> load_rec(const uint64_t *src) {
    Kmer<NW> k;
    if (NW == 2) {
        ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src);
        k.w[0] = v.x; k.w[1] = v.y;
    } else if (NW == 4) {
        ulonglong2 a = reinterpret_cast<const ulonglong2 *>(src)[0], b = reinterpret_cast<const ulonglong2 *>(src)[1];
        k.w[0] = a.x; k.w[1] = a.y; k.w[2] = b.x; k.w[3] = b.y;
    } else {
#pragma unroll
        for (int q = 0; q < NW; ++q) k.w[q] = src[q];
    }
    return k;
}

// ------------------------------------------------------------------------------------------------------------
// level A
// ------------------------------------------------------------------------------------------------------------
struct LevelA {
    int K;
    uint32_t B;
    uint32_t b_lo, b_hi;    // buckets of this pass
    int rA;                 // key bits folded into the partition id
    uint32_t PA;            // (b_hi-b_lo) << rA
};

static const int kATile = 256;        // items (reads) per tile
static const int kAThreads = 1024;

// tile prologue: per-item record counts -> exclusive prefix in shared memory; returns the tile total
template <class Src>
__device__ __forceinline__ uint32_t tile_prefix(const Src &src, int64_t item0, int nitems, uint32_t *pref /*kATile+1*/, uint32_t *uniform = nullptr) {
    // kAThreads >= kATile: thread t owns item t
    uint32_t c = 0;
    if ((int)threadIdx.x < nitems) c = src.nrec(item0 + threadIdx.x);
    // block scan over the first kATile threads (8 warps)
    __shared__ uint32_t wsum[kATile / 32 + 1];
    __shared__ uint32_t s_c0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_c0 = c;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (warp < kATile / 32 && lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < kATile / 32; ++w) { uint32_t t = wsum[w]; wsum[w] = run; run += t; }
        wsum[kATile / 32] = run;
    }
    // every item yields the same number of records (fixed-length reads): item = i / c0 instead of a binary search
    const uint32_t c0 = s_c0;
    const int same = __syncthreads_and((int)threadIdx.x >= nitems || c == c0);
    if ((int)threadIdx.x < kATile) pref[threadIdx.x] = wsum[warp] + inc - c;
    uint32_t total = wsum[kATile / 32];
    if (threadIdx.x == 0) pref[kATile] = total;
    __syncthreads();
    if (uniform) *uniform = (same && c0) ? c0 : 0u;
    return total;
}

__device__ __forceinline__ int find_item_u(const uint32_t *pref, int nitems, uint32_t i, uint32_t uniform);
__device__ __forceinline__ int find_item(const uint32_t *pref, int nitems, uint32_t i) {
    // largest t with pref[t] <= i   (pref is exclusive, nitems <= kATile)
    int lo = 0, hi = nitems - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pref[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Stage a tile's packed reads in shared memory (ncu: the partition kernel's L1 hit rate on the read words dropped to 31 %
// under its own store traffic and "long scoreboard" became its top stall). The reads of a tile are one contiguous word range
// in every layout this library produces; anything else (or very long reads) falls back to global loads.
static const int kStageWords = 2560;      // 20 KB: 256 reads x 10 words (<= 320 bp each)
struct TileStage {
    uint64_t words[kStageWords];
    uint32_t off[kATile];
};
template <class Src>
__device__ __forceinline__ bool tile_stage(const Src &src, int64_t item0, int nitems, TileStage &ts) {
    if (!Src::kStage) return false;
    __shared__ unsigned s_maxend;
    const uint64_t w0 = src.first_word(item0);
    if (threadIdx.x == 0) s_maxend = 0;
    __syncthreads();
    bool ok = true;
    if ((int)threadIdx.x < nitems) {
        const uint64_t a = src.first_word(item0 + threadIdx.x), b = src.end_word(item0 + threadIdx.x);
        ok = a >= w0 && b >= a && b - w0 <= (uint64_t)kStageWords;
        if (ok) { ts.off[threadIdx.x] = (uint32_t)(a - w0); atomicMax(&s_maxend, (unsigned)(b - w0)); }
    }
    const bool staged = __syncthreads_and(ok) != 0;
    if (staged) {
        const uint32_t nwords = s_maxend;
        for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) ts.words[i] = src.stage_word(w0 + i);
    }
    __syncthreads();
    return staged;
}
template <int NW, class Src>
__device__ __forceinline__ Kmer<NW> tile_get(const Src &src, bool staged, const TileStage &ts, int64_t item0, int it, uint32_t j) {
    if (Src::kStage && staged) return src.template get_at<NW>(ts.words + ts.off[it], j);
    return src.template get<NW>(item0 + it, j);
}
// record stores bypass L1 allocation (they are never re-read by this kernel)
template <int NW>
__device__ __forceinline__ void store_rec_stream(uint64_t *dst, const Kmer<NW> &k) {
    if (NW == 2) {
        asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(k.w[0]), "l"(k.w[1]) : "memory");
    } else if (NW == 4) {
        asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(k.w[0]), "l"(k.w[1]) : "memory");
        asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(dst + 2), "l"(k.w[2]), "l"(k.w[3]) : "memory");
    } else {
#pragma unroll
        for (int q = 0; q < NW; ++q) asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(dst + q), "l"(k.w[q]) : "memory");
    }
}

__device__ __forceinline__ int find_item_u(const uint32_t *pref, int nitems, uint32_t i, uint32_t uniform) {
    return uniform ? (int)(i / uniform) : find_item(pref, nitems, i);
}

template <int NW>
__device__ __forceinline__ bool part_of(const LevelA &p, const Kmer<NW> &k, uint32_t *part) {
    uint32_t b = kmer_bucket<NW>(k, p.B);
    if (b < p.b_lo || b >= p.b_hi) return false;
    *part = ((b - p.b_lo) << p.rA) | key_top_bits<NW>(k, p.K, p.rA);
    return true;
}

// records per tile (kATile items), for the per-record partition-id array
template <class Src>
__global__ void tile_totals_k(Src src, int64_t ntiles, uint32_t *__restrict__ out, uint32_t pad) {
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    const int lane = threadIdx.x & 31;
    const int64_t item0 = t * kATile;
    uint32_t s = 0;
    for (int i = lane; i < kATile && item0 + i < src.n; i += 32) s += src.nrec(item0 + i);
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[t] = (s + pad - 1) / pad * pad;       // rows of the id array start 2*pad-byte aligned
}

// items are split statically: CTA g owns tiles [g*tiles_per, ...) so that count and scatter agree
template <int NW, class Src>
__global__ void __launch_bounds__(kAThreads) levelA_count_k(Src src, LevelA p, uint32_t *__restrict__ blk_counts,
                                                            const uint64_t *__restrict__ tile_off, uint16_t *__restrict__ ids) {
    extern __shared__ uint32_t sm_dyn[];
    uint32_t *hist = sm_dyn;                  // PA
    __shared__ uint32_t pref[kATile + 1];
    __shared__ TileStage ts;
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int64_t ntiles = (src.n + kATile - 1) / kATile;
    const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(ntiles, t0 + per);
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t item0 = t * kATile;
        const int nitems = (int)min((int64_t)kATile, src.n - item0);
        uint32_t unif = 0;
        const uint32_t total = tile_prefix(src, item0, nitems, pref, &unif);
        const bool staged = tile_stage(src, item0, nitems, ts);
        const uint64_t toff = ids ? tile_off[t] : 0;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            int it = find_item_u(pref, nitems, i, unif);
            Kmer<NW> k = tile_get<NW>(src, staged, ts, item0, it, i - pref[it]);
            uint32_t part = 0xffffu;
            if (part_of<NW>(p, k, &part)) atomicAdd(&hist[part], 1u); else part = 0xffffu;
            // remember the partition of every record: the scatter passes (one per bucket group) then neither hash nor
            // even extract the records that are not theirs
            if (ids) ids[toff + i] = (uint16_t)part;
        }
        __syncthreads();
    }
    uint32_t *out = blk_counts + (size_t)blockIdx.x * p.PA;
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) out[i] += hist[i];
}

// one thread per partition: totals and per-CTA bases. base[g][part] is the running cursor of CTA g.
__global__ void levelA_totals_k(const uint32_t *__restrict__ blk_counts, uint32_t PA, int G, uint64_t *__restrict__ part_total) {
    uint32_t part = blockIdx.x * blockDim.x + threadIdx.x;
    if (part >= PA) return;
    uint64_t s = 0;
    for (int g = 0; g < G; ++g) s += blk_counts[(size_t)g * PA + part];
    part_total[part] = s;
}
__global__ void levelA_bases_k(const uint32_t *__restrict__ blk_counts, uint32_t stride, uint32_t PA, int G,
                               const uint64_t *__restrict__ part_start, uint64_t *__restrict__ base) {
    uint32_t part = blockIdx.x * blockDim.x + threadIdx.x;
    if (part >= PA) return;
    uint64_t run = part_start[part];
    for (int g = 0; g < G; ++g) {
        base[(size_t)g * PA + part] = run;
        run += blk_counts[(size_t)g * stride + part];
    }
}

template <int NW, class Src>
__global__ void __launch_bounds__(kAThreads) levelA_scatter_k(Src src, LevelA p, uint64_t *__restrict__ base, uint64_t *__restrict__ out,
                                                              const uint64_t *__restrict__ tile_off, const uint16_t *__restrict__ ids, uint32_t id_lo) {
    extern __shared__ uint32_t sm_dyn[];
    uint64_t *cur_base = reinterpret_cast<uint64_t *>(sm_dyn);          // PA u64
    uint32_t *cnt = reinterpret_cast<uint32_t *>(cur_base + p.PA);      // PA u32
    __shared__ uint32_t pref[kATile + 1];
    __shared__ TileStage ts;
    uint64_t *mybase = base + (size_t)blockIdx.x * p.PA;
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) { cur_base[i] = mybase[i]; cnt[i] = 0; }
    __syncthreads();
    const int64_t ntiles = (src.n + kATile - 1) / kATile;
    const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(ntiles, t0 + per);
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t item0 = t * kATile;
        const int nitems = (int)min((int64_t)kATile, src.n - item0);
        uint32_t unif = 0;
        const uint32_t total = tile_prefix(src, item0, nitems, pref, &unif);
        const bool staged = tile_stage(src, item0, nitems, ts);
        if (ids) {
            // ncu (source view) put 32 % of this kernel's stall samples on the consumer of the per-record 2-byte id load and
            // 17 % on the binary search: issue four id loads before touching any of them, divide instead of searching
            const uint16_t *row = ids + tile_off[t];
            constexpr int U = 4;
            for (uint32_t i0 = threadIdx.x; i0 < total; i0 += U * blockDim.x) {
                uint32_t part[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t i = i0 + u * blockDim.x;
                    part[u] = i < total ? (uint32_t)__ldg(row + i) - id_lo : 0xffffffffu;      // 0xffff - id_lo stays >= PA
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (part[u] >= p.PA) continue;
                    const uint32_t i = i0 + u * blockDim.x;
                    const int it = find_item_u(pref, nitems, i, unif);
                    Kmer<NW> k = tile_get<NW>(src, staged, ts, item0, it, i - pref[it]);
                    uint32_t slot = atomicAdd(&cnt[part[u]], 1u);
                    store_rec_stream<NW>(out + (cur_base[part[u]] + slot) * NW, k);
                }
            }
        } else {
            for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
                uint32_t part;
                int it = find_item_u(pref, nitems, i, unif);
                Kmer<NW> k = tile_get<NW>(src, staged, ts, item0, it, i - pref[it]);
                if (part_of<NW>(p, k, &part)) {
                    uint32_t slot = atomicAdd(&cnt[part], 1u);
                    store_rec_stream<NW>(out + (cur_base[part] + slot) * NW, k);
                }
            }
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) mybase[i] = cur_base[i] + cnt[i];   // chained launches continue here
}

// ---- level A, rolling generation ------------------------------------------------------------------------------------
// ncu on the kernels above: 296 (count) and 250 (scatter) thread instructions per window, issue-bound at 74 % / latency-bound
// at 27 % issue utilisation -- every record re-extracted its window from up to three packed words, re-ran FastRC, divided to
// find its read and (scatter) waited for a 2-byte id load. Here the unit of work is a CHUNK of kRollC consecutive windows of
// one read, walked by one thread with a rolling window + rolling reverse complement (kmer_dev.cuh roll_*): one shared-memory
// word load per 32 bases, ~20 integer instructions per window, and in a bucket-group pass the windows of other groups cost
// only the roll. The 2-byte partition ids of a chunk are 48 contiguous bytes: written as six 8-byte words by the count pass,
// fetched as three 16-byte loads before the walk starts by every scatter pass. Reads only (canonical mode).
static const int kRollC = 24;            // windows per chunk
static_assert(kRollC <= 33 && kRollC % 8 == 0, "roll_init fetches a chunk's bases from two words; an id row is a whole number of 16-byte words");
static const int kRollThreads = 512;     // 2-3 CTAs per SM; units of a tile are dealt round-robin to the threads

// Round 2: the tile of the rolling kernels is a WARP's: 32 consecutive reads, staged, scanned and walked by one warp with
// shuffles and __syncwarp only. (The CTA-wide tiles of the first version cost five __syncthreads per 256 reads; ncu showed
// "barrier" and "long scoreboard" as the top stalls of the partition kernel, 2.2 + 2.1 cycles per issued instruction.) A CTA
// still owns a contiguous range of tiles -- the same range in the count and in every scatter launch, which is what makes the
// per-CTA histogram of the count the cursor table of the scatter -- and deals them round-robin to its warps.
static const int kRollTile = 32;                      // reads per warp tile
static const int kRollWarps = kRollThreads / 32;
static const int kRollStageWords = 320;               // per warp: 32 reads x 10 words (<= 320 bp each), else global loads
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
struct RollWarp {
    uint64_t words[kRollStageWords];     // the tile's packed reads
    uint32_t pref[kRollTile + 1];        // exclusive prefix of chunks per read
    uint32_t len[kRollTile];             // read lengths
    uint32_t off[kRollTile];             // first staged word of a read
    uint32_t pad_;
};
static_assert(sizeof(RollWarp) % 8 == 0, "per-warp slices stay 8-byte aligned");
static size_t roll_smem_bytes(uint32_t nslots) { return (((size_t)nslots * 4 + 15) & ~(size_t)15) + (size_t)kRollWarps * sizeof(RollWarp); }
__device__ __forceinline__ RollWarp *roll_warp_slice(unsigned char *raw, uint32_t nslots) {
    return reinterpret_cast<RollWarp *>(raw + (((size_t)nslots * 4 + 15) & ~(size_t)15)) + (threadIdx.x >> 5);
}

// one warp: chunk counts of the tile's reads -> exclusive prefix, the reads' words -> shared memory. Returns the tile's chunk
// total; *uniform = chunks per read when every read has the same (non-zero) count, else 0; *staged = words are in rw.words.
__device__ __forceinline__ uint32_t roll_warp_setup(const ReadsSrc &src, int64_t item0, int nitems, RollWarp &rw, uint32_t *uniform, bool *staged) {
    const int lane = threadIdx.x & 31;
    uint32_t c = 0, L = 0;
    uint64_t a = 0, b = 0;
    if (lane < nitems) {
        L = src.lens[item0 + lane];
        a = src.offs[item0 + lane];
        b = a + (((uint64_t)L + 31) >> 5);
        const uint32_t w = (int)L >= src.K ? (uint32_t)((int)L - src.K + 1) : 0u;
        c = (w + kRollC - 1) / kRollC;
    }
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    rw.len[lane] = L;
    rw.pref[lane] = inc - c;
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 31) rw.pref[kRollTile] = total;
    const uint32_t c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool same = __all_sync(0xffffffffu, lane >= nitems || c == c0);
    *uniform = (same && c0) ? c0 : 0u;
    const uint64_t w0 = __shfl_sync(0xffffffffu, a, 0);
    const bool ok = lane >= nitems || (a >= w0 && b >= a && b - w0 <= (uint64_t)kRollStageWords);
    const bool st = __all_sync(0xffffffffu, ok);
    if (st) {
        rw.off[lane] = (uint32_t)(a - w0);
        uint32_t end = lane < nitems ? (uint32_t)(b - w0) : 0u;
#pragma unroll
        for (int o = 16; o; o >>= 1) end = max(end, __shfl_xor_sync(0xffffffffu, end, o));
        // all loads first, then the stores: the copy costs one memory round trip, not one per 32 words
        uint64_t tmp[kRollStageWords / 32];
#pragma unroll
        for (int j = 0; j < kRollStageWords / 32; ++j) tmp[j] = (uint32_t)(lane + 32 * j) < end ? __ldg(src.words + w0 + lane + 32 * j) : 0ull;
#pragma unroll
        for (int j = 0; j < kRollStageWords / 32; ++j) if ((uint32_t)(lane + 32 * j) < end) rw.words[lane + 32 * j] = tmp[j];
    }
    *staged = st;
    __syncwarp();
    return total;
}

// chunks per warp tile x kRollC = ids per tile (rows of the id array: 48 bytes per chunk, so every row is 16-byte aligned)
__global__ void roll_tile_ids_k(ReadsSrc src, int64_t ntiles, uint32_t *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    const int lane = threadIdx.x & 31;
    const int64_t item = t * kRollTile + lane;
    uint32_t s = 0;
    if (item < src.n) {
        const int L = (int)src.lens[item];
        const uint32_t w = L >= src.K ? (uint32_t)(L - src.K + 1) : 0u;
        s = (w + kRollC - 1) / kRollC;
    }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[t] = s * kRollC;
}

// geometry of one unit (chunk): which read, which windows
struct RollUnit { int it, j0, cnt; };
__device__ __forceinline__ RollUnit roll_unit(const RollWarp &rw, int nitems, uint32_t u, uint32_t unif, int K) {
    RollUnit q;
    q.it = find_item_u(rw.pref, nitems, u, unif);
    q.j0 = (int)(u - rw.pref[q.it]) * kRollC;
    const int nwin = (int)rw.len[q.it] - K + 1;
    q.cnt = nwin - q.j0 < kRollC ? nwin - q.j0 : kRollC;
    return q;
}

template <int NW>
__global__ void __launch_bounds__(kRollThreads, 2) levelA_count_roll_k(ReadsSrc src, LevelA p, uint32_t *__restrict__ blk_counts,
                                                                      const uint64_t *__restrict__ tile_off, uint16_t *__restrict__ ids) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(sm_raw);     // PA
    RollWarp &rw = *roll_warp_slice(sm_raw, p.PA);
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int K = p.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (src.n + kRollTile - 1) / kRollTile;
    const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(ntiles, t0 + per);
    for (int64_t t = t0 + warp; t < t1; t += kRollWarps) {
        const int64_t item0 = t * kRollTile;
        const int nitems = (int)min((int64_t)kRollTile, src.n - item0);
        uint32_t unif = 0;
        bool staged = false;
        // the warp's next tile is kRollWarps tiles ahead: its lengths and offsets go to L2 now, its packed reads (and id rows) at the
        // end of this tile, when the loads of their addresses issued here have long returned
        const bool has_next = t + kRollWarps < t1;
        const int64_t nx = (t + kRollWarps) * kRollTile;
        uint64_t next_w0 = 0;
        if (has_next) {
            next_w0 = src.offs[nx];
            if (lane == 0) prefetch_l2(src.lens + nx);
            else if (lane == 1 && nx + 16 < src.n) prefetch_l2(src.offs + nx + 16);
        }
        const uint32_t nunits = roll_warp_setup(src, item0, nitems, rw, &unif, &staged);
        uint64_t *row = ids ? reinterpret_cast<uint64_t *>(ids + tile_off[t]) : nullptr;
        for (uint32_t u = lane; u < nunits; u += 32) {
            const RollUnit q = roll_unit(rw, nitems, u, unif, K);
            const uint64_t *seq = staged ? static_cast<const uint64_t *>(rw.words + rw.off[q.it]) : src.words + src.offs[item0 + q.it];
            RollState<NW> st;
            roll_init<NW>(st, seq, q.j0, K, q.cnt);
#pragma unroll 1
            for (int g = 0; g < kRollC / 4; ++g) {
                uint64_t acc = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int s = 4 * g + e;
                    uint32_t id = 0xffffu;
                    if (s < q.cnt) {
                        if (s > 0) roll_next<NW>(st, K);
                        const Kmer<NW> k = kmer_is_minimal<NW>(st.f, st.r) ? st.f : st.r;
                        uint32_t part;
                        if (part_of<NW>(p, k, &part)) { atomicAdd(&hist[part], 1u); id = part; }
                    }
                    acc |= (uint64_t)id << (16 * e);
                }
                if (row) row[(size_t)u * (kRollC / 4) + g] = acc;
            }
        }
        if (has_next && lane < 10 && next_w0 + 16 * lane < src.nwords) prefetch_l2(src.words + next_w0 + 16 * lane);   // 10 lines = 32 reads x 5 words
        __syncwarp();                                   // the slice is rewritten by the next tile's setup
    }
    __syncthreads();
    uint32_t *out = blk_counts + (size_t)blockIdx.x * p.PA;
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) out[i] += hist[i];
}

// base rows are `row_stride` cursors long; this launch handles the partitions [q_lo, q_lo + p.PA) of a row (p.PA = the
// sub-range's size, id_lo = id of its first partition): a bucket-group pass may be split into partition sub-ranges so that
// the lines and pages a CTA has open at any time stay few (TLB reach).
// (A sector-pairing variant of this kernel -- two records of a stream leave as one 32-byte store through shared-memory
// mailboxes -- was measured on the B200 in round 2: parity clean but 1.8x SLOWER, 185 -> 336 ms at 100 M reads; the CAS
// traffic on the mailboxes costs more than the full sectors win. profiles/r02a_sweep_variants.log. Removed.)
template <int NW, bool HAS_IDS>
__global__ void __launch_bounds__(kRollThreads, 2) levelA_scatter_roll_k(ReadsSrc src, LevelA p, uint64_t *__restrict__ base, uint64_t *__restrict__ out,
                                                                        const uint64_t *__restrict__ tile_off, const uint16_t *__restrict__ ids,
                                                                        uint32_t id_lo, uint32_t row_stride, uint32_t q_lo, uint64_t ids_len) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    // one 32-bit cursor per partition, relative to the first record this launch may write (a CTA's share of a pass is far below
    // 2^32 records): the slot of a record is ONE shared-memory atomic, no base lookup and no 64-bit add behind it
    uint32_t *cur = reinterpret_cast<uint32_t *>(sm_raw);               // PA u32
    RollWarp &rw = *roll_warp_slice(sm_raw, p.PA);
    uint64_t *mybase = base + (size_t)blockIdx.x * row_stride + q_lo;
    const uint64_t region0 = mybase[0];                                 // cursors of a row ascend with the partition
    for (uint32_t i = threadIdx.x; i < p.PA; i += blockDim.x) cur[i] = (uint32_t)(mybase[i] - region0);
    __syncthreads();
    uint64_t *const out0 = out + region0 * NW;
    const int K = p.K;
    const uint32_t PA = p.PA;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (src.n + kRollTile - 1) / kRollTile;
    const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = min(ntiles, t0 + per);
    for (int64_t t = t0 + warp; t < t1; t += kRollWarps) {
        const int64_t item0 = t * kRollTile;
        const int nitems = (int)min((int64_t)kRollTile, src.n - item0);
        uint32_t unif = 0;
        bool staged = false;
        // the warp's next tile is kRollWarps tiles ahead: its lengths and offsets go to L2 now, its packed reads (and id rows) at the
        // end of this tile, when the loads of their addresses issued here have long returned
        const bool has_next = t + kRollWarps < t1;
        const int64_t nx = (t + kRollWarps) * kRollTile;
        uint64_t next_w0 = 0;
        if (has_next) {
            next_w0 = src.offs[nx];
            if (lane == 0) prefetch_l2(src.lens + nx);
            else if (lane == 1 && nx + 16 < src.n) prefetch_l2(src.offs + nx + 16);
        }
        uint64_t next_row = 0;
        if (HAS_IDS && has_next) next_row = tile_off[t + kRollWarps];
        const uint32_t nunits = roll_warp_setup(src, item0, nitems, rw, &unif, &staged);
        const ulonglong2 *row = HAS_IDS ? reinterpret_cast<const ulonglong2 *>(ids + tile_off[t]) : nullptr;
        for (uint32_t u = lane; u < nunits; u += 32) {
            uint64_t idw[kRollC / 4];
            if (HAS_IDS) {
                // the chunk's 24 ids, in flight while the window is set up
#pragma unroll
                for (int v = 0; v < kRollC / 8; ++v) {
                    const ulonglong2 x = __ldg(row + (size_t)u * (kRollC / 8) + v);
                    idw[2 * v] = x.x; idw[2 * v + 1] = x.y;
                }
            }
            const RollUnit q = roll_unit(rw, nitems, u, unif, K);
            const uint64_t *seq = staged ? static_cast<const uint64_t *>(rw.words + rw.off[q.it]) : src.words + src.offs[item0 + q.it];
            RollState<NW> st;
            roll_init<NW>(st, seq, q.j0, K, q.cnt);
            // (Delaying a window's store by one window, so that the atomic's latency hides behind the next roll, and 3..6 CTAs per
            // SM to even out the tail were both measured in round 2: 160-162 ms either way at 100 M reads. profiles/r02n_*.)
#pragma unroll
            for (int s = 0; s < kRollC; ++s) {
                if (s >= q.cnt) break;
                if (s > 0) roll_next<NW>(st, K);
                uint32_t part;
                bool mine;
                if (HAS_IDS) {
                    // a foreign window costs the roll and this compare; the canonical choice is made for own windows only
                    part = (uint32_t)((idw[s >> 2] >> (16 * (s & 3))) & 0xffffu) - id_lo;         // 0xffff - id_lo stays >= PA
                    mine = part < PA;
                    if (mine) {
                        const uint32_t slot = atomicAdd(&cur[part], 1u);
                        store_rec_stream<NW>(out0 + (size_t)slot * NW, kmer_is_minimal<NW>(st.f, st.r) ? st.f : st.r);
                    }
                } else {
                    const Kmer<NW> k = kmer_is_minimal<NW>(st.f, st.r) ? st.f : st.r;
                    mine = part_of<NW>(p, k, &part);
                    if (mine) {
                        const uint32_t slot = atomicAdd(&cur[part], 1u);
                        store_rec_stream<NW>(out0 + (size_t)slot * NW, k);
                    }
                }
            }
        }
        if (has_next && lane < 10 && next_w0 + 16 * lane < src.nwords) prefetch_l2(src.words + next_w0 + 16 * lane);   // 10 lines = 32 reads x 5 words
        if (HAS_IDS && has_next) {                                              // the next tile's id rows: 48 lines for 128 chunks
            if (next_row + 64 * lane < ids_len) prefetch_l2(ids + next_row + 64 * lane);
            if (lane < 16 && next_row + 64 * (32 + lane) < ids_len) prefetch_l2(ids + next_row + 64 * (32 + lane));
        }
        __syncwarp();                                   // the slice is rewritten by the next tile's setup
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < PA; i += blockDim.x) mybase[i] = region0 + cur[i];   // chained launches continue here
}

// (Round 2 measured a third generation of this kernel -- id sweep, 32-bit keys collected without atomics, ballot-ranked LSD sort
// of the batch in shared memory, run-by-run flush with consecutive threads storing consecutive records of one partition. Parity
// clean, full-sector stores, and SLOWER: 251 vs 180 ms at 100 M reads, 42.8 vs 31.1 ms at 20 M: ~310 thread instructions per
// record at IPC ~1.3 (barriers + dependent ballot / shared-memory chains) against ~170 for roll + atomic + 16-byte store.
// profiles/r02d_sorted_partition_kernel_AB.log. Removed.)

// ------------------------------------------------------------------------------------------------------------
// segments
// ------------------------------------------------------------------------------------------------------------
struct Seg {
    uint64_t start;     // record index in its buffer
    uint64_t len;
    uint32_t bits;      // key bits already fixed by partitioning
    uint32_t bb;        // (bucket << 1) | buffer
};

__global__ void seg_init_k(const uint64_t *__restrict__ part_start, const uint64_t *__restrict__ part_total, uint32_t PA, int rA,
                           uint32_t b_lo, Seg *__restrict__ segs) {
    uint32_t part = blockIdx.x * blockDim.x + threadIdx.x;
    if (part >= PA) return;
    Seg s;
    s.start = part_start[part]; s.len = part_total[part]; s.bits = (uint32_t)rA; s.bb = ((b_lo + (part >> rA)) << 1) | 0u;
    segs[part] = s;
}

// ---- CTA-major staging of the level-A output ------------------------------------------------------------------------------
// Measured on the B200: the partition kernel ran at the same ~0.55 TB/s whether it executed 250 or 50 instructions per
// record and whether a CTA had 128 or 2560 partitions open, while the MSD refinement -- same 16-byte scattered stores, same
// number of open lines -- ran 4x faster. The difference is WHERE a CTA's stores go: with a partition-major output every
// CTA writes all over the 30 GB buffer (15 k 2-MB pages against a 128-entry TLB per SM, B300_MICROARCH.md "TLB"), the
// refinement writes inside one 12 MB segment. So level A writes CTA-major: CTA g owns one contiguous region of the staging
// buffer, partitioned inside ([g][partition]); a partition is then G pieces, and the FIRST refinement round reads its
// segment piece by piece (sequential runs: one page at a time) and writes the children partition-major into the partner
// buffer -- the gather costs no extra pass over the data. Partitions that need no refinement take the same kernel with
// r = 0 (a copy into the partner buffer).
struct Pieces {
    const uint64_t *pbase;     // [G][PA]  first record of piece (g, partition) in the staging buffer
    const uint32_t *cnt;       // blk_counts + p_lo: cnt[g * cnt_stride + partition]
    uint32_t cnt_stride, PA;
    int G;
};
static const int kMaxPieces = 1024;       // CTAs of the level-A grid (G <= 1024)

// row g of cnt -> exclusive prefix inside the row (rel) and the row total
__global__ void stage_rows_k(const uint32_t *__restrict__ cnt, uint32_t cnt_stride, uint32_t PA, uint64_t *__restrict__ rel, uint64_t *__restrict__ row_total) {
    __shared__ uint64_t wsum[8];
    __shared__ uint64_t carry_s;
    const uint32_t *row = cnt + (size_t)blockIdx.x * cnt_stride;
    uint64_t *out = rel + (size_t)blockIdx.x * PA;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;       // 256 threads
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t b = 0; b < PA; b += blockDim.x) {
        const uint32_t i = b + threadIdx.x;
        const uint64_t v = i < PA ? row[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint64_t wb = 0, all = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) wb += wsum[w]; all += wsum[w]; }
        const uint64_t carry = carry_s;
        if (i < PA) out[i] = carry + wb + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_total[blockIdx.x] = carry_s;
}
__global__ void stage_add_k(uint64_t *__restrict__ rel, const uint64_t *__restrict__ row_start, uint32_t PA, int G) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)G * PA) return;
    rel[i] += row_start[i / PA];
}

struct RefinePlan { uint32_t cap, target, rmax; int total_bits; };
__device__ __forceinline__ int plan_r(const RefinePlan &rp, uint64_t len, uint32_t bits) {
    if (len <= rp.cap) return 0;
    int rem = rp.total_bits - (int)bits;
    if (rem <= 0) return 0;
    uint64_t want = (len + rp.target - 1) / rp.target;
    int r = 1;
    while (r < (int)rp.rmax && (1ull << r) < want) ++r;
    return r < rem ? r : rem;
}
// children[i] = 2^r or 1 ; work flag
__global__ void refine_plan_k(const Seg *__restrict__ segs, uint64_t n, RefinePlan rp, uint32_t *__restrict__ nchild, uint32_t *__restrict__ isw, int force_all) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = plan_r(rp, segs[i].len, segs[i].bits);
    nchild[i] = r ? (1u << r) : 1u;
    isw[i] = (r || force_all) ? 1u : 0u;      // force_all: staged level-A output, every segment is gathered by the refinement kernel
}
__global__ void refine_copy_k(const Seg *__restrict__ segs, uint64_t n, const uint32_t *__restrict__ isw, const uint64_t *__restrict__ child_base,
                              const uint64_t *__restrict__ work_pos, Seg *__restrict__ nsegs, uint64_t *__restrict__ worklist) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (isw[i]) worklist[work_pos[i]] = i;
    else nsegs[child_base[i]] = segs[i];
}

// stores of the pairing refinement (pair_mailbox.cuh): one full 32-byte sector for two records of a bin, 16 bytes for a lone record
struct PairSink {
    uint64_t *out;
    __device__ __forceinline__ void pair(uint64_t pos, uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1) {
        asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(out + pos * 2), "l"(a0), "l"(a1), "l"(b0), "l"(b1) : "memory");
    }
    __device__ __forceinline__ void single(uint64_t pos, uint64_t w0, uint64_t w1) {
        asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(out + pos * 2), "l"(w0), "l"(w1) : "memory");
    }
};

static const int kRThreads = 1024;
static const int kRMaxBins = 2048;

// one CTA splits one oversize segment by its next r key bits: buf[bb&1] -> buf[(bb&1)^1]
// PIECED: the segments are level-A partitions whose records lie in G pieces of the CTA-major staging buffer (buf0); segment
// index == partition. The pieces are read one after the other, the children are written contiguously into buf1.
// PAIR (opt-in, 16-byte records): the scatter phase pairs the two records of a 32-byte sector through one mailbox per bin
// (pair_mailbox.cuh) -- at 2048 bins a bin receives a record only every ~8 us and half-written sectors do not survive in L2.
template <int NW, bool PIECED, bool PAIR>
__global__ void __launch_bounds__(kRThreads) refine_k(const Seg *__restrict__ segs, const uint64_t *__restrict__ worklist, uint64_t nwork,
                                                     const uint64_t *__restrict__ child_base, RefinePlan rp, int K,
                                                     uint64_t *__restrict__ buf0, uint64_t *__restrict__ buf1, Seg *__restrict__ nsegs,
                                                     unsigned long long *__restrict__ work_counter, Pieces pc) {
    __shared__ uint32_t hist[kRMaxBins];
    __shared__ uint32_t warp_tot[kRThreads / 32];
    __shared__ unsigned long long s_w;
    __shared__ uint64_t pc_start[PIECED ? kMaxPieces : 1];
    __shared__ uint32_t pc_len[PIECED ? kMaxPieces : 1];
    extern __shared__ uint64_t sm_refine_dyn[];
    PmBox *boxes = reinterpret_cast<PmBox *>(sm_refine_dyn);           // PAIR: one mailbox per bin
    for (;;) {
        if (threadIdx.x == 0) s_w = atomicAdd(work_counter, 1ull);
        __syncthreads();
        const uint64_t wi = s_w;
        __syncthreads();
        if (wi >= nwork) return;
        const uint64_t si = worklist[wi];
        const Seg s = segs[si];
        const int r = plan_r(rp, s.len, s.bits);
        const uint32_t nb = 1u << r;
        const uint64_t *src = ((s.bb & 1) ? buf1 : buf0) + s.start * NW;
        uint64_t *dst = ((s.bb & 1) ? buf0 : buf1) + s.start * NW;
        if (PIECED) {
            for (int g = threadIdx.x; g < pc.G; g += blockDim.x) {
                pc_start[g] = pc.pbase[(size_t)g * pc.PA + si];
                pc_len[g] = pc.cnt[(size_t)g * pc.cnt_stride + si];
            }
        }
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        if (PIECED) {
            // a WARP streams a piece (a few thousand consecutive records), four independent 512-byte loads in flight per warp:
            // with the whole CTA striding over one piece at a time a thread had a single 16-byte load outstanding
            for (int g = (int)(threadIdx.x >> 5); g < pc.G; g += kRThreads / 32) {
                const uint64_t *ps = buf0 + pc_start[g] * NW;
                const uint32_t n = pc_len[g];
                uint32_t i = threadIdx.x & 31;
                for (; i + 96 < n; i += 128) {
                    const Kmer<NW> k0 = load_rec<NW>(ps + (uint64_t)i * NW), k1 = load_rec<NW>(ps + (uint64_t)(i + 32) * NW);
                    const Kmer<NW> k2 = load_rec<NW>(ps + (uint64_t)(i + 64) * NW), k3 = load_rec<NW>(ps + (uint64_t)(i + 96) * NW);
                    atomicAdd(&hist[key_bits<NW>(k0, K, (int)s.bits, r)], 1u);
                    atomicAdd(&hist[key_bits<NW>(k1, K, (int)s.bits, r)], 1u);
                    atomicAdd(&hist[key_bits<NW>(k2, K, (int)s.bits, r)], 1u);
                    atomicAdd(&hist[key_bits<NW>(k3, K, (int)s.bits, r)], 1u);
                }
                for (; i < n; i += 32) {
                    const Kmer<NW> k = load_rec<NW>(ps + (uint64_t)i * NW);
                    atomicAdd(&hist[key_bits<NW>(k, K, (int)s.bits, r)], 1u);
                }
            }
        } else {
            // (Four loads in flight for the contiguous form too -- second-level splits, the multi-GPU path -- were measured: 57 instead of
            // 32 registers halve its residency, no gain at 2 GPUs (200.9 vs 200.6 ms) and +15 ms on one (the second-level splits). Not kept.)
            for (uint64_t i = threadIdx.x; i < s.len; i += blockDim.x) {
                Kmer<NW> k = load_rec<NW>(src + i * NW);
                atomicAdd(&hist[key_bits<NW>(k, K, (int)s.bits, r)], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of hist[0..nb) (nb <= 2048 = 2 per thread)
        uint32_t a = 0, b = 0;
        const uint32_t i0 = 2 * threadIdx.x;
        if (i0 < nb) a = hist[i0];
        if (i0 + 1 < nb) b = hist[i0 + 1];
        uint32_t v = a + b, inc = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_tot[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t ex = warp_tot[warp] + inc - v;
        const uint64_t cb = child_base[si];
        if (i0 < nb) {
            Seg c; c.start = s.start + ex; c.len = a; c.bits = s.bits + (uint32_t)r; c.bb = s.bb ^ 1u;
            nsegs[cb + i0] = c;
        }
        if (i0 + 1 < nb) {
            Seg c; c.start = s.start + ex + a; c.len = b; c.bits = s.bits + (uint32_t)r; c.bb = s.bb ^ 1u;
            nsegs[cb + i0 + 1] = c;
        }
        __syncthreads();
        if (i0 < nb) hist[i0] = ex;
        if (i0 + 1 < nb) hist[i0 + 1] = ex + a;
        if (PAIR)
            for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) boxes[i].state = 0u;
        __syncthreads();
        uint64_t *dbuf = (s.bb & 1) ? buf0 : buf1;                    // dst = dbuf + s.start * NW
        auto put = [&](const Kmer<NW> &k) {
            const uint32_t bin = key_bits<NW>(k, K, (int)s.bits, r);
            const uint32_t slot = atomicAdd(&hist[bin], 1u);
            if constexpr (PAIR && NW == 2) {
                PairSink sink{dbuf};
                pm_put<1>(boxes + bin, s.start, slot, k.w[0], k.w[1], sink);
            } else {
                store_rec<NW>(dst + (uint64_t)slot * NW, k);
            }
        };
        if (PIECED) {
            for (int g = (int)(threadIdx.x >> 5); g < pc.G; g += kRThreads / 32) {
                const uint64_t *ps = buf0 + pc_start[g] * NW;
                const uint32_t n = pc_len[g];
                uint32_t i = threadIdx.x & 31;
                for (; i + 96 < n; i += 128) {
                    const Kmer<NW> k0 = load_rec<NW>(ps + (uint64_t)i * NW), k1 = load_rec<NW>(ps + (uint64_t)(i + 32) * NW);
                    const Kmer<NW> k2 = load_rec<NW>(ps + (uint64_t)(i + 64) * NW), k3 = load_rec<NW>(ps + (uint64_t)(i + 96) * NW);
                    put(k0); put(k1); put(k2); put(k3);
                }
                for (; i < n; i += 32) put(load_rec<NW>(ps + (uint64_t)i * NW));
            }
        } else {
            for (uint64_t i = threadIdx.x; i < s.len; i += blockDim.x) put(load_rec<NW>(src + i * NW));
        }
        __syncthreads();
        if constexpr (PAIR && NW == 2) {
            PairSink sink{dbuf};
            for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) pm_flush_box(&boxes[i], s.start, sink);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// local sort + unique + count
// ------------------------------------------------------------------------------------------------------------
template <int NW> struct SortCfg { static const int CAP = NW <= 2 ? 2048 : 1024; };
static const int kSThreads = 256;
static const int kSWarps = kSThreads / 32;

// stable LSD pass over key bits [pos, pos+width) : A -> Bf
template <int NW>
__device__ __forceinline__ void lsd_pass(const uint64_t *A, uint64_t *Bf, uint32_t n, int K, int pos, int width, uint32_t *cnt /*[kSWarps][256]*/,
                                         uint32_t *tot /*[256]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = ((n + kSWarps * 32 - 1) / (kSWarps * 32)) * 32;   // items per warp, multiple of 32
    const uint32_t w0 = warp * chunk;
    for (int i = threadIdx.x; i < kSWarps * 256; i += kSThreads) cnt[i] = 0;
    __syncthreads();
    for (uint32_t r = 0; r < chunk; r += 32) {
        const uint32_t i = w0 + r + lane;
        uint32_t d = 256;
        if (i < n) d = key_bits<NW>(load_rec<NW>(A + (size_t)i * NW), K, pos, width);
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        if (d < 256 && (int)(__ffs(m) - 1) == lane) cnt[warp * 256 + d] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    // per digit: warp prefix + digit totals
    {
        const uint32_t d = threadIdx.x;     // kSThreads == 256 digits
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSWarps; ++w) { uint32_t t = cnt[w * 256 + d]; cnt[w * 256 + d] = run; run += t; }
        // exclusive scan of run over 256 digits
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) tot[warp] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; ++w) wb += tot[w];
        __syncthreads();
        const uint32_t ex = wb + inc - run;
#pragma unroll
        for (int w = 0; w < kSWarps; ++w) cnt[w * 256 + d] += ex;
    }
    __syncthreads();
    for (uint32_t r = 0; r < chunk; r += 32) {
        const uint32_t i = w0 + r + lane;
        uint32_t d = 256;
        Kmer<NW> k;
        if (i < n) { k = load_rec<NW>(A + (size_t)i * NW); d = key_bits<NW>(k, K, pos, width); }
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        uint32_t rank = 0;
        if (d < 256) rank = cnt[warp * 256 + d] + __popc(m & ((1u << lane) - 1));
        __syncwarp();
        if (d < 256) {
            store_rec<NW>(Bf + (size_t)rank * NW, k);
            if ((int)(__ffs(m) - 1) == lane) cnt[warp * 256 + d] += __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
}

// sort A[0..n) by key bits [lo, hi) (MSB-first positions). Result may end up in A or Bf; returns pointer.
template <int NW>
__device__ __forceinline__ uint64_t *lsd_sort_range(uint64_t *A, uint64_t *Bf, uint32_t n, int K, int lo, int hi, uint32_t *cnt, uint32_t *tot) {
    int p = hi;
    while (p > lo) {
        int w = p - lo >= 8 ? 8 : p - lo;
        lsd_pass<NW>(A, Bf, n, K, p - w, w, cnt, tot);
        uint64_t *t = A; A = Bf; Bf = t;
        p -= w;
    }
    return A;
}

// ---- local sort: representative + residual -----------------------------------------------------------------------------
// After level A + MSD refinement a segment holds ~CAP*3/4 records that agree on their first `bits` key bits. Most of them are
// copies of a few keys (a genomic (k+1)-mer is seen ~coverage times) plus error singletons. Two earlier generations (a full LSD
// radix sort per segment; one counting pass into 2^11 bins with one thread collapsing each bin) were replaced by the kernel
// below; the LSD passes above survive as its exact fallback.
static const int kSortBinBits = 11;   // bins per segment = 2^11: fewer bins = less per-segment bookkeeping (init + two scans over the bins) but more
                                      // keys sharing a bin (B200, 20 M reads: 2^10 bins 51.0 ms vs 2^11 bins 50.2 ms -- no gain, profiles/r02a_sweep_variants.log)

// ncu on the one-thread-per-bin generation showed 8.5 active lanes per instruction and barrier stalls on top: one thread chewing
// through the ~coverage copies of a genomic k-mer held up its whole CTA. Here every bin elects a representative (the
// record with the smallest index, one shared-memory atomicMin per record); records equal to their bin's representative
// only bump a counter. Only the few records that differ from it (bins holding two or more distinct keys) are scattered
// into a small residual buffer and deduplicated by one thread per bin, so the per-bin work is a handful of compares.
// Anything unusual (residual overflow, a bin with many distinct keys) falls back to the exact LSD radix path, which
// uses the segment's region in the partner buffer as scratch.
static const int kResCap = 512;

template <int NW, int kBinBits, int CAP>
__global__ void __launch_bounds__(kSThreads) local_sort3_k(const Seg *__restrict__ segs, uint64_t nsegs, int K, uint64_t *__restrict__ buf0,
                                                          uint64_t *__restrict__ buf1, uint32_t *__restrict__ ndist,
                                                          unsigned long long *__restrict__ work_counter, unsigned long long *__restrict__ stats) {
    constexpr int kBins = 1 << kBinBits;
    constexpr int BPT = kBins / kSThreads;
    static_assert(3 * kBins >= CAP + 1 && 3 * kBins >= kSWarps * 256, "rep/repcnt/rhist double as scratch of the LSD fallback");
    static_assert(CAP % kSThreads == 0 && CAP <= SortCfg<NW>::CAP, "segment capacity: a multiple of the CTA size, at most the default");
    constexpr int IPT = CAP / kSThreads;                  // records per thread
    extern __shared__ uint64_t sm64[];
    uint64_t *A = sm64;                                   // CAP*NW   the segment
    uint64_t *R = A + (size_t)CAP * NW;                   // kResCap*NW residual records, grouped by bin
    uint32_t *rep = reinterpret_cast<uint32_t *>(R + (size_t)kResCap * NW);   // kBins  index of the bin's representative
    uint32_t *repcnt = rep + kBins;                       // kBins  copies of the representative besides itself
    uint32_t *rhist = repcnt + kBins;                     // kBins  residual count -> cursor
    uint16_t *rstart = reinterpret_cast<uint16_t *>(rhist + kBins);           // kBins+2
    uint16_t *rcnt = rstart + kBins + 2;                  // kResCap multiplicity per residual slot
    uint16_t *dres = rcnt + kResCap;                      // kBins  distinct residual keys per bin
    uint32_t *lsdcnt = rep;                               // fallback only (kSWarps*256 <= 3*kBins)
    __shared__ uint32_t tot[kSWarps + 1];
    __shared__ unsigned long long s_w;
    __shared__ int s_flag;
    const int total_bits = 2 * K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) { s_w = atomicAdd(work_counter, 1ull); s_flag = 0; }
        __syncthreads();
        const uint64_t si = s_w;
        if (si >= nsegs) return;
        const Seg s = segs[si];
        uint64_t *gsrc = ((s.bb & 1) ? buf1 : buf0) + s.start * NW;
        uint64_t *gpartner = ((s.bb & 1) ? buf0 : buf1) + s.start * NW;
        uint32_t *gcnt = reinterpret_cast<uint32_t *>(gpartner);
        if (s.len == 0) { if (threadIdx.x == 0) ndist[si] = 0; __syncthreads(); continue; }
        if (s.len > (uint64_t)CAP) {
            if (threadIdx.x == 0) { gcnt[0] = (uint32_t)s.len; ndist[si] = 1; if (s.bits < (uint32_t)total_bits) atomicAdd(&stats[0], 1ull); }
            __syncthreads();
            continue;
        }
        const uint32_t n = (uint32_t)s.len;
        const int lo = (int)s.bits;
        const int r2 = (total_bits - lo) < kBinBits ? (total_bits - lo) : kBinBits;
        const uint32_t nb = 1u << r2;
        for (uint32_t i = threadIdx.x; i < nb; i += kSThreads) { rep[i] = 0xffffffffu; repcnt[i] = 0; rhist[i] = 0; dres[i] = 0; }
        __syncthreads();
        // ---- P1: load, elect representatives
        uint32_t dig[IPT];
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t i = threadIdx.x + j * kSThreads;
            dig[j] = 0;
            if (i < n) {
                Kmer<NW> k = load_rec<NW>(gsrc + (size_t)i * NW);
                store_rec<NW>(A + (size_t)i * NW, k);
                dig[j] = r2 ? key_bits<NW>(k, K, lo, r2) : 0u;
                atomicMin(&rep[dig[j]], i);
            }
        }
        __syncthreads();
        // ---- P2: copies of the representative only count; everything else is residual
        uint32_t resmask = 0;
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t i = threadIdx.x + j * kSThreads;
            if (i < n) {
                const uint32_t r = rep[dig[j]];
                if (r != i) {
                    if (kmer_eq<NW>(load_rec<NW>(A + (size_t)i * NW), load_rec<NW>(A + (size_t)r * NW))) atomicAdd(&repcnt[dig[j]], 1u);
                    else { atomicAdd(&rhist[dig[j]], 1u); resmask |= 1u << j; }
                }
            }
        }
        __syncthreads();
        // ---- P3: exclusive scan of the residual counts
        const uint32_t b0 = threadIdx.x * BPT;
        uint32_t loc[BPT];
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) { loc[q] = (b0 + q < nb) ? rhist[b0 + q] : 0; sum += loc[q]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) tot[warp] = inc;
        __syncthreads();
        uint32_t wb = 0, rtotal = 0;
        for (int w = 0; w < kSWarps; ++w) { if (w < warp) wb += tot[w]; rtotal += tot[w]; }
        {
            uint32_t run = wb + inc - sum;
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                if (b0 + q < nb) { rstart[b0 + q] = (uint16_t)run; rhist[b0 + q] = run; }
                run += loc[q];
            }
            if (threadIdx.x == kSThreads - 1) rstart[nb] = (uint16_t)rtotal;
        }
        bool bad = rtotal > (uint32_t)kResCap;
        __syncthreads();
        // ---- P4: scatter the residual records into their bins
        if (!bad) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                if (resmask & (1u << j)) {
                    const uint32_t i = threadIdx.x + j * kSThreads;
                    const uint32_t pos = atomicAdd(&rhist[dig[j]], 1u);
                    store_rec<NW>(R + (size_t)pos * NW, load_rec<NW>(A + (size_t)i * NW));
                }
            }
        }
        __syncthreads();
        // ---- P5: the thread that owns a bin's representative deduplicates + sorts that bin's residual (tiny) and records
        //          how many distinct residual keys the bin has. Work is per occupied bin, not per bin.
        if (!bad) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const uint32_t i = threadIdx.x + j * kSThreads;
                if (i >= n || rep[dig[j]] != i) continue;
                const uint32_t b = dig[j];
                const uint32_t bs = rstart[b], be = rstart[b + 1];
                if (be == bs) continue;
                uint32_t d = 0;
                if (be - bs == 1) { rcnt[bs] = 1; d = 1; }
                else {
                    uint32_t rem_end = be, p = bs, work = 0;
                    while (p < rem_end) {
                        const Kmer<NW> key = load_rec<NW>(R + (size_t)p * NW);
                        uint32_t c = 1, w = p + 1;
                        for (uint32_t z = p + 1; z < rem_end; ++z) {
                            const Kmer<NW> x = load_rec<NW>(R + (size_t)z * NW);
                            if (kmer_eq<NW>(x, key)) ++c;
                            else { if (w != z) store_rec<NW>(R + (size_t)w * NW, x); ++w; }
                        }
                        work += rem_end - p;
                        rcnt[p] = (uint16_t)c;
                        rem_end = w;
                        ++p;
                        if (work > 1024u || p - bs > 16u) { bad = true; break; }
                    }
                    if (bad) break;
                    d = p - bs;
                    for (uint32_t a = bs + 1; a < bs + d; ++a) {
                        const Kmer<NW> key = load_rec<NW>(R + (size_t)a * NW);
                        const uint16_t kc = rcnt[a];
                        uint32_t z = a;
                        while (z > bs && kmer_word_cmp<NW>(load_rec<NW>(R + (size_t)(z - 1) * NW), key) > 0) {
                            store_rec<NW>(R + (size_t)z * NW, load_rec<NW>(R + (size_t)(z - 1) * NW));
                            rcnt[z] = rcnt[z - 1];
                            --z;
                        }
                        store_rec<NW>(R + (size_t)z * NW, key);
                        rcnt[z] = kc;
                    }
                }
                dres[b] = (uint16_t)d;
            }
        }
        if (bad) s_flag = 1;
        __syncthreads();
        if (s_flag) {
            // exact fallback: LSD radix over every remaining key bit (scratch = the partner buffer's region), run-length unique
            if (threadIdx.x == 0) atomicAdd(&stats[1], 1ull);
            __syncthreads();
            uint64_t *S = lsd_sort_range<NW>(A, gpartner, n, K, lo, total_bits, lsdcnt, tot);
            if (S != A) {
                for (uint32_t i = threadIdx.x; i < n * NW; i += kSThreads) A[i] = S[i];
                __syncthreads();
            }
            uint32_t *heads = rep;      // the sort is done: the 3*kBins u32 of rep/repcnt/rhist (>= CAP+1) are free
            __syncthreads();
            const uint32_t per = (n + kSThreads - 1) / kSThreads;
            const uint32_t i0 = threadIdx.x * per, i1 = min(n, i0 + per);
            uint32_t nh = 0;
            for (uint32_t i = i0; i < i1; ++i)
                if (i == 0 || kmer_word_cmp<NW>(load_rec<NW>(A + (size_t)(i - 1) * NW), load_rec<NW>(A + (size_t)i * NW)) != 0) ++nh;
            uint32_t hinc = nh;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, hinc, o);
                if (lane >= o) hinc += t;
            }
            if (lane == 31) tot[warp] = hinc;
            __syncthreads();
            uint32_t hb = 0, all = 0;
            for (int w = 0; w < kSWarps; ++w) { if (w < warp) hb += tot[w]; all += tot[w]; }
            uint32_t j = hb + hinc - nh;
            for (uint32_t i = i0; i < i1; ++i)
                if (i == 0 || kmer_word_cmp<NW>(load_rec<NW>(A + (size_t)(i - 1) * NW), load_rec<NW>(A + (size_t)i * NW)) != 0) heads[j++] = i;
            if (threadIdx.x == 0) { heads[all] = n; ndist[si] = all; }
            __syncthreads();
            for (uint32_t q = threadIdx.x; q < all; q += kSThreads) {
                const uint32_t h = heads[q];
                store_rec<NW>(gsrc + (size_t)q * NW, load_rec<NW>(A + (size_t)h * NW));
                gcnt[q] = heads[q + 1] - h;
            }
            __syncthreads();
            continue;
        }
        // ---- P6: output offset of every bin (exclusive scan of 1 + residual-distinct over the occupied bins), then the
        //          representative's owner writes its bin in key order (representative merged into the sorted residual)
        uint32_t dsum = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
            loc[q] = (b0 + q < nb && rep[b0 + q] != 0xffffffffu) ? 1u + dres[b0 + q] : 0u;
            dsum += loc[q];
        }
        uint32_t dinc = dsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, dinc, o);
            if (lane >= o) dinc += t;
        }
        if (lane == 31) tot[warp] = dinc;
        __syncthreads();
        uint32_t db = 0, all = 0;
        for (int w = 0; w < kSWarps; ++w) { if (w < warp) db += tot[w]; all += tot[w]; }
        {
            uint32_t run = db + dinc - dsum;
#pragma unroll
            for (int q = 0; q < BPT; ++q) {
                if (b0 + q < nb) rhist[b0 + q] = run;        // rhist is free after P4: output offset of the bin
                run += loc[q];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const uint32_t i = threadIdx.x + j * kSThreads;
            if (i >= n || rep[dig[j]] != i) continue;
            const uint32_t b = dig[j];
            const Kmer<NW> rk = load_rec<NW>(A + (size_t)i * NW);
            const uint32_t rc = repcnt[b] + 1;
            const uint32_t bs = rstart[b], d = dres[b];
            uint32_t o = rhist[b];
            bool placed = false;
            for (uint32_t z = 0; z < d; ++z) {
                const Kmer<NW> x = load_rec<NW>(R + (size_t)(bs + z) * NW);
                if (!placed && kmer_word_cmp<NW>(rk, x) < 0) {
                    store_rec<NW>(gsrc + (size_t)o * NW, rk); gcnt[o] = rc; ++o; placed = true;
                }
                store_rec<NW>(gsrc + (size_t)o * NW, x); gcnt[o] = rcnt[bs + z]; ++o;
            }
            if (!placed) { store_rec<NW>(gsrc + (size_t)o * NW, rk); gcnt[o] = rc; }
        }
        if (threadIdx.x == 0) ndist[si] = all;
        __syncthreads();
    }
}

// (Round 2 measured a fourth generation -- (21-bit digit | index) keys sorted with three ballot-ranked 7-bit LSD passes, equal
// records found as neighbours, no shared-memory atomics at all. Parity clean and 1.7x SLOWER than the kernel above: 427 vs 250 ms
// at 100 M reads, 89 vs 50 ms at 20 M. profiles/r02e_local_sort4_AB.log. Removed together with the sort primitive.)

// compaction: one warp per segment copies its distinct records / counts to the dense output
template <int NW>
__global__ void compact_k(const Seg *__restrict__ segs, uint64_t nsegs, const uint32_t *__restrict__ ndist, const uint64_t *__restrict__ dbase,
                          const uint64_t *__restrict__ buf0, const uint64_t *__restrict__ buf1, int K, int want_counts, int double_selfrc,
                          uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_counts, unsigned long long *__restrict__ bucket_sizes) {
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= nsegs) return;
    const Seg s = segs[wid];
    const uint32_t nd = ndist[wid];
    if (nd == 0) return;
    const uint64_t *src = ((s.bb & 1) ? buf1 : buf0) + s.start * NW;
    const uint32_t *csrc = reinterpret_cast<const uint32_t *>(((s.bb & 1) ? buf0 : buf1) + s.start * NW);
    const uint64_t ob = dbase[wid];
    for (uint32_t q = lane; q < nd; q += 32) {
        Kmer<NW> k = load_rec<NW>(src + (size_t)q * NW);
        store_rec<NW>(out_keys + (ob + q) * NW, k);
        if (want_counts) {
            uint32_t c = csrc[q];
            if (double_selfrc && kmer_eq<NW>(k, kmer_rc<NW>(k, K))) c *= 2u;   // a self-RC (k+1)-mer is seen in the read and in its RC
            out_counts[ob + q] = c;
        }
    }
    if (lane == 0) atomicAdd(&bucket_sizes[s.bb >> 1], (unsigned long long)nd);
}

// ------------------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------------------
struct Timer {
    cudaEvent_t a, b; cudaStream_t s;
    Timer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

// Environment, read ONCE per process and clamped to what the kernels support. User options: SGPU_ARENA_GB (sgpu_internal.h) and
// SGPU_TRACE (per-phase wall clock on stderr). Tuning knobs kept for A/B runs on the GPU box:
//   SGPU_PA_MAX   level-A partitions for the whole job (default 4096; every (CTA, partition) pair is an open write stream)
//   SGPU_RMAX     key bits per refinement round (default 11 = 2048 bins per CTA)
//   SGPU_A_SUB    partition sub-ranges per level-A scatter pass (default 0 = automatic)
struct Tuning {
    uint32_t pa_max = 4096;
    uint32_t rmax = 11;
    int a_sub = 0;
    bool trace = false;
};
static const Tuning &tuning() {
    static const Tuning t = [] {
        Tuning x;
        if (const char *e = getenv("SGPU_PA_MAX")) x.pa_max = (uint32_t)std::min(8192, std::max(1, atoi(e)));
        if (const char *e = getenv("SGPU_RMAX")) x.rmax = (uint32_t)std::min(11, std::max(1, atoi(e)));
        if (const char *e = getenv("SGPU_A_SUB")) x.a_sub = std::min(64, std::max(0, atoi(e)));
        x.trace = getenv("SGPU_TRACE") != nullptr;
        return x;
    }();
    return t;
}

struct Trace {
    bool on; cudaStream_t st; double t0;
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    Trace(cudaStream_t s) : on(tuning().trace), st(s), t0(now()) {}
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        double t = now();
        fprintf(stderr, "[sgpu-trace] %-28s %9.3f ms\n", what, t - t0);
        t0 = t;
    }
};

static int ilog2_floor(uint64_t v) { int r = 0; while (v >>= 1) ++r; return r; }

static const int kLevelAMaxParts = 8192;      // partitions one level-A launch can address (shared-memory histogram / cursor tables)

// mean segment length the refinement aims for: 3/4 of the local-sort capacity
template <int NW> static uint32_t sort_target() { return (uint32_t)SortCfg<NW>::CAP * 6 / 8; }

// refinement + local sort + compaction of one pass: X holds the level-A output (CTA-major pieces when `pieces` is given, else
// partition-major with the given starts/totals), Y is the ping-pong partner of the same size.
template <int NW>
static void sort_pass(Ctx *ctx, int K, DArr<uint64_t> &X, DArr<uint64_t> &Y, const uint64_t *part_start_p, const uint64_t *part_total_p, uint32_t PA,
                      int rA_, uint32_t b_lo, int b_hi, int64_t first, bool want_counts, bool double_selfrc, unsigned long long *d_bsz_p, Chunk &ch_out,
                      Timer &tm, Trace &tr, const Pieces *pieces = nullptr) {
    constexpr int CAP = SortCfg<NW>::CAP;
    const int total_bits = 2 * K;
    cudaStream_t st = ctx->stream;
    // ---- segments + refinement rounds
    uint64_t nsegs = PA;
    DArr<Seg> segs(ctx, nsegs);
    seg_init_k<<<div_up(PA, 256), 256, 0, st>>>(part_start_p, part_total_p, PA, rA_, b_lo, segs.p);
    ctx->launches++;
    RefinePlan rp; rp.cap = CAP; rp.target = sort_target<NW>(); rp.total_bits = total_bits;
    rp.rmax = tuning().rmax;
    DArr<unsigned long long> wcounter(ctx, 4);
    tm.start();
    for (int round = 0; round < 300; ++round) {
        DArr<uint32_t> nchild(ctx, nsegs + 1), isw(ctx, nsegs + 1);
        DArr<uint64_t> cbase(ctx, nsegs + 1), wpos(ctx, nsegs + 1);
        SG_CUDA(cudaMemsetAsync(nchild.p + nsegs, 0, 4, st));
        SG_CUDA(cudaMemsetAsync(isw.p + nsegs, 0, 4, st));
        const bool pieced = pieces && round == 0;       // X holds the CTA-major staging buffer: round 0 gathers every partition
        refine_plan_k<<<div_up(nsegs, 256), 256, 0, st>>>(segs.p, nsegs, rp, nchild.p, isw.p, pieced ? 1 : 0);
        ctx->launches++;
        exclusive_scan_u32_to_u64(ctx, nchild.p, cbase.p, nsegs + 1);
        exclusive_scan_u32_to_u64(ctx, isw.p, wpos.p, nsegs + 1);
        uint64_t tot[2];
        SG_CUDA(cudaMemcpyAsync(&tot[0], cbase.p + nsegs, 8, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(&tot[1], wpos.p + nsegs, 8, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        if (tot[1] == 0) break;
        DArr<Seg> nsegs_arr(ctx, tot[0]);
        DArr<uint64_t> worklist(ctx, tot[1]);
        refine_copy_k<<<div_up(nsegs, 256), 256, 0, st>>>(segs.p, nsegs, isw.p, cbase.p, wpos.p, nsegs_arr.p, worklist.p);
        ctx->launches++;
        SG_CUDA(cudaMemsetAsync(wcounter.p, 0, 8, st));
        const int grid = (int)std::min<uint64_t>(tot[1], (uint64_t)ctx->num_sms * 2);
        // 16-byte records: the scatter phase pairs the two records of a 32-byte sector (B200, 100 M reads: 285 -> 185 ms)
        if (NW == 2) {
            const size_t sm = (size_t)kRMaxBins * sizeof(PmBox);
            if (pieced) {
                SG_CUDA(cudaFuncSetAttribute(refine_k<NW, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                refine_k<NW, true, true><<<grid, kRThreads, sm, st>>>(segs.p, worklist.p, tot[1], cbase.p, rp, K, X.p, Y.p, nsegs_arr.p, wcounter.p, *pieces);
            } else {
                SG_CUDA(cudaFuncSetAttribute(refine_k<NW, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                refine_k<NW, false, true><<<grid, kRThreads, sm, st>>>(segs.p, worklist.p, tot[1], cbase.p, rp, K, X.p, Y.p, nsegs_arr.p, wcounter.p, Pieces());
            }
        } else if (pieced) {
            refine_k<NW, true, false><<<grid, kRThreads, 0, st>>>(segs.p, worklist.p, tot[1], cbase.p, rp, K, X.p, Y.p, nsegs_arr.p, wcounter.p, *pieces);
        } else {
            refine_k<NW, false, false><<<grid, kRThreads, 0, st>>>(segs.p, worklist.p, tot[1], cbase.p, rp, K, X.p, Y.p, nsegs_arr.p, wcounter.p, Pieces());
        }
        ctx->launches++;
        SG_CUDA(cudaGetLastError());
        tr.mark("refine round");
        segs = std::move(nsegs_arr);
        nsegs = tot[0];
    }
    ctx->times.refine += tm.stop();
    // ---- local sort
    DArr<uint32_t> ndist(ctx, nsegs + 1);
    DArr<unsigned long long> stats(ctx, 4);
    SG_CUDA(cudaMemsetAsync(ndist.p, 0, ndist.bytes(), st));
    SG_CUDA(cudaMemsetAsync(stats.p, 0, 32, st));
    SG_CUDA(cudaMemsetAsync(wcounter.p, 0, 8, st));
    tm.start();
    {
        auto launch = [&](auto kernel, size_t smem) {
            SG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 1;
            SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSThreads, smem));
            if (occ < 1) occ = 1;
            int grid = (int)std::min<uint64_t>(nsegs, (uint64_t)ctx->num_sms * occ);
            if (grid < 1) grid = 1;
            kernel<<<grid, kSThreads, smem, st>>>(segs.p, nsegs, K, X.p, Y.p, ndist.p, wcounter.p, stats.p);
        };
        constexpr int bins = 1 << kSortBinBits;
        launch(local_sort3_k<NW, kSortBinBits, CAP>, (size_t)(CAP + kResCap) * NW * sizeof(uint64_t) + (size_t)3 * bins * sizeof(uint32_t) +
                                                         ((size_t)2 * bins + 2 + kResCap) * sizeof(uint16_t) + 16);
        ctx->launches++;
        SG_CUDA(cudaGetLastError());
    }
    DArr<uint64_t> dbase(ctx, nsegs + 1);
    exclusive_scan_u32_to_u64(ctx, ndist.p, dbase.p, nsegs + 1);
    uint64_t D = 0;
    unsigned long long h_stats[4];
    SG_CUDA(cudaMemcpyAsync(&D, dbase.p + nsegs, 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_stats, stats.p, 32, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    ctx->times.local_sort += tm.stop();
    tr.mark("local sort");
    SG_CHECK(h_stats[0] == 0, 6, "internal: oversize segment with unfixed key bits reached the local sort");
    // ---- compaction into the dense chunk
    Chunk ch;
    ch.n = (int64_t)D; ch.b_lo = (int)b_lo; ch.b_hi = b_hi; ch.first = first;
    ch.keys.alloc(ctx, (size_t)D * NW + 2, true);
    if (want_counts) ch.counts.alloc(ctx, (size_t)D + 1, true);
    tm.start();
    if (nsegs) {
        compact_k<NW><<<div_up((int64_t)nsegs * 32, 256), 256, 0, st>>>(segs.p, nsegs, ndist.p, dbase.p, X.p, Y.p, K, want_counts ? 1 : 0,
                                                                     double_selfrc ? 1 : 0, ch.keys.p, ch.counts.p, d_bsz_p);
        ctx->launches++;
        SG_CUDA(cudaGetLastError());
    }
    ctx->times.compact += tm.stop();
    tr.mark("compact");
    ch_out = std::move(ch);
}

// ---- level A as a job: one histogram (+ partition id) pass over the source, then one scatter per bucket-group pass -----------
// Shared by the single-GPU count and the multi-GPU count (where a "pass" scatters this rank's shard for the pass's buckets).
template <int NW, class Src>
struct LevelAJob {
    Ctx *ctx = nullptr;
    std::vector<Src> srcs;
    int K = 0, B = 0, rA = 0, G = 0;
    int s_lo = 0, s_hi = 0;               // buckets covered by this job (one histogram super-range)
    uint32_t PA_all = 0;                  // (s_hi - s_lo) << rA
    bool roll = false, use_ids = false;
    DArr<uint32_t> blk_counts;            // [G][PA_all] records of (CTA, partition)
    DArr<uint64_t> part_total_all;        // [PA_all]
    std::vector<DArr<uint64_t>> tile_off; // per source: first id of every tile
    std::vector<DArr<uint16_t>> ids;      // per source: 2-byte partition id per record slot
    std::vector<uint64_t> h_part;         // host copy of part_total_all
    uint64_t bucket_records(int b) const {
        uint64_t s = 0;
        for (uint32_t q = 0; q < (1u << rA); ++q) s += h_part[((size_t)(b - s_lo) << rA) + q];
        return s;
    }
};

// total fan-out bits wanted for est_records, minus what the bucket function provides, clamped to the tables
// CTAs of the level-A grid per SM: the two that are resident. (More waves -- 3, 4, 6 per SM -- were measured for tail balance:
// 170 / 164 / 156 ms against 162 ms for the partition kernel at 100 M reads, with 3x smaller pieces for the gather. Not kept.)
static int levelA_ctas_per_sm() { return 2; }
static int levelA_key_bits(uint64_t est_records, int B, int total_bits, uint32_t target, uint32_t pa_max) {
    const int want = ilog2_floor(est_records / target + 1) + 1;
    const int bbits = ilog2_floor((uint64_t)B) + (((1u << ilog2_floor((uint64_t)B)) < (uint32_t)B) ? 1 : 0);
    int rA = want - bbits;
    if (rA < 0) rA = 0;
    while (rA > 0 && ((uint64_t)B << rA) > pa_max) --rA;
    if (rA > total_bits) rA = total_bits;
    if (rA > 13) rA = 13;                 // a histogram super-range holds at least one bucket: (1 << rA) <= kLevelAMaxParts
    return rA;
}

template <int NW, class Src>
static void levelA_count(LevelAJob<NW, Src> &job, uint64_t est_records, Timer &tm, Trace &tr) {
    Ctx *ctx = job.ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t PA_all = job.PA_all;
    const int G = job.G;
    SG_CHECK(PA_all >= 1 && PA_all <= (uint32_t)kLevelAMaxParts, 6, "internal: level-A partition table too large");
    LevelA pa_all;
    pa_all.K = job.K; pa_all.B = (uint32_t)job.B; pa_all.b_lo = (uint32_t)job.s_lo; pa_all.b_hi = (uint32_t)job.s_hi; pa_all.rA = job.rA; pa_all.PA = PA_all;
    job.blk_counts.alloc(ctx, (size_t)G * PA_all);
    job.part_total_all.alloc(ctx, (size_t)PA_all + 1);
    job.h_part.assign(PA_all, 0);
    SG_CUDA(cudaMemsetAsync(job.blk_counts.p, 0, job.blk_counts.bytes(), st));
    job.tile_off.clear(); job.ids.clear();
    job.tile_off.resize(job.srcs.size()); job.ids.resize(job.srcs.size());
    // per-record partition ids (2 bytes per record slot): only when they fit comfortably next to the sort buffers
    job.use_ids = PA_all < 0xffffu && (double)est_records * 2.0 < (double)ctx->free_bytes() * 0.20;
    if (job.use_ids) {
        for (size_t si = 0; si < job.srcs.size(); ++si) {
            const Src &src = job.srcs[si];
            if (src.n == 0) continue;
            const int64_t ntiles = (src.n + (job.roll ? kRollTile : kATile) - 1) / (job.roll ? kRollTile : kATile);
            DArr<uint32_t> ttot(ctx, (size_t)ntiles + 1);
            job.tile_off[si].alloc(ctx, (size_t)ntiles + 1);
            SG_CUDA(cudaMemsetAsync(ttot.p + ntiles, 0, 4, st));
            if (job.roll) {
                if constexpr (std::is_same<Src, ReadsSrc>::value) roll_tile_ids_k<<<div_up(ntiles, 8), 256, 0, st>>>(src, ntiles, ttot.p);
            } else {
                tile_totals_k<Src><<<div_up(ntiles, 8), 256, 0, st>>>(src, ntiles, ttot.p, 8u);
            }
            ctx->launches++;
            exclusive_scan_u32_to_u64(ctx, ttot.p, job.tile_off[si].p, (size_t)ntiles + 1);
            uint64_t nrec_src = 0;
            SG_CUDA(cudaMemcpyAsync(&nrec_src, job.tile_off[si].p + ntiles, 8, cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaStreamSynchronize(st));
            job.ids[si].alloc(ctx, (size_t)nrec_src + 8);
        }
    }
    tm.start();
    for (size_t si = 0; si < job.srcs.size(); ++si) {
        const Src &src = job.srcs[si];
        if (src.n == 0) continue;
        if (job.roll) {
            if constexpr (std::is_same<Src, ReadsSrc>::value) {
                SG_CUDA(cudaFuncSetAttribute(levelA_count_roll_k<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roll_smem_bytes(PA_all)));
                levelA_count_roll_k<NW><<<G, kRollThreads, roll_smem_bytes(PA_all), st>>>(src, pa_all, job.blk_counts.p, job.tile_off[si].p, job.ids[si].p);
            }
        } else {
            SG_CUDA(cudaFuncSetAttribute(levelA_count_k<NW, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PA_all * sizeof(uint32_t))));
            levelA_count_k<NW, Src><<<G, kAThreads, PA_all * sizeof(uint32_t), st>>>(src, pa_all, job.blk_counts.p, job.tile_off[si].p, job.ids[si].p);
        }
        ctx->launches++;
    }
    SG_CUDA(cudaGetLastError());
    levelA_totals_k<<<div_up(PA_all, 256), 256, 0, st>>>(job.blk_counts.p, PA_all, G, job.part_total_all.p);
    ctx->launches++;
    SG_CUDA(cudaMemcpyAsync(job.h_part.data(), job.part_total_all.p, (size_t)PA_all * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    ctx->times.extract_count += tm.stop();
    tr.mark("A1 count+totals");
}

// Scatter the records of buckets [b_lo, b_hi) into X, CTA-major (see `Pieces`): CTA g of the level-A grid owns one contiguous
// region, partitioned inside. pbase_buf ([G][PA], caller-owned) receives the first record of every (CTA, partition) piece.
// I_pass / total_records only steer the number of partition sub-ranges.
template <int NW, class Src>
static void levelA_scatter(LevelAJob<NW, Src> &job, int b_lo, int b_hi, uint64_t I_pass, uint64_t total_records, uint64_t *X, uint64_t *pbase_buf,
                           Pieces &pcs, Timer &tm, Trace &tr) {
    Ctx *ctx = job.ctx;
    cudaStream_t st = ctx->stream;
    const int G = job.G;
    const uint32_t PA_all = job.PA_all;
    const uint32_t p_lo = (uint32_t)(b_lo - job.s_lo) << job.rA;
    LevelA pa;
    pa.K = job.K; pa.B = (uint32_t)job.B; pa.b_lo = (uint32_t)b_lo; pa.b_hi = (uint32_t)b_hi; pa.rA = job.rA; pa.PA = (uint32_t)(b_hi - b_lo) << job.rA;
    const uint32_t PA = pa.PA;
    SG_CHECK(PA >= 1 && PA <= (uint32_t)kLevelAMaxParts && G <= kMaxPieces, 6, "internal: level-A pass geometry");
    DArr<uint64_t> base(ctx, (size_t)G * PA);
    {
        DArr<uint64_t> row_total(ctx, (size_t)G + 1), row_start(ctx, (size_t)G + 1);
        SG_CUDA(cudaMemsetAsync(row_total.p + G, 0, 8, st));
        stage_rows_k<<<G, 256, 0, st>>>(job.blk_counts.p + p_lo, PA_all, PA, base.p, row_total.p);
        exclusive_scan_u64(ctx, row_total.p, row_start.p, (size_t)G + 1);
        stage_add_k<<<div_up((int64_t)G * PA, 256), 256, 0, st>>>(base.p, row_start.p, PA, G);
        ctx->launches += 2;
        SG_CUDA(cudaMemcpyAsync(pbase_buf, base.p, (size_t)G * PA * 8, cudaMemcpyDeviceToDevice, st));   // the scatter advances `base`
    }
    pcs.pbase = pbase_buf; pcs.cnt = job.blk_counts.p + p_lo; pcs.cnt_stride = PA_all; pcs.PA = PA; pcs.G = G;
    tm.start();
    if (job.roll) {
        if constexpr (std::is_same<Src, ReadsSrc>::value) {
            const size_t smem = roll_smem_bytes(PA);
            SG_CUDA(cudaFuncSetAttribute(levelA_scatter_roll_k<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SG_CUDA(cudaFuncSetAttribute(levelA_scatter_roll_k<NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            {
            // partition sub-ranges (only with the id array, where a foreign window costs just the roll): fewer lines and pages open
            // per CTA. A sub-range costs one more roll over ALL windows of the source: worth it when the pass holds most of the job's
            // records (measured: 20 M reads / 1 pass: 46.5 / 36.7 / 31.5 ms with 1 / 2 / 4 sub-ranges; 100 M reads / 5 passes: 184 ms
            // with 1, 353 ms with 2)
            int nsub_auto = (int)(4.0 * (double)I_pass / (double)std::max<uint64_t>(1, total_records) + 0.5);
            nsub_auto = std::min(4, std::max(1, nsub_auto));
            uint32_t nsub = job.use_ids ? (uint32_t)(tuning().a_sub ? tuning().a_sub : nsub_auto) : 1u;
            if (nsub > PA) nsub = PA;
            for (uint32_t sb = 0; sb < nsub; ++sb) {
                const uint32_t q_lo = (uint32_t)((uint64_t)PA * sb / nsub), q_hi = (uint32_t)((uint64_t)PA * (sb + 1) / nsub);
                if (q_hi == q_lo) continue;
                LevelA pa_sub = pa;
                pa_sub.PA = q_hi - q_lo;
                const size_t smem_sub = roll_smem_bytes(pa_sub.PA);
                for (size_t si = 0; si < job.srcs.size(); ++si) {
                    const Src &src = job.srcs[si];
                    if (src.n == 0) continue;
                    if (job.use_ids)
                        levelA_scatter_roll_k<NW, true><<<G, kRollThreads, smem_sub, st>>>(src, pa_sub, base.p, X, job.tile_off[si].p, job.ids[si].p, p_lo + q_lo, PA, q_lo, (uint64_t)job.ids[si].n);
                    else
                        levelA_scatter_roll_k<NW, false><<<G, kRollThreads, smem, st>>>(src, pa, base.p, X, nullptr, nullptr, 0u, PA, 0u, 0ull);
                    ctx->launches++;
                }
            }
            }
        }
    } else {
        const size_t smem = (size_t)PA * (sizeof(uint64_t) + sizeof(uint32_t));
        SG_CUDA(cudaFuncSetAttribute(levelA_scatter_k<NW, Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (size_t si = 0; si < job.srcs.size(); ++si) {
            const Src &src = job.srcs[si];
            if (src.n == 0) continue;
            levelA_scatter_k<NW, Src><<<G, kAThreads, smem, st>>>(src, pa, base.p, X, job.use_ids ? job.tile_off[si].p : nullptr,
                                                                 job.use_ids ? job.ids[si].p : nullptr, p_lo);
            ctx->launches++;
        }
    }
    SG_CUDA(cudaGetLastError());
    ctx->times.extract_scatter += tm.stop();       // synchronises: `base` may go out of scope
    tr.mark("A2 scatter");
}

// what one pass needs next to what is already resident: X + Y + its own output (distinct/instances <= 0.6 assumed, checked
// against the arena when the output is allocated)
static double pass_bytes_needed(uint64_t recs, size_t W) { return (double)recs * W * 2.0 + (double)recs * (W + 4) * 0.6 + (64 << 20); }

template <int NW, class Src>
static void run_count(Ctx *ctx, const std::vector<Src> &srcs, int K, int B, bool want_counts, bool double_selfrc, uint64_t est_records, KSet *out) {
    const int total_bits = 2 * K;
    const size_t W = 8 * NW;
    cudaStream_t st = ctx->stream;
    Timer tm(st);
    Trace tr(st);
    constexpr bool kIsReads = std::is_same<Src, ReadsSrc>::value;

    out->bsz.assign(B, 0);
    DArr<unsigned long long> d_bsz(ctx, B);
    SG_CUDA(cudaMemsetAsync(d_bsz.p, 0, B * sizeof(unsigned long long), st));

    // ---- level-A geometry for the whole job. Partition id = (bucket, top rA key bits). ONE histogram pass over the
    // source serves every bucket-group pass (the groups are contiguous partition ranges), so a multi-pass job hashes
    // the source once, and passes are planned from exact per-bucket record counts.
    const int rA = levelA_key_bits(est_records, B, total_bits, sort_target<NW>(), tuning().pa_max);
    const int SR = std::max(1, kLevelAMaxParts >> rA);      // buckets per histogram super-range
    int64_t first = 0;
    for (int s_lo = 0; s_lo < B; s_lo += SR) {
        LevelAJob<NW, Src> job;
        job.ctx = ctx; job.srcs = srcs; job.K = K; job.B = B; job.rA = rA; job.G = ctx->num_sms * levelA_ctas_per_sm();
        job.s_lo = s_lo; job.s_hi = std::min(B, s_lo + SR);
        job.PA_all = (uint32_t)(job.s_hi - job.s_lo) << rA;
        if constexpr (kIsReads) job.roll = !srcs.empty() && !srcs[0].both;
        levelA_count(job, est_records, tm, tr);
        // bucket-group passes: simulate the greedy "as many whole buckets as fit" plan to learn how many passes are needed, then
        // aim for equally sized passes (a tiny last pass still costs a full scan of the source)
        auto current_limit = [&]() {
            return ctx->hbm_budget ? (ctx->hbm_budget > ctx->allocated ? ctx->hbm_budget - ctx->allocated : 0) : (size_t)(ctx->free_bytes() * 0.90);
        };
        uint64_t total_records = 0;
        for (int b = job.s_lo; b < job.s_hi; ++b) total_records += job.bucket_records(b);
        uint64_t pass_target = total_records;
        {
            double lim_sim = (double)current_limit();
            int npass_sim = 0, b = job.s_lo;
            while (b < job.s_hi) {
                uint64_t I = 0; const int b0 = b;
                while (b < job.s_hi) {
                    const uint64_t ib = job.bucket_records(b);
                    if (b > b0 && pass_bytes_needed(I + ib, W) > lim_sim) break;
                    I += ib; ++b;
                }
                lim_sim -= (double)I * (W + 4) * 0.5;            // this pass's output stays resident
                ++npass_sim;
            }
            pass_target = total_records / (uint64_t)npass_sim + total_records / 64 + 1;
        }
        int b_lo = job.s_lo;
        while (b_lo < job.s_hi) {
            // ---- plan this pass: whole buckets that fit next to what is already resident (X + Y + its own output)
            const size_t lim = current_limit();
            int b_hi = b_lo;
            uint64_t I = 0;
            while (b_hi < job.s_hi) {
                const uint64_t ib = job.bucket_records(b_hi);
                if (b_hi > b_lo && (pass_bytes_needed(I + ib, W) > (double)lim || I + ib > pass_target)) break;
                I += ib; ++b_hi;
            }
            const uint32_t p_lo = (uint32_t)(b_lo - job.s_lo) << rA;
            const uint32_t PA = (uint32_t)(b_hi - b_lo) << rA;
            DArr<uint64_t> part_total(ctx, PA + 1), part_start(ctx, PA + 1);
            SG_CUDA(cudaMemcpyAsync(part_total.p, job.part_total_all.p + p_lo, (size_t)PA * 8, cudaMemcpyDeviceToDevice, st));
            SG_CUDA(cudaMemsetAsync(part_total.p + PA, 0, 8, st));
            exclusive_scan_u64(ctx, part_total.p, part_start.p, PA + 1);
            ctx->times.passes++;
            ctx->times.instances += I;
            DArr<uint64_t> X(ctx, (size_t)I * NW + 2), Y(ctx, (size_t)I * NW + 2);
            DArr<uint64_t> pbase(ctx, (size_t)job.G * PA);
            Pieces pcs;
            levelA_scatter(job, b_lo, b_hi, I, total_records, X.p, pbase.p, pcs, tm, tr);
            Chunk ch;
            sort_pass<NW>(ctx, K, X, Y, part_start.p, part_total.p, PA, rA, (uint32_t)b_lo, b_hi, first, want_counts, double_selfrc, d_bsz.p, ch, tm, tr, &pcs);
            first += ch.n;
            out->chunks.push_back(std::move(ch));
            b_lo = b_hi;
        }
    }
    std::vector<unsigned long long> hb(B);
    SG_CUDA(cudaMemcpyAsync(hb.data(), d_bsz.p, B * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    out->bstart.assign(B + 1, 0);
    for (int b = 0; b < B; ++b) { out->bsz[b] = (int64_t)hb[b]; out->bstart[b + 1] = out->bstart[b] + out->bsz[b]; }
    out->n = first;
    SG_CHECK(out->bstart[B] == out->n, 6, "internal: bucket sizes do not add up");
}

__global__ void count_windows_k(const uint32_t *__restrict__ lens, int64_t n, int K, unsigned long long *__restrict__ out) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int L = (int)lens[i];
        if (L >= K) s += (unsigned long long)(L - K + 1);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// windows of the context's read set (exact), for the level-A geometry
static uint64_t count_windows(Ctx *ctx, int K) {
    DArr<unsigned long long> d_w(ctx, 1);
    SG_CUDA(cudaMemsetAsync(d_w.p, 0, 8, ctx->stream));
    if (ctx->n_reads) {
        count_windows_k<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->d_lens, ctx->n_reads, K, d_w.p);
        ctx->launches++;
    }
    unsigned long long wn = 0;
    SG_CUDA(cudaMemcpyAsync(&wn, d_w.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    return (uint64_t)wn;
}

static ReadsSrc reads_source(Ctx *ctx, int K, bool both) {
    ReadsSrc src;
    src.words = ctx->d_words; src.offs = ctx->d_offs; src.lens = ctx->d_lens; src.n = ctx->n_reads; src.nwords = ctx->n_words; src.K = K; src.both = both;
    return src;
}

template <int NW>
static KSet *count_reads_nw(Ctx *ctx, int K, int B, int mode) {
    ensure_reads_on_device(ctx);
    KSet *ks = new KSet();
    ks->ctx = ctx; ks->K = K; ks->nw = NW; ks->B = B; ks->has_counts = (mode == kCanonical);
    try {
        const bool both = (mode == kAllWindows);
        const uint64_t wn = count_windows(ctx, K);
        std::vector<ReadsSrc> srcs{reads_source(ctx, K, both)};
        run_count<NW, ReadsSrc>(ctx, srcs, K, B, mode == kCanonical, mode == kCanonical && (K % 2 == 0), wn * (both ? 2 : 1), ks);
    } catch (...) { delete ks; throw; }
    return ks;
}

KSet *count_from_reads(Ctx *ctx, int K, int B, int mode) {
    SG_CHECK(K >= 1 && K <= 128, 2, "K must be in [1,128]");
    SG_CHECK(B >= 1 && B <= (1 << 20), 2, "num_buckets must be in [1, 2^20]");
    ctx->times = PhaseTimes();
    switch (nwords_of(K)) {
        case 1: return count_reads_nw<1>(ctx, K, B, mode);
        case 2: return count_reads_nw<2>(ctx, K, B, mode);
        case 3: return count_reads_nw<3>(ctx, K, B, mode);
        default: return count_reads_nw<4>(ctx, K, B, mode);
    }
}

template <int NW, int NWS>
static KSet *kmers_from_kpomers_nw(Ctx *ctx, const KSet *kp, int B) {
    const int K = kp->K - 1;
    KSet *ks = new KSet();
    ks->ctx = ctx; ks->K = K; ks->nw = NW; ks->B = B; ks->has_counts = false;
    try {
        std::vector<KpomerSrc<NWS>> srcs;
        for (const Chunk &c : kp->chunks) {
            KpomerSrc<NWS> s; s.keys = c.keys.p; s.n = c.n; s.K = K;
            srcs.push_back(s);
        }
        run_count<NW, KpomerSrc<NWS>>(ctx, srcs, K, B, false, false, (uint64_t)kp->n * 2, ks);
    } catch (...) { delete ks; throw; }
    return ks;
}

KSet *kmers_from_kpomers(Ctx *ctx, const KSet *kp, int B) {
    SG_CHECK(kp->K >= 2, 2, "source k-mers too short");
    SG_CHECK(B >= 1 && B <= (1 << 20), 2, "num_buckets must be in [1, 2^20]");
    const int K = kp->K - 1;
    const int nw = nwords_of(K), nws = kp->nw;
    ctx->times = PhaseTimes();
    if (nw == 1 && nws == 1) return kmers_from_kpomers_nw<1, 1>(ctx, kp, B);
    if (nw == 1 && nws == 2) return kmers_from_kpomers_nw<1, 2>(ctx, kp, B);
    if (nw == 2 && nws == 2) return kmers_from_kpomers_nw<2, 2>(ctx, kp, B);
    if (nw == 2 && nws == 3) return kmers_from_kpomers_nw<2, 3>(ctx, kp, B);
    if (nw == 3 && nws == 3) return kmers_from_kpomers_nw<3, 3>(ctx, kp, B);
    if (nw == 3 && nws == 4) return kmers_from_kpomers_nw<3, 4>(ctx, kp, B);
    if (nw == 4 && nws == 4) return kmers_from_kpomers_nw<4, 4>(ctx, kp, B);
    throw Error(2, "unsupported k-mer word combination");
}

// checksums of a counted set (bench / multi-GPU self check): out = { n, sum of all key words, xor of all key words rotated by
// their word index, sum of multiplicities } -- order independent, so per-rank values of disjoint bucket sets add / xor up
__global__ void kset_checksum_k(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, uint64_t n, int nw,
                                unsigned long long *__restrict__ out) {
    unsigned long long s = 0, x = 0, c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        for (int q = 0; q < nw; ++q) {
            const unsigned long long w = keys[i * nw + q];
            s += w * (unsigned long long)(2 * q + 1);
            x ^= (w << (7 * q + 1)) | (w >> (64 - (7 * q + 1)));
        }
        if (counts) c += counts[i];
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        x ^= __shfl_down_sync(0xffffffffu, x, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&out[1], s); atomicXor(&out[2], x); atomicAdd(&out[3], c); }
}
void kset_checksum(const KSet *ks, uint64_t *out4) {
    Ctx *ctx = ks->ctx;
    DArr<unsigned long long> d(ctx, 4);
    SG_CUDA(cudaMemsetAsync(d.p, 0, 32, ctx->stream));
    for (const Chunk &c : ks->chunks) {
        if (c.n == 0) continue;
        kset_checksum_k<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(c.keys.p, ks->has_counts ? c.counts.p : nullptr, (uint64_t)c.n, ks->nw, d.p);
        ctx->launches++;
    }
    SG_CUDA(cudaGetLastError());
    unsigned long long h[4];
    SG_CUDA(cudaMemcpyAsync(h, d.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    out4[0] = (uint64_t)ks->n; out4[1] = h[1]; out4[2] = h[2]; out4[3] = h[3];
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU count (replaces projects/hpcspades/mpi/stages/construction_mpi.cpp:222-300). Buckets are the unit of
// independence (KMerSegmentPolicy is a pure function of the k-mer), so every bucket has one owner GPU. Each rank histograms its
// own read shard once with the same rolling kernel as the single-GPU count (2-byte partition ids included); the per-partition
// totals are all-gathered (host plumbing, torch.distributed). Per pass every rank scatters its shard's records of the pass's
// buckets into a local staging buffer (CTA-major pieces, like the single-GPU count), and then ONE kernel on the owner does the
// exchange AND the merge: it pulls every (source rank, level-A CTA, partition) piece out of the sources' staging buffers through
// NVLink peer mappings (cudaIpc: every rank maps the peers' memory ARENAS once per process; buffers are offsets inside them) with
// coalesced 16-byte loads and lays the pieces of a partition next to each other, ready for the ordinary refinement / local sort
// / compaction. (A first version pushed each 16-byte record to its owner from inside the partition kernel; remote scattered
// stores are not write-combined and ran at ~1 GB/s.)
// ------------------------------------------------------------------------------------------------------------
struct DistPlan {
    int world = 1, rank = 0, B = 0, rA = 0;
    uint32_t PA_all = 0;
    std::vector<int> pass_b;                   // bucket boundaries of the passes planned so far (starts as {0})
    std::vector<uint64_t> tot;                 // PA_all: records per partition summed over ranks
    std::vector<uint64_t> Tb;                  // B+1: records in buckets < b, all ranks
    std::vector<uint64_t> Ps;                  // world x (B+1): records of rank s in buckets < b
    uint64_t pass_target = 0;                  // records per pass aimed for (equal passes: a tiny last pass still costs a full id sweep)
    int npass() const { return (int)pass_b.size() - 1; }
    static int own_lo_of(int b_lo, int b_hi, int world, int g) { return b_lo + (int)((int64_t)(b_hi - b_lo) * g / world); }
    int own_lo(int p, int g) const { return own_lo_of(pass_b[p], pass_b[p + 1], world, g); }
    // largest number of records an owner receives / a rank sends if [b_lo, b_hi) is one pass
    void maxima(int b_lo, int b_hi, uint64_t *mx, uint64_t *ms) const {
        *mx = 0; *ms = 0;
        for (int g = 0; g < world; ++g) *mx = std::max(*mx, Tb[own_lo_of(b_lo, b_hi, world, g + 1)] - Tb[own_lo_of(b_lo, b_hi, world, g)]);
        for (int s = 0; s < world; ++s) *ms = std::max(*ms, Ps[(size_t)s * (B + 1) + b_hi] - Ps[(size_t)s * (B + 1) + b_lo]);
    }
};

static const double kDistHeadroom = 1.05;      // staging / merged buffers are sized 5 % above the planned maximum

// device bytes one pass needs next to what is already resident: staging buffer (doubles as the sort's ping-pong partner, so
// >= recv), merged buffer, its own output (distinct/instances <= 0.6 assumed, like the single-GPU planner) and the per-pass tables
static double dist_pass_bytes(uint64_t mx, uint64_t ms, size_t W, uint64_t fixed_bytes) {
    return ((double)std::max(mx, ms) + (double)mx) * W * (kDistHeadroom + 0.03) + (double)mx * (W + 4) * 0.6 + (double)fixed_bytes;
}

// pure host functions (also exported for the CPU/gloo tests): identical on every rank given the same inputs
void dist_plan_tables(DistPlan &pl, int world, int rank, int B, int rA, const uint64_t *cnt_all) {
    pl.world = world; pl.rank = rank; pl.B = B; pl.rA = rA; pl.PA_all = (uint32_t)B << rA;
    pl.tot.assign(pl.PA_all, 0);
    pl.Tb.assign((size_t)B + 1, 0);
    pl.Ps.assign((size_t)world * (B + 1), 0);
    for (int s = 0; s < world; ++s) {
        for (uint32_t q = 0; q < pl.PA_all; ++q) pl.tot[q] += cnt_all[(size_t)s * pl.PA_all + q];
        for (int b = 0; b < B; ++b) {
            uint64_t t = 0;
            for (uint32_t q = (uint32_t)b << rA; q < ((uint32_t)(b + 1) << rA); ++q) t += cnt_all[(size_t)s * pl.PA_all + q];
            pl.Ps[(size_t)s * (B + 1) + b + 1] = pl.Ps[(size_t)s * (B + 1) + b] + t;
        }
    }
    for (int b = 0; b <= B; ++b) {           // Tb = sum over ranks of Ps
        uint64_t t = 0;
        for (int s = 0; s < world; ++s) t += pl.Ps[(size_t)s * (B + 1) + b];
        pl.Tb[b] = t;
    }
    pl.pass_b.assign(1, 0);
    pl.pass_target = 0;
}
// the greedy "as many whole buckets as fit" step shared by the simulation and the real planning: returns b_hi > b_lo
static int dist_greedy_pass(const DistPlan &pl, int b_lo, double budget, size_t W, uint64_t fixed_bytes, uint64_t target) {
    int b_hi = b_lo;
    while (b_hi < pl.B) {
        uint64_t mx, ms;
        pl.maxima(b_lo, b_hi + 1, &mx, &ms);
        const uint64_t recs = pl.Tb[b_hi + 1] - pl.Tb[b_lo];
        if (b_hi > b_lo && (dist_pass_bytes(mx, ms, W, fixed_bytes) > budget || (target && recs > target) ||
                            ((size_t)(b_hi + 1 - b_lo) << pl.rA) > (size_t)kLevelAMaxParts)) break;
        ++b_hi;
    }
    return b_hi;
}
// plan the next pass against `budget_bytes` = what every rank can allocate NOW (minimum over ranks). Returns false when all
// buckets are done. The first call also fixes the pass size aimed for by simulating the whole job (outputs of earlier passes
// stay resident: ~half a record's bytes per record, as measured on read sets with errors).
bool dist_next_pass(DistPlan &pl, uint64_t budget_bytes, size_t W, uint64_t fixed_bytes, uint64_t *mx_out, uint64_t *ms_out) {
    const int b_lo = pl.pass_b.back();
    if (b_lo >= pl.B) return false;
    if (pl.pass_target == 0) {
        double lim = (double)budget_bytes;
        int np = 0, b = 0;
        while (b < pl.B) {
            const int e = dist_greedy_pass(pl, b, lim, W, fixed_bytes, 0);
            uint64_t mx, ms;
            pl.maxima(b, e, &mx, &ms);
            lim -= (double)mx * (W + 4) * 0.5;
            b = e; ++np;
        }
        const uint64_t total = pl.Tb[pl.B];
        pl.pass_target = total / (uint64_t)np + total / 64 + 1;
    }
    const int b_hi = dist_greedy_pass(pl, b_lo, (double)budget_bytes, W, fixed_bytes, pl.pass_target);
    pl.pass_b.push_back(b_hi);
    pl.maxima(b_lo, b_hi, mx_out, ms_out);
    return true;
}

struct PullSrc { const uint64_t *sbuf; const uint64_t *pbase; const uint32_t *blk; };     // one source rank, as seen from this GPU

struct DistState {
    Ctx *ctx = nullptr;
    int K = 0, B = 0, mode = 0, nw = 0, G = 0;
    DistPlan plan;
    std::vector<uint64_t> h_cnt_all;         // world x PA_all
    DArr<uint64_t> sbuf, xbuf;               // staging buffer (peers read it; later the ping-pong partner) and the merged buffer
    DArr<uint64_t> pbase;                    // [G][max partitions of a pass]: piece starts of the current pass (peers read it)
    std::vector<PullSrc> peers;              // world entries (own entry = local pointers)
    DArr<unsigned long long> d_bsz;
    KSet *out = nullptr;
    int64_t first = 0;
    bool want_counts = false, double_selfrc = false;
    virtual ~DistState() { delete out; }
    virtual void begin() = 0;
    virtual void local_counts(uint64_t *h_out) = 0;
    virtual const uint32_t *blk_counts_ptr() = 0;
    virtual void scatter(int p) = 0;
    virtual void pull(int p) = 0;
    virtual void sort(int p) = 0;
};

// exchange + merge in one kernel. Work item = (owned partition q, source rank s): the CTA fetches the source's G piece
// descriptors (start in its staging buffer, record count) with one parallel remote read, then streams the pieces into the
// merged buffer back to back with 16-byte loads over NVLink.
static const int kPullThreads = 512;
template <int NW>
__global__ void __launch_bounds__(kPullThreads) dist_pull_k(const PullSrc *__restrict__ srcs, int world, int G, uint32_t PA_all, uint32_t PAp, uint32_t pq0,
                                                            uint32_t q0, uint32_t nq, const uint64_t *__restrict__ dst_off, uint64_t *__restrict__ out,
                                                            unsigned long long *__restrict__ work_counter) {
    __shared__ unsigned long long s_w;
    __shared__ uint64_t p_start[kMaxPieces];      // source record index of piece g
    __shared__ uint64_t p_dst[kMaxPieces + 1];    // destination record index of piece g (exclusive prefix of the counts)
    __shared__ uint32_t wsum[kPullThreads / 32];
    const uint64_t nwork = (uint64_t)nq * (uint64_t)world;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) s_w = atomicAdd(work_counter, 1ull);
        __syncthreads();
        const uint64_t w = s_w;
        __syncthreads();
        if (w >= nwork) return;
        const uint32_t q = q0 + (uint32_t)(w / (uint64_t)world);
        const int s = (int)(w % (uint64_t)world);
        const PullSrc ps = srcs[s];
        // piece counts -> exclusive prefix (G <= kMaxPieces = 2 * kPullThreads: two per thread)
        uint32_t c0 = 0, c1 = 0;
        const int g0 = 2 * threadIdx.x;
        if (g0 < G) { c0 = ps.blk[(size_t)g0 * PA_all + q]; p_start[g0] = ps.pbase[(size_t)g0 * PAp + (q - pq0)]; }
        if (g0 + 1 < G) { c1 = ps.blk[(size_t)(g0 + 1) * PA_all + q]; p_start[g0 + 1] = ps.pbase[(size_t)(g0 + 1) * PAp + (q - pq0)]; }
        uint32_t v = c0 + c1, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for (int x = 0; x < warp; ++x) wb += wsum[x];
        const uint64_t d0 = dst_off[w] + (uint64_t)(wb + inc - v);
        if (g0 < G) p_dst[g0] = d0;
        if (g0 + 1 < G) p_dst[g0 + 1] = d0 + c0;
        if (g0 + 2 >= G && g0 < G) p_dst[G] = d0 + v;      // the thread holding the last piece(s) also writes the end
        __syncthreads();
        for (int g = 0; g < G; ++g) {
            const uint64_t n = p_dst[g + 1] - p_dst[g];
            if (n == 0) continue;
            const uint64_t *src = ps.sbuf + p_start[g] * NW;
            uint64_t *dst = out + p_dst[g] * NW;
            const uint64_t nwords = n * NW;
            if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
                const ulonglong2 *s2 = reinterpret_cast<const ulonglong2 *>(src);
                ulonglong2 *d2 = reinterpret_cast<ulonglong2 *>(dst);
                const uint64_t n2 = nwords / 2;
                uint64_t i = threadIdx.x;
                for (; i + 3ull * kPullThreads < n2; i += 4ull * kPullThreads) {       // four loads in flight per thread
                    const ulonglong2 a = s2[i], b = s2[i + kPullThreads], c = s2[i + 2 * kPullThreads], e = s2[i + 3 * kPullThreads];
                    d2[i] = a; d2[i + kPullThreads] = b; d2[i + 2 * kPullThreads] = c; d2[i + 3 * kPullThreads] = e;
                }
                for (; i < n2; i += kPullThreads) d2[i] = s2[i];
                if ((nwords & 1) && threadIdx.x == 0) dst[nwords - 1] = src[nwords - 1];
            } else {
                for (uint64_t i = threadIdx.x; i < nwords; i += kPullThreads) dst[i] = src[i];
            }
        }
        __syncthreads();      // p_start / p_dst are rewritten by the next work item
    }
}

template <int NW>
struct DistStateNW : DistState {
    LevelAJob<NW, ReadsSrc> job;

    void begin() override {
        Timer tm(ctx->stream);
        Trace tr(ctx->stream);
        const uint64_t wn = count_windows(ctx, K);
        job.ctx = ctx; job.K = K; job.B = B; job.rA = plan.rA; job.G = G;
        job.srcs.assign(1, reads_source(ctx, K, mode == kAllWindows));
        job.s_lo = 0; job.s_hi = B; job.PA_all = plan.PA_all;
        job.roll = (mode == kCanonical);
        levelA_count(job, wn * (mode == kAllWindows ? 2 : 1), tm, tr);
    }
    void local_counts(uint64_t *h_out) override { memcpy(h_out, job.h_part.data(), (size_t)plan.PA_all * 8); }
    const uint32_t *blk_counts_ptr() override { return job.blk_counts.p; }

    void scatter(int p) override {
        // local partition of this rank's shard for the pass's buckets into the staging buffer (CTA-major pieces)
        Timer tm(ctx->stream);
        Trace tr(ctx->stream);
        const int b_lo = plan.pass_b[p], b_hi = plan.pass_b[p + 1];
        uint64_t I = 0, total = 0;
        for (uint32_t q = 0; q < plan.PA_all; ++q) {
            const uint64_t c = h_cnt_all[(size_t)plan.rank * plan.PA_all + q];
            total += c;
            if (q >= ((uint32_t)b_lo << plan.rA) && q < ((uint32_t)b_hi << plan.rA)) I += c;
        }
        SG_CHECK((I * NW + 2) <= sbuf.n, 6, "internal: staging buffer smaller than the pass");
        Pieces pcs;
        levelA_scatter(job, b_lo, b_hi, I, total, sbuf.p, pbase.p, pcs, tm, tr);
        SG_CUDA(cudaStreamSynchronize(ctx->stream));
    }

    void pull(int p) override {
        cudaStream_t st = ctx->stream;
        const int world = plan.world;
        const uint32_t pq0 = (uint32_t)plan.pass_b[p] << plan.rA;
        const uint32_t PAp = (uint32_t)(plan.pass_b[p + 1] - plan.pass_b[p]) << plan.rA;
        const uint32_t q0 = (uint32_t)plan.own_lo(p, plan.rank) << plan.rA, q1 = (uint32_t)plan.own_lo(p, plan.rank + 1) << plan.rA;
        const uint32_t nq = q1 - q0;
        if (nq == 0) return;
        // destination of (q, s): partitions in order, inside a partition the sources in rank order
        std::vector<uint64_t> dst_off((size_t)nq * world);
        uint64_t run = 0;
        for (uint32_t q = q0; q < q1; ++q)
            for (int s = 0; s < world; ++s) {
                dst_off[(size_t)(q - q0) * world + s] = run;
                run += h_cnt_all[(size_t)s * plan.PA_all + q];
            }
        SG_CHECK(run * NW + 2 <= xbuf.n, 6, "internal: merged buffer smaller than the pass");
        if (run == 0) return;
        DArr<uint64_t> d_off(ctx, dst_off.size());
        DArr<PullSrc> d_src(ctx, (size_t)world);
        DArr<unsigned long long> wc(ctx, 1);
        SG_CUDA(cudaMemcpyAsync(d_off.p, dst_off.data(), dst_off.size() * 8, cudaMemcpyHostToDevice, st));
        SG_CUDA(cudaMemcpyAsync(d_src.p, peers.data(), (size_t)world * sizeof(PullSrc), cudaMemcpyHostToDevice, st));
        SG_CUDA(cudaMemsetAsync(wc.p, 0, 8, st));
        Timer tm(st);
        tm.start();
        const uint64_t nwork = (uint64_t)nq * world;
        const int grid = (int)std::min<uint64_t>(nwork, (uint64_t)ctx->num_sms * 4);
        dist_pull_k<NW><<<grid, kPullThreads, 0, st>>>(d_src.p, world, G, plan.PA_all, PAp, pq0, q0, nq, d_off.p, xbuf.p, wc.p);
        ctx->launches++;
        SG_CUDA(cudaGetLastError());
        ctx->times.exchange += tm.stop();
    }

    void sort(int p) override {
        cudaStream_t st = ctx->stream;
        const int my_lo = plan.own_lo(p, plan.rank), my_hi = plan.own_lo(p, plan.rank + 1);
        const uint32_t PA = (uint32_t)(my_hi - my_lo) << plan.rA;
        const size_t Q0 = (size_t)my_lo << plan.rA;
        std::vector<uint64_t> tot(PA + 1, 0), start(PA + 1, 0);
        for (uint32_t q = 0; q < PA; ++q) { tot[q] = plan.tot[Q0 + q]; start[q + 1] = start[q] + tot[q]; }
        ctx->times.passes++;
        ctx->times.instances += start[PA];
        Chunk ch;
        if (PA && start[PA]) {
            DArr<uint64_t> d_tot(ctx, PA + 1), d_start(ctx, PA + 1);
            SG_CUDA(cudaMemcpyAsync(d_tot.p, tot.data(), (size_t)(PA + 1) * 8, cudaMemcpyHostToDevice, st));
            SG_CUDA(cudaMemcpyAsync(d_start.p, start.data(), (size_t)(PA + 1) * 8, cudaMemcpyHostToDevice, st));
            Timer tm(st);
            Trace tr(st);
            sort_pass<NW>(ctx, K, xbuf, sbuf, d_start.p, d_tot.p, PA, plan.rA, (uint32_t)my_lo, my_hi, first, want_counts, double_selfrc, d_bsz.p, ch, tm, tr);
        } else {
            ch.n = 0; ch.b_lo = my_lo; ch.b_hi = my_hi; ch.first = first;
            ch.keys.alloc(ctx, 2, true);
            if (want_counts) ch.counts.alloc(ctx, 1, true);
        }
        first += ch.n;
        out->chunks.push_back(std::move(ch));
    }
};

DistState *dist_begin(Ctx *ctx, int K, int B, int mode, int world, int rank) {
    SG_CHECK(K >= 1 && K <= 128, 2, "K must be in [1,128]");
    SG_CHECK(B >= 1 && B <= kLevelAMaxParts, 2, "distributed count: num_buckets must be in [1, 8192]");
    SG_CHECK(world >= 1 && world <= 255 && rank >= 0 && rank < world, 2, "bad world/rank");
    SG_CHECK(mode == kCanonical || mode == kAllWindows, 2, "bad mode");
    ensure_reads_on_device(ctx);
    ctx->times = PhaseTimes();
    DistState *d = nullptr;
    switch (nwords_of(K)) {
        case 1: d = new DistStateNW<1>(); break;
        case 2: d = new DistStateNW<2>(); break;
        case 3: d = new DistStateNW<3>(); break;
        default: d = new DistStateNW<4>(); break;
    }
    d->ctx = ctx; d->K = K; d->B = B; d->mode = mode; d->nw = nwords_of(K);
    d->G = ctx->num_sms * levelA_ctas_per_sm();
    d->want_counts = (mode == kCanonical); d->double_selfrc = (mode == kCanonical) && (K % 2 == 0);
    // every rank must use the same geometry, so it depends on B only. As many level-A partitions as the shared-memory tables allow:
    // an owner's segment is the union of all ranks' records of a partition, so finer partitions keep refinement at one round
    int rA = 0;
    while (rA < 8 && ((uint64_t)B << (rA + 1)) <= (uint64_t)kLevelAMaxParts && rA + 1 <= 2 * K) ++rA;
    d->plan.world = world; d->plan.rank = rank; d->plan.B = B; d->plan.rA = rA; d->plan.PA_all = (uint32_t)B << rA;
    try { d->begin(); } catch (...) { delete d; throw; }
    return d;
}
uint32_t dist_num_partitions(const DistState *d) { return d->plan.PA_all; }
void dist_local_counts(DistState *d, uint64_t *h_out) { d->local_counts(h_out); }

static const uint64_t kDistFixedBytes = (uint64_t)192 << 20;      // per-pass tables (segments, work lists, scans) next to the big buffers

void dist_plan(DistState *d, const uint64_t *cnt_all, uint64_t *total_records) {
    Ctx *ctx = d->ctx;
    dist_plan_tables(d->plan, d->plan.world, d->plan.rank, d->B, d->plan.rA, cnt_all);
    d->h_cnt_all.assign(cnt_all, cnt_all + (size_t)d->plan.world * d->plan.PA_all);
    d->d_bsz.alloc(ctx, (size_t)d->B);
    SG_CUDA(cudaMemsetAsync(d->d_bsz.p, 0, (size_t)d->B * 8, ctx->stream));
    d->out = new KSet();
    d->out->ctx = ctx; d->out->K = d->K; d->out->nw = d->nw; d->out->B = d->B; d->out->has_counts = d->want_counts;
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    *total_records = d->plan.Tb[d->B];
}

// bytes this rank could allocate for the next pass (the previous pass's staging / merged buffers are given back first)
uint64_t dist_free_bytes(DistState *d) {
    d->sbuf.release(); d->xbuf.release(); d->pbase.release();
    return (uint64_t)d->ctx->free_bytes();
}

// plan the next pass against budget_bytes (the minimum of dist_free_bytes over the ranks, so that every rank takes the same
// decision) and allocate its buffers. Returns the pass index, or -1 when every bucket has been processed.
int dist_next_pass(DistState *d, uint64_t budget_bytes) {
    Ctx *ctx = d->ctx;
    const size_t W = (size_t)8 * d->nw;
    d->sbuf.release(); d->xbuf.release(); d->pbase.release();
    uint64_t mx = 0, ms = 0;
    if (!dist_next_pass(d->plan, budget_bytes, W, kDistFixedBytes, &mx, &ms)) return -1;
    const int p = d->plan.npass() - 1;
    // buffers the peers read must live inside the arena (one driver allocation, mapped by the peers as a whole)
    const uint32_t pa = (uint32_t)(d->plan.pass_b[p + 1] - d->plan.pass_b[p]) << d->plan.rA;
    d->sbuf.alloc(ctx, (size_t)((double)std::max(mx, ms) * kDistHeadroom) * d->nw + 16);
    d->xbuf.alloc(ctx, (size_t)((double)mx * kDistHeadroom) * d->nw + 16);
    d->pbase.alloc(ctx, (size_t)d->G * pa);
    d->peers.clear();
    return p;
}

// descriptor a rank publishes (all_gather) after dist_plan: arena handle + where its peer-readable buffers are inside the arena
struct DistDesc {
    cudaIpcMemHandle_t arena;     // 64 bytes
    uint64_t arena_size, off_sbuf, off_pbase, off_blk;
};
static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(DistDesc) == 96, "descriptor layout (SGPU_IPC_BYTES)");

void dist_ipc_handle(DistState *d, uint8_t *out96) {
    Ctx *ctx = d->ctx;
    auto off_of = [&](const void *p, const char *what) {
        const char *c = (const char *)p;
        SG_CHECK(ctx->arena && c >= ctx->arena && c < ctx->arena + ctx->arena_size, 4, what);
        return (uint64_t)(c - ctx->arena);
    };
    DistDesc ds;
    memset(&ds, 0, sizeof ds);
    ds.off_sbuf = off_of(d->sbuf.p, "distributed count: the staging buffer did not fit the device memory arena");
    ds.off_pbase = off_of(d->pbase.p, "distributed count: the piece table did not fit the device memory arena");
    ds.off_blk = off_of(d->blk_counts_ptr(), "distributed count: the level-A count table did not fit the device memory arena");
    ds.arena_size = ctx->arena_size;
    if (d->plan.world > 1) SG_CUDA(cudaIpcGetMemHandle(&ds.arena, ctx->arena));
    memcpy(out96, &ds, sizeof ds);
}

void dist_open_peers(DistState *d, const uint8_t *descs) {
    Ctx *ctx = d->ctx;
    const int world = d->plan.world;
    d->peers.assign(world, PullSrc{nullptr, nullptr, nullptr});
    if ((int)ctx->peer_arena.size() != world) { ctx->peer_close(); ctx->peer_arena.assign(world, nullptr); ctx->peer_handle.assign((size_t)world * 64, 0); }
    for (int g = 0; g < world; ++g) {
        DistDesc ds;
        memcpy(&ds, descs + (size_t)g * sizeof(DistDesc), sizeof ds);
        char *base = nullptr;
        if (g == d->plan.rank) {
            base = ctx->arena;
        } else {
            // a peer's arena is mapped once per process; a different handle under the same rank (its context was re-created) remaps
            if (ctx->peer_arena[g] && memcmp(&ctx->peer_handle[(size_t)g * 64], &ds.arena, 64) != 0) {
                cudaIpcCloseMemHandle(ctx->peer_arena[g]);
                ctx->peer_arena[g] = nullptr;
            }
            if (!ctx->peer_arena[g]) {
                void *p = nullptr;
                SG_CUDA(cudaIpcOpenMemHandle(&p, ds.arena, cudaIpcMemLazyEnablePeerAccess));
                ctx->peer_arena[g] = (char *)p;
                memcpy(&ctx->peer_handle[(size_t)g * 64], &ds.arena, 64);
            }
            base = ctx->peer_arena[g];
        }
        SG_CHECK(ds.off_sbuf < ds.arena_size && ds.off_pbase < ds.arena_size && ds.off_blk < ds.arena_size, 2, "bad peer descriptor");
        d->peers[g] = PullSrc{(const uint64_t *)(base + ds.off_sbuf), (const uint64_t *)(base + ds.off_pbase), (const uint32_t *)(base + ds.off_blk)};
    }
}

void dist_scatter(DistState *d, int p) {
    SG_CHECK(p >= 0 && p == d->plan.npass() - 1 && d->sbuf.p, 2, "bad pass (sgpu_dist_next_pass decides the current one)");
    d->scatter(p);
    SG_CUDA(cudaStreamSynchronize(d->ctx->stream));        // the staging buffer is complete; peers may read it after the next barrier
}
void dist_exchange(DistState *d, int p) {
    SG_CHECK(p >= 0 && p == d->plan.npass() - 1 && d->sbuf.p, 2, "bad pass (sgpu_dist_next_pass decides the current one)");
    SG_CHECK((int)d->peers.size() == d->plan.world, 2, "sgpu_dist_open_peers has not run");
    d->pull(p);
    SG_CUDA(cudaStreamSynchronize(d->ctx->stream));        // this rank no longer reads any peer's staging buffer
}
void dist_sort(DistState *d, int p) {
    SG_CHECK(p >= 0 && p == d->plan.npass() - 1 && d->sbuf.p, 2, "bad pass (sgpu_dist_next_pass decides the current one)");
    d->sort(p);
    SG_CUDA(cudaStreamSynchronize(d->ctx->stream));
}
KSet *dist_end(DistState *d) {
    Ctx *ctx = d->ctx;
    KSet *ks = d->out;
    const int B = d->B;
    std::vector<unsigned long long> hb(B);
    SG_CUDA(cudaMemcpyAsync(hb.data(), d->d_bsz.p, (size_t)B * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SG_CUDA(cudaStreamSynchronize(ctx->stream));
    ks->bsz.assign(B, 0); ks->bstart.assign(B + 1, 0);
    for (int b = 0; b < B; ++b) { ks->bsz[b] = (int64_t)hb[b]; ks->bstart[b + 1] = ks->bstart[b] + ks->bsz[b]; }
    ks->n = d->first;
    SG_CHECK(ks->bstart[B] == ks->n, 6, "internal: distributed bucket sizes do not add up");
    d->out = nullptr;
    return ks;
}
void dist_free(DistState *d) { delete d; }

// the whole pass sequence for a FIXED budget (CPU tests): every pass is planned against the budget minus the estimated outputs of
// the passes before it
int dist_plan_host(int world, int B, int rA, const uint64_t *cnt_all, uint64_t budget_bytes, int record_bytes, int *pass_b, uint64_t *max_recv) {
    DistPlan pl;
    dist_plan_tables(pl, world, 0, B, rA, cnt_all);
    double lim = (double)budget_bytes;
    uint64_t mx = 0, ms = 0, worst = 0;
    while (dist_next_pass(pl, (uint64_t)std::max(0.0, lim), (size_t)record_bytes, 0, &mx, &ms)) {
        worst = std::max(worst, mx);
        lim -= (double)mx * (record_bytes + 4) * 0.5;
    }
    for (int p = 0; p <= pl.npass(); ++p) pass_b[p] = pl.pass_b[p];
    *max_recv = worst;
    return pl.npass();
}

}  // namespace sg
