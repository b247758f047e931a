// sgpu_internal.h -- internal C++ structures behind the C ABI of include/spades_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "kmer_dev.cuh"

namespace sg {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define SG_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            char b__[512];                                                                             \
            snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            throw sg::Error(5, b__);                                                                   \
        }                                                                                              \
    } while (0)

#define SG_CHECK(cond, code, msg)                                                                      \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            char b__[512];                                                                             \
            snprintf(b__, sizeof b__, "%s:%d: %s", __FILE__, __LINE__, (msg));                         \
            throw sg::Error((code), b__);                                                              \
        }                                                                                              \
    } while (0)

// device allocation with byte accounting (so multi-pass planning can see what is resident)
struct Ctx;
template <class T>
struct DArr {
    T *p = nullptr;
    size_t n = 0;
    size_t blk = 0;     // bytes of the underlying block
    Ctx *ctx = nullptr;
    DArr() {}
    DArr(Ctx *c, size_t n_) { alloc(c, n_); }
    DArr(const DArr &) = delete;
    DArr &operator=(const DArr &) = delete;
    DArr(DArr &&o) noexcept { p = o.p; n = o.n; blk = o.blk; ctx = o.ctx; o.p = nullptr; o.n = 0; o.blk = 0; }
    DArr &operator=(DArr &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; blk = o.blk; ctx = o.ctx; o.p = nullptr; o.n = 0; o.blk = 0; }
        return *this;
    }
    ~DArr() { release(); }
    void alloc(Ctx *c, size_t n_);
    void release();
    size_t bytes() const { return n * sizeof(T); }
};

struct PhaseTimes {   // device milliseconds measured with CUDA events on ctx->stream
    float extract_count = 0, extract_scatter = 0, refine = 0, local_sort = 0, compact = 0, mphf = 0, exchange = 0, total = 0;
    uint64_t launches = 0;
    uint64_t instances = 0;       // records written by the partition kernel
    uint64_t passes = 0;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    size_t hbm_budget = 0;       // 0 = use free memory
    size_t allocated = 0, peak = 0;
    int verbose = 0;
    std::string err;
    PhaseTimes times;
    uint64_t launches = 0;
    // reads
    DArr<uint64_t> r_words;      // owned copy (host-appended) ...
    DArr<uint64_t> r_offs;
    DArr<uint32_t> r_lens;
    const uint64_t *d_words = nullptr;   // ... or adopted device pointers
    const uint64_t *d_offs = nullptr;
    const uint32_t *d_lens = nullptr;
    int64_t n_reads = 0;
    uint64_t n_words = 0;
    std::vector<uint64_t> h_words, h_offs;   // host staging until first use
    std::vector<uint32_t> h_lens;
    bool staged_dirty = false;
    // caching device allocator: freed blocks are kept for reuse (a bench step repeats the same sizes), and are
    // handed back to the driver only under memory pressure or when the context dies
    std::vector<std::pair<void *, size_t>> pool_free;
    size_t pool_cached = 0;
    void *pool_alloc(size_t bytes, size_t *blk);
    void pool_release(void *p, size_t blk);
    void pool_trim();
    size_t free_bytes();     // driver-free + cached
};

inline void Ctx::pool_trim() {
    for (auto &b : pool_free) cudaFree(b.first);
    pool_free.clear();
    pool_cached = 0;
}
inline size_t Ctx::free_bytes() {
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    return f + pool_cached;
}
inline void *Ctx::pool_alloc(size_t bytes, size_t *blk) {
    size_t want = (bytes + 511) & ~(size_t)511;
    int best = -1;
    for (int i = 0; i < (int)pool_free.size(); ++i) {
        size_t sz = pool_free[i].second;
        if (sz >= want && sz <= want + want / 4 + (1 << 20) && (best < 0 || sz < pool_free[best].second)) best = i;
    }
    if (best >= 0) {
        void *p = pool_free[best].first;
        *blk = pool_free[best].second;
        pool_cached -= *blk;
        pool_free.erase(pool_free.begin() + best);
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_trim();
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        char m[256];
        snprintf(m, sizeof m, "cudaMalloc(%zu bytes) failed: %s (resident %zu)", want, cudaGetErrorString(e), allocated);
        throw Error(4, m);
    }
    *blk = want;
    return p;
}
inline void Ctx::pool_release(void *p, size_t blk) {
    pool_free.push_back({p, blk});
    pool_cached += blk;
}

template <class T>
void DArr<T>::alloc(Ctx *c, size_t n_) {
    release();
    ctx = c; n = n_;
    size_t b = (n_ ? n_ : 1) * sizeof(T);
    p = (T *)c->pool_alloc(b, &blk);
    c->allocated += blk;
    if (c->allocated > c->peak) c->peak = c->allocated;
}
template <class T>
void DArr<T>::release() {
    if (p) {
        ctx->allocated -= blk;
        ctx->pool_release(p, blk);
    }
    p = nullptr; n = 0; blk = 0;
}

// A counted k-mer set resident in HBM: what KMerDiskCounter::Count leaves on disk
// (kmer_index_builder.hpp:306-332), bucket-major, strictly increasing inside a bucket.
struct Chunk {
    DArr<uint64_t> keys;     // n * nw words (records of W = 8*nw bytes, the on-disk record format)
    DArr<uint32_t> counts;   // n (canonical mode) or empty
    int64_t n = 0;
    int b_lo = 0, b_hi = 0;  // buckets [b_lo, b_hi)
    int64_t first = 0;       // index of its first record in final_kmers order
};
struct KSet {
    Ctx *ctx = nullptr;
    int K = 0, nw = 0, B = 0;
    int64_t n = 0;
    bool has_counts = false;
    std::vector<Chunk> chunks;
    std::vector<int64_t> bsz;          // B
    std::vector<int64_t> bstart;       // B+1 exclusive prefix (final_kmers order)
};

// boomphf-compatible index resident in HBM (one mphf per bucket, BooPHF.h / kmer_index.hpp)
static const int kLevels = 25;
struct Mphf {
    Ctx *ctx = nullptr;
    int K = 0, nw = 0, B = 0;
    int64_t n = 0;
    std::vector<uint64_t> dom;         // B*25 hash domains
    std::vector<uint64_t> nchar;       // B*25 words per level (1 + dom/64)
    std::vector<uint64_t> woff;        // B*25 word offset of the level in `bits` (levels padded to 8 words)
    std::vector<uint64_t> lastrank;    // B
    std::vector<int64_t> bsz;          // B
    std::vector<uint64_t> starts;      // B+1, as serialized (last entry not accumulated)
    uint64_t total_words = 0;
    DArr<uint64_t> bits;               // all buckets, all levels
    DArr<uint64_t> ranks;              // one per 8 words
    DArr<uint64_t> d_dom, d_woff, d_starts;   // device copies of the tables
};

// scans / utilities (scan.cu)
void exclusive_scan_u64(Ctx *ctx, const uint64_t *in, uint64_t *out, size_t n);
void exclusive_scan_u32_to_u64(Ctx *ctx, const uint32_t *in, uint64_t *out, size_t n);
void ensure_reads_on_device(Ctx *ctx);

// count.cu
enum CountMode { kCanonical = 0, kAllWindows = 1 };
KSet *count_from_reads(Ctx *ctx, int K, int B, int mode);
KSet *kmers_from_kpomers(Ctx *ctx, const KSet *kp, int B);

// distributed count (count.cu)
struct DistState;
struct DistPlan;
DistState *dist_begin(Ctx *ctx, int K, int B, int mode, int world, int rank);
uint32_t dist_num_partitions(const DistState *d);
void dist_local_counts(DistState *d, uint64_t *h_out);
void dist_plan(DistState *d, const uint64_t *cnt_all, uint64_t budget_bytes, int *npass, uint64_t *xchg_records);
void dist_ipc_handle(DistState *d, uint8_t *out64);
void dist_open_peers(DistState *d, const uint8_t *handles);
int dist_adopt(DistState *d, DistState *old);
void dist_scatter(DistState *d, int p);
void dist_exchange(DistState *d, int p);
void dist_sort(DistState *d, int p);
KSet *dist_end(DistState *d);
void dist_free(DistState *d);
// pure host planning (exported for CPU tests): returns npass, fills pass boundaries / owner boundaries / totals
int dist_plan_host(int world, int B, int rA, const uint64_t *cnt_all, uint64_t budget_bytes, int record_bytes, int *pass_b /*B+1 max*/, uint64_t *max_recv);

// mphf.cu
Mphf *mphf_build(Ctx *ctx, const KSet *ks);
size_t mphf_serialized_size(const Mphf *m);
void mphf_serialize_to(const Mphf *m, uint8_t *out, size_t cap);
void mphf_lookup_host_keys(Ctx *ctx, const Mphf *m, const uint64_t *h_keys, int64_t n, uint64_t *h_out);

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace sg
