// sgpu_internal.h -- internal C++ structures behind the C ABI of include/spades_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <iterator>
#include <map>
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
    DArr(Ctx *c, size_t n_, bool persistent = false) { alloc(c, n_, persistent); }
    DArr(const DArr &) = delete;
    DArr &operator=(const DArr &) = delete;
    DArr(DArr &&o) noexcept { p = o.p; n = o.n; blk = o.blk; ctx = o.ctx; o.p = nullptr; o.n = 0; o.blk = 0; }
    DArr &operator=(DArr &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; blk = o.blk; ctx = o.ctx; o.p = nullptr; o.n = 0; o.blk = 0; }
        return *this;
    }
    ~DArr() { release(); }
    void alloc(Ctx *c, size_t n_, bool persistent = false);
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
    // objects created from this context that still hold arena blocks (k-mer sets, indexes, graphs, distributed counts):
    // sgpu_destroy refuses to run while any is alive (their destructors release blocks into this context's arena)
    int live_children = 0;
    bool destroy_pending = false;
    void *owner = nullptr;                   // the sgpu_ctx this context is embedded in
    // multi-GPU: the peers' arenas mapped through cudaIpc, once per process (rank-indexed; own entry unused)
    std::vector<char *> peer_arena;
    std::vector<uint8_t> peer_handle;        // 64 bytes per rank: the handle a mapping was opened from
    void peer_close();
    // device memory: one arena reserved from the driver at first use and sub-allocated with a coalescing free list.
    // cudaMalloc/cudaFree of tens of GB cost 100s of ms and a 100 M-read step turns over ~300 GB of buffers; inside the
    // arena an allocation is a map lookup. Short-lived buffers (X/Y ping-pong, scratch) grow from the bottom, long-lived
    // results (k-mer chunks, MPHF, masks) from the top, so the big short-lived blocks keep finding the same hole.
    char *arena = nullptr;
    size_t arena_size = 0;
    std::map<size_t, size_t> arena_free;                     // offset -> size, coalesced
    std::vector<std::pair<void *, size_t>> direct;           // allocations that did not fit the arena
    size_t pool_cached = 0;                                  // free bytes inside the arena
    void arena_init();
    void *pool_alloc(size_t bytes, size_t *blk, bool persistent);
    void pool_release(void *p, size_t blk);
    void pool_trim();
    size_t free_bytes();                                     // allocatable bytes (arena free list)
};

inline void Ctx::arena_init() {
    if (arena) return;
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    size_t want = hbm_budget ? hbm_budget : (size_t)((double)f * 0.92);
    // SGPU_ARENA_GB: user option, caps the arena (e.g. to leave ncu room for its replay buffers)
    if (const char *e = getenv("SGPU_ARENA_GB")) { const double gb = atof(e); if (gb >= 0.0625) want = std::min(want, (size_t)(gb * (double)(1ull << 30))); }
    want &= ~(size_t)((2u << 20) - 1);
    while (want >= ((size_t)64 << 20)) {
        void *p = nullptr;
        if (cudaMalloc(&p, want) == cudaSuccess) { arena = (char *)p; arena_size = want; break; }
        cudaGetLastError();
        want = (size_t)((double)want * 0.9) & ~(size_t)((2u << 20) - 1);
    }
    if (!arena) throw Error(4, "cannot reserve the device memory arena");
    arena_free.clear();
    arena_free[0] = arena_size;
    pool_cached = arena_size;
}
inline void Ctx::peer_close() {
    for (char *p : peer_arena) if (p) cudaIpcCloseMemHandle(p);
    peer_arena.clear(); peer_handle.clear();
}
inline void Ctx::pool_trim() {
    peer_close();
    if (arena) cudaFree(arena);
    arena = nullptr; arena_size = 0; arena_free.clear(); pool_cached = 0;
    for (auto &d : direct) cudaFree(d.first);
    direct.clear();
}
inline size_t Ctx::free_bytes() {
    if (!arena) arena_init();
    return pool_cached;
}
inline void *Ctx::pool_alloc(size_t bytes, size_t *blk, bool persistent) {
    if (!arena) arena_init();
    const size_t want = (bytes + 511) & ~(size_t)511;
    if (!persistent) {
        for (auto it = arena_free.begin(); it != arena_free.end(); ++it) {
            if (it->second >= want) {                         // first fit from the bottom
                const size_t off = it->first, sz = it->second;
                arena_free.erase(it);
                if (sz > want) arena_free[off + want] = sz - want;
                pool_cached -= want; *blk = want;
                return arena + off;
            }
        }
    } else {
        for (auto it = arena_free.rbegin(); it != arena_free.rend(); ++it) {
            if (it->second >= want) {                         // last fit, carved from the top end of the hole
                const size_t off = it->first, sz = it->second;
                if (sz == want) arena_free.erase(std::next(it).base());
                else it->second = sz - want;
                pool_cached -= want; *blk = want;
                return arena + off + (sz - want);
            }
        }
    }
    // does not fit (fragmentation or a request beyond the arena): ask the driver directly
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        char m[256];
        snprintf(m, sizeof m, "out of device memory: %zu bytes requested, %zu free in the %zu-byte arena (in use %zu)", want, pool_cached, arena_size, allocated);
        throw Error(4, m);
    }
    direct.push_back({p, want});
    *blk = want;
    return p;
}
inline void Ctx::pool_release(void *p, size_t blk) {
    char *c = (char *)p;
    if (arena && c >= arena && c < arena + arena_size) {
        size_t off = (size_t)(c - arena), sz = blk;
        auto nxt = arena_free.lower_bound(off);
        if (nxt != arena_free.end() && off + sz == nxt->first) { sz += nxt->second; nxt = arena_free.erase(nxt); }
        if (nxt != arena_free.begin()) {
            auto prv = std::prev(nxt);
            if (prv->first + prv->second == off) { prv->second += sz; pool_cached += blk; return; }
        }
        arena_free[off] = sz;
        pool_cached += blk;
        return;
    }
    for (size_t i = 0; i < direct.size(); ++i)
        if (direct[i].first == p) { cudaFree(p); direct.erase(direct.begin() + i); return; }
    cudaFree(p);
}
template <class T>
void DArr<T>::alloc(Ctx *c, size_t n_, bool persistent) {
    release();
    ctx = c; n = n_;
    size_t b = (n_ ? n_ : 1) * sizeof(T);
    p = (T *)c->pool_alloc(b, &blk, persistent);
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

// ingest_gpu.cu
void reads_pack_text(Ctx *ctx, const char *text, uint64_t text_bytes, const uint64_t *seq_off, const uint32_t *seq_len, int64_t n, int longest_valid);
void reads_download(Ctx *ctx, uint64_t *words, uint64_t *offs, uint32_t *lens);
void cov_filter(Ctx *ctx, int K, unsigned thr, int apply, uint8_t *keep_out, uint64_t *stats);   // covfilter.cu

// count.cu
enum CountMode { kCanonical = 0, kAllWindows = 1 };
KSet *count_from_reads(Ctx *ctx, int K, int B, int mode);
KSet *kmers_from_kpomers(Ctx *ctx, const KSet *kp, int B);

void kset_checksum(const KSet *ks, uint64_t *out4);

// distributed count (count.cu)
struct DistState;
struct DistPlan;
DistState *dist_begin(Ctx *ctx, int K, int B, int mode, int world, int rank);
uint32_t dist_num_partitions(const DistState *d);
void dist_local_counts(DistState *d, uint64_t *h_out);
void dist_plan(DistState *d, const uint64_t *cnt_all, uint64_t *total_records);
uint64_t dist_free_bytes(DistState *d);
int dist_next_pass(DistState *d, uint64_t budget_bytes);
void dist_ipc_handle(DistState *d, uint8_t *out96);
void dist_open_peers(DistState *d, const uint8_t *descs);
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
