// api.cu -- the extern "C" boundary (include/spades_b200.h). No exceptions cross it; no CPU fallbacks live behind it.
#include <fcntl.h>
#include <stdio.h>
#include <unistd.h>

#include <string>

#include "../../include/spades_b200.h"
#include "graph.h"
#include "host_par.h"
#include "sgpu_internal.h"

using namespace sg;

struct sgpu_ctx { Ctx c; bool own_stream = true; };

// Lifetime: k-mer sets, indexes, graphs and distributed counts keep blocks of their context's arena. sgpu_destroy() while any of
// them is alive only marks the context; the last child to be freed then tears it down (no use-after-free whatever the order,
// e.g. Python finalisers at interpreter exit).
static void ctx_teardown(sgpu_ctx *ctx) {
    cudaSetDevice(ctx->c.device);
    ctx->c.r_words.release(); ctx->c.r_offs.release(); ctx->c.r_lens.release();
    ctx->c.pool_trim();
    if (ctx->c.stream && ctx->own_stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}
static void child_add(Ctx *c) { c->live_children++; }
static void child_release(Ctx *c) {
    if (--c->live_children == 0 && c->destroy_pending) ctx_teardown(static_cast<sgpu_ctx *>(c->owner));
}
struct sgpu_kset { KSet *s; };
struct sgpu_mphf { Mphf *m; };
struct sgpu_graph { Graph *g; };
struct sgpu_edge_index { EdgeIndex *e; };

static int fail(Ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code ? code : SGPU_EINTERNAL;
}
#define API_TRY(ctxptr, ...)                                    \
    try { __VA_ARGS__; return SGPU_OK; }                           \
    catch (const sg::Error &e) { return fail((ctxptr), e.code, e.what()); } \
    catch (const std::bad_alloc &) { return fail((ctxptr), SGPU_ENOMEM, "host allocation failed"); } \
    catch (const std::exception &e) { return fail((ctxptr), SGPU_EINTERNAL, e.what()); }

extern "C" {

int sgpu_create(const sgpu_config *cfg, sgpu_ctx **out) {
    if (!out) return SGPU_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return SGPU_ENODEV; }
    int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= ndev) return SGPU_EINVAL;
    if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return SGPU_ENODEV; }
    sgpu_ctx *h = new sgpu_ctx();
    h->c.owner = h;
    h->c.device = dev;
    h->c.hbm_budget = cfg ? (size_t)cfg->hbm_budget_bytes : 0;
    h->c.verbose = cfg ? cfg->verbose : 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); delete h; return SGPU_ECUDA; }
    if (cfg && cfg->stream) { h->c.stream = (cudaStream_t)(uintptr_t)cfg->stream; h->own_stream = false; }
    else if (cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete h; return SGPU_ECUDA; }
    h->c.num_sms = prop.multiProcessorCount;
    *out = h;
    return SGPU_OK;
}

void sgpu_destroy(sgpu_ctx *ctx) {
    if (!ctx || ctx->c.destroy_pending) return;
    if (ctx->c.live_children > 0) { ctx->c.destroy_pending = true; return; }      // deferred until the last child is freed
    ctx_teardown(ctx);
}

const char *sgpu_last_error(const sgpu_ctx *ctx) { return ctx ? ctx->c.err.c_str() : "no context"; }

int sgpu_get_times(const sgpu_ctx *ctx, sgpu_times *out) {
    if (!ctx || !out) return SGPU_EINVAL;
    const PhaseTimes &t = ctx->c.times;
    out->extract_count_ms = t.extract_count; out->extract_scatter_ms = t.extract_scatter; out->refine_ms = t.refine;
    out->local_sort_ms = t.local_sort; out->compact_ms = t.compact; out->mphf_ms = t.mphf; out->exchange_ms = t.exchange;
    out->instances = t.instances; out->passes = t.passes; out->launches = ctx->c.launches; out->peak_bytes = ctx->c.peak; out->cached_bytes = ctx->c.pool_cached;
    return SGPU_OK;
}

int sgpu_reads_clear(sgpu_ctx *ctx) {
    if (!ctx) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        c->h_words.clear(); c->h_offs.clear(); c->h_lens.clear();
        c->r_words.release(); c->r_offs.release(); c->r_lens.release();
        c->d_words = nullptr; c->d_offs = nullptr; c->d_lens = nullptr; c->n_reads = 0; c->n_words = 0; c->staged_dirty = false;
    })
}

int sgpu_reads_append_packed(sgpu_ctx *ctx, const uint64_t *words, uint64_t nwords, const uint64_t *offs, const uint32_t *lens, int64_t nreads) {
    if (!ctx || nreads < 0 || (nreads && (!words || !offs || !lens))) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        const uint64_t base = c->h_words.size();
        for (int64_t r = 0; r < nreads; ++r) {
            const uint64_t need = ((uint64_t)lens[r] + 31) / 32;
            SG_CHECK(offs[r] + need <= nwords, SGPU_EINVAL, "read extends past the word buffer");
            c->h_offs.push_back(base + offs[r]);
            c->h_lens.push_back(lens[r]);
        }
        c->h_words.insert(c->h_words.end(), words, words + nwords);
        c->staged_dirty = true;
    })
}

int sgpu_reads_append_batch(sgpu_ctx *ctx, const sgpu_read_batch *b) {
    if (!ctx || !b) return SGPU_EINVAL;
    return sgpu_reads_append_packed(ctx, sgpu_read_batch_words(b), sgpu_read_batch_num_words(b), sgpu_read_batch_offs(b), sgpu_read_batch_lens(b),
                                    sgpu_read_batch_num_reads(b));
}

int sgpu_reads_upload(sgpu_ctx *ctx, const uint64_t *words, uint64_t nwords, const uint64_t *offs, const uint32_t *lens, int64_t nreads) {
    if (!ctx || nreads < 0 || (nreads && (!words || !offs || !lens))) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        c->h_words.clear(); c->h_offs.clear(); c->h_lens.clear(); c->staged_dirty = false;
        if (c->r_words.n < nwords + 4) c->r_words.alloc(c, nwords + 4, true);
        if (c->r_offs.n < (size_t)nreads + 1) c->r_offs.alloc(c, (size_t)nreads + 1, true);
        if (c->r_lens.n < (size_t)nreads + 1) c->r_lens.alloc(c, (size_t)nreads + 1, true);
        // one asynchronous copy per array on the context's stream (a chunked copy overlapped with the first pass over the reads was
        // measured slower on the B200: the per-chunk launches cost more than the overlap won). The caller's buffers must stay
        // valid and unmodified until the next call that synchronises (sgpu_count / sgpu_dist_begin), see spades_b200.h.
        if (nwords) SG_CUDA(cudaMemcpyAsync(c->r_words.p, words, nwords * 8, cudaMemcpyHostToDevice, c->stream));
        SG_CUDA(cudaMemsetAsync(c->r_words.p + nwords, 0, 4 * 8, c->stream));       // the padding words the window loads may touch
        if (nreads) {
            SG_CUDA(cudaMemcpyAsync(c->r_offs.p, offs, (size_t)nreads * 8, cudaMemcpyHostToDevice, c->stream));
            SG_CUDA(cudaMemcpyAsync(c->r_lens.p, lens, (size_t)nreads * 4, cudaMemcpyHostToDevice, c->stream));
        }
        c->d_words = c->r_words.p; c->d_offs = c->r_offs.p; c->d_lens = c->r_lens.p; c->n_reads = nreads; c->n_words = nwords;
    })
}

int sgpu_reads_pack_text(sgpu_ctx *ctx, const char *text, uint64_t text_bytes, const uint64_t *seq_off, const uint32_t *seq_len, int64_t nreads, int longest_valid) {
    if (!ctx || nreads < 0 || (nreads && (!text || !seq_off || !seq_len))) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); reads_pack_text(c, text, text_bytes, seq_off, seq_len, nreads, longest_valid); })
}
int sgpu_reads_cov_filter(sgpu_ctx *ctx, int K, unsigned threshold, int apply, uint8_t *keep_out, uint64_t *stats) {
    if (!ctx) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); cov_filter(c, K, threshold, apply, keep_out, stats); })
}
int sgpu_reads_info(sgpu_ctx *ctx, int64_t *nreads, uint64_t *nwords) {
    if (!ctx || !nreads || !nwords) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); ensure_reads_on_device(c); *nreads = c->n_reads; *nwords = c->n_words; })
}
int sgpu_reads_download(sgpu_ctx *ctx, uint64_t *words, uint64_t *offs, uint32_t *lens) {
    if (!ctx) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); reads_download(c, words, offs, lens); })
}

int sgpu_reads_adopt_device(sgpu_ctx *ctx, const uint64_t *d_words, uint64_t nwords, const uint64_t *d_offs, const uint32_t *d_lens, int64_t nreads) {
    if (!ctx || nreads < 0) return SGPU_EINVAL;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        c->h_words.clear(); c->h_offs.clear(); c->h_lens.clear(); c->staged_dirty = false;
        c->r_words.release(); c->r_offs.release(); c->r_lens.release();
        c->d_words = d_words; c->d_offs = d_offs; c->d_lens = d_lens; c->n_reads = nreads; c->n_words = nwords;
    })
}

int sgpu_count(sgpu_ctx *ctx, int K, int num_buckets, int mode, sgpu_kset **out) {
    if (!ctx || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CHECK(mode == SGPU_CANONICAL || mode == SGPU_ALL_WINDOWS, SGPU_EINVAL, "bad mode");
        SG_CUDA(cudaSetDevice(c->device));
        KSet *s = count_from_reads(c, K, num_buckets, mode);
        *out = new sgpu_kset{s};
        child_add(c);
    })
}

int sgpu_kmers_from_kpomers(sgpu_ctx *ctx, const sgpu_kset *kpomers, int num_buckets, sgpu_kset **out) {
    if (!ctx || !kpomers || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        KSet *s = kmers_from_kpomers(c, kpomers->s, num_buckets);
        *out = new sgpu_kset{s};
        child_add(c);
    })
}

int64_t sgpu_kset_size(const sgpu_kset *s) { return s ? s->s->n : -1; }
int sgpu_kset_k(const sgpu_kset *s) { return s ? s->s->K : -1; }
int sgpu_kset_num_buckets(const sgpu_kset *s) { return s ? s->s->B : -1; }
int sgpu_kset_record_bytes(const sgpu_kset *s) { return s ? 8 * s->s->nw : -1; }
int sgpu_kset_bucket_sizes(const sgpu_kset *s, int64_t *out) {
    if (!s || !out) return SGPU_EINVAL;
    for (int b = 0; b < s->s->B; ++b) out[b] = s->s->bsz[b];
    return SGPU_OK;
}

}  // extern "C"

template <class T, class Get>
static void download_range(const KSet *ks, int64_t first, int64_t n, T *out, size_t per, Get get) {
    SG_CHECK(first >= 0 && n >= 0 && first + n <= ks->n, SGPU_EINVAL, "range outside the k-mer set");
    Ctx *c = ks->ctx;
    SG_CUDA(cudaSetDevice(c->device));
    int64_t done = 0;
    for (const Chunk &ch : ks->chunks) {
        const int64_t lo = std::max(first, ch.first), hi = std::min(first + n, ch.first + ch.n);
        if (lo >= hi) continue;
        SG_CUDA(cudaMemcpyAsync(out + (size_t)(lo - first) * per, get(ch) + (size_t)(lo - ch.first) * per, (size_t)(hi - lo) * per * sizeof(T),
                                cudaMemcpyDeviceToHost, c->stream));
        done += hi - lo;
    }
    SG_CUDA(cudaStreamSynchronize(c->stream));
    SG_CHECK(done == n, SGPU_EINTERNAL, "chunk table does not cover the range");
}

static void write_range(const KSet *ks, int64_t first, int64_t n, FILE *f) {
    const size_t W = (size_t)ks->nw;
    const int64_t step = 1 << 22;
    std::vector<uint64_t> buf;
    for (int64_t o = 0; o < n; o += step) {
        const int64_t m = std::min(step, n - o);
        buf.resize((size_t)m * W);
        download_range<uint64_t>(ks, first + o, m, buf.data(), W, [](const Chunk &ch) { return ch.keys.p; });
        SG_CHECK(fwrite(buf.data(), 8 * W, (size_t)m, f) == (size_t)m, SGPU_EIO, "short write");
    }
}

extern "C" {

int sgpu_kset_checksum(const sgpu_kset *s, uint64_t *out4) {
    if (!s || !out4) return SGPU_EINVAL;
    Ctx *c = s->s->ctx;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); kset_checksum(s->s, out4); })
}
int sgpu_kset_download_keys(const sgpu_kset *s, int64_t first, int64_t n, uint64_t *out) {
    if (!s || (n && !out)) return SGPU_EINVAL;
    Ctx *c = s->s->ctx;
    API_TRY(c, { download_range<uint64_t>(s->s, first, n, out, (size_t)s->s->nw, [](const Chunk &ch) { return ch.keys.p; }); })
}
int sgpu_kset_download_counts(const sgpu_kset *s, int64_t first, int64_t n, uint32_t *out) {
    if (!s || (n && !out)) return SGPU_EINVAL;
    Ctx *c = s->s->ctx;
    API_TRY(c, {
        SG_CHECK(s->s->has_counts, SGPU_EINVAL, "this k-mer set carries no multiplicities");
        download_range<uint32_t>(s->s, first, n, out, 1, [](const Chunk &ch) { return ch.counts.p; });
    })
}

int sgpu_kset_write_buckets(const sgpu_kset *s, const char *prefix) {
    if (!s || !prefix) return SGPU_EINVAL;
    Ctx *c = s->s->ctx;
    API_TRY(c, {
        for (int b = 0; b < s->s->B; ++b) {
            std::string p = std::string(prefix) + "." + std::to_string(b);
            FILE *f = fopen(p.c_str(), "wb");
            SG_CHECK(f, SGPU_EIO, "cannot open bucket file for writing");
            try { write_range(s->s, s->s->bstart[b], s->s->bsz[b], f); } catch (...) { fclose(f); throw; }
            fclose(f);
        }
    })
}
int sgpu_kset_write_final(const sgpu_kset *s, const char *path) {
    if (!s || !path) return SGPU_EINVAL;
    Ctx *c = s->s->ctx;
    API_TRY(c, {
        FILE *f = fopen(path, "wb");
        SG_CHECK(f, SGPU_EIO, "cannot open final_kmers for writing");
        try { write_range(s->s, 0, s->s->n, f); } catch (...) { fclose(f); throw; }
        fclose(f);
    })
}
void sgpu_kset_free(sgpu_kset *s) {
    if (!s) return;
    if (s->s) { Ctx *c = s->s->ctx; cudaSetDevice(c->device); delete s->s; child_release(c); }
    delete s;
}

int sgpu_mphf_build(sgpu_ctx *ctx, const sgpu_kset *s, sgpu_mphf **out) {
    if (!ctx || !s || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        Mphf *m = mphf_build(c, s->s);
        *out = new sgpu_mphf{m};
        child_add(c);
    })
}
int64_t sgpu_mphf_serialized_size(const sgpu_mphf *m) { return m ? (int64_t)mphf_serialized_size(m->m) : -1; }
int sgpu_mphf_serialize(const sgpu_mphf *m, uint8_t *out, int64_t cap) {
    if (!m || !out || cap < 0) return SGPU_EINVAL;
    Ctx *c = m->m->ctx;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        mphf_serialize_to(m->m, out, (size_t)cap);
    })
}
int sgpu_mphf_lookup(const sgpu_mphf *m, const uint64_t *keys, int64_t n, uint64_t *out_idx) {
    if (!m || (n && (!keys || !out_idx))) return SGPU_EINVAL;
    Ctx *c = m->m->ctx;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); mphf_lookup_host_keys(c, m->m, keys, n, out_idx); })
}
void sgpu_mphf_free(sgpu_mphf *m) {
    if (!m) return;
    if (m->m) { Ctx *c = m->m->ctx; cudaSetDevice(c->device); delete m->m; child_release(c); }
    delete m;
}

int sgpu_graph_build(sgpu_ctx *ctx, const sgpu_kset *kpomers, const sgpu_kset *kmers, const sgpu_mphf *kmer_index, const sgpu_mphf *kpomer_index,
                     int keep_perfect_loops, sgpu_graph **out) {
    if (!ctx || !kpomers || !kmers || !kmer_index || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        GraphOptions opt;
        opt.keep_perfect_loops = keep_perfect_loops != 0;
        Graph *g = graph_build(c, kpomers->s, kmers->s, kmer_index->m, kpomer_index ? kpomer_index->m : nullptr, opt);
        *out = new sgpu_graph{g};
        child_add(c);
    })
}
int sgpu_graph_build_ex(sgpu_ctx *ctx, const sgpu_kset *kpomers, const sgpu_kset *kmers, const sgpu_mphf *kmer_index, const sgpu_mphf *kpomer_index,
                        int keep_perfect_loops, uint64_t early_tip_length_bound, sgpu_graph **out) {
    if (!ctx || !kpomers || !kmers || !kmer_index || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        GraphOptions opt;
        opt.keep_perfect_loops = keep_perfect_loops != 0;
        opt.early_tip_length_bound = early_tip_length_bound;
        Graph *g = graph_build(c, kpomers->s, kmers->s, kmer_index->m, kpomer_index ? kpomer_index->m : nullptr, opt);
        *out = new sgpu_graph{g};
        child_add(c);
    })
}
int sgpu_graph_build_opts(sgpu_ctx *ctx, const sgpu_kset *kpomers, const sgpu_kset *kmers, const sgpu_mphf *kmer_index, const sgpu_mphf *kpomer_index,
                          const sgpu_graph_options *o, sgpu_graph **out) {
    if (!ctx || !kpomers || !kmers || !kmer_index || !o || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CHECK(!o->early_at_clipper || (o->at_ratio > 0.0 && o->at_ratio <= 1.0 && o->at_max_length >= 1), SGPU_EINVAL, "bad A/T clipper parameters");
        SG_CUDA(cudaSetDevice(c->device));
        GraphOptions opt;
        opt.keep_perfect_loops = o->keep_perfect_loops != 0;
        opt.early_tip_length_bound = o->early_tip_length_bound;
        opt.early_at = o->early_at_clipper != 0;
        opt.at_ratio = o->at_ratio; opt.at_min_len = o->at_min_length; opt.at_max_len = o->at_max_length;
        Graph *g = graph_build(c, kpomers->s, kmers->s, kmer_index->m, kpomer_index ? kpomer_index->m : nullptr, opt);
        *out = new sgpu_graph{g};
        child_add(c);
    })
}
int sgpu_graph_at_clipper_stats(const sgpu_graph *g, uint64_t *out4) {
    if (!g || !out4) return SGPU_EINVAL;
    for (int i = 0; i < 4; ++i) out4[i] = g->g->at_stats[i];
    return SGPU_OK;
}
int sgpu_graph_tip_clipper_stats(const sgpu_graph *g, uint64_t *out3) {
    if (!g || !out3) return SGPU_EINVAL;
    for (int i = 0; i < 3; ++i) out3[i] = g->g->tc_stats[i];
    return SGPU_OK;
}
int sgpu_graph_masks(const sgpu_graph *g, uint8_t *out, int64_t n) {
    if (!g || (n && !out)) return SGPU_EINVAL;
    Ctx *c = g->g->ctx;
    API_TRY(c, {
        SG_CHECK(n == g->g->km->n, SGPU_EINVAL, "mask array size != number of k-mers");
        SG_CUDA(cudaSetDevice(c->device));
        if (n) SG_CUDA(cudaMemcpy(out, g->g->masks_final.p, (size_t)n, cudaMemcpyDeviceToHost));
    })
}
int sgpu_graph_coverage(const sgpu_graph *g, uint32_t *out, int64_t n) {
    if (!g || (n && !out)) return SGPU_EINVAL;
    Ctx *c = g->g->ctx;
    API_TRY(c, {
        SG_CHECK(g->g->cov.p, SGPU_EINVAL, "graph was built without a (k+1)-mer index: no coverage");
        SG_CHECK(n == g->g->kp->n, SGPU_EINVAL, "coverage array size != number of (k+1)-mers");
        SG_CUDA(cudaSetDevice(c->device));
        if (n) SG_CUDA(cudaMemcpy(out, g->g->cov.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    })
}
int64_t sgpu_graph_histogram(const sgpu_graph *g, uint64_t *out, int64_t cap) {
    if (!g) return -1;
    Ctx *c = g->g->ctx;
    try {
        cudaSetDevice(c->device);
        std::vector<uint64_t> h = graph_histogram(c, g->g);
        if (out) for (int64_t i = 0; i < (int64_t)h.size() && i < cap; ++i) out[i] = h[i];
        return (int64_t)h.size();
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
}
int64_t sgpu_graph_num_unitigs(const sgpu_graph *g) { return g ? (int64_t)g->g->edge_len.size() : -1; }
int64_t sgpu_graph_unitig_bases(const sgpu_graph *g) { return g ? (int64_t)g->g->seq.size() : -1; }
int sgpu_graph_unitigs(const sgpu_graph *g, char *out, uint32_t *lens) {
    if (!g) return SGPU_EINVAL;
    if (out && !g->g->seq.empty()) memcpy(out, g->g->seq.data(), g->g->seq.size());
    if (lens) for (size_t i = 0; i < g->g->edge_len.size(); ++i) lens[i] = g->g->edge_len[i];
    return SGPU_OK;
}
int64_t sgpu_graph_gfa(const sgpu_graph *g, const char *version, char *out, int64_t cap) {
    if (!g) return -1;
    try {
        std::string t = graph_gfa(g->g, version ? version : "SPAdes-4.3.0-dev");
        if (out && (int64_t)t.size() <= cap) memcpy(out, t.data(), t.size());
        return (int64_t)t.size();
    } catch (const std::exception &e) { g->g->ctx->err = e.what(); return -1; }
}
int sgpu_graph_write_gfa(const sgpu_graph *g, const char *version, const char *path) {
    if (!g || !path) return SGPU_EINVAL;
    Ctx *c = g->g->ctx;
    API_TRY(c, {
        std::vector<std::string> ch = graph_gfa_chunks(g->g, version ? version : "SPAdes-4.3.0-dev");
        // the pieces go to their offsets of the file concurrently (a 5-25 GB text: one writer is the bottleneck of the whole path)
        const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        SG_CHECK(fd >= 0, SGPU_EIO, "cannot open GFA file for writing");
        std::vector<uint64_t> off(ch.size() + 1, 0);
        for (size_t i = 0; i < ch.size(); ++i) off[i + 1] = off[i] + ch[i].size();
        std::vector<int> okv(ch.size(), 1);
        par_chunks(ch.size(), (int)std::min<size_t>(ch.size(), 32), [&](int, size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const char *p = ch[i].data();
                uint64_t left = ch[i].size(), at = off[i];
                while (left) {
                    const ssize_t w = pwrite(fd, p, (size_t)std::min<uint64_t>(left, 1ull << 30), (off_t)at);
                    if (w <= 0) { okv[i] = 0; break; }
                    p += w; at += (uint64_t)w; left -= (uint64_t)w;
                }
                std::string().swap(ch[i]);
            }
        });
        bool ok = close(fd) == 0;
        for (int v : okv) ok = ok && v;
        SG_CHECK(ok, SGPU_EIO, "short write");
    })
}
int sgpu_edge_index_build(sgpu_ctx *ctx, const sgpu_graph *g, int K, int num_buckets, sgpu_edge_index **out) {
    if (!ctx || !g || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        EdgeIndex *e = edge_index_build(c, g->g, K, num_buckets);
        *out = new sgpu_edge_index{e};
        child_add(c);
    })
}
int sgpu_edge_index_k(const sgpu_edge_index *e) { return e ? e->e->K : -1; }
int64_t sgpu_edge_index_size(const sgpu_edge_index *e) { return e ? e->e->ks->n : -1; }
int64_t sgpu_edge_index_serialized_size(const sgpu_edge_index *e) { return e ? (int64_t)mphf_serialized_size(e->e->m) : -1; }
int sgpu_edge_index_serialize(const sgpu_edge_index *e, uint8_t *out, int64_t cap) {
    if (!e || !out || cap < 0) return SGPU_EINVAL;
    Ctx *c = e->e->ctx;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        mphf_serialize_to(e->e->m, out, (size_t)cap);
        if (e->e->single_segment) memset(out + mphf_serialized_size(e->e->m) - 8, 0, 8);     // see edge_index.cu: segment_starts_[1] of the single-index branch
    })
}
int sgpu_edge_index_values(const sgpu_edge_index *e, uint64_t *edge_ids, uint32_t *offsets, int64_t n) {
    if (!e || (n && (!edge_ids || !offsets))) return SGPU_EINVAL;
    Ctx *c = e->e->ctx;
    API_TRY(c, {
        SG_CHECK(n == e->e->ks->n, SGPU_EINVAL, "value array size != number of K-mers in the edge index");
        SG_CUDA(cudaSetDevice(c->device));
        if (n) {
            SG_CUDA(cudaMemcpy(edge_ids, e->e->edge_id.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
            SG_CUDA(cudaMemcpy(offsets, e->e->offset.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        }
    })
}
int sgpu_edge_index_lookup(const sgpu_edge_index *e, const uint64_t *keys, int64_t n, uint64_t *out_idx) {
    if (!e || (n && (!keys || !out_idx))) return SGPU_EINVAL;
    Ctx *c = e->e->ctx;
    API_TRY(c, { SG_CUDA(cudaSetDevice(c->device)); mphf_lookup_host_keys(c, e->e->m, keys, n, out_idx); })
}
void sgpu_edge_index_free(sgpu_edge_index *e) {
    if (!e) return;
    if (e->e) { Ctx *c = e->e->ctx; cudaSetDevice(c->device); delete e->e; child_release(c); }
    delete e;
}
void sgpu_graph_free(sgpu_graph *g) {
    if (!g) return;
    if (g->g) { Ctx *c = g->g->ctx; cudaSetDevice(c->device); delete g->g; child_release(c); }
    delete g;
}

}  // extern "C"

struct sgpu_dist { DistState *d; Ctx *c; };
extern "C" {
int sgpu_dist_begin(sgpu_ctx *ctx, int K, int num_buckets, int mode, int world, int rank, sgpu_dist **out) {
    if (!ctx || !out) return SGPU_EINVAL;
    *out = nullptr;
    Ctx *c = &ctx->c;
    API_TRY(c, {
        SG_CUDA(cudaSetDevice(c->device));
        DistState *d = dist_begin(c, K, num_buckets, mode, world, rank);
        *out = new sgpu_dist{d, c};
        child_add(c);
    })
}
int64_t sgpu_dist_num_partitions(const sgpu_dist *d) { return d ? (int64_t)dist_num_partitions(d->d) : -1; }
int sgpu_dist_local_counts(sgpu_dist *d, uint64_t *out) {
    if (!d || !out) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_local_counts(d->d, out); })
}
int sgpu_dist_plan(sgpu_dist *d, const uint64_t *all_counts, uint64_t *total_records) {
    if (!d || !all_counts || !total_records) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_plan(d->d, all_counts, total_records); })
}
int sgpu_dist_free_bytes(sgpu_dist *d, uint64_t *out) {
    if (!d || !out) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); *out = dist_free_bytes(d->d); })
}
int sgpu_dist_next_pass(sgpu_dist *d, uint64_t budget_bytes, int *pass) {
    if (!d || !pass) return SGPU_EINVAL;
    *pass = -1;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); *pass = dist_next_pass(d->d, budget_bytes); })
}
int sgpu_dist_ipc_handle(sgpu_dist *d, uint8_t *out) {
    if (!d || !out) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_ipc_handle(d->d, out); })
}
int sgpu_dist_open_peers(sgpu_dist *d, const uint8_t *handles) {
    if (!d || !handles) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_open_peers(d->d, handles); })
}
int sgpu_dist_scatter(sgpu_dist *d, int pass) {
    if (!d) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_scatter(d->d, pass); })
}
int sgpu_dist_exchange(sgpu_dist *d, int pass) {
    if (!d) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_exchange(d->d, pass); })
}
int sgpu_dist_sort(sgpu_dist *d, int pass) {
    if (!d) return SGPU_EINVAL;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); dist_sort(d->d, pass); })
}
int sgpu_dist_end(sgpu_dist *d, sgpu_kset **out) {
    if (!d || !out) return SGPU_EINVAL;
    *out = nullptr;
    API_TRY(d->c, { SG_CUDA(cudaSetDevice(d->c->device)); KSet *s = dist_end(d->d); *out = new sgpu_kset{s}; child_add(d->c); })
}
void sgpu_dist_free(sgpu_dist *d) {
    if (!d) return;
    Ctx *c = d->c;
    cudaSetDevice(c->device);
    dist_free(d->d);
    delete d;
    child_release(c);
}
int sgpu_dist_plan_host(int world, int num_buckets, int key_bits_in_partition, const uint64_t *all_counts, uint64_t budget_bytes, int record_bytes,
                        int *pass_bounds, uint64_t *max_recv) {
    if (world < 1 || num_buckets < 1 || !all_counts || !pass_bounds || !max_recv) return -1;
    try { return dist_plan_host(world, num_buckets, key_bits_in_partition, all_counts, budget_bytes, record_bytes, pass_bounds, max_recv); }
    catch (...) { return -1; }
}
}  // extern "C"

// ---- self test of kmer_dev.cuh on host and device -----------------------------------------------------------------------
template <int NW>
__host__ __device__ uint64_t selftest_one(int op, int K, uint64_t arg, const uint64_t *key) {
    Kmer<NW> k;
    for (int q = 0; q < NW; ++q) k.w[q] = key[q];
    switch (op) {
        case 0: return xxh3_64<NW>(k);
        case 1: return xxh3_128<NW>(k).lo;
        case 2: return xxh3_128<NW>(k).hi;
        case 3: return kmer_bucket<NW>(k, (uint32_t)arg);
        case 4: return kmer_is_minimal<NW>(k, kmer_rc<NW>(k, K)) ? 1 : 0;
        case 9: return key_bits<NW>(k, K, (int)(arg >> 8), (int)(arg & 255));
        default: {
            int j = op - 5;
            Kmer<NW> r = kmer_rc<NW>(k, K);
            return j < NW ? r.w[j] : 0;
        }
    }
}
// op 10: the rolling window of the level-A kernels against direct extraction. keys = one packed sequence of n*NW words,
// arg = its length in bases; unit u walks windows [24u, 24u+24). out[u] = (windows walked << 32) | mismatches.
template <int NW>
__host__ __device__ uint64_t selftest_roll_unit(int K, int L, const uint64_t *seq, int64_t u) {
    const int nwin = L - K + 1;
    const int64_t j0 = u * 24;
    if (j0 >= nwin) return 0;
    const int cnt = nwin - j0 < 24 ? (int)(nwin - j0) : 24;
    RollState<NW> st;
    roll_init<NW>(st, seq, (int)j0, K, cnt);
    uint64_t bad = 0;
    for (int s = 0; s < cnt; ++s) {
        if (s) roll_next<NW>(st, K);
        const Kmer<NW> f = kmer_window<NW>(seq, j0 + s, K);
        const Kmer<NW> r = kmer_rc<NW>(f, K);
        if (!kmer_eq<NW>(f, st.f) || !kmer_eq<NW>(r, st.r)) ++bad;
    }
    return ((uint64_t)cnt << 32) | bad;
}
template <int NW>
__global__ void selftest_roll_k(int K, int L, const uint64_t *seq, int64_t n, uint64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = selftest_roll_unit<NW>(K, L, seq, i);
}
template <int NW>
__global__ void selftest_k(int op, int K, uint64_t arg, const uint64_t *keys, int64_t n, uint64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = selftest_one<NW>(op, K, arg, keys + i * NW);
}
template <int NW>
static void selftest_nw(Ctx *c, int on_device, int op, int K, uint64_t arg, const uint64_t *keys, int64_t n, uint64_t *out) {
    if (op == 10) SG_CHECK((int64_t)arg >= K && (int64_t)arg <= 32 * n * NW, SGPU_EINVAL, "roll self test: bad sequence length");
    if (!on_device) {
        for (int64_t i = 0; i < n; ++i) out[i] = op == 10 ? selftest_roll_unit<NW>(K, (int)arg, keys, i) : selftest_one<NW>(op, K, arg, keys + i * NW);
        return;
    }
    SG_CHECK(c, SGPU_EINVAL, "device self test needs a context");
    DArr<uint64_t> dk(c, (size_t)n * NW), dout(c, (size_t)n);
    SG_CUDA(cudaMemcpyAsync(dk.p, keys, (size_t)n * NW * 8, cudaMemcpyHostToDevice, c->stream));
    if (op == 10) selftest_roll_k<NW><<<div_up(n, 256), 256, 0, c->stream>>>(K, (int)arg, dk.p, n, dout.p);
    else selftest_k<NW><<<div_up(n, 256), 256, 0, c->stream>>>(op, K, arg, dk.p, n, dout.p);
    c->launches++;
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    SG_CUDA(cudaStreamSynchronize(c->stream));
}
// op 11 (host only): the sector-pairing mailbox protocol of pair_mailbox.cuh under real concurrency, see selftest_host.cpp
extern "C" int sg_selftest_pair_mailbox(uint64_t arg, int64_t per_thread, uint64_t *out);

extern "C" int sgpu_selftest(sgpu_ctx *ctx, int on_device, int op, int K, uint64_t arg, const uint64_t *keys, int64_t n, uint64_t *out) {
    if (K < 1 || K > 128 || n < 0 || (n && (!keys || !out))) return SGPU_EINVAL;
    Ctx *c = ctx ? &ctx->c : nullptr;
    if (on_device && !c) return SGPU_EINVAL;
    if (op == 11) {
        if (on_device || n < 3) return SGPU_EINVAL;
        return sg_selftest_pair_mailbox(arg, (int64_t)keys[0], out);
    }
    API_TRY(c, {
        if (on_device) SG_CUDA(cudaSetDevice(c->device));
        switch (nwords_of(K)) {
            case 1: selftest_nw<1>(c, on_device, op, K, arg, keys, n, out); break;
            case 2: selftest_nw<2>(c, on_device, op, K, arg, keys, n, out); break;
            case 3: selftest_nw<3>(c, on_device, op, K, arg, keys, n, out); break;
            default: selftest_nw<4>(c, on_device, op, K, arg, keys, n, out); break;
        }
    })
}
