// host_par.h -- the host halves of the graph phase (link records, GFA text, unitig packing) on the host threads.
#pragma once
#include <stddef.h>
#include <stdlib.h>

#include <algorithm>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

namespace sg {

// std::vector whose resize() leaves trivially-constructible elements uninitialised: a GB-sized array is then first touched (page
// faults included) by the threads that fill it, not zeroed by the one that allocates it
template <class T>
struct default_init_alloc : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_alloc<U>; };
    default_init_alloc() = default;
    template <class U> default_init_alloc(const default_init_alloc<U> &) {}
    template <class U, class... A>
    void construct(U *p, A &&...a) {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T> using raw_vector = std::vector<T, default_init_alloc<T>>;

// Config 3 has ~10^8 edges and a 10-25 GB GFA: everything below runs on the host threads in contiguous chunks (the text of a chunk
// depends only on its own edges / vertices), the chunks are emitted in order.
// SGPU_HOST_THREADS caps the count (1 = the sequential code path; tests compare the two)
inline int host_threads_for(size_t n) {
    if (n < (size_t)1 << 15) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    if (const char *e = getenv("SGPU_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) hw = std::min<unsigned>(hw, (unsigned)v); }
    return (int)std::min<size_t>(std::min<unsigned>(hw, 64u), n >> 13);
}
template <class F>
inline void par_chunks(size_t n, int T, F f) {               // f(chunk, lo, hi)
    if (T <= 1) { f(0, (size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([=] { f(t, n * (size_t)t / (size_t)T, n * (size_t)(t + 1) / (size_t)T); });
    for (auto &x : th) x.join();
}
// sample sort: P chunks sorted concurrently, P-1 splitters from their samples, then thread t gathers the t-th key range of every chunk
// and sorts it -- no serial merge at the top (the pairwise-merge version spent 1.6 s on 67 M link records, most of it in its last rounds)
template <class Vec, class Less>
inline void par_sort(Vec &v, Less less) {
    typedef typename Vec::value_type T;
    const size_t n = v.size();
    const int P = host_threads_for(n);
    if (P <= 1) { std::sort(v.begin(), v.end(), less); return; }
    std::vector<size_t> cut(P + 1);
    for (int i = 0; i <= P; ++i) cut[i] = n * (size_t)i / (size_t)P;
    par_chunks((size_t)P, P, [&](int, size_t lo, size_t hi) { for (size_t c = lo; c < hi; ++c) std::sort(v.begin() + cut[c], v.begin() + cut[c + 1], less); });
    std::vector<T> samples;
    samples.reserve((size_t)P * (size_t)(P - 1));
    for (int c = 0; c < P; ++c)
        for (int q = 1; q < P; ++q) {
            const size_t len = cut[c + 1] - cut[c];
            if (len) samples.push_back(v[cut[c] + len * (size_t)q / (size_t)P]);
        }
    std::sort(samples.begin(), samples.end(), less);
    std::vector<T> split;
    for (int q = 1; q < P; ++q) split.push_back(samples[samples.size() * (size_t)q / (size_t)P]);
    // pos[c][t] = first element of chunk c that belongs to range t (range t = [split[t-1], split[t]))
    std::vector<size_t> pos((size_t)P * (size_t)(P + 1));
    par_chunks((size_t)P, P, [&](int, size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; ++c) {
            pos[c * (P + 1)] = cut[c];
            for (int t = 1; t < P; ++t) pos[c * (P + 1) + t] = (size_t)(std::lower_bound(v.begin() + cut[c], v.begin() + cut[c + 1], split[t - 1], less) - v.begin());
            pos[c * (P + 1) + P] = cut[c + 1];
        }
    });
    std::vector<size_t> out_off(P + 1, 0);
    for (int t = 0; t < P; ++t) {
        size_t sz = 0;
        for (int c = 0; c < P; ++c) sz += pos[(size_t)c * (P + 1) + t + 1] - pos[(size_t)c * (P + 1) + t];
        out_off[t + 1] = out_off[t] + sz;
    }
    raw_vector<T> tmp(n);
    par_chunks((size_t)P, P, [&](int, size_t lo, size_t hi) {
        for (size_t t = lo; t < hi; ++t) {
            size_t o = out_off[t];
            for (int c = 0; c < P; ++c) {
                const size_t a = pos[(size_t)c * (P + 1) + t], b = pos[(size_t)c * (P + 1) + t + 1];
                std::copy(v.begin() + a, v.begin() + b, tmp.begin() + o);
                o += b - a;
            }
            std::sort(tmp.begin() + out_off[t], tmp.begin() + out_off[t + 1], less);
        }
    });
    par_chunks(n, P, [&](int, size_t lo, size_t hi) { std::copy(tmp.begin() + lo, tmp.begin() + hi, v.begin() + lo); });
}

}  // namespace sg
