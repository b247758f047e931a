// host_par.h -- the host halves of the graph phase (link records, GFA text, unitig packing) on the host threads.
#pragma once
#include <stddef.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace sg {

// Config 3 has ~10^8 edges and a 10-25 GB GFA: everything below runs on the host threads in contiguous chunks (the text of a chunk
// depends only on its own edges / vertices), the chunks are emitted in order.
inline int host_threads_for(size_t n) {
    if (n < (size_t)1 << 15) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    return (int)std::min<size_t>(std::min<unsigned>(hw, 64u), n >> 13);
}
template <class F>
inline void par_chunks(size_t n, int T, F f) {               // f(chunk, lo, hi)
    if (T <= 1) { f(0, (size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([=] { f(t, n * (size_t)t / (size_t)T, n * (size_t)(t + 1) / (size_t)T); });
    for (auto &x : th) x.join();
}
// chunk-sort + pairwise merges, every merge of a round on its own thread
template <class T, class Less>
inline void par_sort(std::vector<T> &v, Less less) {
    const size_t n = v.size();
    int P = host_threads_for(n);
    if (P <= 1) { std::sort(v.begin(), v.end(), less); return; }
    int pow2 = 1;
    while (pow2 * 2 <= P) pow2 *= 2;
    P = pow2;
    std::vector<size_t> cut(P + 1);
    for (int i = 0; i <= P; ++i) cut[i] = n * (size_t)i / (size_t)P;
    par_chunks((size_t)P, P, [&](int, size_t lo, size_t hi) { for (size_t c = lo; c < hi; ++c) std::sort(v.begin() + cut[c], v.begin() + cut[c + 1], less); });
    std::vector<T> tmp(n);
    std::vector<T> *src = &v, *dst = &tmp;
    for (int width = 1; width < P; width *= 2) {
        const int pairs = P / (2 * width);
        std::vector<std::thread> th;
        for (int q = 0; q < pairs; ++q) {
            const size_t a = cut[2 * width * q], m = cut[2 * width * q + width], b = cut[2 * width * (q + 1)];
            th.emplace_back([=] { std::merge(src->begin() + a, src->begin() + m, src->begin() + m, src->begin() + b, dst->begin() + a, less); });
        }
        for (auto &x : th) x.join();
        std::swap(src, dst);
    }
    if (src != &v) v.swap(tmp);
}

}  // namespace sg
