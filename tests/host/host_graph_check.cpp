// CPU check of the threaded host half of the graph phase (spades_b200/csrc/host_graph.cpp, host_par.h): a synthetic Graph (edges, link
// records, coverages -- the fields graph_gfa reads) large enough for the threaded code path; prints the FNV-1a hash and the size of the
// GFA text. tests/test_host_graph.py runs it with SGPU_HOST_THREADS=1 and with all threads and compares the two.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <string>
#include <vector>

#include "../../spades_b200/csrc/graph.h"
#include "../../spades_b200/csrc/host_par.h"

int main(int argc, char **argv) {
    const size_t E = argc > 1 ? (size_t)atoll(argv[1]) : 200000;
    const int K = 21;
    std::mt19937_64 rng(12345);
    sg::Graph g;
    g.k = K;
    const uint64_t V = E / 3 + 7;                       // vertex k-mer indices: several edges meet at a vertex
    for (size_t i = 0; i < E; ++i) {
        const uint32_t L = (uint32_t)(K + 1 + rng() % 60);
        g.edge_off.push_back(g.seq.size());
        g.edge_len.push_back(L);
        for (uint32_t p = 0; p < L; ++p) g.seq.push_back("ACGT"[rng() & 3]);
        g.raw_cov.push_back((uint32_t)(rng() % 5000));
        // LinkRecord::hash_and_mask_ = (vertex k-mer index << 2) | is_rc << 1 | is_start; ~0 = the placeholder of a self-conjugate edge
        g.link_start.push_back(((rng() % V) << 2) | ((rng() & 1) << 1) | 1u);
        g.link_end.push_back((rng() % 97 == 0) ? ~0ull : (((rng() % V) << 2) | ((rng() & 1) << 1) | 0u));
    }
    const std::string t = sg::graph_gfa(&g, "check");
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : t) { h ^= c; h *= 1099511628211ull; }
    size_t lines = 0;
    for (char c : t) lines += c == '\n';
    // the sort primitive on its own: duplicates, all sizes around the threading threshold
    int bad = 0;
    for (size_t n : {(size_t)0, (size_t)1, (size_t)40000, (size_t)1000003}) {
        sg::raw_vector<uint64_t> v(n);
        for (auto &x : v) x = rng() % (n / 5 + 1);
        std::vector<uint64_t> w(v.begin(), v.end());
        std::sort(w.begin(), w.end());
        sg::par_sort(v, [](uint64_t a, uint64_t b) { return a < b; });
        if (!std::equal(v.begin(), v.end(), w.begin())) ++bad;
    }
    printf("%016llx %zu %zu %d %d\n", (unsigned long long)h, t.size(), lines, sg::host_threads_for(2 * E), bad);
    return bad;
}
