// host_graph.cpp -- graph linking and GFA text on the host (string-heavy, a few million records).
//   FastGraphFromSequencesConstructor::ConstructGraph (src/common/assembly_graph/construction/debruijn_graph_constructor.hpp:506-567)
//   ids: GraphCore ID_BIAS=3 (assembly_graph/core/graph_core.hpp:233), edge i -> 3+2i, conjugate +1 (:514-531), vertex v -> 3+2v (:459-479)
//   PairedVertex::AddOutgoingEdge keeps outgoing edges sorted by id (graph_core.hpp:206-209)
//   GFAWriter::WriteSegments / WriteLinks / WriteVertexLinks (src/common/io/graph/gfa_writer.cpp:19-116)
// Inputs are the per-edge link records and raw coverages computed on the GPU (graph.cu).
#include <stdio.h>

#include <algorithm>
#include <string>
#include <vector>

#include "graph.h"

namespace sg {

namespace {
struct Rec {
    uint64_t hm;     // hash_and_mask_
    uint64_t edge;   // EdgeId
    uint64_t edge_and_mask() const { return (edge << 2) | (hm & 3); }
};
const uint64_t kMinId = 3;

void put_u(std::string &s, uint64_t v) {
    char tmp[24];
    int n = snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)v);
    s.append(tmp, n);
}
}  // namespace

std::string graph_gfa(const Graph *g, const char *version) {
    const size_t E = g->edge_len.size();
    const int K = g->k;
    std::vector<Rec> recs(2 * E);
    std::vector<uint8_t> selfc(E ? E : 1, 0);
    for (size_t i = 0; i < E; ++i) {
        const uint64_t e = kMinId + 2 * i;
        recs[2 * i] = Rec{g->link_start[i], e};
        if (g->link_end[i] == ~0ull) { selfc[i] = 1; recs[2 * i + 1] = Rec{~0ull, 0}; }      // LinkRecord() of a self-conjugate edge
        else recs[2 * i + 1] = Rec{g->link_end[i], e};
    }
    std::sort(recs.begin(), recs.end(), [](const Rec &a, const Rec &b) {                  // CompareByVertexKMerEdgeIdAndMask
        const uint64_t ha = a.hm >> 2, hb = b.hm >> 2;
        if (ha != hb) return ha < hb;
        return a.edge_and_mask() < b.edge_and_mask();
    });
    std::vector<size_t> groups;
    for (size_t i = 0; i < recs.size(); ++i) {
        if (i == 0 || (recs[i].hm >> 2) != (recs[i - 1].hm >> 2)) {
            const bool invalid = (recs[i].hm + 1 == 0) && recs[i].edge == 0;
            if (!invalid) groups.push_back(i);
        }
    }
    std::sort(groups.begin(), groups.end(), [&](size_t a, size_t b) { return recs[a].edge_and_mask() < recs[b].edge_and_mask(); });
    const size_t V = groups.size();
    // outgoing edge lists of vertex v (slot 2v) and of its conjugate (slot 2v+1)
    std::vector<std::vector<uint64_t>> out(2 * V);
    for (size_t vn = 0; vn < V; ++vn) {
        const size_t i = groups[vn];
        for (size_t j = i; j < recs.size() && (recs[j].hm >> 2) == (recs[i].hm >> 2); ++j) {
            const bool is_start = recs[j].hm & 1, is_rc = recs[j].hm & 2;
            const uint64_t e = recs[j].edge;
            const size_t ei = (size_t)((e - kMinId) / 2);
            const uint64_t ce = selfc[ei] ? e : e + 1;
            const int side = is_rc ? 1 : 0;                                     // LinkEdge: v1 = is_rc ? conjugate(v) : v
            if (is_start) out[2 * vn + side].push_back(e);                      // LinkOutgoingEdge(v1, e)
            else out[2 * vn + (side ^ 1)].push_back(ce);                        // LinkIncomingEdge(v1, e): cvertex(v1) gets conjugate(e)
        }
        std::sort(out[2 * vn].begin(), out[2 * vn].end());
        std::sort(out[2 * vn + 1].begin(), out[2 * vn + 1].end());
    }
    std::string t;
    t.reserve(g->seq.size() + 64 * E + 32 * 4 * V + 64);
    t += "H\tsp:Z:"; t += version; t += "\n";
    for (size_t i = 0; i < E; ++i) {
        t += "S\t"; put_u(t, kMinId + 2 * i); t += "\t";
        t.append(g->seq, g->edge_off[i], g->edge_len[i]);
        const uint32_t raw = g->raw_cov[i];
        const double c = (double)raw / (double)(g->edge_len[i] - K);            // CoverageIndex::coverage, core/coverage.hpp:59-61
        char tmp[64];
        int n = snprintf(tmp, sizeof tmp, "\tDP:f:%g\tKC:i:%u\n", (double)(float)c, raw);   // `os << float(cov)`, gfa_writer.cpp:24
        t.append(tmp, n);
    }
    for (size_t vn = 0; vn < V; ++vn) {
        for (uint64_t x : out[2 * vn + 1]) {                                     // IncomingEdges(v) = conjugates of OutgoingEdges(conj v)
            const size_t xi = (size_t)((x - kMinId) / 2);
            const uint64_t inc = selfc[xi] ? x : (((x - kMinId) & 1) ? x - 1 : x + 1);
            for (uint64_t oe : out[2 * vn]) {
                const uint64_t ends[2] = {inc, oe};
                t += "L";
                for (int q = 0; q < 2; ++q) {
                    const uint64_t e = ends[q];
                    const size_t ei = (size_t)((e - kMinId) / 2);
                    const bool canon = selfc[ei] || (((e - kMinId) & 1) == 0);
                    t += "\t"; put_u(t, kMinId + 2 * ei); t += canon ? "\t+" : "\t-";
                }
                char tmp[32];
                int n = snprintf(tmp, sizeof tmp, "\t%dM\n", K);
                t.append(tmp, n);
            }
        }
    }
    return t;
}

}  // namespace sg
