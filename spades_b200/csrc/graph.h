// graph.h -- condensed de Bruijn graph artefacts (device masks/coverage + host unitigs and link records)
#pragma once
#include <string>
#include <vector>

#include "sgpu_internal.h"

namespace sg {

struct Graph {
    Ctx *ctx = nullptr;
    int k = 0;
    const KSet *kp = nullptr, *km = nullptr;
    const Mphf *mk = nullptr, *mkp = nullptr;
    DArr<uint8_t> masks;         // working copy (mutated by RemoveSequences semantics), k-mer MPHF order
    DArr<uint8_t> masks_final;   // extension masks as the reference's DeBruijnExtensionIndex holds them
    DArr<uint32_t> cov;          // (k+1)-mer multiplicities, (k+1)-mer MPHF order (the reference's coverage_map)
    // edges in the reference's order: unbranching paths in final_kmers x out-edge order, then perfect loops
    std::string seq;             // ASCII, concatenated
    std::vector<uint64_t> edge_off;
    std::vector<uint32_t> edge_len;
    std::vector<uint64_t> link_start, link_end;   // LinkRecord::hash_and_mask_ (debruijn_graph_constructor.hpp:422-452)
    std::vector<uint32_t> raw_cov;                // CoverageIndex raw coverage per edge
    uint64_t tc_stats[3] = {0, 0, 0};             // early tip clipper: removed k-mers, tipped junctions, clipped links
};

Graph *graph_build(Ctx *ctx, const KSet *kp, const KSet *km, const Mphf *mk, const Mphf *mkp, bool keep_loops, uint64_t early_tc_bound = 0);
std::vector<uint64_t> graph_histogram(Ctx *ctx, const Graph *g);

// host_graph.cpp : FastGraphFromSequencesConstructor::ConstructGraph + GFAWriter
std::string graph_gfa(const Graph *g, const char *version);

}  // namespace sg
