// integration/construction_gpu_phase.cpp -- the pipeline seam of INTEGRATION.md as code that COMPILES against the unmodified
// reference: Construction::Phase subclasses that put the hot path on the GPU inside the `Construction` stage of spades-core.
//
// Construction::Phase and ConstructionStorage are private to src/common/stages/construction.cpp (:25-40, :215-257), so a SPAdes
// maintainer adds these classes to that file and swaps them in inside Construction::Construction() (:439-453):
//         add<KMerCountingGpu>();            // instead of add<KMerCounting>()
//         add<ExtensionIndexBuilder>();       // unchanged: consumes storage().kmers (bucket files written from HBM)
// To keep this file honest without patching the reference, it textually includes the UNMODIFIED translation unit and is compiled
// (object only; linking spades-core needs the rest of the assembler) by `make -C integration phase`.
#include "stages/construction.cpp"

#include "gpu_kmer_counter.hpp"

namespace debruijn_graph {
namespace {

// process-wide GPU context of the stage (one GPU per spades-core process; created on first use, never a CPU fallback)
inline sgpu_ctx *StageGpuContext() {
    static sgpu_ctx *ctx = [] {
        sgpu_config cfg = {0, 0, 0, 0};
        sgpu_ctx *c = nullptr;
        if (int rc = sgpu_create(&cfg, &c)) FATAL_ERROR("spades_b200: cannot create a GPU context (error " << rc << "); there is no CPU fallback");
        return c;
    }();
    return ctx;
}

// replaces KMerCounting (construction.cpp:215-257): (k+1)-mer counting of reads (+ trusted contigs) on the GPU; leaves
// storage().kmers exactly as the CPU phase does (KMerDiskStorage<RtSeq> with 10 x nthreads buckets, construction.cpp:242)
class KMerCountingGpu : public Construction::Phase {
public:
    KMerCountingGpu() : Construction::Phase("k+1-mer counting (GPU)", "kpomer_counting") { }
    virtual ~KMerCountingGpu() = default;

    void run(graph_pack::GraphPack &, const char*) override {
        auto &read_streams = storage().read_streams;
        auto &contigs_streams = storage().contigs_streams;
        const auto &index = storage().ext_index;
        VERIFY_MSG(read_streams.size(), "No input streams specified");
        io::ReadStreamList<io::SingleReadSeq> merge_streams = temp_merge_read_streams(read_streams, contigs_streams);
        const unsigned nthreads = (unsigned)merge_streams.size();

        // The streams of Construction::init are RC-wrapped (read, then its reverse complement). The GPU splitter canonicalises every
        // window itself, so the RC copies only repeat work: the (k+1)-mer SET is the same either way, and the multiplicities are
        // not taken from this phase (PHMCoverageFiller recounts through the MPHF, construction.cpp:371-420).
        kmers::GpuKMerDiskCounter counter(storage().workdir, index.k() + 1, StageGpuContext(), SGPU_CANONICAL);
        merge_streams.reset();
        for (size_t i = 0; i < merge_streams.size(); ++i) counter.AddStream(merge_streams[i]);
        auto kmers = counter.Count(10 * nthreads, nthreads);
        storage().kmers.reset(new kmers::KMerDiskStorage<RtSeq>(std::move(kmers)));
    }

    void load(graph_pack::GraphPack&, const std::filesystem::path &, const char*) override { VERIFY_MSG(false, "implement me"); }
    void save(const graph_pack::GraphPack&, const std::filesystem::path &, const char*) const override { }
};

}  // namespace

// the registration a maintainer writes in Construction::Construction(); an external function here so that the phase's code is emitted
void RegisterGpuPhases(Construction &stage) { stage.add<KMerCountingGpu>(); }

}  // namespace debruijn_graph
