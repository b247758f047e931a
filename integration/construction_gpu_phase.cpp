// integration/construction_gpu_phase.cpp -- the pipeline seam of INTEGRATION.md as code that COMPILES against the unmodified
// reference: Construction::Phase subclasses that put the hot path on the GPU inside the `Construction` stage of spades-core.
//
// Construction::Phase and ConstructionStorage are private to src/common/stages/construction.cpp (:25-40, :215-257), so a SPAdes
// maintainer adds these classes to that file and swaps them in inside Construction::Construction() (:439-453):
//         add<CoverageFilterGpu>();          // instead of add<CoverageFilter>()   (only when read_cov_threshold > 0, :446)
//         add<KMerCountingGpu>();            // instead of add<KMerCounting>()
//         add<ExtensionIndexBuilder>();       // unchanged: consumes storage().kmers (bucket files written from HBM)
// To keep this file honest without patching the reference, it textually includes the UNMODIFIED translation unit and is compiled
// (object only; linking spades-core needs the rest of the assembler) by `make -C integration phase`.
#include "stages/construction.cpp"

#include "gpu_kmer_counter.hpp"
#include "io/reads/vector_reader.hpp"
#include "io/reads/rc_reader_wrapper.hpp"

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

// a read of the GPU's packed read set as something Sequence's generic constructor accepts (sequence.hpp:380-420: operator[] returning
// 0..3 selects its "digit string" branch)
struct PackedReadView {
    const uint64_t *w;
    size_t n;
    size_t size() const { return n; }
    char operator[](size_t i) const { return (char)((w[i >> 5] >> ((i & 31) << 1)) & 3); }
};

// replaces CoverageFilter (construction.cpp:167-198: EstimateCardinalityUpperBound -> qf::cqf -> FillCoverageHistogram ->
// io::CovFilteringWrap): one library call computes the same verdict for every read; the survivors replace the stage's read streams
// as in-memory streams, RC-wrapped like the ones Construction::init installs, so every later phase (GPU or CPU) sees exactly the
// reads the CPU wrapper would have let through.
class CoverageFilterGpu : public Construction::Phase {
public:
    CoverageFilterGpu() : Construction::Phase("k-mer multiplicity estimation (GPU)", "cqf_filter") { }
    virtual ~CoverageFilterGpu() = default;

    void run(graph_pack::GraphPack &, const char*) override {
        auto &read_streams = storage().read_streams;
        const auto &index = storage().ext_index;
        VERIFY_MSG(read_streams.size(), "No input streams specified");
        const unsigned rthr = storage().params.read_cov_threshold;
        const unsigned kplusone = index.k() + 1;
        sgpu_ctx *ctx = StageGpuContext();

        // hand the FORWARD reads over: the stage's streams yield every read followed by its reverse complement
        // (io/reads/rc_reader_wrapper.hpp:34-43); the filter's hash is symmetric and its counts are per genomic window
        kmers::GpuKMerDiskCounter feeder(storage().workdir, kplusone, ctx, SGPU_CANONICAL);
        const size_t nstreams = read_streams.size();
        read_streams.reset();
        for (size_t i = 0; i < nstreams; ++i) {
            io::SingleReadSeq r;
            for (size_t pos = 0; !read_streams[i].eof(); ++pos) {
                read_streams[i] >> r;
                if ((pos & 1) == 0) feeder.AddRead(r.sequence());
            }
        }
        feeder.Flush();
        uint64_t st[4] = {0, 0, 0, 0};
        if (sgpu_reads_cov_filter(ctx, (int)kplusone, rthr, /* apply */ 1, nullptr, st)) FATAL_ERROR("spades_b200: " << sgpu_last_error(ctx));
        INFO("Estimated " << st[0] << " distinct kmers");
        INFO("Counting threshold " << rthr << ": " << st[3] << " reads pass the coverage filter");

        // survivors -> in-memory streams, as many as before
        int64_t n = 0;
        uint64_t nw = 0;
        if (sgpu_reads_info(ctx, &n, &nw)) FATAL_ERROR("spades_b200: " << sgpu_last_error(ctx));
        std::vector<uint64_t> words(nw + 1), offs((size_t)n + 1);
        std::vector<uint32_t> lens((size_t)n + 1);
        if (sgpu_reads_download(ctx, words.data(), offs.data(), lens.data())) FATAL_ERROR("spades_b200: " << sgpu_last_error(ctx));
        io::ReadStreamList<io::SingleReadSeq> filtered;
        for (size_t i = 0; i < nstreams; ++i) {
            std::vector<io::SingleReadSeq> chunk;
            for (size_t r = (size_t)n * i / nstreams; r < (size_t)n * (i + 1) / nstreams; ++r)
                chunk.emplace_back(Sequence(PackedReadView{words.data() + offs[r], lens[r]}));
            filtered.push_back(io::RCWrap<io::SingleReadSeq>(io::VectorReadStream<io::SingleReadSeq>(chunk)));
        }
        storage().read_streams = std::move(filtered);
    }

    void load(graph_pack::GraphPack&, const std::filesystem::path &, const char*) override { VERIFY_MSG(false, "implement me"); }
    void save(const graph_pack::GraphPack&, const std::filesystem::path &, const char*) const override { }
};

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
void RegisterGpuPhases(Construction &stage, bool with_coverage_filter) {
    if (with_coverage_filter) stage.add<CoverageFilterGpu>();
    stage.add<KMerCountingGpu>();
}

}  // namespace debruijn_graph
