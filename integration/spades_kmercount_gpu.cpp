// integration/spades_kmercount_gpu.cpp -- spades-kmercount (projects/spades_tools/kmercount.cpp:191-230) with the counter
// swapped for the GPU adapter: REFERENCE host code (KMerDiskStorage, KMerIndexBuilder, KMerIndex, RtSeq, Sequence, fs::TmpDir,
// logger -- all unmodified, compiled where they lie) calling hand-written sm_100a CUDA through the C ABI.
//
//   spades_kmercount_gpu <reads.txt> <k> <workdir> [num_buckets=16]
//
// reads: FASTA / FASTQ, plain or gzip (parsed by the library's ingest with the original tool's semantics: kseq records +
// LongestValid), or one ACGT read per line (ref_probe's format). Output: <workdir>/final_kmers, byte-identical to the original
// tool's. It then proves the GPU-written storage is a drop-in for the rest of SPAdes:
//   1. the reference's OWN KMerIndexBuilder::BuildIndex runs over the GPU-written bucket files,
//   2. the GPU-built MPHF goes through the reference's OWN KMerIndex::deserialize,
//   3. both indices map every k-mer of final_kmers to the same slot (and that map is a bijection onto [0, n)).
// Exit code 0 only if all of that holds.
#include "gpu_kmer_counter.hpp"
#include <iterator>

#include "kmer_index/kmer_mph/kmer_index_traits.hpp"
#include "utils/logger/log_writers.hpp"
#include "utils/filesystem/temporary.hpp"

#include <fstream>
#include <iostream>
#include <vector>

static void create_console_logger() {
    using namespace logging;
    logger *lg = create_logger("");
    lg->add_writer(std::make_shared<console_writer>());
    attach_logger(lg);
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s reads.txt k workdir [num_buckets]\n", argv[0]); return 2; }
    const std::string reads_path = argv[1];
    const unsigned K = (unsigned)atoi(argv[2]);
    const std::filesystem::path workdir = argv[3];
    const unsigned B = argc > 4 ? (unsigned)atoi(argv[4]) : 16;      // kmercount.cpp:220
    create_console_logger();
    std::filesystem::create_directories(workdir);

    sgpu_config cfg = {0, 0, 0, 0};
    sgpu_ctx *ctx = nullptr;
    if (int rc = sgpu_create(&cfg, &ctx)) {
        fprintf(stderr, "spades_kmercount_gpu: cannot create a GPU context (error %d): there is no CPU fallback\n", rc);
        return 3;
    }
    INFO("K-mer length set to " << K);
    typedef kmers::KMerIndex<kmers::kmer_index_traits<RtSeq>> Index;
    int bad = 0;
    {
        kmers::GpuKMerDiskCounter counter(fs::tmp::make_temp_dir(workdir, "kmer_counter"), K, ctx, SGPU_ALL_WINDOWS);
        {
            std::ifstream is(reads_path, std::ios::binary);
            const int c0 = is.get(), c1 = is.get();
            is.seekg(0);
            if (c0 == '>' || c0 == '@' || (c0 == 0x1f && c1 == 0x8b)) {       // FASTA / FASTQ / gzip: what the original tool takes
                is.close();
                INFO("Parsed " << counter.AddFile(reads_path) << " reads from " << reads_path);
            } else {                                                            // one ACGT read per line (ref_probe's format)
                std::string line;
                while (std::getline(is, line))
                    if (!line.empty()) counter.AddRead(Sequence(line));
            }
        }
        auto storage = counter.Count(B, 1);                         // KMerDiskStorage<RtSeq>, buckets written from HBM
        const size_t total = storage.total_kmers();
        if (!storage.is_unique_and_sorted()) { ERROR("GPU-written buckets are not sorted/unique"); ++bad; }

        Index ref_index, gpu_index;
        kmers::KMerIndexBuilder<Index>(1).BuildIndex(ref_index, storage);         // 1. reference builder over GPU-written files
        kmers::BuildIndexOnGpu(gpu_index, ctx, counter.device_set());             // 2. GPU MPHF through the reference's deserialize
        if (ref_index.size() != total || gpu_index.size() != total) { ERROR("index sizes differ from the storage"); ++bad; }

        // 3. same slot for every k-mer, bijection
        std::vector<char> seen(total, 0);
        size_t checked = 0;
        for (unsigned b = 0; b < B && !bad; ++b) {
            for (auto it = storage.bucket_begin(b), e = storage.bucket_end(b); it != e; ++it) {
                const RtSeq kmer(K, (*it).first);
                const size_t a = ref_index.seq_idx(kmer), g = gpu_index.seq_idx(kmer);
                if (a != g || a >= total || seen[a]) { ERROR("index mismatch at k-mer " << kmer.str() << ": " << a << " vs " << g); ++bad; break; }
                seen[a] = 1; ++checked;
            }
        }
        if (checked != total) ++bad;
        INFO("Checked " << checked << " k-mers: reference-built and GPU-built KMerIndex agree");

        storage.merge();                                             // the reference's own merge (kmer_index_builder.hpp:190-203)
        auto final_kmers = storage.final_kmers();
        const std::filesystem::path out = workdir / "final_kmers";
        std::rename(final_kmers->file().c_str(), out.c_str());       // kmercount.cpp:222-223
        INFO("K-mer counting done, kmers saved to " << out);

        // 4. the splitter-level seam: the reference's OWN KMerDiskCounter over kmers::GpuKMerSplitter (one sorted-unique run + .idx per
        //    bucket) must arrive at the same final_kmers
        {
            kmers::KMerDiskCounter<RtSeq> ref_counter(workdir, kmers::GpuKMerSplitter(workdir, K, ctx, SGPU_ALL_WINDOWS));
            auto st2 = ref_counter.CountAll(B, 1, /* merge */ true);
            std::ifstream a(out, std::ios::binary), b2(st2.final_kmers()->file(), std::ios::binary);
            const std::string sa((std::istreambuf_iterator<char>(a)), std::istreambuf_iterator<char>());
            const std::string sb((std::istreambuf_iterator<char>(b2)), std::istreambuf_iterator<char>());
            if (st2.total_kmers() != total || sa != sb) { ERROR("KMerDiskCounter over GpuKMerSplitter differs from GpuKMerDiskCounter"); ++bad; }
            else INFO("reference KMerDiskCounter over GpuKMerSplitter: identical final_kmers (" << st2.total_kmers() << " k-mers)");
        }
    }
    sgpu_destroy(ctx);
    return bad ? 1 : 0;
}

namespace llvm {      // see oracle/ref_probe.cpp: the profiler hooks are no-ops unless the profiler was initialised
TimeTraceProfiler *getTimeTraceProfilerInstance() { return nullptr; }
void timeTraceProfilerBegin(StringRef, StringRef) {}
void timeTraceProfilerEnd() {}
}
