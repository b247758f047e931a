// oracle/ref_fastx.cpp -- TEST INFRASTRUCTURE ONLY. Ingest pin: the UNMODIFIED reference's own FASTA/FASTQ(.gz) parser
// (io::FastaFastqGzParser over its vendored kseq + zlib) followed by io::LongestValid -- what io::EasyStream does for the tools
// (io/reads/io_helper.cpp:21-35). A separate program because the reference cannot include this parser and ireadstream.hpp (which
// ref_probe needs) in one translation unit: both instantiate kseq at namespace scope.
//
// usage: ref_fastx <reads.fa|fq[.gz]> <out.txt>     one surviving read per line; last line "#records <n>"
#include "io/reads/fasta_fastq_gz_parser.hpp"
#include "io/reads/longest_valid_wrapper.hpp"

#include <fstream>

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s reads out.txt\n", argv[0]); return 2; }
    io::FastaFastqGzParser parser(argv[1]);
    std::ofstream os(argv[2]);
    io::SingleRead r;
    size_t records = 0;
    while (!parser.eof()) {
        parser >> r;
        ++records;
        io::LongestValid(r);
        if (r.size()) os << r.GetSequenceString() << "\n";
    }
    os << "#records " << records << "\n";
    return 0;
}
