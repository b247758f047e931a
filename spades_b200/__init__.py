"""spades_b200 -- B200-native k-mer counting / de Bruijn graph construction path of SPAdes."""
from .packing import longest_valid, pack_reads, pack_fixed, synthetic_reads, unpack_kmers  # noqa: F401
