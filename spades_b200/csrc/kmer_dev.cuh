// kmer_dev.cuh -- host/device arithmetic shared by every kernel of the path.
//
// 2-bit k-mer words, reverse complement, minimality, XXH3-64/128 short paths, bucket function,
// MSD digit extraction. Everything is __host__ __device__ so the same code is unit-tested on the
// CPU (tests/test_hostdev_helpers.py through sgpu_selftest_*) before it ever runs on the GPU.
//
// Reference semantics (paths relative to the SPAdes tree):
//   packing           src/common/sequence/rtseq.hpp:379-382 (nucleotide i at bits 2(i%32) of word i/32)
//   FastRC            src/common/sequence/rtseq.hpp:81-117
//   IsMinimal         src/common/sequence/rtseq.hpp:409-417
//   GetHash           src/common/sequence/rtseq.hpp:690-696 -> XXH3_64bits_withSeed(words, 8*nw, 0)
//   bucket            src/common/kmer_index/kmer_mph/kmer_buckets.hpp:32-34,47-52 ; adt/lemiere_mod_reduce.hpp:18-21
//   XXH3 64           ext/include/xxh/xxhash.h:4537-4567,4606-4675 ; 128: :6449-6625 ; secret :4239
//   sort order        ext/include/pdqsort/pdqsort_pod.h:725-734 (word 0 most significant)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD inline
#endif

namespace sg {

template <int NW>
struct Kmer {
    uint64_t w[NW];
};

SG_HD constexpr int nwords_of(int K) { return (K + 31) >> 5; }

// ---- bit tricks ---------------------------------------------------------------------------------
SG_HD uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
SG_HD uint64_t bswap64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    lo = __byte_perm(lo, 0, 0x0123);
    hi = __byte_perm(hi, 0, 0x0123);
    return ((uint64_t)lo << 32) | hi;
#else
    return __builtin_bswap64(x);
#endif
}
SG_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
SG_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
SG_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// reverse the order of the 32 nucleotides of a word and complement them (A<->T, C<->G == 3-c == ~c)
SG_HD uint64_t rc_word(uint64_t x) {
    uint64_t r = brev64(~x);
    return ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
}
SG_HD uint64_t shr_safe(uint64_t x, int s) { return s >= 64 ? 0 : (x >> s); }
SG_HD uint64_t shl_safe(uint64_t x, int s) { return s >= 64 ? 0 : (x << s); }

template <int NW>
SG_HD uint64_t last_word_mask(int K) {
    int bits = 2 * K - 64 * (NW - 1);            // 2..64
    return bits >= 64 ? ~0ULL : ((1ULL << bits) - 1);
}

// reverse complement of a K-mer held in NW words
template <int NW>
SG_HD Kmer<NW> kmer_rc(const Kmer<NW> &a, int K) {
    Kmer<NW> t, r;
#pragma unroll
    for (int j = 0; j < NW; ++j) t.w[j] = rc_word(a.w[NW - 1 - j]);
    const int sh = 64 * NW - 2 * K;               // 0..62
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        uint64_t lo = t.w[j] >> sh;
        uint64_t hi = (j + 1 < NW && sh) ? (t.w[j + 1] << (64 - sh)) : 0;
        r.w[j] = lo | hi;
    }
    return r;
}

// nucleotide-lexicographic "a < b" from position 0 (RtSeq operator<, rtseq.hpp:740-748); ties -> false
template <int NW>
SG_HD bool kmer_nuc_less(const Kmer<NW> &a, const Kmer<NW> &b) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        uint64_t d = a.w[j] ^ b.w[j];
        if (d) {
            int p = ctz64(d) & ~1;
            return ((a.w[j] >> p) & 3) < ((b.w[j] >> p) & 3);
        }
    }
    return false;
}
template <int NW>
SG_HD bool kmer_eq(const Kmer<NW> &a, const Kmer<NW> &b) {
    bool e = true;
#pragma unroll
    for (int j = 0; j < NW; ++j) e = e && (a.w[j] == b.w[j]);
    return e;
}
// word-lexicographic compare (bucket sort order): <0, 0, >0
template <int NW>
SG_HD int kmer_word_cmp(const Kmer<NW> &a, const Kmer<NW> &b) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        if (a.w[j] != b.w[j]) return a.w[j] < b.w[j] ? -1 : 1;
    }
    return 0;
}
// IsMinimal: fwd <= rc in nucleotide order (self-RC counts as minimal). r MUST be the reverse complement of f. Then the
// nucleotide order from position 0 equals the INTEGER order of the packed words (word NW-1 most significant): the integer
// compare decides at the highest digit i with f[i] != r[i] = 3 - f[K-1-i], the nucleotide compare at the lowest digit j with
// f[j] != r[j] = 3 - f[K-1-j]; these are the same pair of positions (j = K-1-i) and both say "f first" iff f[i] + f[j] < 3.
// A multiword unsigned compare is 2 instructions per word; kmer_nuc_less (first differing digit via ctz) is ~45 and branches.
template <int NW>
SG_HD bool kmer_is_minimal(const Kmer<NW> &f, const Kmer<NW> &r) {
    bool le = true;
#pragma unroll
    for (int j = 0; j < NW; ++j) le = (f.w[j] < r.w[j]) || (f.w[j] == r.w[j] && le);
    return le;
}

// K-mer window starting at base `pos` of a packed sequence
template <int NW, typename Ptr>
SG_HD Kmer<NW> kmer_window(Ptr seq, int64_t pos, int K) {
    Kmer<NW> k;
    const int64_t wi = pos >> 5;
    const int64_t lastw = (pos + K - 1) >> 5;    // never read past the last word the window touches
    const int s = (int)(pos & 31) << 1;
    uint64_t cur = seq[wi];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        uint64_t nxt = (wi + j + 1 <= lastw) ? seq[wi + j + 1] : 0;
        k.w[j] = (cur >> s) | ((nxt << 1) << (63 - s));
        cur = nxt;
    }
    k.w[NW - 1] &= last_word_mask<NW>(K);
    return k;
}
// append nucleotide c at the end, dropping the first (RtSeq::operator<<=, rtseq.hpp:459-476)
template <int NW>
SG_HD void kmer_shl(Kmer<NW> &k, int K, int c) {
#pragma unroll
    for (int j = 0; j < NW - 1; ++j) k.w[j] = (k.w[j] >> 2) | ((k.w[j + 1] & 3) << 62);
    const int sh = ((K + 31) & 31) << 1;
    k.w[NW - 1] = (k.w[NW - 1] >> 2) | ((uint64_t)c << sh);
}
template <int NW>
SG_HD int kmer_nuc(const Kmer<NW> &k, int i) { return (int)((k.w[i >> 5] >> ((i & 31) << 1)) & 3); }

// ---- rolling window ---------------------------------------------------------------------------------
// One thread walks consecutive windows of one read keeping the window AND its reverse complement: appending base c is
// RtSeq::operator<<= on the forward strand (rtseq.hpp:459-476) and "complement enters at position 0, everything moves up
// one nucleotide, the last one drops out" on the reverse strand (what RtSeq::operator>> of !kmer does, rtseq.hpp:569-588).
// ~20 integer instructions per window for K <= 64 instead of re-extracting the window and re-running FastRC.
template <int NW>
struct RollState {
    Kmer<NW> f, r;      // window, reverse complement of the window
    uint64_t cw;        // the bases that follow the window, next one in bits 0..1 (up to 32 of them: a whole chunk's worth)
};
// window at base j0 of a chunk of `cnt` windows (1 <= cnt <= 33): the cnt-1 bases the walk will append are fetched here, from the
// one or two packed words that hold them -- the walk itself touches no memory (the first version reloaded a word every 32 bases
// behind a per-window test; ncu charged that load's latency to nearly every step of a warp, its lanes being in different phases)
template <int NW, typename Ptr>
SG_HD void roll_init(RollState<NW> &st, Ptr seq, int j0, int K, int cnt) {
    st.f = kmer_window<NW>(seq, (int64_t)j0, K);
    st.r = kmer_rc<NW>(st.f, K);
    st.cw = 0;
    if (cnt > 1) {
        const int p = j0 + K, last = p + cnt - 2;            // first and last base appended
        const int sh = (p & 31) << 1;
        st.cw = seq[p >> 5] >> sh;
        if ((last >> 5) != (p >> 5)) st.cw |= (seq[(p >> 5) + 1] << 1) << (63 - sh);
    }
}
// advance to the next window; only legal cnt-1 times after roll_init(.., cnt)
template <int NW>
SG_HD void roll_next(RollState<NW> &st, int K) {
    const uint64_t c = st.cw & 3;
    st.cw >>= 2;
    kmer_shl<NW>(st.f, K, (int)c);
#pragma unroll
    for (int j = NW - 1; j >= 1; --j) st.r.w[j] = (st.r.w[j] << 2) | (st.r.w[j - 1] >> 62);
    st.r.w[0] = (st.r.w[0] << 2) | (3 - c);
    st.r.w[NW - 1] &= last_word_mask<NW>(K);
}

// prefix (drop last nucleotide) / suffix (drop first) of a (K+1)-mer as K-mers. NWS = words of the source.
template <int NW, int NWS>
SG_HD Kmer<NW> kmer_prefix(const Kmer<NWS> &x, int K) {
    Kmer<NW> k;
#pragma unroll
    for (int j = 0; j < NW; ++j) k.w[j] = x.w[j];
    k.w[NW - 1] &= last_word_mask<NW>(K);
    return k;
}
template <int NW, int NWS>
SG_HD Kmer<NW> kmer_suffix(const Kmer<NWS> &x, int K) {
    Kmer<NW> k;
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        uint64_t hi = (j + 1 < NWS) ? (x.w[j + 1] << 62) : 0;
        k.w[j] = (x.w[j] >> 2) | hi;
    }
    k.w[NW - 1] &= last_word_mask<NW>(K);
    return k;
}

// ---- XXH3 (xxHash 0.8.2) short paths, seed 0 -----------------------------------------------------
// readLE64(kSecret + off) for the offsets the 8..32-byte paths touch (xxhash.h:4239-4252)
#define SG_SEC0   0xbe4ba423396cfeb8ULL
#define SG_SEC8   0x1cad21f72c81017cULL
#define SG_SEC16  0xdb979083e96dd4deULL
#define SG_SEC24  0x1f67b3b7a4a44072ULL
#define SG_SEC32  0x78e5c0cc4ee679cbULL
#define SG_SEC40  0x2172ffcc7dd05a82ULL
#define SG_SEC48  0x8e2443f7744608b8ULL
#define SG_SEC56  0x4c263a81e69035e0ULL
#define SG_P64_1  0x9E3779B185EBCA87ULL
#define SG_P64_2  0xC2B2AE3D27D4EB4FULL
#define SG_P64_4  0x85EBCA77C2B2AE63ULL
#define SG_P32_2  0x85EBCA77ULL
#define SG_PMX1   0x165667919E3779F9ULL
#define SG_PMX2   0x9FB21C651E98DF25ULL

SG_HD uint64_t xxh_fold(uint64_t a, uint64_t b) { return (a * b) ^ mulhi64(a, b); }
SG_HD uint64_t xxh_aval(uint64_t h) { h ^= h >> 37; h *= SG_PMX1; return h ^ (h >> 32); }
SG_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

template <int NW>
SG_HD uint64_t xxh3_64(const Kmer<NW> &k) {
    if (NW == 1) {                                           // XXH3_len_4to8_64b
        uint64_t in64 = (k.w[0] >> 32) | (k.w[0] << 32);
        uint64_t h = in64 ^ (SG_SEC8 ^ SG_SEC16);
        h ^= rotl64(h, 49) ^ rotl64(h, 24);
        h *= SG_PMX2;
        h ^= (h >> 35) + 8;
        h *= SG_PMX2;
        return h ^ (h >> 28);
    } else if (NW == 2) {                                    // XXH3_len_9to16_64b
        uint64_t lo = k.w[0] ^ (SG_SEC24 ^ SG_SEC32);
        uint64_t hi = k.w[1] ^ (SG_SEC40 ^ SG_SEC48);
        uint64_t acc = 16 + bswap64(lo) + hi + xxh_fold(lo, hi);
        return xxh_aval(acc);
    } else {                                                 // XXH3_len_17to128_64b, len 24 / 32
        uint64_t acc = (uint64_t)(8 * NW) * SG_P64_1;
        acc += xxh_fold(k.w[0] ^ SG_SEC0, k.w[1] ^ SG_SEC8);
        acc += xxh_fold(k.w[NW - 2] ^ SG_SEC16, k.w[NW - 1] ^ SG_SEC24);
        return xxh_aval(acc);
    }
}

struct Hash128 { uint64_t lo, hi; };

template <int NW>
SG_HD Hash128 xxh3_128(const Kmer<NW> &k) {
    Hash128 r;
    if (NW == 1) {                                           // XXH3_len_4to8_128b
        uint64_t keyed = k.w[0] ^ (SG_SEC16 ^ SG_SEC24);
        uint64_t mul = SG_P64_1 + (8ULL << 2);
        uint64_t lo = keyed * mul, hi = mulhi64(keyed, mul);
        hi += (lo << 1);
        lo ^= (hi >> 3);
        lo ^= lo >> 35; lo *= SG_PMX2; lo ^= lo >> 28;
        r.lo = lo; r.hi = xxh_aval(hi);
    } else if (NW == 2) {                                    // XXH3_len_9to16_128b
        uint64_t bitflipl = SG_SEC32 ^ SG_SEC40, bitfliph = SG_SEC48 ^ SG_SEC56;
        uint64_t ilo = k.w[0], ihi = k.w[1];
        uint64_t x = ilo ^ ihi ^ bitflipl;
        uint64_t mlo = x * SG_P64_1, mhi = mulhi64(x, SG_P64_1);
        mlo += (uint64_t)(16 - 1) << 54;
        ihi ^= bitfliph;
        mhi += ihi + (uint64_t)(uint32_t)ihi * (SG_P32_2 - 1);
        mlo ^= bswap64(mhi);
        uint64_t hlo = mlo * SG_P64_2, hhi = mulhi64(mlo, SG_P64_2);
        hhi += mhi * SG_P64_2;
        r.lo = xxh_aval(hlo); r.hi = xxh_aval(hhi);
    } else {                                                 // XXH3_len_17to128_128b, len 24 / 32
        uint64_t alo = (uint64_t)(8 * NW) * SG_P64_1, ahi = 0;
        alo += xxh_fold(k.w[0] ^ SG_SEC0, k.w[1] ^ SG_SEC8);
        alo ^= k.w[NW - 2] + k.w[NW - 1];
        ahi += xxh_fold(k.w[NW - 2] ^ SG_SEC16, k.w[NW - 1] ^ SG_SEC24);
        ahi ^= k.w[0] + k.w[1];
        uint64_t hlo = alo + ahi;
        uint64_t hhi = alo * SG_P64_1 + ahi * SG_P64_4 + (uint64_t)(8 * NW) * SG_P64_2;
        r.lo = xxh_aval(hlo);
        r.hi = (uint64_t)0 - xxh_aval(hhi);
    }
    return r;
}

template <int NW>
SG_HD uint32_t kmer_bucket(const Kmer<NW> &k, uint32_t B) {
    if (B == 1) return 0;
    // mulhi64(hash, B) with a 32-bit B: two 32x32->64 products instead of a full 64x64->128
    const uint64_t h = xxh3_64<NW>(k);
    const uint64_t t = (uint64_t)(uint32_t)h * B;
    return (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * B + (t >> 32)) >> 32);
}

// ---- boomphf level hashes (BooPHF.h:606-613, :94-100): s0 = high64, s1 = low64 -------------------
struct LevelHasher {
    uint64_t s0, s1;
    int level;
    SG_HD LevelHasher(const Hash128 &h) : s0(h.hi), s1(h.lo), level(0) {}
    // returns the hash of the current level and advances
    SG_HD uint64_t next() {
        uint64_t r;
        if (level == 0) r = s0;
        else if (level == 1) r = s1;
        else {
            uint64_t a = s0; const uint64_t b = s1;
            s0 = b;
            a ^= a << 23;
            s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
            r = s1 + b;
        }
        ++level;
        return r;
    }
};

// ---- MSD digit extraction over the sort key -------------------------------------------------------
// The bucket order is word-lexicographic with word 0 most significant. The key bit string T is
// w0[63..0] w1[63..0] ... with the last word contributing only its 2K-64(NW-1) valid low bits.
// get_bits(k, pos, r) returns T[pos .. pos+r) as an integer (bits past the end read as 0). r <= 32.
template <int NW>
SG_HD uint32_t key_bits_generic(const Kmer<NW> &k, int K, int pos, int r);
template <int NW>
SG_HD uint32_t key_bits(const Kmer<NW> &k, int K, int pos, int r) {
    if (NW == 2) {
        // two words: left-align the second one, take 64 bits of the 128-bit string at `pos`, keep the top r. Branch-free apart from the
        // (segment-uniform) pos < 64 test; the generic loop below costs ~35 instructions with data-dependent branches.
        if (r <= 0) return 0u;
        const int lastbits = 2 * K - 64;                                   // 2..64
        const uint64_t w1l = k.w[1] << (64 - lastbits);
        uint64_t x;
        if (pos < 64) x = (k.w[0] << pos) | ((w1l >> 1) >> (63 - pos));      // pos == 0: the second term is w1l >> 64 == 0
        else x = pos < 128 ? (w1l << (pos - 64)) : 0;
        return (uint32_t)(x >> (64 - r));
    }
    return key_bits_generic<NW>(k, K, pos, r);
}
template <int NW>
SG_HD uint32_t key_bits_generic(const Kmer<NW> &k, int K, int pos, int r) {
    const int lastbits = 2 * K - 64 * (NW - 1);
    uint64_t acc = 0;
    int got = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const int wb = (j == NW - 1) ? lastbits : 64;         // valid bits of this word
        const int base = 64 * j;                              // T offset of this word's first bit
        // overlap of [pos+got, pos+r) with [base, base+wb)
        int lo = pos + got - base;
        if (lo >= 0 && lo < wb && got < r) {
            int take = r - got;
            if (take > wb - lo) take = wb - lo;
            // bits [lo, lo+take) counted from the word's top valid bit
            uint64_t v = (k.w[j] >> (wb - lo - take)) & ((take >= 64) ? ~0ULL : ((1ULL << take) - 1));
            acc = (acc << take) | v;
            got += take;
        }
    }
    return (uint32_t)(acc << (r - got));
}
// key_bits(k, K, 0, r), r <= 32: the partition digit of level A. With two or more words the first word is full: one shift.
template <int NW>
SG_HD uint32_t key_top_bits(const Kmer<NW> &k, int K, int r) {
    if (NW >= 2) return r ? (uint32_t)(k.w[0] >> (64 - r)) : 0u;
    return key_bits<NW>(k, K, 0, r);
}
SG_HD int key_total_bits(int K) { return 2 * K; }

}  // namespace sg
