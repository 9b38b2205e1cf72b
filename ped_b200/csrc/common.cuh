// Shared host/device helpers for the pedk k-mer hot path.
//
// Conventions:
//   * alphabet order A<C<G<T (ped.pl:91 sets LANG=C, so GNU sort is bytewise);
//     2-bit code A=0,C=1,G=2,T=3 keeps that order for fixed-length strings.
//   * a packed read is W = ceil(L/32) 64-bit words, base i in word i>>5 at bit
//     shift 62-2*(i&31): the first base is most significant, so comparing the
//     word tuple numerically == comparing the strings bytewise.  Tail bits are 0.
//   * a k-mer (k<=32) is a right-aligned 2k-bit integer, first base most
//     significant, so numeric order == string order of equal-length k-mers.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define PEDK_HD __host__ __device__ __forceinline__
#else
#define PEDK_HD static inline
#endif

#define PEDK_MAX_WORDS 16  // reads up to 512 bases

namespace pedk {

// ASCII -> 2-bit code; valid only for 'A','C','G','T'.
PEDK_HD uint32_t base_code(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }

PEDK_HD bool is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

PEDK_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

PEDK_HD uint64_t bitrev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    return (x >> 32) | (x << 32);
#endif
}

// Reverse the order of the 32 two-bit groups of x and complement each base.
// (ped.pl:3464-3484 `complement`: reverse, A<->T, C<->G; code c -> 3-c == c^3.)
PEDK_HD uint64_t revcomp_groups64(uint64_t x) {
    uint64_t r = bitrev64(~x);
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

// Reverse complement of a right-aligned k-mer (k<=32).
PEDK_HD uint64_t kmer_revcomp(uint64_t kmer, int k) {
    // complement also turns the unused high bits into ones; after the group
    // reversal they sit in the low bits and the shift drops them.
    return revcomp_groups64(kmer) >> (64 - 2 * k);
}

// 32 bases of a packed read starting at (possibly negative) base index s;
// positions outside [0, 32*W) read as 0.
PEDK_HD uint64_t read_window64(const uint64_t* words, int W, int s) {
    int w = s >> 5;  // arithmetic shift: floor division
    int o = s & 31;
    uint64_t hi = (w >= 0 && w < W) ? words[w] : 0ull;
    if (o == 0) return hi;
    uint64_t lo = (w + 1 >= 0 && w + 1 < W) ? words[w + 1] : 0ull;
    return (hi << (2 * o)) | (lo >> (64 - 2 * o));
}

// k-mer starting at base i (0 <= i <= L-k) of a packed read.
PEDK_HD uint64_t kmer_at(const uint64_t* words, int W, int i, int k) {
    return read_window64(words, W, i) >> (64 - 2 * k);
}

// Mask that keeps the valid bases of word w of an L-base read.
PEDK_HD uint64_t tail_mask(int L, int w) {
    int valid = L - 32 * w;
    if (valid >= 32) return ~0ull;
    if (valid <= 0) return 0ull;
    return ~0ull << (64 - 2 * valid);
}

// Reverse complement of a packed L-base read into out[0..W).
PEDK_HD void read_revcomp(const uint64_t* words, int W, int L, uint64_t* out) {
    for (int w = 0; w < W; ++w) {
        // output bases 32w..32w+31 come from source bases L-1-32w down to L-32-32w
        uint64_t x = read_window64(words, W, L - 32 * w - 32);
        out[w] = revcomp_groups64(x) & tail_mask(L, w);
    }
}

// <0, 0, >0 like memcmp on the unpacked strings.
PEDK_HD int words_cmp(const uint64_t* a, const uint64_t* b, int W) {
    for (int w = 0; w < W; ++w) {
        if (a[w] != b[w]) return a[w] < b[w] ? -1 : 1;
    }
    return 0;
}

PEDK_HD uint64_t words_hash(const uint64_t* a, int W) {
    uint64_t h = 0x243F6A8885A308D3ull;
    for (int w = 0; w < W; ++w) h = mix64(h ^ a[w]);
    return h;
}

// Bijection on 2k-bit integers that spreads canonical k-mers evenly over their high bits
// (canonical k-mers crowd the low prefixes, and genomes are not uniform): multiply by an odd
// constant, fold the high half onto the low half, multiply again, all modulo 2^(2k).  The
// counting path bins a k-mer by the TOP bits of kmer_hash -- level-A bin, level-B bin -- and
// keeps only the remaining low bits (32 of them at k = 20); kmer_unhash gives the k-mer back.
PEDK_HD uint64_t kmer_hash(uint64_t kmer, int nbits) {
    const uint64_t mask = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1);
    const int s = (nbits + 1) >> 1;
    uint64_t h = (kmer * 0x9E3779B97F4A7C15ull) & mask;
    h ^= h >> s;
    return (h * 0xD6E8FEB86659FD93ull) & mask;
}
PEDK_HD uint64_t kmer_unhash(uint64_t h, int nbits) {
    const uint64_t mask = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1);
    const int s = (nbits + 1) >> 1;
    h = (h * 0xCFEE444D8B59A89Bull) & mask;  // inverse of the second constant modulo 2^64
    h ^= h >> s;                             // 2s >= nbits: the fold is its own inverse
    return (h * 0xF1DE83E19937733Dull) & mask;
}

// Bits of the two binning levels for a given k: a + b <= 2k, at most 8 each.
PEDK_HD int kmer_bits_a(int k) { return 2 * k / 2 < 8 ? 2 * k / 2 : 8; }
PEDK_HD int kmer_bits_b(int k) { int r = 2 * k - kmer_bits_a(k); return r < 8 ? r : 8; }

// Owner rank of a canonical k-mer in a multi-GPU job: the ranks own contiguous ranges of the
// level-A bins (the top bits of kmer_hash), identical on host and device.
PEDK_HD int kmer_owner_rank(uint64_t canon, int k, int nranks) {
    const int a = kmer_bits_a(k);
    const uint64_t bin = kmer_hash(canon, 2 * k) >> (2 * k - a);
    return (int)((bin * (uint64_t)nranks) >> a);
}

// Owner rank of a packed canonical read.
PEDK_HD int read_owner_rank(const uint64_t* words, int W, int nranks) {
    return (int)(((words_hash(words, W) >> 32) * (uint64_t)nranks) >> 32);
}

}  // namespace pedk
