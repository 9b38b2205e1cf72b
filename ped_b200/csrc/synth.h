// Deterministic synthetic FASTQ generator shared by the host and the device
// (SURVEY.md section 8d "Synthetic inputs").  Counter-based: every byte of
// read i depends only on (seed, i), so the host loop and the CUDA kernel
// produce identical files and any slice can be generated independently.
#pragma once
#include "common.cuh"

namespace pedk {

struct SynthParams {
    uint64_t seed;        // per-sample seed (3000+cfg control, 4000+cfg target)
    uint64_t genome_len;  // G
    uint32_t read_len;    // L
    uint32_t err_ppm;     // substitution errors per million bases (0.2 % -> 2000)
    uint32_t n_ppm;       // reads that get one 'N', per million reads (0.1 % -> 1000)
    uint32_t dup_ppm;     // reads that repeat the start/strand of read i-1 (exercises dedup)
};

PEDK_HD uint64_t synth_rnd(uint64_t seed, uint64_t idx, uint64_t stream) {
    return mix64(mix64(seed * 0x9E3779B97F4A7C15ull + stream) ^ (idx * 0xD1342543DE82EF95ull + stream));
}

PEDK_HD char synth_genome_base(uint64_t seed, uint64_t i) {
    const char acgt[4] = {'A', 'C', 'G', 'T'};
    return acgt[synth_rnd(seed, i, 7) & 3];
}

PEDK_HD uint32_t synth_digits(uint64_t i) {
    uint32_t d = 1;
    while (i >= 10) { i /= 10; ++d; }
    return d;
}

// Byte offset of record i: sum_{j<i} (2L + 7 + digits(j)).
// Record i is "@r<i>\n" SEQ "\n+\n" QUAL "\n".
PEDK_HD uint64_t synth_record_offset(uint64_t i, uint32_t L) {
    uint64_t off = i * (2ull * L + 7ull);
    // sum of digits(j) for j in [0, i)
    uint64_t lo = 0, hi = 10, d = 1;
    while (lo < i) {
        uint64_t top = hi < i ? hi : i;
        off += (top - lo) * d;
        lo = hi;
        hi = hi * 10;
        ++d;
    }
    return off;
}

PEDK_HD uint64_t synth_total_bytes(uint64_t n_reads, uint32_t L) { return synth_record_offset(n_reads, L); }

// (start, strand) of read i; a "dup" read copies the placement of read i-1
// but draws its own errors, so only error-free copies are exact duplicates.
PEDK_HD void synth_placement(const SynthParams& p, uint64_t i, uint64_t* start, uint32_t* strand) {
    uint64_t j = i;
    // at most a short chain: follow dup links backwards
    while (j > 0 && p.dup_ppm && (synth_rnd(p.seed, j, 5) % 1000000ull) < p.dup_ppm) --j;
    uint64_t r = synth_rnd(p.seed, j, 0);
    *start = r % (p.genome_len - p.read_len + 1);
    *strand = (uint32_t)(synth_rnd(p.seed, j, 1) & 1);
}

// Base j of read i, given the genome and placement.
PEDK_HD char synth_read_base(const SynthParams& p, const char* genome, uint64_t i, uint32_t j,
                             uint64_t start, uint32_t strand, int64_t n_pos) {
    if ((int64_t)j == n_pos) return 'N';
    char c;
    if (!strand) {
        c = genome[start + j];
    } else {
        char g = genome[start + p.read_len - 1 - j];
        c = g == 'A' ? 'T' : g == 'C' ? 'G' : g == 'G' ? 'C' : 'A';
    }
    uint64_t e = synth_rnd(p.seed, i * (uint64_t)p.read_len + j, 2);
    if ((e % 1000000ull) < p.err_ppm) {
        const char acgt[4] = {'A', 'C', 'G', 'T'};
        c = acgt[(e >> 40) & 3];  // may equal the original base
    }
    return c;
}

PEDK_HD int64_t synth_n_pos(const SynthParams& p, uint64_t i) {
    uint64_t r = synth_rnd(p.seed, i, 3);
    if ((r % 1000000ull) < p.n_ppm) return (int64_t)((r >> 32) % p.read_len);
    return -1;
}

// Write record i at out (which must have room for 2L+7+digits(i) bytes).
PEDK_HD void synth_write_record(const SynthParams& p, const char* genome, uint64_t i, char* out) {
    uint32_t L = p.read_len;
    uint32_t d = synth_digits(i);
    out[0] = '@';
    out[1] = 'r';
    uint64_t v = i;
    for (uint32_t t = 0; t < d; ++t) { out[1 + d - t] = (char)('0' + (v % 10)); v /= 10; }
    out[2 + d] = '\n';
    char* seq = out + 3 + d;
    uint64_t start; uint32_t strand;
    synth_placement(p, i, &start, &strand);
    int64_t n_pos = synth_n_pos(p, i);
    for (uint32_t j = 0; j < L; ++j) seq[j] = synth_read_base(p, genome, i, j, start, strand, n_pos);
    seq[L] = '\n';
    seq[L + 1] = '+';
    seq[L + 2] = '\n';
    char* q = seq + L + 3;
    for (uint32_t j = 0; j < L; ++j) q[j] = 'I';
    q[L] = '\n';
}

}  // namespace pedk
