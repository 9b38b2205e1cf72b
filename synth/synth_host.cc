// Host-side synthetic genomes and FASTQ reads (tests, scripts and bench.py only).
// Plain C++, no CUDA: built into synth/libpedk_synth.so, outside the product package, so that the CPU arms of bench.py and the golden
// scripts can make their inputs without mapping the product library.  The device generator in
// synth.cu writes the same bytes (both include synth.h).
#include <string.h>

#include "../include/pedk_synth.h"
#include "../ped_b200/csrc/synth.h"

using namespace pedk;

#define PEDK_OK 0
#define PEDK_E_INVAL (-1)
#define PEDK_E_NOMEM (-3)

static SynthParams to_params(const pedk_synth_params* p) {
    SynthParams s;
    s.seed = p->seed;
    s.genome_len = p->genome_len;
    s.read_len = p->read_len;
    s.err_ppm = p->err_ppm;
    s.n_ppm = p->n_ppm;
    s.dup_ppm = p->dup_ppm;
    return s;
}

extern "C" int pedk_synth_genome(uint64_t seed, uint64_t genome_len, char* out) {
    if (!out) return PEDK_E_INVAL;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)genome_len; ++i) out[i] = synth_genome_base(seed, (uint64_t)i);
    return PEDK_OK;
}

extern "C" int pedk_synth_mutate(const char* genome, uint64_t genome_len, uint64_t seed, uint32_t snp_ppm,
                                 uint32_t indel_ppm, char* out, uint64_t cap, uint64_t* out_len) {
    if (!genome || !out || !out_len) return PEDK_E_INVAL;
    static const char acgt[4] = {'A', 'C', 'G', 'T'};
    uint64_t o = 0;
    for (uint64_t i = 0; i < genome_len; ++i) {
        if (o + 12 > cap) return PEDK_E_NOMEM;
        uint64_t r = synth_rnd(seed, i, 11);
        uint64_t r2 = synth_rnd(seed, i, 12);
        if ((r % 1000000ull) < snp_ppm) {
            uint32_t c = base_code((uint32_t)genome[i]);
            out[o++] = acgt[(c + 1 + ((r >> 40) % 3)) & 3];  // always a different base
        } else if ((r2 % 1000000ull) < indel_ppm) {
            uint32_t len = 1 + (uint32_t)((r2 >> 40) % 10);
            if ((r2 >> 60) & 1) {
                i += len - 1;  // deletion of len bases starting here
            } else {
                out[o++] = genome[i];
                for (uint32_t j = 0; j < len; ++j) out[o++] = acgt[synth_rnd(seed, i * 16 + j, 13) & 3];
            }
        } else {
            out[o++] = genome[i];
        }
    }
    *out_len = o;
    return PEDK_OK;
}

extern "C" uint64_t pedk_synth_fastq_bytes(uint64_t first_read, uint64_t n_reads, uint32_t read_len) {
    return synth_record_offset(first_read + n_reads, read_len) - synth_record_offset(first_read, read_len);
}

extern "C" int pedk_synth_fastq_host(const pedk_synth_params* p, const char* genome, uint64_t first_read,
                                     uint64_t n_reads, char* out) {
    if (!p || !genome || !out || p->genome_len < p->read_len) return PEDK_E_INVAL;
    SynthParams s = to_params(p);
    uint64_t origin = synth_record_offset(first_read, s.read_len);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)n_reads; ++j) {
        const uint64_t i = first_read + (uint64_t)j;
        synth_write_record(s, genome, i, out + (synth_record_offset(i, s.read_len) - origin));
    }
    return PEDK_OK;
}
