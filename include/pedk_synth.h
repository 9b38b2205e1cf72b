/*
 * pedk_synth.h -- host-only synthetic inputs (libpedk_synth.so; tests, scripts and bench.py).
 *
 * Not part of the drop-in boundary: the reference has no generator.  SURVEY.md 8d defines the
 * inputs (uniform random genome, planted SNPs/indels, fixed-length reads with substitution errors
 * and a few reads holding an N).  Counter-based, so any slice of a file can be generated
 * independently, and libpedk.so's pedk_synth_fastq_device writes the same bytes in HBM.
 */
#ifndef PEDK_SYNTH_H
#define PEDK_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PEDK_SYNTH_PARAMS_DEFINED
#define PEDK_SYNTH_PARAMS_DEFINED
typedef struct pedk_synth_params {
    uint64_t seed;
    uint64_t genome_len;
    uint32_t read_len;
    uint32_t err_ppm;   /* substitution errors per million bases */
    uint32_t n_ppm;     /* reads with one N per million reads */
    uint32_t dup_ppm;   /* reads re-using the placement of the previous read, per million */
} pedk_synth_params;
#endif

int pedk_synth_genome(uint64_t seed, uint64_t genome_len, char *out);
int pedk_synth_mutate(const char *genome, uint64_t genome_len, uint64_t seed, uint32_t snp_ppm, uint32_t indel_ppm,
                      char *out, uint64_t cap, uint64_t *out_len);
uint64_t pedk_synth_fastq_bytes(uint64_t first_read, uint64_t n_reads, uint32_t read_len);
int pedk_synth_fastq_host(const pedk_synth_params *p, const char *genome, uint64_t first_read, uint64_t n_reads,
                          char *out);

#ifdef __cplusplus
}
#endif
#endif /* PEDK_SYNTH_H */
