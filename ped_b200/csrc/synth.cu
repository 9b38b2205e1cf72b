// Synthetic genomes and FASTQ reads (tests and bench.py only).
// The host functions make no CUDA call, so they also run in the GPU-less build
// container to feed the reference (scripts/make_golden.py).
#include <string.h>

#include "../../include/pedk.h"
#include "kernels.h"
#include "synth.h"
#include "util.cuh"

namespace pedk {

namespace {
__global__ void __launch_bounds__(128) k_synth_fastq(SynthParams p, const char* __restrict__ genome, uint64_t first,
                                                    uint64_t n, uint64_t out_origin, char* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t i = first + t;
    synth_write_record(p, genome, i, out + (synth_record_offset(i, p.read_len) - out_origin));
}
}  // namespace

void launch_synth_fastq(const SynthParams& p, const char* genome_dev, uint64_t first, uint64_t n, char* out_dev,
                        cudaStream_t st) {
    if (!n) return;
    uint64_t origin = synth_record_offset(first, p.read_len);
    k_synth_fastq<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, genome_dev, first, n, origin, out_dev);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk

using namespace pedk;

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
    for (uint64_t i = 0; i < genome_len; ++i) out[i] = synth_genome_base(seed, i);
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
    for (uint64_t i = first_read; i < first_read + n_reads; ++i)
        synth_write_record(s, genome, i, out + (synth_record_offset(i, s.read_len) - origin));
    return PEDK_OK;
}

namespace pedk {
cudaStream_t ctx_stream(pedk_ctx* ctx);
int ctx_fail(pedk_ctx* ctx, int code, const std::string& msg);
void ctx_count_launch(pedk_ctx* ctx, int stage, uint64_t n);
}  // namespace pedk

extern "C" int pedk_synth_fastq_device(pedk_ctx* ctx, const pedk_synth_params* p, const char* genome_dev,
                                       uint64_t first_read, uint64_t n_reads, char* out_dev) {
    if (!ctx || !p || !genome_dev || !out_dev || p->genome_len < p->read_len) return PEDK_E_INVAL;
    try {
        launch_synth_fastq(to_params(p), genome_dev, first_read, n_reads, out_dev, ctx_stream(ctx));
        PEDK_CUDA_TRY(cudaStreamSynchronize(ctx_stream(ctx)));
    } catch (const CudaError& e) {
        return ctx_fail(ctx, PEDK_E_CUDA, e.msg);
    }
    return PEDK_OK;
}
