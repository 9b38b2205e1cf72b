// Synthetic FASTQ reads generated in HBM (tests and bench.py only).  The host-side
// generator of the same bytes lives in synth_host.cc -> libpedk_synth.so (include/pedk_synth.h), so
// that feeding a CPU arm never maps the product library.
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
