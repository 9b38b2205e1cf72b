// Microbenchmark: shared-memory atomic throughput on sm_100a (design input for bucket.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int THREADS = 512;
constexpr int ITEMS = 16;
constexpr int ROUNDS = 64;

__device__ __forceinline__ uint32_t rnd(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int MODE, int LOGBINS>
__global__ void __launch_bounds__(THREADS, 2) k(uint32_t* out, uint32_t seed) {
    extern __shared__ uint32_t sm[];
    constexpr uint32_t BINS = 1u << LOGBINS;
    for (uint32_t i = threadIdx.x; i < 2 * BINS; i += THREADS) sm[i] = MODE == 3 || MODE == 4 ? 0xffffffffu : 0;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t x = rnd(seed + blockIdx.x * THREADS + threadIdx.x);
    for (int r = 0; r < ROUNDS; ++r) {
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            x = x * 1664525u + 1013904223u;
            uint32_t b = (x >> 8) & (BINS - 1);
            if (MODE == 0) acc += b;                                  // ALU only
            if (MODE == 1) atomicAdd(&sm[b], 1u);                      // no return
            if (MODE == 2) acc += atomicAdd(&sm[b], 1u);               // return used
            if (MODE == 3) {                                           // CAS claim + add
                uint32_t old = atomicCAS(&sm[b], 0xffffffffu, b);
                if (old == 0xffffffffu || old == b) atomicAdd(&sm[BINS + b], 1u);
            }
            if (MODE == 4) {                                           // LDS check, CAS only when empty, add
                uint32_t cur = sm[b];
                if (cur != b) cur = atomicCAS(&sm[b], 0xffffffffu, b);
                atomicAdd(&sm[BINS + b], 1u);
                acc += cur;
            }
            if (MODE == 5) {                                           // match.any rank
                unsigned m = __match_any_sync(0xffffffffu, b);
                acc += __popc(m & ((1u << (threadIdx.x & 31)) - 1));
            }
            if (MODE == 6) {                                           // plain STS + LDS (no atomics)
                sm[b] = x; acc += sm[(b + 1) & (BINS - 1)];
            }
        }
    }
    __syncthreads();
    if (acc == 0x12345678u) out[0] = acc + sm[threadIdx.x & (BINS - 1)];
}

template <int MODE, int LOGBINS>
void run(const char* name) {
    uint32_t* out; cudaMalloc(&out, 4);
    size_t smem = (size_t)8 << LOGBINS;
    cudaFuncSetAttribute(k<MODE, LOGBINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = 148 * 2 * 4;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, LOGBINS><<<grid, THREADS, smem>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE, LOGBINS><<<grid, THREADS, smem>>>(out, 2);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)grid * THREADS * ITEMS * ROUNDS;
    printf("%-40s bins=2^%-2d  %8.3f ms  %8.1f Gops/s  (%s)\n", name, LOGBINS, ms, ops / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0, 8>("ALU only");
    run<1, 8>("atomicAdd no return");
    run<1, 13>("atomicAdd no return");
    run<2, 8>("atomicAdd with return");
    run<2, 13>("atomicAdd with return");
    run<3, 13>("CAS + add");
    run<4, 13>("LDS check, CAS if empty, add");
    run<5, 8>("match.any rank");
    run<6, 8>("STS + LDS");
    run<6, 13>("STS + LDS");
    return 0;
}
