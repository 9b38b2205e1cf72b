// Internal utilities for libpedk.so: error plumbing, warp/block primitives and
// a 3-phase exclusive scan.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "common.cuh"

namespace pedk {

struct Err {
    int code = 0;
    std::string msg;
};

#define PEDK_CUDA_TRY(expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _b[512];                                                                       \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                     cudaGetErrorString(_e));                                                   \
            throw ::pedk::CudaError(_b);                                                        \
        }                                                                                       \
    } while (0)

struct CudaError {
    std::string msg;
    explicit CudaError(const char* m) : msg(m) {}
};
struct PedkError {
    int code;
    std::string msg;
    PedkError(int c, std::string m) : code(c), msg(std::move(m)) {}
};

constexpr int kNumSMs = 148;  // B200

// Opt a kernel in to `bytes` of dynamic shared memory (> 48 KB needs it).  The attribute belongs to
// the CURRENT device's context, so it is remembered per (kernel, device) -- a process may hold
// contexts on several GPUs (ctx.cu).
void set_max_dyn_smem(const void* kernel, size_t bytes);
template <typename F>
inline void ensure_dyn_smem(F* kernel, size_t bytes) {
    set_max_dyn_smem(reinterpret_cast<const void*>(kernel), bytes);
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- look-back words: 2 flag bits + 62 value bits, one 64-bit store ----
constexpr uint64_t LB_FLAG_AGG = 1ull << 62;  // tile aggregate available
constexpr uint64_t LB_FLAG_INC = 2ull << 62;  // inclusive prefix available
constexpr uint64_t LB_VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ void lb_store(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t lb_load(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Walk predecessors of `tile` in a chained scan with one word per tile
// (stride words apart) and return the exclusive prefix.  `err` is raised and 0
// returned if a predecessor never shows up (bounded spin, so a logic error
// cannot hang the GPU).
__device__ __forceinline__ uint64_t lb_exclusive_prefix(const uint64_t* lb, long long tile, size_t stride,
                                                        int* err) {
    uint64_t sum = 0;
    for (long long t = tile - 1; t >= 0; --t) {
        uint64_t v;
        unsigned spins = 0;
        while (((v = lb_load(lb + (size_t)t * stride)) >> 62) == 0) {
            __nanosleep(20);
            if (++spins > (1u << 24)) {
                atomicExch(err, 1);
                return 0;
            }
        }
        sum += v & LB_VAL_MASK;
        if (v & LB_FLAG_INC) break;
    }
    return sum;
}

// Block-wide exclusive scan of one u64 per thread.  THREADS multiple of 32, <= 1024.
template <int THREADS>
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total) {
    __shared__ uint64_t warp_sums[THREADS / 32];
    __shared__ uint64_t block_total;
    unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint64_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < THREADS / 32 ? warp_sums[lane] : 0;
        uint64_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += o;
        }
        if (lane < THREADS / 32) warp_sums[lane] = winc - w;
        if (lane == 31) block_total = winc;
    }
    __syncthreads();
    uint64_t r = warp_sums[warp] + inc - v;
    if (total) *total = block_total;
    __syncthreads();
    return r;
}

// out[i] = sum_{j<i} in[j]; *total_dev = sum of all.  ws must hold
// scan_workspace_words(n) u64 words.  Three launches, no spin waits.
size_t scan_workspace_words(size_t n);
void exclusive_scan_u32(const uint32_t* in, uint64_t* out, size_t n, unsigned long long* total_dev, uint64_t* ws,
                        cudaStream_t st);
void exclusive_scan_u8(const uint8_t* in, uint64_t* out, size_t n, unsigned long long* total_dev, uint64_t* ws,
                       cudaStream_t st);

}  // namespace pedk
