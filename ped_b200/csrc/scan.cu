// 3-phase exclusive scan (reduce per block -> scan block sums -> downsweep).
// Used for bookkeeping arrays (newline counts, chunk counts, unique flags); the
// k-mer stream itself never goes through here.
#include "util.cuh"

namespace pedk {

namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const T* __restrict__ in, size_t n,
                                                             uint64_t* __restrict__ block_sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        size_t i = base + (size_t)j * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint64_t total;
    block_exclusive_scan<SCAN_THREADS>(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_block_sums(uint64_t* __restrict__ block_sums, size_t nb,
                                                         unsigned long long* __restrict__ total_dev) {
    uint64_t carry = 0;
    for (size_t base = 0; base < nb; base += 1024) {
        size_t i = base + threadIdx.x;
        uint64_t v = i < nb ? block_sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_scan<1024>(v, &total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_dev) *total_dev = carry;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const T* __restrict__ in, size_t n,
                                                           const uint64_t* __restrict__ block_sums,
                                                           uint64_t* __restrict__ out) {
    // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        size_t i = base + j;
        v[j] = i < n ? (uint64_t)in[i] : 0;
        s += v[j];
    }
    uint64_t ex = block_exclusive_scan<SCAN_THREADS>(s, nullptr) + block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        size_t i = base + j;
        if (i < n) out[i] = ex;
        ex += v[j];
    }
}

template <typename T>
void scan_impl(const T* in, uint64_t* out, size_t n, unsigned long long* total_dev, uint64_t* ws, cudaStream_t st) {
    if (n == 0) {
        if (total_dev) PEDK_CUDA_TRY(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), st));
        return;
    }
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_reduce<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, ws);
    k_scan_block_sums<<<1, 1024, 0, st>>>(ws, nb, total_dev);
    k_scan_down<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, ws, out);
    PEDK_CUDA_TRY(cudaGetLastError());
}
}  // namespace

size_t scan_workspace_words(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

void exclusive_scan_u32(const uint32_t* in, uint64_t* out, size_t n, unsigned long long* total_dev, uint64_t* ws,
                        cudaStream_t st) {
    scan_impl<uint32_t>(in, out, n, total_dev, ws, st);
}
void exclusive_scan_u8(const uint8_t* in, uint64_t* out, size_t n, unsigned long long* total_dev, uint64_t* ws,
                       cudaStream_t st) {
    scan_impl<uint8_t>(in, out, n, total_dev, ws, st);
}

}  // namespace pedk
