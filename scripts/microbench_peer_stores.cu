// Microbenchmark: how fast can one B200 store short runs into a peer's HBM over NVLink 5?
// (design input for the fused k-mer exchange: k_part_scatter stages a tile by bin in shared
// memory and stores every bin's run -- a few hundred bytes -- into the owner's buffer.)
//
//   mode 0  st.global from the staged tile, thread i stores element i of the tile (as k_part_scatter)
//   mode 1  cp.async.bulk shared -> global, one bulk copy per run (16-byte aligned runs)
// for run lengths of 160 B .. 16 KiB, to local memory and to the peer, with 1024 x 1 and 512 x 2
// threads x CTAs per SM.  Usage: ./microbench_peer_stores   (needs two GPUs with peer access)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// TILE_BYTES of shared memory per CTA, written out `rounds` times as runs of run_bytes each to
// pseudo-random run slots of dst (dst_runs slots).  spin: ALU work between two store phases
// (mimics the extract + hash phase of the real kernel: no stores in flight meanwhile).
template <int MODE>
__global__ void k_store(uint8_t* __restrict__ dst, uint64_t dst_runs, int run_bytes, int tile_bytes, int rounds,
                        int spin, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t tile[];
    for (int i = threadIdx.x * 4; i < tile_bytes; i += blockDim.x * 4) *reinterpret_cast<uint32_t*>(tile + i) = i;
    __syncthreads();
    const int runs = tile_bytes / run_bytes;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        for (int s = 0; s < spin; ++s) acc = mix(acc + s);
        const uint32_t seed = (blockIdx.x * 1000003u + r) * 2654435761u;
        if (MODE == 0) {
            // thread i stores 4-byte element i of the staged tile: consecutive threads, consecutive
            // addresses inside a run; a warp-store may straddle two runs
            for (int i = threadIdx.x; i < tile_bytes / 4; i += blockDim.x) {
                const int run = (i * 4) / run_bytes, off = (i * 4) % run_bytes;
                const uint64_t slot = mix(seed + run) % dst_runs;
                *reinterpret_cast<uint32_t*>(dst + slot * (uint64_t)run_bytes + off) =
                    *reinterpret_cast<const uint32_t*>(tile + i * 4);
            }
        } else {
            // one thread per run issues a bulk copy shared -> global; the CTA waits for the group
            // to have READ shared memory before the tile is reused (here: immediately)
            for (int run = threadIdx.x; run < runs; run += blockDim.x) {
                const uint64_t slot = mix(seed + run) % dst_runs;
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile + (size_t)run * run_bytes);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + slot * (uint64_t)run_bytes),
                             "r"(src), "r"(run_bytes)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE>
double run(uint8_t* dst, uint64_t dst_bytes, int run_bytes, int threads, int ctas_per_sm, int tile_bytes, int rounds, int spin,
           uint32_t* sink) {
    const uint64_t dst_runs = dst_bytes / run_bytes;
    const int grid = 148 * ctas_per_sm;
    CK(cudaFuncSetAttribute(k_store<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    k_store<MODE><<<grid, threads, tile_bytes>>>(dst, dst_runs, run_bytes, tile_bytes, 2, spin, sink);  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k_store<MODE><<<grid, threads, tile_bytes>>>(dst, dst_runs, run_bytes, tile_bytes, rounds, spin, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return (double)grid * rounds * (tile_bytes / run_bytes * run_bytes) / (ms * 1e-3) / 1e9;
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    const uint64_t BYTES = 4ull << 30;
    uint8_t *local = nullptr, *peer = nullptr;
    uint32_t* sink;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&local, BYTES));
    CK(cudaMalloc(&sink, 4));
    if (n >= 2) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, 0, 1));
        if (can) {
            CK(cudaSetDevice(1));
            CK(cudaMalloc(&peer, BYTES));
            CK(cudaSetDevice(0));
            CK(cudaDeviceEnablePeerAccess(1, 0));
        }
    }
    printf("devices %d, peer buffer %s\n", n, peer ? "yes" : "NO (local numbers only)");
    const int runs[] = {160, 320, 640, 1280, 2560, 16384};
    struct Shape { int threads, ctas, tile; } shapes[] = {{1024, 1, 81920}, {512, 2, 40960}, {1024, 2, 81920}};
    for (int where = 0; where < 2; ++where) {
        uint8_t* dst = where ? peer : local;
        if (!dst) continue;
        for (const Shape& sh : shapes)
            for (int spin : {0, 2000})
                for (int rb : runs) {
                    const double g0 = run<0>(dst, BYTES, rb, sh.threads, sh.ctas, sh.tile, 200, spin, sink);
                    const double g1 = rb % 16 == 0 ? run<1>(dst, BYTES, rb, sh.threads, sh.ctas, sh.tile, 200, spin, sink) : 0.0;
                    printf("%-5s %4d thr x %d CTA/SM tile %6d B  spin %4d  run %6d B   st.global %8.1f GB/s   cp.async.bulk %8.1f GB/s\n",
                           where ? "peer" : "local", sh.threads, sh.ctas, sh.tile, spin, rb, g0, g1);
                    fflush(stdout);
                }
    }
    return 0;
}
