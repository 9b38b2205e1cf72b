// S1: exact de-duplication of the strand-closed read set.
//
// ped.pl writes every read and its reverse complement to 64 bucket files and
// runs `sort | uniq` on each (ped.pl:1460-1464, 1416).  The set is closed under
// reverse complement, so one representative min(read, revcomp(read)) per pair
// carries the same information; equality of representatives is decided by a
// full-key comparison inside an open-addressing table (hash picks the slot,
// the packed words decide), so the result is exact, never probabilistic.
#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

__global__ void __launch_bounds__(256) k_dedup(const uint64_t* __restrict__ words, int W, uint32_t n,
                                              uint32_t* __restrict__ table, uint32_t mask,
                                              const uint8_t* __restrict__ valid, uint8_t* __restrict__ uniq) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (valid && !valid[i]) {
        uniq[i] = 0;
        return;
    }
    const uint64_t* me = words + (uint64_t)i * W;
    uint64_t key[PEDK_MAX_WORDS];
#pragma unroll
    for (int w = 0; w < PEDK_MAX_WORDS; ++w) key[w] = w < W ? me[w] : 0ull;
    uint64_t h = 0x243F6A8885A308D3ull;
#pragma unroll
    for (int w = 0; w < PEDK_MAX_WORDS; ++w)
        if (w < W) h = mix64(h ^ key[w]);
    uint32_t slot = (uint32_t)(h >> 20) & mask;
    uint8_t is_new = 0;
    for (;;) {
        uint32_t cur = table[slot];
        if (cur == 0xFFFFFFFFu) {
            cur = atomicCAS(&table[slot], 0xFFFFFFFFu, i);
            if (cur == 0xFFFFFFFFu) {
                is_new = 1;
                break;
            }
        }
        if (cur == i) {  // cannot happen (each index inserted once); guard against livelock
            is_new = 1;
            break;
        }
        const uint64_t* other = words + (uint64_t)cur * W;
        bool same = true;
#pragma unroll
        for (int w = 0; w < PEDK_MAX_WORDS; ++w)
            if (w < W) same &= (other[w] == key[w]);
        if (same) break;  // duplicate of an already inserted read
        slot = (slot + 1) & mask;
    }
    uniq[i] = is_new;
}

__global__ void __launch_bounds__(256) k_compact_reads(const uint64_t* __restrict__ words,
                                                      const uint8_t* __restrict__ pal,
                                                      const uint8_t* __restrict__ uniq,
                                                      const uint64_t* __restrict__ pos, int W, uint32_t n,
                                                      uint64_t* __restrict__ out_words,
                                                      uint8_t* __restrict__ out_pal) {
    // one thread per (read, word)
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t i = t / (unsigned)W;
    int w = (int)(t % (unsigned)W);
    if (i >= n || !uniq[i]) return;
    uint64_t o = pos[i];
    out_words[o * W + w] = words[i * W + w];
    if (w == 0) out_pal[o] = pal[i];
}

__global__ void __launch_bounds__(256) k_count_flags(const uint8_t* __restrict__ f, uint64_t n,
                                                    unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        c += f[i] != 0;
    c = __reduce_add_sync(0xffffffffu, (unsigned)c);
    if (lane_id() == 0 && c) atomicAdd(out, c);
}

}  // namespace

void launch_count_flags(const uint8_t* flags, uint64_t n, unsigned long long* out, cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 256 * 64 - 1) / (256 * 64);
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_count_flags<<<grid, 256, 0, st>>>(flags, n, out);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_dedup(const uint64_t* words, int W, uint32_t n, uint32_t* table, uint32_t capacity, const uint8_t* valid,
                  uint8_t* uniq, cudaStream_t st) {
    if (!n) return;
    k_dedup<<<(n + 255) / 256, 256, 0, st>>>(words, W, n, table, capacity - 1, valid, uniq);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_compact_reads(const uint64_t* words, const uint8_t* pal, const uint8_t* uniq, const uint64_t* pos,
                          int W, uint32_t n, uint64_t* out_words, uint8_t* out_pal, cudaStream_t st) {
    if (!n) return;
    uint64_t threads = (uint64_t)n * W;
    k_compact_reads<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(words, pal, uniq, pos, W, n, out_words,
                                                                        out_pal);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
