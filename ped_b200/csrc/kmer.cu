// S2: k-mer extraction and S4: run-length count.
//
// S2 replaces the Perl substr loop of countKmerSub (ped.pl:909-920): one warp
// per unique read, lane i takes windows i, i+32, ... out of the packed words
// (two warp shuffles + a funnel shift), writes min(K, revcomp K) and feeds the
// per-digit histograms the radix sort needs.  The reference slices both the
// read and its reverse complement; every window of the reverse complement is
// the reverse complement of a window of the read, so the canonical stream holds
// exactly half of the reference's instances (see edges.cu for how the strand
// table is rebuilt).
//
// S4 replaces `uniq -c | awk` and the 64-round `join` merge of countKmerMerge
// (ped.pl:928, 862-889): one pass over the sorted stream, chained scan over the
// number of run heads per tile.
#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

constexpr int EX_THREADS = 256;
constexpr int EX_MAX_PASS = 8;

__global__ void __launch_bounds__(EX_THREADS) k_extract(const uint64_t* __restrict__ words, int W, int L, int k,
                                                       uint64_t n_reads, uint64_t* __restrict__ out,
                                                       unsigned long long* __restrict__ digit_hist, int n_pass) {
    __shared__ unsigned h[EX_MAX_PASS][256];
    for (int i = threadIdx.x; i < EX_MAX_PASS * 256; i += EX_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * EX_THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * EX_THREADS) >> 5;
    const int nwin = L - k + 1;
    const int kshift = 64 - 2 * k;
    for (uint64_t r = warp; r < n_reads; r += n_warps) {
        uint64_t mine = (int)lane < W ? words[r * (uint64_t)W + lane] : 0ull;
        uint64_t* dst = out + r * (uint64_t)nwin;
        for (int t = 0; 32 * t < nwin; ++t) {
            uint64_t hi = __shfl_sync(0xffffffffu, mine, t);
            uint64_t lo = __shfl_sync(0xffffffffu, mine, t + 1);
            int i = 32 * t + (int)lane;
            uint64_t x = lane ? ((hi << (2 * lane)) | (lo >> (64 - 2 * lane))) : hi;
            uint64_t km = x >> kshift;
            uint64_t rc = revcomp_groups64(km) >> kshift;
            uint64_t canon = km < rc ? km : rc;
            if (i < nwin) {
                dst[i] = canon;
                for (int p = 0; p < n_pass; ++p) atomicAdd(&h[p][(canon >> (8 * p)) & 255u], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pass * 256; i += EX_THREADS) {
        unsigned c = (&h[0][0])[i];
        if (c) atomicAdd(&digit_hist[i], (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(EX_THREADS) k_extract_pal(const uint64_t* __restrict__ words,
                                                           const uint8_t* __restrict__ pal, int W, int L, int k,
                                                           uint64_t n_reads, uint64_t* __restrict__ out,
                                                           unsigned long long* __restrict__ cursor) {
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * EX_THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * EX_THREADS) >> 5;
    const int nwin = L - k + 1;
    const int kshift = 64 - 2 * k;
    for (uint64_t r0 = warp * 32; r0 < n_reads; r0 += n_warps * 32) {
        uint64_t rr = r0 + lane;
        unsigned m = __ballot_sync(0xffffffffu, rr < n_reads && pal[rr] != 0);
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            uint64_t r = r0 + b;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)nwin);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint64_t mine = (int)lane < W ? words[r * (uint64_t)W + lane] : 0ull;
            for (int t = 0; 32 * t < nwin; ++t) {
                uint64_t hi = __shfl_sync(0xffffffffu, mine, t);
                uint64_t lo = __shfl_sync(0xffffffffu, mine, t + 1);
                int i = 32 * t + (int)lane;
                uint64_t x = lane ? ((hi << (2 * lane)) | (lo >> (64 - 2 * lane))) : hi;
                uint64_t km = x >> kshift;
                uint64_t rc = revcomp_groups64(km) >> kshift;
                if (i < nwin) out[base + i] = km < rc ? km : rc;
            }
        }
    }
}

constexpr int RL_THREADS = 256;
constexpr int RL_ITEMS = 16;
constexpr int RL_TILE = RL_THREADS * RL_ITEMS;
constexpr int RL_WARPS = RL_THREADS / 32;

__global__ void __launch_bounds__(RL_THREADS) k_rle(const uint64_t* __restrict__ keys, size_t n,
                                                   uint64_t* __restrict__ ukey, uint32_t* __restrict__ ustart,
                                                   uint64_t* __restrict__ lookback,
                                                   unsigned int* __restrict__ tile_ctr, unsigned n_tiles,
                                                   int* __restrict__ err, unsigned long long* __restrict__ n_runs) {
    __shared__ unsigned s_tile;
    __shared__ unsigned warp_heads[RL_WARPS];
    __shared__ unsigned warp_excl[RL_WARPS];
    __shared__ uint64_t s_tile_excl;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_ctr, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t seg = (size_t)tile * RL_TILE + (size_t)warp * (32 * RL_ITEMS);

    uint64_t key[RL_ITEMS];
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        size_t i = seg + 32 * j + lane;
        key[j] = i < n ? keys[i] : 0ull;
    }
    unsigned heads[RL_ITEMS];  // ballot of head flags per item (warp-uniform)
    unsigned wcount = 0;
    uint64_t carry = (seg > 0 && seg <= n) ? keys[seg - 1] : 0ull;  // key just before the segment
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        size_t i = seg + 32 * j + lane;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key[j], 1);
        if (lane == 0) prev = carry;
        bool head = i < n && (i == 0 || key[j] != prev);
        heads[j] = __ballot_sync(0xffffffffu, head);
        wcount += __popc(heads[j]);
        carry = __shfl_sync(0xffffffffu, key[j], 31);
    }
    if (lane == 0) warp_heads[warp] = wcount;
    __syncthreads();
    if (tid == 0) {
        unsigned total = 0;
#pragma unroll
        for (int w = 0; w < RL_WARPS; ++w) {
            warp_excl[w] = total;
            total += warp_heads[w];
        }
        lb_store(lookback + tile, (tile == 0 ? LB_FLAG_INC : LB_FLAG_AGG) | (uint64_t)total);
        uint64_t excl = 0;
        if (tile > 0) {
            excl = lb_exclusive_prefix(lookback, (long long)tile, 1, err);
            lb_store(lookback + tile, LB_FLAG_INC | (excl + total));
        }
        s_tile_excl = excl;
        if (tile == n_tiles - 1) {
            ustart[excl + total] = (uint32_t)n;
            *n_runs = excl + total;
        }
    }
    __syncthreads();
    uint64_t r = s_tile_excl + warp_excl[warp];
#pragma unroll
    for (int j = 0; j < RL_ITEMS; ++j) {
        unsigned b = heads[j];
        if (b & (1u << lane)) {
            uint64_t o = r + __popc(b & lanemask_lt());
            ukey[o] = key[j];
            ustart[o] = (uint32_t)(seg + 32 * j + lane);
        }
        r += __popc(b);
    }
}

}  // namespace

void launch_extract(const uint64_t* words, int W, int L, int k, uint64_t n_reads, uint64_t* out,
                    unsigned long long* digit_hist, int n_pass, cudaStream_t st) {
    if (!n_reads || L < k) return;
    uint64_t blocks = (n_reads + 7) / 8;
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_extract<<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist, n_pass);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_extract_pal(const uint64_t* words, const uint8_t* pal, int W, int L, int k, uint64_t n_reads,
                        uint64_t* out, unsigned long long* cursor, cudaStream_t st) {
    if (!n_reads || L < k) return;
    uint64_t blocks = (n_reads + 255) / 256;
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_extract_pal<<<grid, EX_THREADS, 0, st>>>(words, pal, W, L, k, n_reads, out, cursor);
    PEDK_CUDA_TRY(cudaGetLastError());
}

size_t rle_tiles(size_t n) { return (n + RL_TILE - 1) / RL_TILE; }

void launch_rle(const uint64_t* keys, size_t n, uint64_t* ukey, uint32_t* ustart, uint64_t* lookback,
                unsigned int* tile_ctr, int* err, unsigned long long* n_runs, cudaStream_t st) {
    if (!n) {
        PEDK_CUDA_TRY(cudaMemsetAsync(n_runs, 0, sizeof(unsigned long long), st));
        PEDK_CUDA_TRY(cudaMemsetAsync(ustart, 0, sizeof(uint32_t), st));
        return;
    }
    size_t tiles = rle_tiles(n);
    PEDK_CUDA_TRY(cudaMemsetAsync(lookback, 0, tiles * sizeof(uint64_t), st));
    PEDK_CUDA_TRY(cudaMemsetAsync(tile_ctr, 0, sizeof(unsigned), st));
    k_rle<<<(unsigned)tiles, RL_THREADS, 0, st>>>(keys, n, ukey, ustart, lookback, tile_ctr, (unsigned)tiles, err,
                                                  n_runs);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
