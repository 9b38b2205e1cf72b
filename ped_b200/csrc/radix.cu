// S3: least-significant-digit radix sort of 64-bit k-mer words (optionally with a
// 32-bit payload), 8-bit digits, one "onesweep" pass per digit.
//
// Replaces `sort -S 1M` (ped.pl:928, 1416).  Per pass every key is read once
// and written once (16 B/key of HBM traffic, 24 B/row with a payload):
//   * the global digit histograms are produced upstream (by the k-mer
//     extraction kernel, or launch_digit_hist for row tables), so a pass is a
//     single kernel;
//   * a tile is 256 threads x 16 keys, four tiles resident per SM.  Phase 1
//     counts digits per warp with shared-memory atomics; the tile aggregate is
//     published for the decoupled look-back right away.  Phase 2 ranks the keys
//     stably inside each warp (eight vote.ballot rounds find the lanes holding
//     the same digit; match.any is data dependent and 2x slower on sm_100, and
//     an atomicOr mask table measured the same as the ballots -- profiles/) and
//     stores them straight into a shared-memory staging buffer in digit order.
//     The look-back over the 256 digit counters comes after phase 2, so the
//     predecessors have had that long to publish.  Phase 3 writes each digit
//     run contiguously;
//   * tiles take their index from an atomic ticket, so a tile only ever waits
//     on tiles that are already running; the wait is bounded (err flag);
//   * the digit position is a template parameter: digit and ballot predicates
//     are single instructions on one 32-bit half of the key.
// Look-back words are 32 bit (2 flag + 30 value bits); inputs longer than
// PORTION keys are sorted in portions, the last tile of a portion folding its
// inclusive digit counts into bin_base for the next portion.
#include <stdlib.h>

#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr size_t RS_PORTION = ((size_t(1) << 30) / RS_TILE - 1) * RS_TILE;

constexpr uint32_t LB32_AGG = 1u << 30;
constexpr uint32_t LB32_INC = 2u << 30;
constexpr uint32_t LB32_MASK = (1u << 30) - 1;

__device__ __forceinline__ void lb32_store(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lb32_load(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <bool PAIRS>
struct RadixSmem {
    uint64_t keys[RS_TILE];
    uint32_t warp_hist[RS_WARPS][256];
    uint64_t gbase[256];
    unsigned tile;
};
template <>
struct RadixSmem<true> {
    uint64_t keys[RS_TILE];
    uint32_t vals[RS_TILE];
    uint32_t warp_hist[RS_WARPS][256];
    uint64_t gbase[256];
    unsigned tile;
};

// the 32-bit half of the key that holds the digit, and the digit itself
template <int SHIFT>
__device__ __forceinline__ uint32_t digit_word(uint64_t key) {
    return SHIFT < 32 ? (uint32_t)key : (uint32_t)(key >> 32);
}
template <int SHIFT>
__device__ __forceinline__ unsigned digit_of(uint64_t key) {
    return (digit_word<SHIFT>(key) >> (SHIFT & 31)) & 255u;
}

// Lanes whose digit differs from this lane's in bit POS: ballot of the bit, flipped for the
// lanes that have it set.  Written in PTX so that each bit costs a predicate-setting LOP3,
// the VOTE, a SELP and one fused LOP3 (the C++ form compiled to seven instructions per bit).
template <int POS>
__device__ __forceinline__ unsigned bit_differs(uint32_t w) {
    unsigned r;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, m;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.b32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
        "selp.b32 t, -1, 0, p;\n\t"
        "xor.b32 %0, m, t;\n\t"
        "}"
        : "=r"(r)
        : "r"(w), "n"(1u << POS));
    return r;
}
template <int S>
__device__ __forceinline__ unsigned differing_lanes(uint32_t w) {
    return bit_differs<S>(w) | bit_differs<S + 1>(w) | bit_differs<S + 2>(w) | bit_differs<S + 3>(w) |
           bit_differs<S + 4>(w) | bit_differs<S + 5>(w) | bit_differs<S + 6>(w) | bit_differs<S + 7>(w);
}

template <bool PAIRS, int SHIFT>
__global__ void __launch_bounds__(RS_THREADS, 4) k_radix_pass(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                             const uint32_t* __restrict__ vin,
                                                             uint32_t* __restrict__ vout, size_t n,
                                                             unsigned long long* __restrict__ bin_base,
                                                             uint32_t* __restrict__ lookback,
                                                             unsigned int* __restrict__ tile_ctr, unsigned n_tiles,
                                                             int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RadixSmem<PAIRS>& sm = *reinterpret_cast<RadixSmem<PAIRS>*>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr int S = SHIFT & 31;

    if (tid == 0) sm.tile = atomicAdd(tile_ctr, 1u);
#pragma unroll
    for (int i = 0; i < 256 / 32; ++i) sm.warp_hist[warp][i * 32 + lane] = 0;
    // read the global base before anything is published (see portion hand-over below)
    const unsigned long long my_base = bin_base[tid];
    __syncthreads();
    const unsigned tile = sm.tile;
    const size_t tile_off = (size_t)tile * RS_TILE;
    const unsigned tile_n = (unsigned)((n - tile_off) < (size_t)RS_TILE ? (n - tile_off) : (size_t)RS_TILE);

    // ---- load (warp-striped: item j of lane l is key warp_off + 32j + l) ----
    uint64_t key[RS_ITEMS];
    uint32_t val[PAIRS ? RS_ITEMS : 1];
    const unsigned warp_off = warp * (32 * RS_ITEMS);
    if (tile_n == RS_TILE) {
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            key[j] = in[tile_off + warp_off + 32 * j + lane];
            if constexpr (PAIRS) val[j] = vin[tile_off + warp_off + 32 * j + lane];
        }
    } else {
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            unsigned i = warp_off + 32 * j + lane;
            key[j] = i < tile_n ? in[tile_off + i] : ~0ull;  // padding ranks last in the last bin
            if constexpr (PAIRS) val[j] = i < tile_n ? vin[tile_off + i] : 0u;
        }
    }

    // ---- phase 1: per-warp digit counts (order does not matter yet) ----
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) atomicAdd(&sm.warp_hist[warp][digit_of<SHIFT>(key[j])], 1u);
    __syncthreads();

    // ---- per digit: tile count -> publish the aggregate early, then turn the per-warp
    //      counts into absolute positions in the staging buffer ----
    unsigned dcount = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) dcount += sm.warp_hist[w][tid];
    lb32_store(lookback + (size_t)tile * 256 + tid, (tile == 0 ? LB32_INC : LB32_AGG) | dcount);
    const uint32_t ex = (uint32_t)block_exclusive_scan<RS_THREADS>(dcount, nullptr);
    {
        unsigned running = ex;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            unsigned c = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = running;
            running += c;
        }
    }
    __syncthreads();

    // ---- phase 2: stable rank inside the warp, straight into the staging buffer ----
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const uint32_t w = digit_word<SHIFT>(key[j]);
        // peers = lanes whose item j has the same digit: one ballot per digit bit;
        // acc collects the lanes that differ from this lane in some bit
        const unsigned acc = differing_lanes<S>(w);
        const unsigned peers = ~acc;
        const unsigned below = peers & lt;
        unsigned* ctr = &sm.warp_hist[warp][(w >> S) & 255u];
        const unsigned base = *ctr;  // every peer reads the same word
        __syncwarp();
        if (below == 0) *ctr = base + __popc(peers);  // the lowest peer moves the counter on
        __syncwarp();
        const unsigned pos = base + __popc(below);
        sm.keys[pos] = key[j];
        if constexpr (PAIRS) sm.vals[pos] = val[j];
    }

    // ---- look-back: predecessors had the whole of phase 2 to publish ----
    {
        uint32_t excl = 0;
        if (tile > 0) {
            long long t = (long long)tile - 1;
            unsigned spins = 0;
            while (t >= 0) {
                // up to four predecessors in flight at once
                uint32_t v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[q] = (t - q >= 0) ? lb32_load(lookback + (size_t)(t - q) * 256 + tid) : LB32_INC;
                bool done = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (done) break;
                    if ((v[q] >> 30) == 0) {  // not published yet: poll again from here
                        done = true;
                        if (++spins > (1u << 22)) {
                            atomicExch(err, 1);
                            t = -1;
                        }
                        break;
                    }
                    excl += v[q] & LB32_MASK;
                    --t;
                    if (v[q] & LB32_INC) {
                        t = -1;
                        done = true;
                    }
                }
            }
            lb32_store(lookback + (size_t)tile * 256 + tid, LB32_INC | (excl + dcount));
        }
        sm.gbase[tid] = my_base + excl - ex;
        // last tile of the portion: hand the running digit totals to the next portion
        if (tile == n_tiles - 1) bin_base[tid] = my_base + excl + dcount;
    }
    __syncthreads();

    // ---- phase 3: write each digit run contiguously ----
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        unsigned i = j * RS_THREADS + tid;
        if (i < tile_n) {
            uint64_t k = sm.keys[i];
            size_t g = (size_t)(sm.gbase[digit_of<SHIFT>(k)] + i);
            out[g] = k;
            if constexpr (PAIRS) vout[g] = sm.vals[i];
        }
    }
}

constexpr int DH_THREADS = 256;
constexpr int DH_MAX_PASS = 8;

template <int NPASS>
__global__ void __launch_bounds__(DH_THREADS) k_digit_hist(const uint64_t* __restrict__ keys, size_t n,
                                                          unsigned long long* __restrict__ digit_hist) {
    __shared__ unsigned h[NPASS][256];
    for (int i = threadIdx.x; i < NPASS * 256; i += DH_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * DH_THREADS;
    for (size_t i = (size_t)blockIdx.x * DH_THREADS + threadIdx.x; i < n; i += stride) {
        uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < NPASS; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * 256; i += DH_THREADS) {
        unsigned c = (&h[0][0])[i];
        if (c) atomicAdd(&digit_hist[i], (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(256) k_hist_to_base(unsigned long long* __restrict__ digit_hist) {
    unsigned long long* h = digit_hist + (size_t)blockIdx.x * 256;
    uint64_t v = h[threadIdx.x];
    uint64_t ex = block_exclusive_scan<256>(v, nullptr);
    h[threadIdx.x] = ex;
}

template <bool PAIRS, int SHIFT>
void launch_pass_variant(const uint64_t* in, uint64_t* out, const uint32_t* vin, uint32_t* vout, size_t pn,
                         unsigned long long* bin_base, RadixWorkspace ws, unsigned tiles, cudaStream_t st) {
    ensure_dyn_smem(k_radix_pass<PAIRS, SHIFT>, sizeof(RadixSmem<PAIRS>));
    k_radix_pass<PAIRS, SHIFT><<<tiles, RS_THREADS, sizeof(RadixSmem<PAIRS>), st>>>(
        in, out, vin, vout, pn, bin_base, reinterpret_cast<uint32_t*>(ws.lookback), ws.tile_ctr, tiles, ws.err);
}

template <bool PAIRS>
void launch_pass_shift(int pass, const uint64_t* in, uint64_t* out, const uint32_t* vin, uint32_t* vout, size_t pn,
                       unsigned long long* bin_base, RadixWorkspace ws, unsigned tiles, cudaStream_t st) {
    switch (pass) {
    case 0: launch_pass_variant<PAIRS, 0>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 1: launch_pass_variant<PAIRS, 8>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 2: launch_pass_variant<PAIRS, 16>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 3: launch_pass_variant<PAIRS, 24>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 4: launch_pass_variant<PAIRS, 32>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 5: launch_pass_variant<PAIRS, 40>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 6: launch_pass_variant<PAIRS, 48>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    case 7: launch_pass_variant<PAIRS, 56>(in, out, vin, vout, pn, bin_base, ws, tiles, st); break;
    default: throw PedkError(-8, "radix pass out of range");
    }
}

}  // namespace

size_t radix_tiles(size_t n) {
    size_t p = n < RS_PORTION ? n : RS_PORTION;
    return (p + RS_TILE - 1) / RS_TILE;
}

void launch_radix_pass(const uint64_t* in, uint64_t* out, const uint32_t* vin, uint32_t* vout, size_t n, int pass,
                       const unsigned long long* digit_hist, RadixWorkspace ws, cudaStream_t st) {
    if (!n) return;
    unsigned long long* bin_base = const_cast<unsigned long long*>(digit_hist) + (size_t)pass * 256;
    for (size_t off = 0; off < n; off += RS_PORTION) {
        size_t pn = (n - off) < RS_PORTION ? (n - off) : RS_PORTION;
        unsigned tiles = (unsigned)((pn + RS_TILE - 1) / RS_TILE);
        if (tiles > ws.tiles_cap) throw PedkError(-8, "radix workspace too small");
        PEDK_CUDA_TRY(cudaMemsetAsync(ws.lookback, 0, (size_t)tiles * 256 * sizeof(uint32_t), st));
        PEDK_CUDA_TRY(cudaMemsetAsync(ws.tile_ctr, 0, sizeof(unsigned), st));
        if (vin)
            launch_pass_shift<true>(pass, in + off, out, vin + off, vout, pn, bin_base, ws, tiles, st);
        else
            launch_pass_shift<false>(pass, in + off, out, nullptr, nullptr, pn, bin_base, ws, tiles, st);
        PEDK_CUDA_TRY(cudaGetLastError());
    }
}

void launch_digit_hist(const uint64_t* keys, size_t n, unsigned long long* digit_hist, int n_pass,
                       cudaStream_t st) {
    if (!n) return;
    size_t blocks = (n + DH_THREADS * 16 - 1) / (DH_THREADS * 16);
    unsigned grid = (unsigned)(blocks < (size_t)kNumSMs * 8 ? blocks : (size_t)kNumSMs * 8);
    switch (n_pass) {
    case 1: k_digit_hist<1><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    case 2: k_digit_hist<2><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    case 3: k_digit_hist<3><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    case 4: k_digit_hist<4><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    case 5: k_digit_hist<5><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    case 6: k_digit_hist<6><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    case 7: k_digit_hist<7><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    default: k_digit_hist<DH_MAX_PASS><<<grid, DH_THREADS, 0, st>>>(keys, n, digit_hist); break;
    }
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_hist_to_base(unsigned long long* digit_hist, int n_pass, cudaStream_t st) {
    k_hist_to_base<<<n_pass, 256, 0, st>>>(digit_hist);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
