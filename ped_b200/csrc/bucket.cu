// S2-S4, bucketed: k-mer extraction, two-level binning and per-bucket counting.
//
// Replaces the Perl substr loop, the 64x64 matrix of partial files, `sort -S 1M |
// uniq -c` and the 64-round `join` merge of countKmerSub / countKmerMerge
// (ped.pl:891-932, 862-889).  The reference needs its k-mers *sorted* only so that
// `uniq -c` sees equal lines next to each other; what the table needs is that equal
// k-mers meet.  So the instance stream is never sorted:
//
//   k_part_count     packed reads -> 2^a-bin histogram of the top bits of kmer_hash(K)
//                    (reads only, no stores): sizes the level-A bins exactly
//   k_part_scatter   packed reads -> canonical k-mers -> hash -> level-A bin; a CTA
//                    stages its tile by bin in shared memory and appends every bin's
//                    run with ONE global atomic per (tile, bin).  The bin is the top
//                    a bits, so only the low 2k-a bits are stored: 4 bytes per
//                    instance at k <= 20 instead of 8.  On several GPUs the owner of
//                    a bin is a rank and the run is stored straight into that rank's
//                    buffer over NVLink (peer memory), so this kernel IS the exchange.
//   k_sub_hist       level-A bins -> exact sizes of the 2^a x 2^b buckets
//   k_sub_scatter    same staging, one level down; a tile never leaves its bin
//   k_bucket_count   one CTA per bucket: a shared-memory hash table keyed by the
//                    remaining bits counts the instances of every distinct k-mer
//                    (a k-mer with millions of instances costs one table slot); the
//                    occupied slots become the strand rows (K, c), (revcomp K, c).
//                    A bucket with more distinct k-mers than the table holds is
//                    split by further hash bits and re-read (rare: sizes are even
//                    because the hash is a bijection of the k-mer, not of the genome).
//
// Order inside a bin is arbitrary (atomic ranks), which is fine: nothing downstream
// depends on it, and the rows are sorted afterwards anyway.  Per instance the stream
// is written 2x and read 3x at 4 bytes (20 B) against 8 + 5 x 16 + 8 = 96 B for the
// extract + LSD sort + run-length path it replaces (radix.cu keeps sorting the rows).
#include <stdlib.h>

#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

constexpr int NB = 256;  // bins per level (stride of the bucket index even when 2^b < 256)
constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------ level A: count
constexpr int PC_THREADS = 256;

__global__ void __launch_bounds__(PC_THREADS) k_part_count(const uint64_t* __restrict__ words, int W, int L, int k,
                                                          uint64_t n_reads, int shiftA,
                                                          unsigned long long* __restrict__ histA) {
    __shared__ unsigned h[NB];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * PC_THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * PC_THREADS) >> 5;
    const int nwin = L - k + 1, kshift = 64 - 2 * k, nbits = 2 * k;
    const int nt = (nwin + 31) >> 5;
    for (uint64_t r = warp; r < n_reads; r += n_warps) {
        const uint64_t mine = (int)lane < W ? words[r * (uint64_t)W + lane] : 0ull;
        for (int t = 0; t < nt; ++t) {
            const uint64_t hi = __shfl_sync(FULL, mine, t);
            const uint64_t lo = __shfl_sync(FULL, mine, t + 1);
            const int i = 32 * t + (int)lane;
            const uint64_t x = lane ? ((hi << (2 * lane)) | (lo >> (64 - 2 * lane))) : hi;
            const uint64_t km = x >> kshift;
            const uint64_t rc = revcomp_groups64(km) >> kshift;
            const uint64_t canon = km < rc ? km : rc;
            if (i < nwin) atomicAdd(&h[kmer_hash(canon, nbits) >> shiftA], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&histA[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ------------------------------------------------------------------ level A: scatter
constexpr int PT_THREADS = 512;
constexpr int PT_WARPS = PT_THREADS / 32;

template <typename KT>
struct KeyBudget;
template <>
struct KeyBudget<uint32_t> {
    static constexpr int ITEMS = 20;  // keys a thread keeps in registers between the two phases
};
template <>
struct KeyBudget<uint64_t> {
    static constexpr int ITEMS = 12;
};

// NT = windows per lane per read (ceil((L-k+1)/32) rounded up to an instantiated value),
// RPW = reads per warp per tile.
template <typename KT, int NT, int RPW>
__global__ void __launch_bounds__(PT_THREADS, 2)
    k_part_scatter(const uint64_t* __restrict__ words, int W, int L, int k, uint64_t n_reads, int a, int nranks,
                   unsigned long long* __restrict__ cur, PeerPtrs dst) {
    constexpr int ITEMS = NT * RPW;
    constexpr int TILE = PT_THREADS * ITEMS;
    static_assert(TILE <= 65536, "ranks are kept in 16 bits");
    extern __shared__ __align__(16) unsigned char pt_smem[];
    KT* stage = reinterpret_cast<KT*>(pt_smem);
    uint8_t* sbin = reinterpret_cast<uint8_t*>(stage + TILE);
    __shared__ unsigned hist[NB];
    __shared__ unsigned off[NB];
    __shared__ unsigned long long gdelta[NB];
    __shared__ KT* s_ptr[NB];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int nwin = L - k + 1, kshift = 64 - 2 * k, nbits = 2 * k;
    const int shiftA = nbits - a;
    const uint64_t maskS = (1ull << shiftA) - 1;  // shiftA <= 56
    if (tid < NB) {
        const int owner = (int)tid < (1 << a) ? (int)((tid * (unsigned)nranks) >> a) : 0;
        s_ptr[tid] = reinterpret_cast<KT*>(dst.p[owner]);
    }
    constexpr uint64_t READS_PER_TILE = (uint64_t)PT_WARPS * RPW;
    const uint64_t n_tiles = (n_reads + READS_PER_TILE - 1) / READS_PER_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid < NB) hist[tid] = 0;
        __syncthreads();
        // ---- extract, hash, take a rank inside (tile, bin) ----
        KT keyv[ITEMS];
        uint32_t meta[ITEMS];  // bin << 16 | rank, ~0 = no key
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const uint64_t r = tile * READS_PER_TILE + (uint64_t)warp * RPW + rr;
            const bool live = r < n_reads;
            const uint64_t mine = (live && (int)lane < W) ? words[r * (uint64_t)W + lane] : 0ull;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int idx = rr * NT + t;
                const uint64_t hi = __shfl_sync(FULL, mine, t);
                const uint64_t lo = __shfl_sync(FULL, mine, t + 1);
                const int i = 32 * t + (int)lane;
                const uint64_t x = lane ? ((hi << (2 * lane)) | (lo >> (64 - 2 * lane))) : hi;
                const uint64_t km = x >> kshift;
                const uint64_t rc = revcomp_groups64(km) >> kshift;
                const uint64_t canon = km < rc ? km : rc;
                meta[idx] = 0xffffffffu;
                keyv[idx] = 0;
                if (live && i < nwin) {
                    const uint64_t h = kmer_hash(canon, nbits);
                    const unsigned bin = (unsigned)(h >> shiftA);
                    const unsigned rank = atomicAdd(&hist[bin], 1u);
                    keyv[idx] = (KT)(h & maskS);
                    meta[idx] = (bin << 16) | rank;
                }
            }
        }
        __syncthreads();
        // ---- bin offsets inside the tile; one global atomic per (tile, bin) ----
        const unsigned c = tid < NB ? hist[tid] : 0u;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan<PT_THREADS>(c, &total);
        if (tid < NB) {
            const unsigned long long g = c ? atomicAdd(&cur[tid], (unsigned long long)c) : 0ull;
            off[tid] = (unsigned)ex;
            gdelta[tid] = g - ex;
        }
        __syncthreads();
#pragma unroll
        for (int idx = 0; idx < ITEMS; ++idx) {
            if (meta[idx] != 0xffffffffu) {
                const unsigned bin = meta[idx] >> 16;
                const unsigned pos = off[bin] + (meta[idx] & 0xffffu);
                stage[pos] = keyv[idx];
                sbin[pos] = (uint8_t)bin;
            }
        }
        __syncthreads();
        // ---- every bin's run goes out contiguously (to the owner's memory) ----
        const unsigned n_staged = (unsigned)total;
        for (unsigned i = tid; i < n_staged; i += PT_THREADS) {
            const unsigned bin = sbin[i];
            s_ptr[bin][gdelta[bin] + i] = stage[i];
        }
    }
}

// ------------------------------------------------------------------ level B
constexpr int SUB_TILE = 8192;  // keys per tile; a tile lies inside one level-A bin
constexpr int SH_THREADS = 256;
constexpr int SH_GROUP = 8;  // tiles per CTA of the histogram kernel
constexpr int SS_THREADS = 512;
constexpr int SS_ITEMS = SUB_TILE / SS_THREADS;

// largest bin with tile_prefix[bin] <= t (bins without tiles are never returned)
__device__ __forceinline__ int bin_of_tile(const unsigned* __restrict__ tile_prefix, int n_bins, unsigned t) {
    int lo = 0, hi = n_bins;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_prefix[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename KT>
__global__ void __launch_bounds__(SH_THREADS) k_sub_hist(const KT* __restrict__ keys,
                                                        const unsigned long long* __restrict__ offA,
                                                        const unsigned* __restrict__ tile_prefix, int n_bins,
                                                        unsigned n_tiles, int shiftB, unsigned maskB,
                                                        unsigned* __restrict__ histB) {
    __shared__ unsigned h[NB];
    const unsigned tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    int cur_bin = -1;
    const unsigned t0 = blockIdx.x * SH_GROUP;
    const unsigned t1 = t0 + SH_GROUP < n_tiles ? t0 + SH_GROUP : n_tiles;
    for (unsigned t = t0; t < t1; ++t) {
        const int bin = bin_of_tile(tile_prefix, n_bins, t);
        if (bin != cur_bin) {
            if (cur_bin >= 0) {
                __syncthreads();
                if (h[tid]) atomicAdd(&histB[(size_t)cur_bin * NB + tid], h[tid]);
                h[tid] = 0;
                __syncthreads();
            }
            cur_bin = bin;
        }
        const unsigned long long start = offA[bin] + (unsigned long long)(t - tile_prefix[bin]) * SUB_TILE;
        const unsigned long long lim = offA[bin + 1];
        const unsigned long long end = start + SUB_TILE < lim ? start + SUB_TILE : lim;
        for (unsigned long long i = start + tid; i < end; i += SH_THREADS)
            atomicAdd(&h[(unsigned)(keys[i] >> shiftB) & maskB], 1u);
    }
    __syncthreads();
    if (cur_bin >= 0 && h[tid]) atomicAdd(&histB[(size_t)cur_bin * NB + tid], h[tid]);
}

template <typename KT>
__global__ void __launch_bounds__(SS_THREADS, 2) k_sub_scatter(const KT* __restrict__ in, KT* __restrict__ out,
                                                              const unsigned long long* __restrict__ offA,
                                                              const unsigned* __restrict__ tile_prefix, int n_bins,
                                                              unsigned n_tiles, int shiftB, unsigned maskB,
                                                              unsigned long long* __restrict__ curB) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    KT* stage = reinterpret_cast<KT*>(ss_smem);
    uint8_t* ssub = reinterpret_cast<uint8_t*>(stage + SUB_TILE);
    __shared__ unsigned hist[NB];
    __shared__ unsigned off[NB];
    __shared__ unsigned long long gdelta[NB];
    const unsigned tid = threadIdx.x;
    const unsigned t = blockIdx.x;
    if (t >= n_tiles) return;
    const int bin = bin_of_tile(tile_prefix, n_bins, t);
    const unsigned long long start = offA[bin] + (unsigned long long)(t - tile_prefix[bin]) * SUB_TILE;
    const unsigned long long lim = offA[bin + 1];
    const unsigned n = (unsigned)(start + SUB_TILE < lim ? SUB_TILE : lim - start);
    if (tid < NB) hist[tid] = 0;
    __syncthreads();
    KT key[SS_ITEMS];
    uint32_t meta[SS_ITEMS];
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        const unsigned i = j * SS_THREADS + tid;
        key[j] = i < n ? in[start + i] : KT(0);
    }
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        const unsigned i = j * SS_THREADS + tid;
        meta[j] = 0xffffffffu;
        if (i < n) {
            const unsigned sub = (unsigned)(key[j] >> shiftB) & maskB;
            meta[j] = (sub << 16) | atomicAdd(&hist[sub], 1u);
        }
    }
    __syncthreads();
    const unsigned c = tid < NB ? hist[tid] : 0u;
    const uint64_t ex = block_exclusive_scan<SS_THREADS>(c, nullptr);
    if (tid < NB) {
        const unsigned long long g = c ? atomicAdd(&curB[(size_t)bin * NB + tid], (unsigned long long)c) : 0ull;
        off[tid] = (unsigned)ex;
        gdelta[tid] = g - ex;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        if (meta[j] != 0xffffffffu) {
            const unsigned sub = meta[j] >> 16;
            const unsigned pos = off[sub] + (meta[j] & 0xffffu);
            stage[pos] = key[j];
            ssub[pos] = (uint8_t)sub;
        }
    }
    __syncthreads();
    for (unsigned i = tid; i < n; i += SS_THREADS) {
        const unsigned sub = ssub[i];
        out[gdelta[sub] + i] = stage[i];
    }
}

// ------------------------------------------------------------------ buckets -> rows
constexpr int BC_THREADS = 256;
constexpr int BC_WARPS = BC_THREADS / 32;
constexpr int BC_UNROLL = 4;

__device__ __forceinline__ uint32_t fold32(uint32_t x) { return x; }
__device__ __forceinline__ uint32_t fold32(uint64_t x) { return (uint32_t)x ^ ((uint32_t)(x >> 32) * 0x9E3779B1u); }

__device__ __forceinline__ uint32_t cas_key(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t cas_key(uint64_t* p, uint64_t cmp, uint64_t v) {
    return atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)v);
}

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t* a, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// strand != 0: rows (K, c) and (revcomp K, c), c = the reference's count (SURVEY.md Appendix
//              A.3): with I instances and C windows of palindromic reads among them (corr
//              table), H = 2I - C half-units, c = H if K == revcomp(K) else H / 2.
// strand == 0: rows (K, I).
// counters[0] = rows appended (keeps counting past cap so the caller can retry with room),
// counters[1] += distinct canonical k-mers.  err is raised if a bucket cannot be split further.
template <typename KT>
__global__ void __launch_bounds__(BC_THREADS) k_bucket_count(const KT* __restrict__ keys,
                                                            const unsigned long long* __restrict__ offB,
                                                            unsigned n_buckets, int k, int a, int b, int log_slots,
                                                            int strand, const uint64_t* __restrict__ corr_key,
                                                            const uint32_t* __restrict__ corr_cnt, uint64_t n_corr,
                                                            uint64_t* __restrict__ row_key,
                                                            uint32_t* __restrict__ row_val, uint64_t cap,
                                                            unsigned long long* __restrict__ counters,
                                                            unsigned* __restrict__ ticket, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char bc_smem[];
    const unsigned slots = 1u << log_slots;
    KT* tkey = reinterpret_cast<KT*>(bc_smem);
    uint32_t* tcnt = reinterpret_cast<uint32_t*>(tkey + slots);
    __shared__ unsigned s_bucket, s_distinct, s_overflow;
    __shared__ unsigned w_nh[BC_WARPS], w_nr[BC_WARPS];
    __shared__ unsigned long long s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    const int nbits = 2 * k, rs = nbits - a, r = rs - b;
    const KT EMPTY = ~KT(0);  // r < bits of KT, so no key is all ones
    const KT maskR = (KT)(((KT)1 << r) - 1);
    const unsigned limit = slots - (slots >> 2);
    const unsigned per_warp = slots / BC_WARPS;  // >= 32: log_slots >= 8

    for (;;) {
        if (tid == 0) s_bucket = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned bk = s_bucket;
        if (bk >= n_buckets) break;
        const unsigned long long lo = offB[bk], hi = offB[bk + 1];
        if (hi == lo) {
            __syncthreads();
            continue;
        }
        const uint64_t hbase = ((uint64_t)(bk >> 8) << rs) | ((uint64_t)(bk & 255u) << r);
        // The keys of the bucket are processed in `R` rounds, round `rd` taking the keys whose
        // round hash is rd modulo R.  R is 1 unless the table overflows; an overflowing round
        // (R, rd) is replaced by (2R, rd) and (2R, rd + R).
        unsigned R = 1, rd = 0;
        for (;;) {
            for (unsigned i = tid; i < slots; i += BC_THREADS) {
                tkey[i] = EMPTY;
                tcnt[i] = 0;
            }
            if (tid == 0) {
                s_distinct = 0;
                s_overflow = 0;
            }
            __syncthreads();
            for (unsigned long long base = lo; base < hi; base += BC_THREADS * BC_UNROLL) {
                if (__any_sync(FULL, *(volatile unsigned*)&s_overflow != 0)) break;
                KT x[BC_UNROLL];
                bool v[BC_UNROLL];
#pragma unroll
                for (int u = 0; u < BC_UNROLL; ++u) {
                    const unsigned long long i = base + (unsigned long long)u * BC_THREADS + tid;
                    v[u] = i < hi;
                    x[u] = v[u] ? (KT)(keys[i] & maskR) : EMPTY;
                }
#pragma unroll
                for (int u = 0; u < BC_UNROLL; ++u) {
                    bool want = v[u];
                    if (R > 1) want = want && (((fold32(x[u]) * 0x85EBCA77u) >> 10) & (R - 1)) == rd;
                    // a warp full of one k-mer (repeats, low-complexity sequence): one update
                    const KT x0 = __shfl_sync(FULL, x[u], 0);
                    unsigned add = 1;
                    if (__all_sync(FULL, want && x[u] == x0)) {
                        want = lane == 0;
                        add = 32;
                    }
                    if (want) {
                        unsigned slot = ((fold32(x[u]) * 0x9E3779B1u) >> (32 - log_slots)) & (slots - 1);
                        unsigned probes = 0;
                        for (;;) {
                            const KT old = cas_key(&tkey[slot], EMPTY, x[u]);
                            if (old == EMPTY) {
                                if (atomicAdd(&s_distinct, 1u) >= limit) s_overflow = 1;
                            }
                            if (old == EMPTY || old == x[u]) {
                                atomicAdd(&tcnt[slot], add);
                                break;
                            }
                            slot = (slot + 1) & (slots - 1);
                            if (++probes > slots) {
                                s_overflow = 1;
                                break;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (s_overflow) {
                __syncthreads();  // everybody has seen the flag before it is cleared again
                if (R >= (1u << 20)) {
                    if (tid == 0) atomicExch(err, 2);
                    break;
                }
                R <<= 1;  // rd stays: the first half of the round that overflowed
                continue;
            }
            // ---- rows of this round: count per warp, reserve once per block, write ----
            const unsigned s_lo = warp * per_warp, s_hi = s_lo + per_warp;
            unsigned nh = 0, nr = 0;
            for (unsigned s0 = s_lo; s0 < s_hi; s0 += 32) {
                const KT x = tkey[s0 + lane];
                const bool occ = x != EMPTY;
                nh += __popc(__ballot_sync(FULL, occ));
                if (strand) {
                    const uint64_t K = kmer_unhash(hbase | (uint64_t)x, nbits);
                    nr += __popc(__ballot_sync(FULL, occ && kmer_revcomp(K, k) != K));
                }
            }
            if (lane == 0) {
                w_nh[warp] = nh;
                w_nr[warp] = nr;
            }
            __syncthreads();
            if (tid == 0) {
                unsigned rows = 0, heads = 0;
                for (int w = 0; w < BC_WARPS; ++w) {
                    const unsigned cw = w_nh[w] + w_nr[w];
                    heads += w_nh[w];
                    w_nr[w] = rows;  // becomes the warp's offset inside the block's reservation
                    rows += cw;
                }
                s_base = rows ? atomicAdd(&counters[0], (unsigned long long)rows) : 0ull;
                if (heads) atomicAdd(&counters[1], (unsigned long long)heads);
            }
            __syncthreads();
            unsigned long long of = s_base + w_nr[warp];
            unsigned long long orv = of + nh;
            for (unsigned s0 = s_lo; s0 < s_hi; s0 += 32) {
                const KT x = tkey[s0 + lane];
                const bool occ = x != EMPTY;
                uint64_t K = 0, rcK = 0;
                uint32_t val = 0;
                bool rev = false;
                if (occ) {
                    K = kmer_unhash(hbase | (uint64_t)x, nbits);
                    const uint64_t inst = tcnt[s0 + lane];
                    if (strand) {
                        rcK = kmer_revcomp(K, k);
                        rev = rcK != K;
                        uint64_t corr = 0;
                        if (n_corr) {
                            const uint64_t q = lower_bound_u64(corr_key, n_corr, K);
                            if (q < n_corr && corr_key[q] == K) corr = corr_cnt[q];
                        }
                        const uint64_t Hh = 2 * inst - corr;
                        uint64_t c = rev ? (Hh >> 1) : Hh;
                        if (c > 0x7fffffffull) c = 0x7fffffffull;
                        val = (uint32_t)c;
                    } else {
                        val = (uint32_t)inst;
                    }
                }
                const unsigned hm = __ballot_sync(FULL, occ);
                if (occ) {
                    const unsigned long long pf = of + __popc(hm & lt);
                    if (pf < cap) {
                        row_key[pf] = K;
                        row_val[pf] = val;
                    }
                }
                of += __popc(hm);
                if (strand) {
                    const unsigned rm = __ballot_sync(FULL, rev);
                    if (rev) {
                        const unsigned long long pr = orv + __popc(rm & lt);
                        if (pr < cap) {
                            row_key[pr] = rcK;
                            row_val[pr] = val;
                        }
                    }
                    orv += __popc(rm);
                }
            }
            __syncthreads();  // the table is cleared next
            // next round: the sibling of this one, or of the closest ancestor that is a first half
            while (R > 1 && rd >= (R >> 1)) {
                rd -= R >> 1;
                R >>= 1;
            }
            if (R == 1) break;
            rd += R >> 1;
        }
    }
}

template <typename KT, int NT>
void launch_part_scatter_nt(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks,
                            unsigned long long* cur, const PeerPtrs& dst, cudaStream_t st) {
    constexpr int RPW = KeyBudget<KT>::ITEMS / NT > 0 ? KeyBudget<KT>::ITEMS / NT : 1;
    constexpr int TILE = PT_THREADS * NT * RPW;
    constexpr size_t SMEM = (size_t)TILE * (sizeof(KT) + 1);
    static bool attr_set = false;
    if (!attr_set) {
        PEDK_CUDA_TRY(cudaFuncSetAttribute(k_part_scatter<KT, NT, RPW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)SMEM));
        attr_set = true;
    }
    const uint64_t tiles = (n_reads + (uint64_t)PT_WARPS * RPW - 1) / ((uint64_t)PT_WARPS * RPW);
    const unsigned grid = (unsigned)(tiles < (uint64_t)kNumSMs * 2 ? tiles : (uint64_t)kNumSMs * 2);
    k_part_scatter<KT, NT, RPW><<<grid, PT_THREADS, SMEM, st>>>(words, W, L, k, n_reads, a, nranks, cur, dst);
}

template <typename KT>
void launch_part_scatter_t(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks,
                           unsigned long long* cur, const PeerPtrs& dst, cudaStream_t st) {
    const int nt = (L - k + 1 + 31) / 32;
#define PEDK_PS(N) launch_part_scatter_nt<KT, N>(words, W, L, k, n_reads, a, nranks, cur, dst, st)
    if (nt <= 1) PEDK_PS(1);
    else if (nt <= 2) PEDK_PS(2);
    else if (nt <= 3) PEDK_PS(3);
    else if (nt <= 4) PEDK_PS(4);
    else if (nt <= 5) PEDK_PS(5);
    else if (nt <= 6) PEDK_PS(6);
    else if (nt <= 8) PEDK_PS(8);
    else if (nt <= 12) PEDK_PS(12);
    else PEDK_PS(16);
#undef PEDK_PS
}

}  // namespace

void launch_part_count(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a,
                       unsigned long long* histA, cudaStream_t st) {
    if (!n_reads || L < k) return;
    const uint64_t blocks = (n_reads + 7) / 8;
    const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_part_count<<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_part_scatter(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks, int wide,
                         unsigned long long* cur, const PeerPtrs& dst, cudaStream_t st) {
    if (!n_reads || L < k) return;
    if (wide)
        launch_part_scatter_t<uint64_t>(words, W, L, k, n_reads, a, nranks, cur, dst, st);
    else
        launch_part_scatter_t<uint32_t>(words, W, L, k, n_reads, a, nranks, cur, dst, st);
    PEDK_CUDA_TRY(cudaGetLastError());
}

unsigned sub_tile_keys() { return SUB_TILE; }

void launch_sub_hist(const void* keys, int wide, const unsigned long long* offA, const unsigned* tile_prefix,
                     int n_bins, unsigned n_tiles, int k, int a, int b, unsigned* histB, cudaStream_t st) {
    if (!n_tiles) return;
    const int shiftB = 2 * k - a - b;
    const unsigned maskB = (1u << b) - 1;
    const unsigned grid = (n_tiles + SH_GROUP - 1) / SH_GROUP;
    if (wide)
        k_sub_hist<uint64_t><<<grid, SH_THREADS, 0, st>>>(static_cast<const uint64_t*>(keys), offA, tile_prefix,
                                                          n_bins, n_tiles, shiftB, maskB, histB);
    else
        k_sub_hist<uint32_t><<<grid, SH_THREADS, 0, st>>>(static_cast<const uint32_t*>(keys), offA, tile_prefix,
                                                          n_bins, n_tiles, shiftB, maskB, histB);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_sub_scatter(const void* in, void* out, int wide, const unsigned long long* offA,
                        const unsigned* tile_prefix, int n_bins, unsigned n_tiles, int k, int a, int b,
                        unsigned long long* curB, cudaStream_t st) {
    if (!n_tiles) return;
    const int shiftB = 2 * k - a - b;
    const unsigned maskB = (1u << b) - 1;
    if (wide) {
        constexpr size_t SMEM = (size_t)SUB_TILE * (sizeof(uint64_t) + 1);
        static bool attr_set = false;
        if (!attr_set) {
            PEDK_CUDA_TRY(cudaFuncSetAttribute(k_sub_scatter<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)SMEM));
            attr_set = true;
        }
        k_sub_scatter<uint64_t><<<n_tiles, SS_THREADS, SMEM, st>>>(static_cast<const uint64_t*>(in),
                                                                   static_cast<uint64_t*>(out), offA, tile_prefix,
                                                                   n_bins, n_tiles, shiftB, maskB, curB);
    } else {
        constexpr size_t SMEM = (size_t)SUB_TILE * (sizeof(uint32_t) + 1);
        k_sub_scatter<uint32_t><<<n_tiles, SS_THREADS, SMEM, st>>>(static_cast<const uint32_t*>(in),
                                                                   static_cast<uint32_t*>(out), offA, tile_prefix,
                                                                   n_bins, n_tiles, shiftB, maskB, curB);
    }
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_bucket_count(const void* keys, int wide, const unsigned long long* offB, unsigned n_buckets, int k, int a,
                         int b, int log_slots, int strand, const uint64_t* corr_key, const uint32_t* corr_cnt,
                         uint64_t n_corr, uint64_t* row_key, uint32_t* row_val, uint64_t cap,
                         unsigned long long* counters, unsigned* ticket, int* err, cudaStream_t st) {
    if (!n_buckets) return;
    if (log_slots < 8) log_slots = 8;
    if (log_slots > 13) log_slots = 13;
    const size_t smem = ((size_t)1 << log_slots) * ((wide ? sizeof(uint64_t) : sizeof(uint32_t)) + sizeof(uint32_t));
    PEDK_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    // enough resident CTAs to fill the machine; buckets are handed out by ticket
    const unsigned per_sm = (unsigned)(200 * 1024 / (smem + 1024));
    unsigned grid = kNumSMs * (per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm);
    if (grid > n_buckets) grid = n_buckets;
    if (wide) {
        static bool attr_set = false;
        if (!attr_set) {
            PEDK_CUDA_TRY(cudaFuncSetAttribute(k_bucket_count<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(((size_t)1 << 13) * 12)));
            attr_set = true;
        }
        k_bucket_count<uint64_t><<<grid, BC_THREADS, smem, st>>>(static_cast<const uint64_t*>(keys), offB, n_buckets,
                                                                 k, a, b, log_slots, strand, corr_key, corr_cnt,
                                                                 n_corr, row_key, row_val, cap, counters, ticket, err);
    } else {
        static bool attr_set = false;
        if (!attr_set) {
            PEDK_CUDA_TRY(cudaFuncSetAttribute(k_bucket_count<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(((size_t)1 << 13) * 8)));
            attr_set = true;
        }
        k_bucket_count<uint32_t><<<grid, BC_THREADS, smem, st>>>(static_cast<const uint32_t*>(keys), offB, n_buckets,
                                                                 k, a, b, log_slots, strand, corr_key, corr_cnt,
                                                                 n_corr, row_key, row_val, cap, counters, ticket, err);
    }
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
