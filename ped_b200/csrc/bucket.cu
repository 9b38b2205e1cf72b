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

#include <algorithm>

#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

constexpr int NB = 256;  // bins per level (stride of the bucket index even when 2^b < 256)
constexpr unsigned FULL = 0xffffffffu;


// Canonical k-mers of one read for one lane.  `mine` = word `lane` of the packed read (0 beyond
// the read).  Two layouts of the windows over the lanes:
//   CONTIG  lane l takes windows l*NT .. l*NT+NT-1: ONE 32-base window is cut out of the words
//           (two shuffles, one funnel shift) and reverse-complemented once; the NT k-mers and
//           their reverse complements are shifts of those two words.  Needs NT - 1 + k <= 32.
//   else    lane l takes windows l, l+32, ...: every k-mer is cut out and reversed on its own.
// The alu pipe (shifts, logic) is what bounds these kernels, so CONTIG is ~2x faster.
template <int NT, bool CONTIG>
__device__ __forceinline__ void lane_kmers(uint64_t mine, unsigned lane, const int k, int nwin, bool live,
                                           uint64_t (&canon)[NT], bool (&valid)[NT]) {
    const int kshift = 64 - 2 * k;
    if (CONTIG) {
        const int p = (int)lane * NT;
        const int w = p >> 5, o = p & 31;
        const uint64_t hi = __shfl_sync(FULL, mine, w);
        const uint64_t lo = __shfl_sync(FULL, mine, w + 1);
        const uint64_t x = o ? ((hi << (2 * o)) | (lo >> (64 - 2 * o))) : hi;
        const uint64_t rcx = revcomp_groups64(x);
        const uint64_t mask = kshift ? (~0ull >> kshift) : ~0ull;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const uint64_t km = (x >> (kshift - 2 * j)) & mask;
            const uint64_t rc = (rcx >> (2 * j)) & mask;
            canon[j] = km < rc ? km : rc;
            valid[j] = live && p + j < nwin;
        }
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const uint64_t hi = __shfl_sync(FULL, mine, t);
            const uint64_t lo = __shfl_sync(FULL, mine, t + 1);
            const uint64_t x = lane ? ((hi << (2 * lane)) | (lo >> (64 - 2 * lane))) : hi;
            const uint64_t km = x >> kshift;
            const uint64_t rc = revcomp_groups64(km) >> kshift;
            canon[t] = km < rc ? km : rc;
            valid[t] = live && 32 * t + (int)lane < nwin;
        }
    }
}

// ------------------------------------------------------------------ level A: count
constexpr int PC_THREADS = 256;

// KC = k when it is known at compile time (20, the reference's only k), else 0: with a constant k
// the shifts and masks of the extraction and of kmer_hash fold into immediates.
template <int NT, bool CONTIG, int KC>
__global__ void __launch_bounds__(PC_THREADS) k_part_count(const uint64_t* __restrict__ words, int W, int L, int k_rt,
                                                          uint64_t n_reads, int shiftA_rt,
                                                          unsigned long long* __restrict__ histA) {
    const int k = KC ? KC : k_rt;
    const int shiftA = KC ? 2 * KC - 8 : shiftA_rt;  // KC >= 8: eight level-A bits
    __shared__ unsigned h[NB];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * PC_THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * PC_THREADS) >> 5;
    const int nwin = L - k + 1, nbits = 2 * k;
    for (uint64_t r = warp; r < n_reads; r += n_warps) {
        const uint64_t mine = (int)lane < W ? words[r * (uint64_t)W + lane] : 0ull;
        uint64_t canon[NT];
        bool valid[NT];
        lane_kmers<NT, CONTIG>(mine, lane, k, nwin, true, canon, valid);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const unsigned bin = (unsigned)(kmer_hash(canon[t], nbits) >> shiftA);
            if (valid[t]) atomicAdd(&h[bin], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&histA[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// ------------------------------------------------------------------ level A: scatter
// 512 threads x 20 keys per tile (two CTAs per SM); on several ranks 1024 x 20, one CTA per SM:
// the runs a tile stores over NVLink are twice as long (~320 bytes per bin)

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

// a 1024-thread tile must keep its ranks in 16 bits and its staging area in shared memory
template <typename KT, int NT>
struct PT_BIG_OK {
    static constexpr int RPW = KeyBudget<KT>::ITEMS / NT > 0 ? KeyBudget<KT>::ITEMS / NT : 1;
    static constexpr bool value = 1024 * NT * RPW <= 65536 && (size_t)1024 * NT * RPW * (sizeof(KT) + 1) <= 200 * 1024;
};

// NT = windows per lane per read (ceil((L-k+1)/32) rounded up to an instantiated value),
// RPW = reads per warp per tile.
// BULK: a bin's run leaves shared memory as ONE bulk copy (cp.async.bulk shared::cta -> global, the
// TMA unit) issued by one thread instead of element-wise stores by all: the copy-out costs no issue
// slots, needs no per-element bin array, and drains -- to the owner's HBM over NVLink on several
// ranks -- while the CTA already extracts and hashes its next tile (the staging area is only
// written again after the rank phase, so one buffer is enough).  Bulk copies want 16-byte aligned
// source, destination and size: a run is staged at the same misalignment as its destination, its
// aligned body goes by bulk copy, the <= 3 elements before and after by plain stores.
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename KT, int NT, int RPW, bool CONTIG, int KC, int PT_THREADS, bool BULK>
__global__ void __launch_bounds__(PT_THREADS, PT_THREADS > 512 ? 1 : 2)
    k_part_scatter(const uint64_t* __restrict__ words, int W, int L, int k_rt, uint64_t n_reads, int a_rt, int nranks,
                   unsigned long long* __restrict__ cur, const unsigned long long* __restrict__ lim,
                   int* __restrict__ overflow, PeerPtrs dst, unsigned pass_mask, unsigned pass_id) {
    const int k = KC ? KC : k_rt;
    const int a = KC ? 8 : a_rt;
    constexpr int PT_WARPS = PT_THREADS / 32;
    constexpr int ITEMS = NT * RPW;
    constexpr int TILE = PT_THREADS * ITEMS;
    constexpr unsigned E16 = 16 / sizeof(KT);  // elements per 16 bytes
    static_assert(TILE <= 65536, "ranks are kept in 16 bits");
    extern __shared__ __align__(16) unsigned char pt_smem[];
    KT* stage = reinterpret_cast<KT*>(pt_smem);
    uint8_t* sbin = reinterpret_cast<uint8_t*>(stage + TILE);  // (!BULK)
    __shared__ unsigned hist[NB];
    __shared__ unsigned off[NB];
    __shared__ KT* s_out[NB];  // per tile: where stage[0] would go if the whole tile were this bin's run
    __shared__ KT* s_ptr[NB];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int nwin = L - k + 1, nbits = 2 * k;
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
            uint64_t canon[NT];
            bool valid[NT];
            lane_kmers<NT, CONTIG>(mine, lane, k, nwin, live, canon, valid);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int idx = rr * NT + t;
                const uint64_t h = kmer_hash(canon[t], nbits);
                const unsigned bin = (unsigned)(h >> shiftA);
                keyv[idx] = (KT)(h & maskS);
                meta[idx] = 0xffffffffu;
                // counting in several passes (HBM capacity): only the bins of this pass are kept
                if (valid[t] && (bin & pass_mask) == pass_id) meta[idx] = (bin << 16) | atomicAdd(&hist[bin], 1u);
            }
        }
        __syncthreads();
        // ---- bin offsets inside the tile; one global atomic per (tile, bin) ----
        const unsigned c = tid < NB ? hist[tid] : 0u;
        if (BULK) {
            unsigned long long g = 0;
            bool fits = true;
            if (c) {
                g = atomicAdd(&cur[tid], (unsigned long long)c);
                // optimistic sizing (lim != null): a run that does not fit its segment is dropped and
                // the flag raised -- the caller then sizes the bins exactly and runs again
                fits = !lim || g + c <= lim[tid];
                if (!fits) *overflow = 1;
            }
            const unsigned mis = (unsigned)g & (E16 - 1);  // the run starts this far into a 16-byte line of its destination
            const unsigned need = c ? (mis + c + E16 - 1) & ~(E16 - 1) : 0u;
            const uint64_t ex = block_exclusive_scan<PT_THREADS>(need, nullptr);
            if (tid < NB) {
                off[tid] = (unsigned)ex + mis;
                s_out[tid] = (fits && c) ? s_ptr[tid] + g : nullptr;
                bulk_wait_read();  // the copies of the previous tile have read the staging area
            }
            __syncthreads();
#pragma unroll
            for (int idx = 0; idx < ITEMS; ++idx)
                if (meta[idx] != 0xffffffffu) stage[off[meta[idx] >> 16] + (meta[idx] & 0xffffu)] = keyv[idx];
            fence_async_smem();  // the staged keys are visible to the bulk-copy engine
            __syncthreads();
            if (tid < NB && s_out[tid]) {
                KT* p = s_out[tid];
                const KT* sp = stage + off[tid];
                const unsigned head = c < ((E16 - mis) & (E16 - 1)) ? c : ((E16 - mis) & (E16 - 1));
                const unsigned body = (c - head) & ~(E16 - 1);
                for (unsigned i = 0; i < head; ++i) p[i] = sp[i];
                if (body) bulk_store(p + head, sp + head, body * (unsigned)sizeof(KT));
                for (unsigned i = head + body; i < c; ++i) p[i] = sp[i];
                bulk_commit();
            }
        } else {
            uint64_t total;
            const uint64_t ex = block_exclusive_scan<PT_THREADS>(c, &total);
            if (tid < NB) {
                const unsigned long long g = c ? atomicAdd(&cur[tid], (unsigned long long)c) : 0ull;
                off[tid] = (unsigned)ex;
                const bool fits = !lim || g + c <= lim[tid];
                if (!fits) *overflow = 1;
                s_out[tid] = fits ? s_ptr[tid] + (g - ex) : nullptr;
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
                KT* p = s_out[sbin[i]];
                if (p) p[i] = stage[i];
            }
        }
    }
    if (BULK && tid < NB) bulk_wait_all();  // shared memory must outlive the copies that read it
}

// ------------------------------------------------------------------ level B
// keys per tile of the level-B passes (a tile lies inside one segment of a level-A bin) and the CTA
// shape of the scatter (threads x keys per thread).  The kernel is bound by shared-memory latency
// between its barriers, not by issue slots or HBM.  Measured at configs[1], 32-bit keys: 256 x 16 (four
// CTAs per SM, 64 registers) 14.9 ms/step, 512 x 16 14.8, 512 x 8 (32 registers, 2048 threads per SM)
// 14.2, 256 x 8 (tiles of 2048 keys: runs too short) 22.6; 64-bit keys (72 bytes of spills at 32
// registers): 256 x 16 3.7, 512 x 8 3.95.  Default: 512 x 8 for 32-bit keys, 256 x 16 for 64-bit keys --
// both are tiles of 4096 keys.  PEDK_SUB_THREADS / PEDK_SUB_ITEMS force a shape for both.
constexpr int SH_THREADS = 256;
constexpr int SH_GROUP_KEYS = 65536;  // keys per CTA of the histogram kernel
bool sub_scatter_forced() { return getenv("PEDK_SUB_THREADS") || getenv("PEDK_SUB_ITEMS"); }
int sub_scatter_threads() {
    const int t = getenv("PEDK_SUB_THREADS") ? atoi(getenv("PEDK_SUB_THREADS")) : 256;
    return t >= 512 ? 512 : 256;
}
int sub_scatter_items() {
    const int t = getenv("PEDK_SUB_ITEMS") ? atoi(getenv("PEDK_SUB_ITEMS")) : 16;
    return t == 8 ? 8 : 16;
}

// A level-A bin arrives as one or more SEGMENTS (one per source rank, or one per bin when the
// bins were sized exactly): seg[0][s] = first element, seg[1][s] = end, seg[2][s] = bin.
// largest segment with tile_prefix[seg] <= t (segments without tiles are never returned)
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
                                                        const unsigned long long* __restrict__ seg,
                                                        const unsigned* __restrict__ tile_prefix, int n_bins,
                                                        unsigned n_tiles, unsigned sub_tile, int shiftB,
                                                        unsigned maskB, unsigned* __restrict__ histB) {
    __shared__ unsigned h[NB];
    const unsigned tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    int cur_bin = -1;
    const unsigned group = SH_GROUP_KEYS / sub_tile;
    const unsigned t0 = blockIdx.x * group;
    const unsigned t1 = t0 + group < n_tiles ? t0 + group : n_tiles;
    for (unsigned t = t0; t < t1; ++t) {
        const int sg = bin_of_tile(tile_prefix, n_bins, t);
        const int bin = (int)seg[2 * (size_t)n_bins + sg];
        if (bin != cur_bin) {
            if (cur_bin >= 0) {
                __syncthreads();
                if (h[tid]) atomicAdd(&histB[(size_t)cur_bin * NB + tid], h[tid]);
                h[tid] = 0;
                __syncthreads();
            }
            cur_bin = bin;
        }
        const unsigned long long start = seg[sg] + (unsigned long long)(t - tile_prefix[sg]) * sub_tile;
        const unsigned long long lim = seg[(size_t)n_bins + sg];
        const unsigned long long end = start + sub_tile < lim ? start + sub_tile : lim;
        for (unsigned long long i = start + tid; i < end; i += SH_THREADS)
            atomicAdd(&h[(unsigned)(keys[i] >> shiftB) & maskB], 1u);
    }
    __syncthreads();
    if (cur_bin >= 0 && h[tid]) atomicAdd(&histB[(size_t)cur_bin * NB + tid], h[tid]);
}

// One tile of a level-A bin -> its 2^b sub-bins.  WHOLE: the tile holds SUB_TILE keys (all but
// the last tile of a bin), so no index is tested.  A thread keeps its 16 keys and their ranks
// (two 16-bit ranks per register; the sub-bin is recomputed from the key) between the phases.
template <typename KT, bool WHOLE, int SS_THREADS, int SS_ITEMS>
__device__ __forceinline__ void sub_scatter_tile(const KT* __restrict__ src, unsigned n, KT* __restrict__ out,
                                                 unsigned long long* __restrict__ cur,
                                                 const unsigned long long* __restrict__ lim, int* __restrict__ overflow,
                                                 int shiftB, unsigned maskB, KT* stage, unsigned* hist, unsigned* off,
                                                 KT** s_out) {
    const unsigned tid = threadIdx.x;
    KT key[SS_ITEMS];
    uint32_t rank2[SS_ITEMS / 2];
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        const unsigned i = j * SS_THREADS + tid;
        key[j] = (WHOLE || i < n) ? src[i] : KT(0);
    }
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        const unsigned i = j * SS_THREADS + tid;
        unsigned r = 0xffffu;
        if (WHOLE || i < n) r = atomicAdd(&hist[(unsigned)(key[j] >> shiftB) & maskB], 1u);
        if (j & 1) rank2[j / 2] |= r << 16;
        else rank2[j / 2] = r;
    }
    __syncthreads();
    const unsigned c = tid < NB ? hist[tid] : 0u;
    const uint64_t ex = block_exclusive_scan<SS_THREADS>(c, nullptr);
    if (tid < NB) {
        const unsigned long long g = c ? atomicAdd(&cur[tid], (unsigned long long)c) : 0ull;
        off[tid] = (unsigned)ex;
        // buckets sized optimistically (lim != null): a run that does not fit is dropped and the flag
        // raised -- the caller then sizes the buckets exactly (k_sub_hist) and scatters again
        const bool fits = !lim || g + c <= lim[tid];
        if (!fits) *overflow = 1;
        s_out[tid] = fits ? out + (g - ex) : nullptr;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SS_ITEMS; ++j) {
        const unsigned r = (j & 1) ? rank2[j / 2] >> 16 : rank2[j / 2] & 0xffffu;
        if (WHOLE || r != 0xffffu) {
            const unsigned sub = (unsigned)(key[j] >> shiftB) & maskB;
            stage[off[sub] + r] = key[j];
        }
    }
    __syncthreads();
    // the sub-bin of a staged key is in the key itself: no side array (shared-memory wavefronts,
    // not HBM, bound this kernel)
    if (WHOLE) {
#pragma unroll
        for (int j = 0; j < SS_ITEMS; ++j) {
            const unsigned i = j * SS_THREADS + tid;
            const KT kk = stage[i];
            KT* p = s_out[(unsigned)(kk >> shiftB) & maskB];
            if (p) p[i] = kk;
        }
    } else {
        for (unsigned i = tid; i < n; i += SS_THREADS) {
            const KT kk = stage[i];
            KT* p = s_out[(unsigned)(kk >> shiftB) & maskB];
            if (p) p[i] = kk;
        }
    }
}

template <typename KT, int SS_THREADS, int SS_ITEMS>
__global__ void __launch_bounds__(SS_THREADS, 16384 / (SS_THREADS * SS_ITEMS)) k_sub_scatter(const KT* __restrict__ in, KT* __restrict__ out,
                                                              const unsigned long long* __restrict__ seg,
                                                              const unsigned* __restrict__ tile_prefix, int n_bins,
                                                              unsigned n_tiles, int shiftB, unsigned maskB,
                                                              unsigned long long* __restrict__ curB,
                                                              const unsigned long long* __restrict__ limB,
                                                              int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    KT* stage = reinterpret_cast<KT*>(ss_smem);
    __shared__ unsigned hist[NB];
    __shared__ unsigned off[NB];
    __shared__ KT* s_out[NB];
    const unsigned tid = threadIdx.x;
    const unsigned t = blockIdx.x;
    if (t >= n_tiles) return;
    const int sg = bin_of_tile(tile_prefix, n_bins, t);
    const int bin = (int)seg[2 * (size_t)n_bins + sg];
    constexpr unsigned SUB_TILE = SS_THREADS * SS_ITEMS;
    const unsigned long long start = seg[sg] + (unsigned long long)(t - tile_prefix[sg]) * SUB_TILE;
    const unsigned long long lim = seg[(size_t)n_bins + sg];
    const unsigned n = (unsigned)(start + SUB_TILE < lim ? SUB_TILE : lim - start);
    if (tid < NB) hist[tid] = 0;
    __syncthreads();
    if (n == SUB_TILE)
        sub_scatter_tile<KT, true, SS_THREADS, SS_ITEMS>(in + start, n, out, curB + (size_t)bin * NB,
                                               limB ? limB + (size_t)bin * NB : nullptr, overflow, shiftB, maskB, stage,
                                               hist, off, s_out);
    else
        sub_scatter_tile<KT, false, SS_THREADS, SS_ITEMS>(in + start, n, out, curB + (size_t)bin * NB,
                                                limB ? limB + (size_t)bin * NB : nullptr, overflow, shiftB, maskB, stage,
                                                hist, off, s_out);
}

// ------------------------------------------------------------------ buckets -> rows
// CTA shapes of k_bucket_count: 1024 threads x 2 CTAs per SM (32 registers per thread) or
// 512 x 2 (64 registers: fewer re-computed loop invariants, half the warps); PEDK_BC_THREADS picks
constexpr int BC_MAX_WARPS = 32;
constexpr int BC_UNROLL = 4;  // (the slow path of k_bucket_count selects among exactly four keys)

__device__ __forceinline__ uint32_t fold32(uint32_t x) { return x; }
__device__ __forceinline__ uint32_t fold32(uint64_t x) { return (uint32_t)x ^ ((uint32_t)(x >> 32) * 0x9E3779B1u); }

__device__ __forceinline__ uint32_t cas_key(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t cas_key(uint64_t* p, uint64_t cmp, uint64_t v) {
    return atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)v);
}

// the GS keys of one 16-byte group of the table, read in one go and never from a register copy
__device__ __forceinline__ void load_group(const uint32_t* p, uint32_t (&q)[4]) {
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3])
                 : "r"((unsigned)__cvta_generic_to_shared(p))
                 : "memory");
}
__device__ __forceinline__ void load_group(const uint64_t* p, uint64_t (&q)[2]) {
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];"
                 : "=l"(q[0]), "=l"(q[1])
                 : "r"((unsigned)__cvta_generic_to_shared(p))
                 : "memory");
}

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t* a, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// 16 bytes of keys in one load (the instance stream is read with 128-bit loads)
__device__ __forceinline__ void load_keys16(const uint32_t* p, uint32_t (&x)[4]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ void load_keys16(const uint64_t* p, uint64_t (&x)[2]) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    x[0] = v.x; x[1] = v.y;
}

// tcnt[slot] += 1 if p: one predicated shared-memory reduction, no branch
__device__ __forceinline__ void red_add_if(uint32_t* addr, bool p) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "@q red.shared.add.u32 [%0], 1;\n\t"
        "}"
        :
        : "r"((unsigned)__cvta_generic_to_shared(addr)), "r"((unsigned)p)
        : "memory");
}

constexpr int BC_QUEUE = 32 + 128;  // per warp: < 32 left over + at most 4 x 32 pushed per iteration

// One key that is not in its home group yet (the first instance of a k-mer, or a home group full
// of other k-mers): probe group by group.  A slot only ever goes EMPTY -> key, so everybody
// agrees on the first free slot of a group.  Returns 1 if the key opened a new slot.
template <typename KT>
__device__ __forceinline__ unsigned slow_insert(KT xs, unsigned add, KT* tkey, uint32_t* tcnt, int gshift,
                                                unsigned gmask, unsigned n_groups, unsigned* s_overflow) {
    constexpr int GS = 16 / sizeof(KT);
    const KT EMPTY = ~KT(0);
    unsigned g = ((fold32(xs) * 0x9E3779B1u) >> gshift) & gmask;
    unsigned probes = 0;
    for (;;) {
        KT q[GS];
        load_group(tkey + g * GS, q);
        int hit = -1, free_slot = -1;
#pragma unroll
        for (int j = GS - 1; j >= 0; --j) {
            if (q[j] == EMPTY) free_slot = j;
            if (q[j] == xs) hit = j;
        }
        unsigned is_new = 0;
        if (hit < 0 && free_slot >= 0) {
            // not here yet and the group has room: claim its FIRST free slot
            const KT old = cas_key(&tkey[g * GS + free_slot], EMPTY, xs);
            if (old == EMPTY) {
                is_new = 1;
                hit = free_slot;
            } else if (old == xs) {
                hit = free_slot;
            } else {
                continue;  // another k-mer took it: look at the group again
            }
        }
        if (hit >= 0) {
            atomicAdd(&tcnt[g * GS + hit], add);
            return is_new;
        }
        g = (g + 1) & gmask;  // group full of other k-mers
        if (++probes > n_groups) {
            *s_overflow = 1;
            return 0;
        }
    }
}

// Insert the keys [lo, hi) of `keys` into the table (round rd of R, see k_bucket_count).
//  (1) dense: every lane looks at the home group of each key of its 16-byte load: a k-mer that is
//      already there (~25 of 26 instances at 30x) costs a vector load of the group, GS compares
//      and one predicated fire-and-forget add -- no divergent loop, no branch per key;
//  (2) the rest goes to the warp's queue in shared memory and is inserted 32 keys at a time, every
//      lane busy (slow_insert).  Handling them where they occur made the whole warp walk the
//      probe loop for one or two lanes: as many instructions again as the dense part.
// New k-mers are counted when a queue is drained; beyond `limit` of them the round is abandoned
// (s_overflow).  wq: this warp's queue (BC_QUEUE keys).
template <typename KT, bool MULTI, int BC_THREADS>
__device__ __forceinline__ void bucket_insert(const KT* __restrict__ keys, unsigned long long lo,
                                              unsigned long long hi, KT* tkey, uint32_t* tcnt, KT* wq, int log_slots,
                                              KT maskR, unsigned R, unsigned rd, unsigned limit, unsigned* s_distinct,
                                              unsigned* s_overflow) {
    constexpr int GS = 16 / sizeof(KT);  // slots per 16-byte group = keys per 16-byte load
    constexpr int GS_LOG = sizeof(KT) == 4 ? 2 : 1;
    const KT EMPTY = ~KT(0);
    const unsigned n_groups = (1u << log_slots) >> GS_LOG;
    const unsigned gmask = n_groups - 1;
    const int gshift = 32 - log_slots + GS_LOG;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const unsigned lt = lanemask_lt();
    // the loads are aligned to 16 bytes: the first and the last vector may reach outside [lo, hi)
    const unsigned long long a0 = lo & ~(unsigned long long)(GS - 1);
    const unsigned long long n_vec = (hi - a0 + GS - 1) / GS;
    unsigned n_q = 0;  // keys in this warp's queue (the same in every lane)
    unsigned my_new = 0;
    auto drain = [&](unsigned n_take) {
        // the last n_take (<= 32) keys of the queue, one per lane
        n_q -= n_take;
        __syncwarp();
        if (lane < n_take) my_new += slow_insert<KT>(wq[n_q + lane], 1u, tkey, tcnt, gshift, gmask, n_groups, s_overflow);
        const unsigned wn = __reduce_add_sync(FULL, my_new);
        my_new = 0;
        if (wn && lane == 0 && atomicAdd(s_distinct, wn) + wn > limit) *s_overflow = 1;
        __syncwarp();
    };
    for (unsigned long long v0 = 0; v0 < n_vec; v0 += BC_THREADS) {
        if (__any_sync(FULL, *(volatile unsigned*)s_overflow != 0)) break;
        const unsigned long long v = v0 + tid;
        KT x[GS];
        if (v < n_vec) {
            load_keys16(keys + a0 + v * GS, x);
            if (v == 0 || v + 1 == n_vec) {
#pragma unroll
                for (int j = 0; j < GS; ++j) {
                    const unsigned long long i = a0 + v * GS + j;
                    x[j] = (i >= lo && i < hi) ? (KT)(x[j] & maskR) : EMPTY;  // keys are <= maskR < EMPTY
                }
            } else {
#pragma unroll
                for (int j = 0; j < GS; ++j) x[j] = (KT)(x[j] & maskR);
            }
        } else {
#pragma unroll
            for (int j = 0; j < GS; ++j) x[j] = EMPTY;
        }
        // a warp full of ONE k-mer (repeats, low-complexity sequence): a single update
        {
            const KT s0 = __shfl_sync(FULL, x[0], 0);
            bool same = x[0] == s0 && s0 != EMPTY;
#pragma unroll
            for (int j = 1; j < GS; ++j) same = same && x[j] == s0;
            if (__all_sync(FULL, same)) {
                bool mine = lane == 0;
                if (MULTI) mine = mine && (((fold32(s0) * 0x85EBCA77u) >> 10) & (R - 1)) == rd;
                if (mine) my_new += slow_insert<KT>(s0, 32u * GS, tkey, tcnt, gshift, gmask, n_groups, s_overflow);
                continue;
            }
        }
        unsigned pm[GS];  // per key: the lanes whose key goes to the queue
#pragma unroll
        for (int j = 0; j < GS; ++j) {
            bool want = x[j] != EMPTY;
            if (MULTI) want = want && (((fold32(x[j]) * 0x85EBCA77u) >> 10) & (R - 1)) == rd;
            const unsigned g = ((fold32(x[j]) * 0x9E3779B1u) >> gshift) & gmask;
            KT q[GS];
            load_group(tkey + g * GS, q);
            // the slot of the group that holds the key, if any: selects and ONE predicated add,
            // no branch (a branch per slot cost more instructions than the compares)
            unsigned off = 0;
            bool hit = q[0] == x[j];
#pragma unroll
            for (int jj = 1; jj < GS; ++jj) {
                const bool h = q[jj] == x[j];
                off = h ? (unsigned)jj : off;
                hit = hit || h;
            }
            hit = hit && want;
            red_add_if(&tcnt[g * GS + off], hit);
            pm[j] = __ballot_sync(FULL, want && !hit);
        }
        // pending keys -> the warp's queue (positions from the ballots: no atomics, no branches)
        {
            unsigned at = n_q;
#pragma unroll
            for (int j = 0; j < GS; ++j) {
                if ((pm[j] >> lane) & 1u) wq[at + __popc(pm[j] & lt)] = x[j];
                at += __popc(pm[j]);
            }
            n_q = at;
        }
        while (n_q >= 32u) drain(32u);
    }
    if (n_q) drain(n_q);
    const unsigned wn = __reduce_add_sync(FULL, my_new);
    if (wn && lane == 0 && atomicAdd(s_distinct, wn) + wn > limit) *s_overflow = 1;
}

// strand != 0: rows (K, c) and (revcomp K, c), c = the reference's count (SURVEY.md Appendix
//              A.3): with I instances and C windows of palindromic reads among them (corr
//              table), H = 2I - C half-units, c = H if K == revcomp(K) else H / 2.
// strand == 0: rows (K, I).
// counters[0] = rows appended (keeps counting past cap so the caller can retry with room),
// counters[1] += distinct canonical k-mers.  err is raised if a bucket cannot be split further.
//
// Shared memory: tkey[slots], tcnt[slots], then one area that holds the warps' queues of pending
// keys while a round inserts and list[slots] (u16: the occupied slots, compacted) afterwards.
// A slot only ever goes EMPTY -> key, so the common case (the k-mer is already in the table:
// ~25 of 26 instances at 30x) is a plain load, a compare and one fire-and-forget atomic add;
// only the first instance of a k-mer pays for a compare-and-swap.
template <typename KT, int BC_THREADS>
__global__ void __launch_bounds__(BC_THREADS, 2) k_bucket_count(const KT* __restrict__ keys,
                                                               const unsigned long long* __restrict__ offB,
                                                               const unsigned long long* __restrict__ endB,
                                                               unsigned n_buckets, int k, int a, int b,
                                                               int log_slots, int strand,
                                                               const uint64_t* __restrict__ corr_key,
                                                               const uint32_t* __restrict__ corr_cnt,
                                                               uint64_t n_corr, uint64_t* __restrict__ row_key,
                                                               uint32_t* __restrict__ row_val, uint64_t cap,
                                                               unsigned long long* __restrict__ counters,
                                                               unsigned* __restrict__ ticket,
                                                               int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char bc_smem[];
    const unsigned slots = 1u << log_slots;
    KT* tkey = reinterpret_cast<KT*>(bc_smem);
    uint32_t* tcnt = reinterpret_cast<uint32_t*>(tkey + slots);
    uint16_t* list = reinterpret_cast<uint16_t*>(tcnt + slots);
    KT* wq = reinterpret_cast<KT*>(tcnt + slots) + (threadIdx.x >> 5) * BC_QUEUE;  // aliases list
    __shared__ unsigned s_bucket, s_distinct, s_overflow, s_nh, s_nr, s_rcur;
    constexpr int BC_WARPS = BC_THREADS / 32;
    __shared__ unsigned w_cnt[BC_MAX_WARPS];
    __shared__ unsigned long long s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    const int nbits = 2 * k, rs = nbits - a, r = rs - b;
    const KT EMPTY = ~KT(0);  // r < bits of KT, so no key is all ones
    const KT maskR = (KT)(((KT)1 << r) - 1);
    const unsigned limit = slots - (slots >> 2);
    const unsigned per_warp = slots / BC_WARPS;  // >= 32: log_slots >= 10
    unsigned r_hint = 1;  // rounds the previous bucket of this CTA needed

    for (;;) {
        if (tid == 0) s_bucket = atomicAdd(ticket, 1u);
        __syncthreads();
        const unsigned bk = s_bucket;
        if (bk >= n_buckets) break;
        const unsigned long long lo = offB[bk], hi = endB[bk];  // (exact sizes: endB = offB + 1)
        if (hi <= lo) {
            __syncthreads();
            continue;
        }
        const uint64_t hbase = ((uint64_t)(bk >> 8) << rs) | ((uint64_t)(bk & 255u) << r);
        // The keys of the bucket are processed in `R` rounds, round `rd` taking the keys whose
        // round hash is rd modulo R.  R is 1 unless the table overflows; an overflowing round
        // (R, rd) is replaced by (2R, rd) and (2R, rd + R).  Buckets of one run are alike (the hash
        // spreads the k-mers evenly), so a bucket starts with the split the last one needed
        // instead of finding it again by overflowing (a genome several times the size the 2^16
        // buckets were laid out for: every bucket needs the same few rounds).
        unsigned R = 1, rd = 0, r_used = 1, d_max = 0;
        for (;;) {
            while (R < r_hint) R <<= 1;  // descend to the first half, as an overflow would
            for (unsigned i = tid; i < slots; i += BC_THREADS) {
                tkey[i] = EMPTY;
                tcnt[i] = 0;
            }
            if (tid == 0) {
                s_distinct = 0;
                s_overflow = 0;
                s_nr = 0;    // palindromic k-mers met in this round
                s_rcur = 0;  // reverse rows placed so far (only used when there is a palindrome)
            }
            __syncthreads();
            // ---- insert ----
            if (R == 1)
                bucket_insert<KT, false, BC_THREADS>(keys, lo, hi, tkey, tcnt, wq, log_slots, maskR, R, rd, limit, &s_distinct,
                                         &s_overflow);
            else
                bucket_insert<KT, true, BC_THREADS>(keys, lo, hi, tkey, tcnt, wq, log_slots, maskR, R, rd, limit, &s_distinct,
                                        &s_overflow);
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
            r_used = R > r_used ? R : r_used;
            d_max = s_distinct > d_max ? s_distinct : d_max;
            // ---- compact the occupied slots: count per warp, offsets, write the slot list ----
            const unsigned s_lo = warp * per_warp, s_hi = s_lo + per_warp;
            unsigned nh_w = 0;
            for (unsigned s0 = s_lo; s0 < s_hi; s0 += 32) nh_w += __popc(__ballot_sync(FULL, tkey[s0 + lane] != EMPTY));
            if (lane == 0) w_cnt[warp] = nh_w;
            __syncthreads();
            if (warp == 0) {
                const unsigned c = lane < (unsigned)BC_WARPS ? w_cnt[lane] : 0u;
                unsigned inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned o = __shfl_up_sync(FULL, inc, d);
                    if (lane >= (unsigned)d) inc += o;
                }
                if (lane < (unsigned)BC_WARPS) w_cnt[lane] = inc - c;
                if (lane == 31) s_nh = inc;
            }
            __syncthreads();
            {
                unsigned run = w_cnt[warp];
                for (unsigned s0 = s_lo; s0 < s_hi; s0 += 32) {
                    const bool occ = tkey[s0 + lane] != EMPTY;
                    const unsigned m = __ballot_sync(FULL, occ);
                    if (occ) list[run + __popc(m & lt)] = (uint16_t)(s0 + lane);
                    run += __popc(m);
                }
            }
            __syncthreads();
            const unsigned nh = s_nh;
            // ---- reverse rows: every head that is not its own reverse complement (odd k: all) ----
            unsigned nr = 0;
            if (strand) {
                nr = nh;
                if ((k & 1) == 0) {
                    unsigned pal = 0;
                    for (unsigned e0 = 0; e0 < nh; e0 += BC_THREADS) {
                        const unsigned e = e0 + tid;
                        bool p = false;
                        if (e < nh) {
                            const uint64_t K = kmer_unhash(hbase | (uint64_t)tkey[list[e]], nbits);
                            p = kmer_revcomp(K, k) == K;
                        }
                        pal += __popc(__ballot_sync(FULL, p));
                    }
                    if (lane == 0 && pal) atomicAdd(&s_nr, pal);
                    __syncthreads();
                    nr = nh - s_nr;
                }
            }
            if (tid == 0) {
                const unsigned rows = nh + nr;
                s_base = rows ? atomicAdd(&counters[0], (unsigned long long)rows) : 0ull;
                if (nh) atomicAdd(&counters[1], (unsigned long long)nh);
            }
            __syncthreads();
            const unsigned long long base_f = s_base, base_r = s_base + nh;
            for (unsigned e0 = 0; e0 < nh; e0 += BC_THREADS) {
                const unsigned e = e0 + tid;
                if (e < nh) {
                    const unsigned slot = list[e];
                    const uint64_t K = kmer_unhash(hbase | (uint64_t)tkey[slot], nbits);
                    const uint64_t inst = tcnt[slot];
                    uint32_t val;
                    if (strand) {
                        const uint64_t rcK = kmer_revcomp(K, k);
                        const bool rev = rcK != K;
                        uint64_t corr = 0;
                        if (n_corr) {
                            const uint64_t q = lower_bound_u64(corr_key, n_corr, K);
                            if (q < n_corr && corr_key[q] == K) corr = corr_cnt[q];
                        }
                        const uint64_t Hh = 2 * inst - corr;
                        uint64_t c = rev ? (Hh >> 1) : Hh;
                        if (c > 0x7fffffffull) c = 0x7fffffffull;
                        val = (uint32_t)c;
                        if (rev) {
                            // no palindrome in this round (almost always): the reverse row of head e
                            // sits at e; otherwise the reverse rows take the next free position
                            const unsigned long long pr = base_r + (nr == nh ? e : atomicAdd(&s_rcur, 1u));
                            if (pr < cap) {
                                row_key[pr] = rcK;
                                row_val[pr] = val;
                            }
                        }
                    } else {
                        val = inst > 0xffffffffull ? 0xffffffffu : (uint32_t)inst;
                    }
                    const unsigned long long pf = base_f + e;
                    if (pf < cap) {
                        row_key[pf] = K;
                        row_val[pf] = val;
                    }
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
        // (one odd bucket must not tax the rest of the run: halve the split again when it was roomy)
        r_hint = (r_used > 1 && 2 * d_max <= limit) ? r_used >> 1 : r_used;
    }
}

int part_scatter_bulk() {
    const char* v = getenv("PEDK_PART_BULK");
    return v && *v ? atoi(v) : 0;
}

template <typename KT, int NT, bool CONTIG, int KC, int PT_THREADS>
void launch_part_scatter_th(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks,
                            unsigned long long* cur, const unsigned long long* lim, int* overflow,
                            const PeerPtrs& dst, unsigned pass_mask, unsigned pass_id, cudaStream_t st) {
    constexpr int PT_WARPS = PT_THREADS / 32;
    constexpr int RPW = KeyBudget<KT>::ITEMS / NT > 0 ? KeyBudget<KT>::ITEMS / NT : 1;
    constexpr int TILE = PT_THREADS * NT * RPW;
    const uint64_t tiles = (n_reads + (uint64_t)PT_WARPS * RPW - 1) / ((uint64_t)PT_WARPS * RPW);
    const uint64_t resident = (uint64_t)kNumSMs * (PT_THREADS > 512 ? 1 : 2);
    const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
    if (part_scatter_bulk()) {
        // every run is staged at the misalignment of its destination and ends on a 16-byte line:
        // up to 2 x (16 - sizeof(KT)) bytes of padding per bin
        constexpr size_t SMEM = (size_t)TILE * sizeof(KT) + (size_t)NB * 32 + 16;
        ensure_dyn_smem(k_part_scatter<KT, NT, RPW, CONTIG, KC, PT_THREADS, true>, SMEM);
        k_part_scatter<KT, NT, RPW, CONTIG, KC, PT_THREADS, true>
            <<<grid, PT_THREADS, SMEM, st>>>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id);
    } else {
        constexpr size_t SMEM = (size_t)TILE * (sizeof(KT) + 1);
        ensure_dyn_smem(k_part_scatter<KT, NT, RPW, CONTIG, KC, PT_THREADS, false>, SMEM);
        k_part_scatter<KT, NT, RPW, CONTIG, KC, PT_THREADS, false>
            <<<grid, PT_THREADS, SMEM, st>>>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id);
    }
}

template <typename KT, int NT, bool CONTIG, int KC>
void launch_part_scatter_nt(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks,
                            unsigned long long* cur, const unsigned long long* lim, int* overflow,
                            const PeerPtrs& dst, unsigned pass_mask, unsigned pass_id, cudaStream_t st) {
    // the k = 20 kernel with 32-bit words also comes with long tiles for several ranks
    const int big = getenv("PEDK_PART_BIG_TILE") ? atoi(getenv("PEDK_PART_BIG_TILE")) : 1;
    if (KC == 20 && sizeof(KT) == 4 && (nranks > 1 || big == 2) && big && PT_BIG_OK<KT, NT>::value)
        launch_part_scatter_th<KT, NT, CONTIG, KC, (KC == 20 ? 1024 : 512)>(words, W, L, k, n_reads, a, nranks, cur,
                                                                           lim, overflow, dst, pass_mask, pass_id, st);
    else
        launch_part_scatter_th<KT, NT, CONTIG, KC, 512>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst,
                                                        pass_mask, pass_id, st);
}

// windows per lane, rounded up to an instantiated value
int nt_of(int L, int k) {
    const int nt = (L - k + 1 + 31) / 32;
    return nt <= 6 ? nt : nt <= 8 ? 8 : nt <= 12 ? 12 : 16;
}
// lane l can take nt consecutive windows out of one 32-base word
bool contiguous_ok(int nt, int k) { return nt - 1 + k <= 32; }

template <typename KT>
void launch_part_scatter_t(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks,
                           unsigned long long* cur, const unsigned long long* lim, int* overflow, const PeerPtrs& dst,
                           unsigned pass_mask, unsigned pass_id, cudaStream_t st) {
    const int nt = nt_of(L, k);
    const bool contig = contiguous_ok(nt, k);
    if (k == 20 && sizeof(KT) == 4 && contig && nt >= 2 && nt <= 6) {
        // the reference's k, short reads: k folded into the kernel
        switch (nt) {
        case 2: launch_part_scatter_nt<KT, 2, true, 20>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st); break;
        case 3: launch_part_scatter_nt<KT, 3, true, 20>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st); break;
        case 4: launch_part_scatter_nt<KT, 4, true, 20>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st); break;
        case 5: launch_part_scatter_nt<KT, 5, true, 20>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st); break;
        default: launch_part_scatter_nt<KT, 6, true, 20>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st); break;
        }
        return;
    }
#define PEDK_PS(N)                                                                                        \
    case N:                                                                                               \
        if (contig) launch_part_scatter_nt<KT, N, true, 0>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st); \
        else launch_part_scatter_nt<KT, N, false, 0>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st);   \
        break;
    switch (nt) {
        PEDK_PS(1) PEDK_PS(2) PEDK_PS(3) PEDK_PS(4) PEDK_PS(5) PEDK_PS(6) PEDK_PS(8) PEDK_PS(12)
    default:
        launch_part_scatter_nt<KT, 16, false, 0>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st);
    }
#undef PEDK_PS
}

}  // namespace

void launch_part_count(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a,
                       unsigned long long* histA, cudaStream_t st) {
    if (!n_reads || L < k) return;
    const uint64_t blocks = (n_reads + 7) / 8;
    const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    const int nt = nt_of(L, k);
    const bool contig = contiguous_ok(nt, k);
    if (k == 20 && contig && nt >= 2 && nt <= 6) {
        switch (nt) {
        case 2: k_part_count<2, true, 20><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA); break;
        case 3: k_part_count<3, true, 20><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA); break;
        case 4: k_part_count<4, true, 20><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA); break;
        case 5: k_part_count<5, true, 20><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA); break;
        default: k_part_count<6, true, 20><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA); break;
        }
        PEDK_CUDA_TRY(cudaGetLastError());
        return;
    }
#define PEDK_PC(N)                                                                                             \
    case N:                                                                                                    \
        if (contig) k_part_count<N, true, 0><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA); \
        else k_part_count<N, false, 0><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA);  \
        break;
    switch (nt) {
        PEDK_PC(1) PEDK_PC(2) PEDK_PC(3) PEDK_PC(4) PEDK_PC(5) PEDK_PC(6) PEDK_PC(8) PEDK_PC(12)
    default:
        k_part_count<16, false, 0><<<grid, PC_THREADS, 0, st>>>(words, W, L, k, n_reads, 2 * k - a, histA);
    }
#undef PEDK_PC
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_part_scatter(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks, int wide,
                         unsigned long long* cur, const unsigned long long* lim, int* overflow, const PeerPtrs& dst,
                         int n_pass, int pass, cudaStream_t st) {
    if (!n_reads || L < k) return;
    const unsigned pass_mask = (unsigned)n_pass - 1u, pass_id = (unsigned)pass;  // n_pass is a power of two
    if (wide)
        launch_part_scatter_t<uint64_t>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st);
    else
        launch_part_scatter_t<uint32_t>(words, W, L, k, n_reads, a, nranks, cur, lim, overflow, dst, pass_mask, pass_id, st);
    PEDK_CUDA_TRY(cudaGetLastError());
}

unsigned sub_tile_keys() { return (unsigned)(sub_scatter_threads() * sub_scatter_items()); }

namespace {
__global__ void k_cursor_counts(const unsigned long long* __restrict__ cur, const unsigned long long* __restrict__ lim,
                                unsigned long long cap, int n, unsigned long long* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (lim[i] == 0) {  // a bin outside this pass: no segment, nothing stored
        cnt[i] = 0;
        return;
    }
    const unsigned long long start = lim[i] - cap;
    const unsigned long long c = cur[i] - start;
    cnt[i] = c > cap ? cap : c;  // beyond cap only after an overflow (the flag is raised then)
}
}  // namespace

void launch_cursor_counts(const unsigned long long* cur, const unsigned long long* lim, unsigned long long cap, int n,
                          unsigned long long* cnt, cudaStream_t st) {
    k_cursor_counts<<<(n + 255) / 256, 256, 0, st>>>(cur, lim, cap, n, cnt);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_sub_hist_bits(const void* keys, int wide, const unsigned long long* seg, const unsigned* tile_prefix,
                          int n_seg, unsigned n_tiles, int shiftB, unsigned maskB, unsigned* histB, cudaStream_t st) {
    const unsigned long long* offA = seg;
    const int n_bins = n_seg;
    if (!n_tiles) return;
    const unsigned sub_tile = sub_tile_keys();
    const unsigned group = SH_GROUP_KEYS / sub_tile;
    const unsigned grid = (n_tiles + group - 1) / group;
    if (wide)
        k_sub_hist<uint64_t><<<grid, SH_THREADS, 0, st>>>(static_cast<const uint64_t*>(keys), offA, tile_prefix,
                                                          n_bins, n_tiles, sub_tile, shiftB, maskB, histB);
    else
        k_sub_hist<uint32_t><<<grid, SH_THREADS, 0, st>>>(static_cast<const uint32_t*>(keys), offA, tile_prefix,
                                                          n_bins, n_tiles, sub_tile, shiftB, maskB, histB);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_sub_hist(const void* keys, int wide, const unsigned long long* seg, const unsigned* tile_prefix,
                     int n_seg, unsigned n_tiles, int k, int a, int b, unsigned* histB, cudaStream_t st) {
    launch_sub_hist_bits(keys, wide, seg, tile_prefix, n_seg, n_tiles, 2 * k - a - b, (1u << b) - 1, histB, st);
}

void launch_sub_scatter_bits(const void* in, void* out, int wide, const unsigned long long* seg,
                             const unsigned* tile_prefix, int n_seg, unsigned n_tiles, int shiftB, unsigned maskB,
                             unsigned long long* curB, const unsigned long long* limB, int* overflow,
                             cudaStream_t st) {
    const unsigned long long* offA = seg;
    const int n_bins = n_seg;
    if (!n_tiles) return;
    int threads = sub_scatter_threads(), items = sub_scatter_items();
    if (!sub_scatter_forced() && !wide) {
        threads = 512;  // the same 4096-key tile with twice the threads
        items = 8;
    }
    const size_t smem = (size_t)threads * items * (wide ? sizeof(uint64_t) : sizeof(uint32_t));
#define PEDK_SS(KT, TH, IT)                                                                                        \
    do {                                                                                                           \
        ensure_dyn_smem(k_sub_scatter<KT, TH, IT>, smem);                                                          \
        k_sub_scatter<KT, TH, IT><<<n_tiles, TH, smem, st>>>(static_cast<const KT*>(in), static_cast<KT*>(out),    \
                                                             offA, tile_prefix, n_bins, n_tiles, shiftB, maskB,    \
                                                             curB, limB, overflow);                               \
    } while (0)
#define PEDK_SS_SHAPE(KT)                                  \
    do {                                                   \
        if (threads == 512 && items == 8) PEDK_SS(KT, 512, 8);  \
        else if (threads == 512) PEDK_SS(KT, 512, 16);     \
        else if (items == 8) PEDK_SS(KT, 256, 8);          \
        else PEDK_SS(KT, 256, 16);                         \
    } while (0)
    if (wide) PEDK_SS_SHAPE(uint64_t);
    else PEDK_SS_SHAPE(uint32_t);
#undef PEDK_SS_SHAPE
#undef PEDK_SS
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_sub_scatter(const void* in, void* out, int wide, const unsigned long long* seg,
                        const unsigned* tile_prefix, int n_seg, unsigned n_tiles, int k, int a, int b,
                        unsigned long long* curB, const unsigned long long* limB, int* overflow, cudaStream_t st) {
    launch_sub_scatter_bits(in, out, wide, seg, tile_prefix, n_seg, n_tiles, 2 * k - a - b, (1u << b) - 1, curB, limB,
                            overflow, st);
}

namespace {
// optimistic bucket layout: bucket (bin, sub) of an owned bin starts at (index of the bin among this
// rank's bins of the pass * 256 + sub) * cap; the other bins get empty buckets at 0
__global__ void k_bucket_layout(const int* __restrict__ bin_index, unsigned n_buckets, unsigned long long cap,
                                unsigned long long* __restrict__ offB, unsigned long long* __restrict__ limB) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_buckets) return;
    const int bi = bin_index[i >> 8];
    const unsigned long long o = bi < 0 ? 0ull : ((unsigned long long)bi * NB + (i & 255u)) * cap;
    offB[i] = o;
    limB[i] = bi < 0 ? 0ull : o + cap;
}
}  // namespace

void launch_bucket_layout(const int* bin_index, unsigned n_buckets, unsigned long long cap, unsigned long long* offB,
                          unsigned long long* limB, cudaStream_t st) {
    k_bucket_layout<<<(n_buckets + 255) / 256, 256, 0, st>>>(bin_index, n_buckets, cap, offB, limB);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_bucket_count(const void* keys, int wide, const unsigned long long* offB, const unsigned long long* endB,
                         unsigned n_buckets, int k, int a,
                         int b, int log_slots, int strand, const uint64_t* corr_key, const uint32_t* corr_cnt,
                         uint64_t n_corr, uint64_t* row_key, uint32_t* row_val, uint64_t cap,
                         unsigned long long* counters, unsigned* ticket, int* err, cudaStream_t st) {
    if (!n_buckets) return;
    if (log_slots < 10) log_slots = 10;
    if (log_slots > 13) log_slots = 13;
    const size_t ksz = wide ? sizeof(uint64_t) : sizeof(uint32_t);
    const int threads_env = getenv("PEDK_BC_THREADS") ? atoi(getenv("PEDK_BC_THREADS")) : 512;
    const int threads = threads_env <= 512 ? 512 : 1024;
    const size_t area = std::max(((size_t)1 << log_slots) * sizeof(uint16_t), (size_t)(threads / 32) * BC_QUEUE * ksz);
    const size_t smem = ((size_t)1 << log_slots) * (ksz + sizeof(uint32_t)) + area;
    PEDK_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    // two resident CTAs per SM while two tables fit its shared memory; buckets are handed out by ticket
    unsigned grid = kNumSMs * (smem > 100 * 1024 ? 1 : 2);
    if (grid > n_buckets) grid = n_buckets;
#define PEDK_BC(KT, TH)                                                                                            \
    do {                                                                                                           \
        ensure_dyn_smem(k_bucket_count<KT, TH>, smem);                                                             \
        k_bucket_count<KT, TH><<<grid, TH, smem, st>>>(static_cast<const KT*>(keys), offB, endB, n_buckets, k, a, b, log_slots, \
                                                       strand, corr_key, corr_cnt, n_corr, row_key, row_val, cap,  \
                                                       counters, ticket, err);                                     \
    } while (0)
    if (wide) {
        if (threads == 512) PEDK_BC(uint64_t, 512);
        else PEDK_BC(uint64_t, 1024);
    } else {
        if (threads == 512) PEDK_BC(uint32_t, 512);
        else PEDK_BC(uint32_t, 1024);
    }
#undef PEDK_BC
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
