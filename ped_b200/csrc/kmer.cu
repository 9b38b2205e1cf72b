// S2: k-mer extraction and S4: run-length count -- the SORT form of the counting path.
//
// Since round 1d the instance stream of a sample goes through bucket.cu (hash bins ->
// buckets -> shared-memory tables) and is never sorted.  The kernels here still (a) build
// the small correction table of the windows of palindromic reads (extract -> LSD sort ->
// run-length), and (b) are the whole counting path under PEDK_COUNT_MODE=sort, which the
// parity tests keep running against the bucketed path.
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
// (ped.pl:928, 862-889): one pass over the sorted stream that turns every run head
// straight into the rows of the strand table (no scan, no intermediate table).
#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

constexpr int EX_THREADS = 256;

template <int NPASS>
__global__ void __launch_bounds__(EX_THREADS) k_extract(const uint64_t* __restrict__ words, int W, int L, int k,
                                                       uint64_t n_reads, uint64_t* __restrict__ out,
                                                       unsigned long long* __restrict__ digit_hist) {
    __shared__ unsigned h[NPASS ? NPASS : 1][256];
    for (int i = threadIdx.x; i < NPASS * 256; i += EX_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * EX_THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * EX_THREADS) >> 5;
    const int nwin = L - k + 1;
    const int kshift = 64 - 2 * k;
    const int nt = (nwin + 31) >> 5;
    for (uint64_t r = warp; r < n_reads; r += n_warps) {
        uint64_t mine = (int)lane < W ? words[r * (uint64_t)W + lane] : 0ull;
        uint64_t* dst = out + r * (uint64_t)nwin;
        for (int t = 0; t < nt; ++t) {
            uint64_t hi = __shfl_sync(0xffffffffu, mine, t);
            uint64_t lo = __shfl_sync(0xffffffffu, mine, t + 1);
            int i = 32 * t + (int)lane;
            uint64_t x = lane ? ((hi << (2 * lane)) | (lo >> (64 - 2 * lane))) : hi;
            uint64_t km = x >> kshift;
            uint64_t rc = revcomp_groups64(km) >> kshift;
            uint64_t canon = km < rc ? km : rc;
            if (i < nwin) {
                dst[i] = canon;
                const uint32_t clo = (uint32_t)canon, chi = (uint32_t)(canon >> 32);
#pragma unroll
                for (int p = 0; p < NPASS; ++p)
                    atomicAdd(&h[p][((p < 4 ? clo : chi) >> (8 * (p & 3))) & 255u], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * 256; i += EX_THREADS) {
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

constexpr int CR_THREADS = 256;
constexpr int CR_ITEMS = 8;
constexpr int CR_WARPS = CR_THREADS / 32;
constexpr int CR_SEG = 32 * CR_ITEMS;        // keys per warp
constexpr int CR_TILE = CR_SEG * CR_WARPS;   // keys per block

__device__ __forceinline__ uint64_t lower_bound_dev(const uint64_t* a, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// first index >= from whose key differs from K (or n): a short linear probe, then a
// gallop + bisection, so one k-mer with millions of instances costs ~60 loads, not millions
__device__ __forceinline__ size_t run_end_from(const uint64_t* __restrict__ keys, size_t n, size_t from, uint64_t K) {
    size_t j = from;
    for (int s = 0; s < 8 && j < n && keys[j] == K; ++s) ++j;
    if (j >= n || keys[j] != K) return j;
    size_t lo = j, step = 32, hi;
    for (;;) {
        hi = lo + step;
        if (hi >= n) { hi = n; break; }
        if (keys[hi] != K) break;
        lo = hi;
        step <<= 1;
    }
    while (hi - lo > 1) {  // keys[lo] == K, (hi == n or keys[hi] != K)
        size_t mid = lo + ((hi - lo) >> 1);
        if (keys[mid] == K) lo = mid; else hi = mid;
    }
    return hi;
}

// One pass over the sorted canonical instances: every run head turns into table rows.
//   STRAND == false: rows (K, run length)                      -- raw run-length count
//   STRAND == true : rows (K, c | sample) and (revcomp K, c | sample), c = the reference's
//                    count: with I instances and C windows of
//                    palindromic reads among them (corr table), H = 2I - C half-units,
//                    c = H if K == revcomp(K) else H / 2.
// Rows are appended through a block-aggregated cursor, in no particular order: the row
// table is sorted afterwards anyway, so no scan (and no chained look-back) is needed.
// counters[0] = rows appended (continues past cap so the caller can retry), counters[1] += heads.
template <bool STRAND>
__global__ void __launch_bounds__(CR_THREADS) k_count_rows(const uint64_t* __restrict__ keys, size_t n, int k,
                                                          uint32_t sample_bit,
                                                          const uint64_t* __restrict__ corr_key,
                                                          const uint32_t* __restrict__ corr_cnt, uint64_t n_corr,
                                                          uint64_t* __restrict__ row_key,
                                                          uint32_t* __restrict__ row_val, uint64_t cap,
                                                          unsigned long long* __restrict__ counters) {
    // the run heads of a warp's 256 keys, compacted: (offset inside the segment, key)
    __shared__ uint64_t s_key[CR_WARPS][CR_SEG];
    __shared__ uint16_t s_pos[CR_WARPS][CR_SEG];
    __shared__ unsigned w_rows[CR_WARPS], w_heads[CR_WARPS];
    __shared__ unsigned long long s_base;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const size_t seg = (size_t)blockIdx.x * CR_TILE + (size_t)warp * CR_SEG;
    const size_t seg_end = seg + CR_SEG < n ? seg + CR_SEG : n;
    const unsigned lt = lanemask_lt();

    // ---- A: head flags for the whole segment; heads drop (offset, key) into the compact list ----
    unsigned nh = 0;
    {
        uint64_t carry = (seg > 0 && seg < n) ? keys[seg - 1] : 0ull;  // key just before the segment
        uint64_t key[CR_ITEMS];
#pragma unroll
        for (int j = 0; j < CR_ITEMS; ++j) {
            size_t i = seg + 32 * j + lane;
            key[j] = i < n ? keys[i] : 0ull;
        }
#pragma unroll
        for (int j = 0; j < CR_ITEMS; ++j) {
            size_t i = seg + 32 * j + lane;
            uint64_t prev = __shfl_up_sync(0xffffffffu, key[j], 1);
            if (lane == 0) prev = carry;
            const bool head = i < n && (i == 0 || key[j] != prev);
            const unsigned h = __ballot_sync(0xffffffffu, head);
            if (head) {
                unsigned r = nh + __popc(h & lt);
                s_key[warp][r] = key[j];
                s_pos[warp][r] = (uint16_t)(32 * j + lane);
            }
            nh += __popc(h);
            carry = __shfl_sync(0xffffffffu, key[j], 31);
        }
    }
    __syncwarp();

    // ---- B1: how many reverse rows (heads that are not their own reverse complement) ----
    unsigned nr = 0;
    if (STRAND) {
        for (unsigned r0 = 0; r0 < nh; r0 += 32) {
            const unsigned r = r0 + lane;
            const bool rev = r < nh && kmer_revcomp(s_key[warp][r], k) != s_key[warp][r];
            nr += __popc(__ballot_sync(0xffffffffu, rev));
        }
    }
    if (lane == 0) {
        w_rows[warp] = nh + nr;
        w_heads[warp] = nh;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned rows = 0, heads = 0;
#pragma unroll
        for (int w = 0; w < CR_WARPS; ++w) {
            unsigned c = w_rows[w];
            w_rows[w] = rows;
            rows += c;
            heads += w_heads[w];
        }
        s_base = rows ? atomicAdd(&counters[0], (unsigned long long)rows) : 0ull;
        if (heads) atomicAdd(&counters[1], (unsigned long long)heads);
    }
    __syncthreads();

    // ---- B2: one lane per head: run length -> count -> rows (forward rows, then reverse rows) ----
    const unsigned long long of = s_base + w_rows[warp];
    unsigned long long orv = of + nh;
    for (unsigned r0 = 0; r0 < nh; r0 += 32) {
        const unsigned r = r0 + lane;
        const bool act = r < nh;
        uint64_t K = 0;
        uint32_t v = 0;
        bool rev = false;
        if (act) {
            K = s_key[warp][r];
            const size_t i = seg + s_pos[warp][r];
            const size_t end = r + 1 < nh ? seg + s_pos[warp][r + 1] : run_end_from(keys, n, seg_end, K);
            const uint64_t inst = end - i;
            if (STRAND) {
                rev = kmer_revcomp(K, k) != K;
                uint64_t corr = 0;
                if (n_corr) {
                    uint64_t q = lower_bound_dev(corr_key, n_corr, K);
                    if (q < n_corr && corr_key[q] == K) corr = corr_cnt[q];
                }
                uint64_t Hh = 2 * inst - corr;
                uint64_t c = rev ? (Hh >> 1) : Hh;
                if (c > 0x7fffffffull) c = 0x7fffffffull;
                v = (uint32_t)c | sample_bit;
            } else {
                v = inst > 0xffffffffull ? 0xffffffffu : (uint32_t)inst;
            }
            const unsigned long long pf = of + r;
            if (pf < cap) {
                row_key[pf] = K;
                row_val[pf] = v;
            }
        }
        if (STRAND) {
            const unsigned m = __ballot_sync(0xffffffffu, rev);
            if (rev) {
                const unsigned long long pr = orv + __popc(m & lt);
                if (pr < cap) {
                    row_key[pr] = kmer_revcomp(K, k);
                    row_val[pr] = v;
                }
            }
            orv += __popc(m);
        }
    }
}

}  // namespace

void launch_extract(const uint64_t* words, int W, int L, int k, uint64_t n_reads, uint64_t* out,
                    unsigned long long* digit_hist, int n_pass, cudaStream_t st) {
    if (!n_reads || L < k) return;
    uint64_t blocks = (n_reads + 7) / 8;
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    switch (n_pass) {
    case 0: k_extract<0><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 1: k_extract<1><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 2: k_extract<2><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 3: k_extract<3><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 4: k_extract<4><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 5: k_extract<5><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 6: k_extract<6><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    case 7: k_extract<7><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    default: k_extract<8><<<grid, EX_THREADS, 0, st>>>(words, W, L, k, n_reads, out, digit_hist); break;
    }
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

void launch_count_rows(const uint64_t* keys, size_t n, int k, int strand, int sample, const uint64_t* corr_key,
                       const uint32_t* corr_cnt, uint64_t n_corr, uint64_t* row_key, uint32_t* row_val, uint64_t cap,
                       unsigned long long* counters, cudaStream_t st) {
    if (!n) return;
    unsigned grid = (unsigned)((n + CR_TILE - 1) / CR_TILE);
    if (strand)
        k_count_rows<true><<<grid, CR_THREADS, 0, st>>>(keys, n, k, sample ? 0x80000000u : 0u, corr_key, corr_cnt,
                                                        n_corr, row_key, row_val, cap, counters);
    else
        k_count_rows<false><<<grid, CR_THREADS, 0, st>>>(keys, n, k, 0u, nullptr, nullptr, 0, row_key, row_val, cap,
                                                         counters);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
