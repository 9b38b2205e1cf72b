// S5/S6: last-base-count groups and the target/control last-base-difference test.
//
// lbcKmerSub (ped.pl:833-860) groups the sorted 20-mer table by its 19-mer
// prefix; joinKmerSub (ped.pl:703-782) inner-joins the control and target lbc
// tables on that prefix and applies the polymorphic-edge test.  Here both
// samples' strand tables are written into ONE row array (K, count | sample<<31)
// and sorted together (radix.cu, pairs), so the at most eight rows of a prefix
// group (4 last bases x 2 samples) are adjacent and the join is a windowed read:
// no merge-path, no binary searches.
#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t* a, uint64_t n, uint64_t key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int NPASS>
__global__ void __launch_bounds__(256) k_copy_rows(const uint64_t* __restrict__ src_key,
                                                  const uint32_t* __restrict__ src_val, uint64_t n,
                                                  uint64_t* __restrict__ dst_key, uint32_t* __restrict__ dst_val,
                                                  uint32_t or_bits, unsigned long long* __restrict__ digit_hist) {
    __shared__ unsigned h[NPASS ? NPASS : 1][256];
    for (int i = threadIdx.x; i < NPASS * 256; i += 256) (&h[0][0])[i] = 0;
    if (NPASS) __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = src_key[i];
        dst_key[i] = k;
        dst_val[i] = src_val[i] | or_bits;
#pragma unroll
        for (int p = 0; p < NPASS; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 255u], 1u);
    }
    if (NPASS) {
        __syncthreads();
        for (int i = threadIdx.x; i < NPASS * 256; i += 256) {
            const unsigned c = (&h[0][0])[i];
            if (c) atomicAdd(&digit_hist[i], (unsigned long long)c);
        }
    }
}

// One block per (sample, 3-nt bucket): the first row of that sample inside the bucket of
// the merged, sorted table.  Rows of both samples interleave, so the hit is normally in
// the first window of 256 rows; a bucket held by one sample only costs one sweep over it.
__global__ void __launch_bounds__(256) k_first_of_bucket(const uint64_t* __restrict__ row_key,
                                                        const uint32_t* __restrict__ row_val, uint64_t n, int k,
                                                        unsigned long long* __restrict__ first) {
    __shared__ unsigned long long s_lo, s_hi, s_min;
    const unsigned s = blockIdx.x >> 6, tag = blockIdx.x & 63u;
    if (threadIdx.x == 0) {
        const int sh = 2 * k - 6;
        s_lo = lower_bound_u64(row_key, n, (uint64_t)tag << sh);
        s_hi = tag == 63u ? n : lower_bound_u64(row_key, n, (uint64_t)(tag + 1) << sh);
        s_min = ~0ull;
    }
    __syncthreads();
    const uint64_t lo = s_lo, hi = s_hi;
    for (uint64_t base = lo; base < hi; base += 256) {
        uint64_t i = base + threadIdx.x;
        bool hit = i < hi && (row_val[i] >> 31) == s;
        if (hit) atomicMin(&s_min, (unsigned long long)i);
        if (__syncthreads_or(hit)) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) first[blockIdx.x] = s_min == ~0ull ? ~0ull : row_key[s_min];
}

__device__ __forceinline__ bool edge_test(const uint32_t* c, const uint32_t* t, uint8_t* cont, uint8_t* alt) {
    // pass 1, ped.pl:720-744 (cutoff is always 3)
    int rep = 0, nohita = 0, nohitb = 0, pola = 0, polb = 0;
    unsigned a = 0, b = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        uint32_t cv = c[x], tv = t[x];
        if (cv > 100) ++rep;
        else if (cv >= 3) { a |= 1u << x; if (tv <= 1) ++pola; }
        else if (cv <= 1) ++nohita;
        if (tv > 100) ++rep;
        else if (tv >= 3) { b |= 1u << x; if (cv <= 1) ++polb; }
        else if (tv <= 1) ++nohitb;
    }
    if (!(a != b && a && b && rep == 0 && nohita >= 2 && nohitb >= 2 && (pola == 1 || polb == 1))) return false;
    // pass 2, ped.pl:750-781; c/total > 0.25 is exactly 4c > total on integers
    uint64_t ct = (uint64_t)c[0] + c[1] + c[2] + c[3];
    uint64_t tt = (uint64_t)t[0] + t[1] + t[2] + t[3];
    if (ct == 0 || tt == 0) return false;
    unsigned co = 0, al = 0;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        if (4ull * c[x] > ct) co |= 1u << x;
        if (4ull * t[x] > tt) al |= 1u << x;
    }
    if (co == al) return false;
    *cont = (uint8_t)co;
    *alt = (uint8_t)al;
    return true;
}

constexpr int ED_THREADS = 256;
constexpr int ED_HALO = 8;  // a group has at most 4 bases x 2 samples rows

// Pass 1 of joinKmerSub (ped.pl:720-744) looks at every count only through four classes:
//   0: <= 1 (or absent) -> "nohit",  1: == 2,  2: 3..100 -> allele,  3: > 100 -> repeat.
// A (k-1)-mer group therefore boils down to 16 bits, two per (sample, last base) slot, and
// whether it passes is a 65536-entry truth table (8 KiB, built once on the host from the same
// rules, L1-resident).  Every row contributes its two bits; the head of a group ORs the <= 8
// words of its group and looks the answer up.  Only the few groups that pass (and are present
// in both samples) gather their counts for pass 2 (ped.pl:750-781).
__host__ __device__ inline unsigned count_class(uint32_t c) { return c > 100 ? 3u : c >= 3 ? 2u : c <= 1 ? 0u : 1u; }

__global__ void __launch_bounds__(ED_THREADS) k_edges(const uint64_t* __restrict__ row_key,
                                                     const uint32_t* __restrict__ row_val, uint64_t n, int k,
                                                     const unsigned long long* __restrict__ first,
                                                     const uint32_t* __restrict__ pass1_bits,
                                                     EdgeRec* __restrict__ edges, uint64_t cap,
                                                     unsigned long long* __restrict__ n_edges) {
    // per staged row: bits 0-15 class bits of its slot, bit 16/17 control/target present,
    // bit 31 = same (k-1)-mer as the row before
    __shared__ uint32_t s_w[ED_THREADS + ED_HALO];
    const unsigned tid = threadIdx.x;
    const uint64_t base = (uint64_t)blockIdx.x * ED_THREADS;
    const uint64_t left = n - base;
    const unsigned loaded = left < (uint64_t)(ED_THREADS + ED_HALO) ? (unsigned)left : (unsigned)(ED_THREADS + ED_HALO);
    uint64_t myK = 0;
    for (unsigned j = tid; j < loaded; j += ED_THREADS) {
        const uint64_t K = row_key[base + j];
        const uint32_t v = row_val[base + j];
        const uint64_t prevK = base + j ? row_key[base + j - 1] : ~K;
        const unsigned s = v >> 31, x = (unsigned)(K & 3);
        uint32_t w = (count_class(v & 0x7fffffffu) << (2 * (4 * s + x))) | (1u << (16 + s));
        if (base + j && (prevK >> 2) == (K >> 2)) w |= 0x80000000u;
        s_w[j] = w;
        if (j == tid) myK = K;
    }
    __syncthreads();
    if (tid >= loaded) return;
    uint32_t acc = s_w[tid];
    if (acc >> 31) return;  // not the head of its group
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        if (tid + j >= loaded) break;
        const uint32_t w = s_w[tid + j];
        if (!(w >> 31)) break;
        acc |= w;
    }
    if (((acc >> 16) & 3u) != 3u) return;  // `join` is an inner join (ped.pl:709)
    const unsigned code = acc & 0xffffu;
    if (!((__ldg(pass1_bits + (code >> 5)) >> (code & 31u)) & 1u)) return;
    // ---- rare from here on: the full evaluation on the counts ----
    const uint64_t P = myK >> 2;
    const unsigned tag = (unsigned)(myK >> (2 * k - 6));
    const unsigned long long fc = first[tag], ft = first[64 + tag];
    if ((fc != ~0ull && (fc >> 2) == P) || (ft != ~0ull && (ft >> 2) == P)) return;  // F7 rows: host
    uint32_t c[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
    const uint64_t i = base + tid;
    for (uint64_t j = i; j < n && j < i + 8; ++j) {
        const uint64_t Kj = row_key[j];
        if ((Kj >> 2) != P) break;
        const uint32_t v = row_val[j];
        const unsigned x = (unsigned)(Kj & 3);
        if (v >> 31) t[x] = v & 0x7fffffffu;
        else c[x] = v;
    }
    uint8_t cont, alt;
    if (!edge_test(c, t, &cont, &alt)) return;
    const unsigned long long o = atomicAdd(n_edges, 1ull);
    if (o < cap) {
        EdgeRec e;
        e.p = P;
#pragma unroll
        for (int x = 0; x < 4; ++x) { e.c[x] = c[x]; e.t[x] = t[x]; }
        e.cont = cont;
        e.alt = alt;
#pragma unroll
        for (int x = 0; x < 6; ++x) e.pad[x] = 0;
        edges[o] = e;
    }
}

__global__ void k_lookup_groups(const uint64_t* __restrict__ row_key, const uint32_t* __restrict__ row_val,
                                uint64_t n, const uint64_t* __restrict__ p_list, int n_p,
                                GroupRec* __restrict__ out) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_p) return;
    uint64_t P = p_list[q];
    GroupRec g;
    g.p = P;
    for (int x = 0; x < 4; ++x) { g.c[x] = g.t[x] = 0; g.hasc[x] = g.hast[x] = 0; }
    uint64_t j = lower_bound_u64(row_key, n, P << 2);
    for (; j < n && (row_key[j] >> 2) == P; ++j) {
        uint32_t v = row_val[j];
        unsigned x = (unsigned)(row_key[j] & 3);
        if (v >> 31) { g.t[x] = v & 0x7fffffffu; g.hast[x] = 1; }
        else { g.c[x] = v; g.hasc[x] = 1; }
    }
    out[q] = g;
}

}  // namespace

void launch_copy_rows(const uint64_t* src_key, const uint32_t* src_val, uint64_t n, uint64_t* dst_key,
                      uint32_t* dst_val, uint32_t or_bits, unsigned long long* digit_hist, int n_pass,
                      cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 16 ? blocks : (uint64_t)kNumSMs * 16);
#define PEDK_CR(N) \
    case N: k_copy_rows<N><<<grid, 256, 0, st>>>(src_key, src_val, n, dst_key, dst_val, or_bits, digit_hist); break;
    switch (digit_hist ? n_pass : 0) {
        PEDK_CR(0) PEDK_CR(1) PEDK_CR(2) PEDK_CR(3) PEDK_CR(4) PEDK_CR(5) PEDK_CR(6) PEDK_CR(7)
    default: k_copy_rows<8><<<grid, 256, 0, st>>>(src_key, src_val, n, dst_key, dst_val, or_bits, digit_hist);
    }
#undef PEDK_CR
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_first_of_bucket(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int k,
                            unsigned long long* first, cudaStream_t st) {
    if (!n) return;
    k_first_of_bucket<<<128, 256, 0, st>>>(row_key, row_val, n, k, first);
    PEDK_CUDA_TRY(cudaGetLastError());
}

// truth table of pass 1 over the 16 class bits of a group (see k_edges)
void build_edge_pass1_table(uint32_t* bits /* 2048 words */) {
    for (unsigned w = 0; w < 2048; ++w) bits[w] = 0;
    for (unsigned code = 0; code < 65536; ++code) {
        int rep = 0, nohita = 0, nohitb = 0, pola = 0, polb = 0;
        unsigned a = 0, b = 0;
        for (int x = 0; x < 4; ++x) {
            const unsigned cv = (code >> (2 * x)) & 3u, tv = (code >> (2 * (4 + x))) & 3u;
            if (cv == 3) ++rep;
            else if (cv == 2) { a |= 1u << x; if (tv == 0) ++pola; }
            else if (cv == 0) ++nohita;
            if (tv == 3) ++rep;
            else if (tv == 2) { b |= 1u << x; if (cv == 0) ++polb; }
            else if (tv == 0) ++nohitb;
        }
        if (a != b && a && b && rep == 0 && nohita >= 2 && nohitb >= 2 && (pola == 1 || polb == 1))
            bits[code >> 5] |= 1u << (code & 31u);
    }
}

void launch_edges(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int k,
                  const unsigned long long* first, const uint32_t* pass1_bits, EdgeRec* edges, uint64_t cap,
                  unsigned long long* n_edges, cudaStream_t st) {
    if (!n) return;
    k_edges<<<(unsigned)((n + ED_THREADS - 1) / ED_THREADS), ED_THREADS, 0, st>>>(row_key, row_val, n, k, first,
                                                                                 pass1_bits, edges, cap, n_edges);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_lookup_groups(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, const uint64_t* p_list,
                          int n_p, GroupRec* out, cudaStream_t st) {
    if (!n_p) return;
    k_lookup_groups<<<(n_p + 127) / 128, 128, 0, st>>>(row_key, row_val, n, p_list, n_p, out);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
