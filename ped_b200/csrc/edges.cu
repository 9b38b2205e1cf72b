// S5/S6: strand expansion of the canonical count tables, last-base-count
// groups and the target/control last-base-difference test.
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

// Canonical row r -> strand rows.  count[K] for the reference's table
// (SURVEY.md Appendix A.3): with I = canonical instances, C = windows of
// palindromic reads among them (corr table), H = 2I - C half-units,
//   count = H      if K == revcomp(K)   (both strands of a read hit the same row)
//   count = H / 2  otherwise            (and revcomp(K) gets the same count)
__global__ void __launch_bounds__(256) k_expand(const uint64_t* __restrict__ ukey, const uint32_t* __restrict__ ustart,
                                               uint64_t D, int k, uint32_t sample_bit,
                                               const uint64_t* __restrict__ corr_key,
                                               const uint32_t* __restrict__ corr_start, uint64_t n_corr,
                                               uint64_t* __restrict__ row_key, uint32_t* __restrict__ row_val,
                                               unsigned long long* __restrict__ cursor) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = r < D;
    uint64_t K = 0, rc = 0;
    uint32_t v = 0;
    bool emit_rc = false;
    if (active) {
        K = ukey[r];
        uint64_t inst = (uint32_t)(ustart[r + 1] - ustart[r]);
        uint64_t corr = 0;
        if (n_corr) {
            uint64_t q = lower_bound_u64(corr_key, n_corr, K);
            if (q < n_corr && corr_key[q] == K) corr = (uint32_t)(corr_start[q + 1] - corr_start[q]);
        }
        rc = kmer_revcomp(K, k);
        uint64_t H = 2 * inst - corr;
        uint64_t cnt = (rc == K) ? H : (H >> 1);
        if (cnt > 0x7fffffffull) cnt = 0x7fffffffull;
        v = (uint32_t)cnt | sample_bit;
        row_key[r] = K;
        row_val[r] = v;
        emit_rc = rc != K;
    }
    unsigned m = __ballot_sync(0xffffffffu, emit_rc);
    if (m) {
        unsigned long long base = 0;
        unsigned lane = lane_id();
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (emit_rc) {
            unsigned long long o = base + __popc(m & lanemask_lt());
            row_key[o] = rc;
            row_val[o] = v;
        }
    }
}

__global__ void __launch_bounds__(256) k_first_of_bucket(const uint64_t* __restrict__ row_key,
                                                        const uint32_t* __restrict__ row_val, uint64_t n, int k,
                                                        unsigned long long* __restrict__ first) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t K = row_key[i];
    unsigned s = row_val[i] >> 31;
    unsigned tag = (unsigned)(K >> (2 * k - 6));
    unsigned long long* f = first + s * 64 + tag;
    // rows are sorted: only the first few rows of a bucket ever pass this filter
    if (K < *(volatile unsigned long long*)f) atomicMin(f, (unsigned long long)K);
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

__global__ void __launch_bounds__(256) k_edges(const uint64_t* __restrict__ row_key,
                                              const uint32_t* __restrict__ row_val, uint64_t n, int k,
                                              const unsigned long long* __restrict__ first,
                                              EdgeRec* __restrict__ edges, uint64_t cap,
                                              unsigned long long* __restrict__ n_edges) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t K = row_key[i];
    uint64_t P = K >> 2;
    if (i > 0 && (row_key[i - 1] >> 2) == P) return;  // not the head of its group
    uint32_t c[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
    bool hasc = false, hast = false;
    for (uint64_t j = i; j < n && j < i + 8; ++j) {
        uint64_t Kj = row_key[j];
        if ((Kj >> 2) != P) break;
        uint32_t v = row_val[j];
        unsigned x = (unsigned)(Kj & 3);
        if (v >> 31) { t[x] = v & 0x7fffffffu; hast = true; }
        else { c[x] = v; hasc = true; }
    }
    if (!(hasc && hast)) return;  // `join` is an inner join (ped.pl:709)
    unsigned tag = (unsigned)(K >> (2 * k - 6));
    unsigned long long fc = first[tag], ft = first[64 + tag];
    if ((fc != ~0ull && (fc >> 2) == P) || (ft != ~0ull && (ft >> 2) == P)) return;  // F7 rows: host
    uint8_t cont, alt;
    if (!edge_test(c, t, &cont, &alt)) return;
    unsigned long long o = atomicAdd(n_edges, 1ull);
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

void launch_expand(const uint64_t* ukey, const uint32_t* ustart, uint64_t D, int k, int sample,
                   const uint64_t* corr_key, const uint32_t* corr_start, uint64_t n_corr, uint64_t* row_key,
                   uint32_t* row_val, unsigned long long* cursor, cudaStream_t st) {
    if (!D) return;
    k_expand<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(ukey, ustart, D, k, sample ? 0x80000000u : 0u, corr_key,
                                                           corr_start, n_corr, row_key, row_val, cursor);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_first_of_bucket(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int k,
                            unsigned long long* first, cudaStream_t st) {
    if (!n) return;
    k_first_of_bucket<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(row_key, row_val, n, k, first);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_edges(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int k,
                  const unsigned long long* first, EdgeRec* edges, uint64_t cap, unsigned long long* n_edges,
                  cudaStream_t st) {
    if (!n) return;
    k_edges<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(row_key, row_val, n, k, first, edges, cap, n_edges);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_lookup_groups(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, const uint64_t* p_list,
                          int n_p, GroupRec* out, cudaStream_t st) {
    if (!n_p) return;
    k_lookup_groups<<<(n_p + 127) / 128, 128, 0, st>>>(row_key, row_val, n, p_list, n_p, out);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
