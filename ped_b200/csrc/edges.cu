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

constexpr unsigned FULL_MASK = 0xffffffffu;

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


// ------------------------------------------------------------------------------------------
// Hash join (the pedk_join path).  The reference sorts both lbc tables only so that `join` sees
// equal 19-mers next to each other (ped.pl:709); what the test needs is that the <= 8 rows of a
// (k-1)-mer P = K >> 2 (4 last bases x 2 samples) MEET.  So the strand rows are not sorted:
//   k_rows_count    rows -> 2^a-bin histogram of the top bits of kmer_hash(P), plus the smallest
//                   k-mer of every (sample, 3-nt bucket): the first lbc row of that file (F7)
//   k_rows_scatter  rows -> 8-byte join records, appended to the level-A' bin of hash(P) (in the
//                   owner rank's buffer, peer memory over NVLink on several GPUs)
//   k_sub_hist / k_sub_scatter (bucket.cu)   the same second binning level as the instances
//   k_join_bucket   one CTA per bucket: a shared-memory table keyed by the remaining hash bits
//                   ORs every row's 7-bit count into its (sample, base) field; a pass over the
//                   occupied slots applies joinKmerSub (pass-1 truth table, then the counts)
// A join record: (low bits of hash(P)) << 10 | last base << 8 | sample << 7 | min(count, 127).
// Counts above 100 only ever mean "repeat" (ped.pl:721, 732), and a row that is printed has
// none, so saturating at 127 loses nothing.
constexpr int JR_PAYLOAD_BITS = 10;

__global__ void __launch_bounds__(256) k_rows_count(const uint64_t* __restrict__ key, uint64_t n, unsigned sample,
                                                   int k, int nbitsP, int shiftA,
                                                   unsigned long long* __restrict__ hist,
                                                   unsigned long long* __restrict__ first) {
    __shared__ unsigned h[256];
    __shared__ unsigned long long smin[64];
    const unsigned tid = threadIdx.x;
    h[tid] = 0;
    if (tid < 64) smin[tid] = ~0ull;
    __syncthreads();
    const int tag_shift = 2 * k - 6;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + tid; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t K = key[i];
        atomicAdd(&h[(unsigned)(kmer_hash(K >> 2, nbitsP) >> shiftA)], 1u);
        const unsigned tag = (unsigned)(K >> tag_shift) & 63u;
        // most rows are not the smallest of their bucket: look before paying for the atomic
        if (K < *(volatile unsigned long long*)&smin[tag]) atomicMin(&smin[tag], (unsigned long long)K);
    }
    __syncthreads();
    if (h[tid]) atomicAdd(&hist[tid], (unsigned long long)h[tid]);
    if (tid < 64 && smin[tid] != ~0ull) atomicMin(&first[sample * 64 + tid], smin[tid]);
}

constexpr int JS_THREADS = 512;
constexpr int JS_ITEMS = 8;
constexpr int JS_TILE = JS_THREADS * JS_ITEMS;

// BULK: a bin's run leaves shared memory as one bulk copy (cp.async.bulk, the TMA unit), as in
// k_part_scatter (bucket.cu): staged at the misalignment of its destination, the aligned body by
// bulk copy, at most one record before and after it by plain stores.
__device__ __forceinline__ void js_bulk_store(void* dst, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}

template <bool BULK>
__global__ void __launch_bounds__(JS_THREADS, 2) k_rows_scatter(const uint64_t* __restrict__ key,
                                                               const uint32_t* __restrict__ val, uint64_t n,
                                                               unsigned sample, int nbitsP, int a, int nranks,
                                                               unsigned long long* __restrict__ cur, PeerPtrs dst) {
    __shared__ __align__(16) uint64_t stage[JS_TILE + 512];  // BULK: up to two records of padding per bin
    __shared__ uint8_t sbin[BULK ? 16 : JS_TILE];
    __shared__ unsigned hist[256];
    __shared__ unsigned off[256];
    __shared__ uint64_t* s_out[256];
    const unsigned tid = threadIdx.x;
    const int shiftA = nbitsP - a;
    const uint64_t maskS = (1ull << shiftA) - 1;
    const uint64_t n_tiles = (n + JS_TILE - 1) / JS_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        uint64_t rec[JS_ITEMS];
        uint32_t meta[JS_ITEMS];  // bin << 16 | rank inside (tile, bin); ~0: no row
#pragma unroll
        for (int j = 0; j < JS_ITEMS; ++j) {
            const uint64_t i = tile * JS_TILE + (uint64_t)j * JS_THREADS + tid;
            meta[j] = 0xffffffffu;
            rec[j] = 0;
            if (i < n) {
                const uint64_t K = key[i];
                const uint32_t c = val[i] & 0x7fffffffu;
                const uint64_t hp = kmer_hash(K >> 2, nbitsP);
                const unsigned bin = (unsigned)(hp >> shiftA);
                rec[j] = ((hp & maskS) << JR_PAYLOAD_BITS) | ((K & 3) << 8) | ((uint64_t)sample << 7) |
                         (c > 127u ? 127u : c);
                meta[j] = (bin << 16) | atomicAdd(&hist[bin], 1u);
            }
        }
        __syncthreads();
        const unsigned c = tid < 256 ? hist[tid] : 0u;
        if (BULK) {
            unsigned long long g = 0;
            if (c) g = atomicAdd(&cur[tid], (unsigned long long)c);
            const unsigned mis = (unsigned)g & 1u;  // two records per 16 bytes
            const unsigned need = c ? (mis + c + 1u) & ~1u : 0u;
            const uint64_t ex = block_exclusive_scan<JS_THREADS>(need, nullptr);
            if (tid < 256) {
                off[tid] = (unsigned)ex + mis;
                const int owner = (int)tid < (1 << a) ? (int)((tid * (unsigned)nranks) >> a) : 0;
                s_out[tid] = c ? dst.p[owner] + g : nullptr;
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous tile's copies have read the staging area
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < JS_ITEMS; ++j)
                if (meta[j] != 0xffffffffu) stage[off[meta[j] >> 16] + (meta[j] & 0xffffu)] = rec[j];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < 256 && s_out[tid]) {
                uint64_t* p = s_out[tid];
                const uint64_t* sp = stage + off[tid];
                const unsigned head = c < mis ? c : mis;
                const unsigned body = (c - head) & ~1u;
                if (head) p[0] = sp[0];
                if (body) js_bulk_store(p + head, sp + head, body * 8u);
                if (head + body < c) p[c - 1] = sp[c - 1];
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            uint64_t total;
            const uint64_t ex = block_exclusive_scan<JS_THREADS>(c, &total);
            if (tid < 256) {
                const unsigned long long g = c ? atomicAdd(&cur[tid], (unsigned long long)c) : 0ull;
                off[tid] = (unsigned)ex;
                const int owner = (int)tid < (1 << a) ? (int)((tid * (unsigned)nranks) >> a) : 0;
                s_out[tid] = dst.p[owner] + (g - ex);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < JS_ITEMS; ++j) {
                if (meta[j] != 0xffffffffu) {
                    const unsigned bin = meta[j] >> 16;
                    const unsigned pos = off[bin] + (meta[j] & 0xffffu);
                    stage[pos] = rec[j];
                    sbin[pos] = (uint8_t)bin;
                }
            }
            __syncthreads();
            const unsigned n_staged = (unsigned)total;
            for (unsigned i = tid; i < n_staged; i += JS_THREADS) s_out[sbin[i]][i] = stage[i];
            __syncthreads();
        }
    }
    if (BULK && tid < 256) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

constexpr int JB_UNROLL = 4;  // records a thread has in flight

__device__ __forceinline__ uint32_t jb_fold(uint32_t x) { return x; }
__device__ __forceinline__ uint32_t jb_fold(uint64_t x) { return (uint32_t)x ^ ((uint32_t)(x >> 32) * 0x9E3779B1u); }
__device__ __forceinline__ uint32_t jb_cas(uint32_t* p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t jb_cas(uint64_t* p, uint64_t cmp, uint64_t v) {
    return atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)cmp, (unsigned long long)v);
}

// KT: type of the table key = the r = 2(k-1) - a - b hash bits a bucket does not fix (u32 while
// r <= 31, so that ~0 is free to mean "empty").  tc/tt: the four counts of the control / target,
// one byte each (a count is at most 127), so that pass 1 of joinKmerSub (ped.pl:720-744) is a few
// byte-parallel adds: (x + 0x7E) & 0x80 per byte says "count >= 2", + 0x7D ">= 3", + 0x1B "> 100".
// Linear probing; about half of the records open a new group (a group is mostly one row per
// sample), so there is no separate fast path.  The pass over the slots that applies the test also
// resets them, so a bucket costs one sweep of the table, not two.  Buckets are dealt round-robin
// and the first records of the next bucket are requested before the sweep of the current one.
// A bucket whose groups do not fit the table is processed in rounds of a secondary hash, like
// k_bucket_count (bucket.cu).
template <typename KT, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS <= 512 ? 2 : 1)
    k_join_bucket(const uint64_t* __restrict__ recs, const unsigned long long* __restrict__ offB, unsigned n_buckets,
                  int kp, int a, int b, int log_slots, const unsigned long long* __restrict__ first,
                  EdgeRec* __restrict__ edges, uint64_t cap, unsigned long long* __restrict__ n_edges,
                  int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char jb_smem[];
    const unsigned slots = 1u << log_slots;
    KT* tkey = reinterpret_cast<KT*>(jb_smem);
    uint32_t* tc = reinterpret_cast<uint32_t*>(tkey + slots);
    uint32_t* tt = tc + slots;
    __shared__ unsigned s_distinct, s_overflow;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const int nbitsP = 2 * kp, rs = nbitsP - a, r = rs - b;
    const KT EMPTY = ~KT(0);
    const KT maskR = r ? (KT)((((uint64_t)1) << r) - 1) : KT(0);
    const unsigned smask = slots - 1;
    const unsigned limit = slots - (slots >> 2);
    const int hshift = 32 - log_slots;
    constexpr unsigned long long STEP = (unsigned long long)THREADS * JB_UNROLL;
    for (unsigned i = tid; i < slots; i += THREADS) {
        tkey[i] = EMPTY;
        tc[i] = 0;
        tt[i] = 0;
    }
    if (tid == 0) {
        s_distinct = 0;
        s_overflow = 0;
    }
    __syncthreads();
    auto load_recs = [&](unsigned long long i0, unsigned long long hi, uint64_t (&rec)[JB_UNROLL]) {
#pragma unroll
        for (int u = 0; u < JB_UNROLL; ++u) {
            const unsigned long long i = i0 + (unsigned long long)u * THREADS + tid;
            rec[u] = i < hi ? recs[i] : ~0ull;  // a record never has all bits set (count <= 127)
        }
    };
    unsigned bk = blockIdx.x;
    unsigned long long lo = 0, hi = 0;
    uint64_t pre[JB_UNROLL];  // the first records of the bucket that comes next
    bool pre_ok = false;
    if (bk < n_buckets) {
        lo = offB[bk];
        hi = offB[bk + 1];
        load_recs(lo, hi, pre);
        pre_ok = true;
    }
    while (bk < n_buckets) {
        const unsigned nbk = bk + gridDim.x;
        unsigned long long nlo = 0, nhi = 0;
        if (nbk < n_buckets) {
            nlo = offB[nbk];
            nhi = offB[nbk + 1];
        }
        const uint64_t hbase = ((uint64_t)(bk >> 8) << rs) | ((uint64_t)(bk & 255u) << r);
        // a group is one to eight rows, mostly two: a bucket with more rows than the table takes
        // groups -- twice the expected number of groups, so that a round rarely outgrows the table
        // and has to be done again -- is split into rounds of a secondary hash from the start (a
        // genome several times the size the 2^16 buckets were laid out for); an overflowing round
        // is split again
        unsigned R = 1, rd = 0;
        while ((hi - lo) > (unsigned long long)limit * R && R < (1u << 20)) R <<= 1;
        const unsigned r_start = R;
        bool use_pre = pre_ok;
        pre_ok = false;
        while (hi > lo) {
            while (R < r_start) R <<= 1;  // descend to the first half, as an overflow would
            // ---- every row ORs its count into the (sample, base) byte of its group ----
            unsigned my_new = 0;
            bool stuck = false;
            for (unsigned long long i0 = lo; i0 < hi; i0 += STEP) {
                // a round that has outgrown the table is abandoned at once, not after filling it
                if (__any_sync(FULL_MASK, *(volatile unsigned*)&s_overflow != 0)) break;
                uint64_t rec[JB_UNROLL];
                if (use_pre && i0 == lo) {
#pragma unroll
                    for (int u = 0; u < JB_UNROLL; ++u) rec[u] = pre[u];
                } else {
                    load_recs(i0, hi, rec);
                }
#pragma unroll
                for (int u = 0; u < JB_UNROLL; ++u) {
                    if (rec[u] == ~0ull) continue;
                    const KT x = (KT)((KT)(rec[u] >> JR_PAYLOAD_BITS) & maskR);
                    if (R > 1 && (((jb_fold(x) * 0x85EBCA77u) >> 10) & (R - 1)) != rd) continue;
                    unsigned h = (jb_fold(x) * 0x9E3779B1u) >> hshift;
                    unsigned probes = 0;
                    for (;;) {
                        KT q = *(volatile KT*)&tkey[h];
                        if (q == EMPTY) {
                            q = jb_cas(&tkey[h], EMPTY, x);
                            if (q == EMPTY) {
                                ++my_new;
                                q = x;
                            }
                        }
                        if (q == x) {
                            const unsigned cnt = (unsigned)rec[u] & 127u, base = (unsigned)(rec[u] >> 8) & 3u;
                            atomicOr(((rec[u] >> 7) & 1u) ? &tt[h] : &tc[h], cnt << (8 * base));
                            break;
                        }
                        h = (h + 1) & smask;
                        if (++probes > 512u) {  // a table this crowded is abandoned (rounds below)
                            stuck = true;
                            break;
                        }
                    }
                }
                const unsigned wn = __reduce_add_sync(FULL_MASK, my_new);
                my_new = 0;
                if (lane == 0 && wn && atomicAdd(&s_distinct, wn) + wn > limit) s_overflow = 1;
                if (stuck) s_overflow = 1;
            }
            use_pre = false;
            __syncthreads();
            if (s_overflow || s_distinct > limit) {
                __syncthreads();  // everybody has seen the flags before they are cleared again
                for (unsigned i = tid; i < slots; i += THREADS) {
                    tkey[i] = EMPTY;
                    tc[i] = 0;
                    tt[i] = 0;
                }
                if (tid == 0) {
                    s_distinct = 0;
                    s_overflow = 0;
                }
                __syncthreads();
                if (R >= (1u << 20)) {
                    if (tid == 0) atomicExch(err, 3);
                    break;
                }
                R <<= 1;
                continue;
            }
            const bool last_round = R == 1 || rd == R - 1;
            if (last_round && nhi > nlo) {  // in flight during the sweep below
                load_recs(nlo, nhi, pre);
                pre_ok = true;
            }
            // ---- joinKmerSub (ped.pl:703-782) on every group; the slot is empty again afterwards ----
            for (unsigned sl = tid; sl < slots; sl += THREADS) {
                const KT q = tkey[sl];
                if (q == EMPTY) continue;
                const uint32_t pc = tc[sl], pt = tt[sl];
                tkey[sl] = EMPTY;
                tc[sl] = 0;
                tt[sl] = 0;
                // pass 1, bytewise (bit 7 of a byte = the condition for that base)
                const uint32_t HI = 0x80808080u;
                const uint32_t ge3c = (pc + 0x7D7D7D7Du) & HI, ge3t = (pt + 0x7D7D7D7Du) & HI;
                const uint32_t repc = (pc + 0x1B1B1B1Bu) & HI, rept = (pt + 0x1B1B1B1Bu) & HI;
                const uint32_t sa = ge3c & ~repc, sb = ge3t & ~rept;  // allele sets: 3 <= count <= 100
                // an absent sample has no allele, so this also makes `join` an inner join (ped.pl:709)
                if (sa == sb || !sa || !sb) continue;
                if (repc | rept) continue;  // a count above 100 anywhere: repeat
                const uint32_t ge2c = (pc + 0x7E7E7E7Eu) & HI, ge2t = (pt + 0x7E7E7E7Eu) & HI;
                if (__popc(ge2c) > 2 || __popc(ge2t) > 2) continue;  // fewer than two bases with count <= 1
                if (__popc(sa & ~ge2t) != 1 && __popc(sb & ~ge2c) != 1) continue;
                // ---- rare from here on: pass 2 on the counts ----
                uint32_t c[4], t[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    c[x] = (pc >> (8 * x)) & 255u;
                    t[x] = (pt >> (8 * x)) & 255u;
                }
                const uint64_t P = kmer_unhash(hbase | (uint64_t)q, nbitsP);
                const unsigned tag = (unsigned)(P >> (nbitsP - 6)) & 63u;
                const unsigned long long fc = first[tag], ft = first[64 + tag];
                if ((fc != ~0ull && (fc >> 2) == P) || (ft != ~0ull && (ft >> 2) == P)) continue;  // F7 rows: host
                uint8_t cont, alt;
                if (!edge_test(c, t, &cont, &alt)) continue;
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
            if (tid == 0) s_distinct = 0;
            __syncthreads();  // the table is empty again
            while (R > 1 && rd >= (R >> 1)) {
                rd -= R >> 1;
                R >>= 1;
            }
            if (R == 1) break;
            rd += R >> 1;
        }
        if (hi == lo && nhi > nlo) {  // an empty bucket: nothing was swept
            load_recs(nlo, nhi, pre);
            pre_ok = true;
        }
        bk = nbk;
        lo = nlo;
        hi = nhi;
    }
}

// The groups of n_p given (k-1)-mers, read from their buckets (host-side F7 rows).  One CTA per P.
__global__ void __launch_bounds__(256) k_join_lookup(const uint64_t* __restrict__ recs,
                                                    const unsigned long long* __restrict__ offB, int kp, int a, int b,
                                                    const uint64_t* __restrict__ p_list, GroupRec* __restrict__ out) {
    __shared__ GroupRec g;
    const int nbitsP = 2 * kp, rs = nbitsP - a, r = rs - b;
    const uint64_t P = p_list[blockIdx.x];
    const uint64_t h = kmer_hash(P, nbitsP);
    const unsigned bk = (unsigned)(h >> rs) * 256u + ((unsigned)(h >> r) & ((1u << b) - 1));
    const uint64_t want = h & ((1ull << rs) - 1);
    if (threadIdx.x == 0) {
        g.p = P;
        for (int x = 0; x < 4; ++x) { g.c[x] = g.t[x] = 0; g.hasc[x] = g.hast[x] = 0; }
    }
    __syncthreads();
    for (unsigned long long i = offB[bk] + threadIdx.x; i < offB[bk + 1]; i += blockDim.x) {
        const uint64_t rec = recs[i];
        if ((rec >> JR_PAYLOAD_BITS) != want) continue;
        const unsigned x = (unsigned)(rec >> 8) & 3u, cnt = (unsigned)rec & 127u;
        if ((rec >> 7) & 1u) { g.t[x] = cnt; g.hast[x] = 1; }
        else { g.c[x] = cnt; g.hasc[x] = 1; }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = g;
}


// Order-independent digest of a strand table: sum over rows of mix64(K * phi + count) modulo 2^64.
__global__ void __launch_bounds__(256) k_rows_digest(const uint64_t* __restrict__ key, const uint32_t* __restrict__ val,
                                                    uint64_t n, unsigned long long* __restrict__ out) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        acc += mix64(key[i] * 0x9E3779B97F4A7C15ull + (uint64_t)(val[i] & 0x7fffffffu));
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, d);
    if ((threadIdx.x & 31u) == 0 && acc) atomicAdd(out, acc);
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


void launch_rows_count(const uint64_t* row_key, uint64_t n, int sample, int k, unsigned long long* hist256,
                       unsigned long long* first, cudaStream_t st) {
    if (!n) return;
    const int kp = k - 1, a = kmer_bits_a(kp);
    const uint64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_rows_count<<<grid, 256, 0, st>>>(row_key, n, (unsigned)sample, k, 2 * kp, 2 * kp - a, hist256, first);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_rows_scatter(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int sample, int k, int nranks,
                         unsigned long long* cur, const PeerPtrs& dst, cudaStream_t st) {
    if (!n) return;
    const int kp = k - 1, a = kmer_bits_a(kp);
    const uint64_t tiles = (n + JS_TILE - 1) / JS_TILE;
    const unsigned grid = (unsigned)(tiles < (uint64_t)kNumSMs * 2 ? tiles : (uint64_t)kNumSMs * 2);
    const char* bulk = getenv("PEDK_PART_BULK");
    if (bulk && atoi(bulk))
        k_rows_scatter<true><<<grid, JS_THREADS, 0, st>>>(row_key, row_val, n, (unsigned)sample, 2 * kp, a, nranks, cur, dst);
    else
        k_rows_scatter<false><<<grid, JS_THREADS, 0, st>>>(row_key, row_val, n, (unsigned)sample, 2 * kp, a, nranks, cur, dst);
    PEDK_CUDA_TRY(cudaGetLastError());
}

int join_record_shift_b(int k, int b) {
    const int kp = k - 1, a = kmer_bits_a(kp);
    return 2 * kp - a - b + JR_PAYLOAD_BITS;
}

void launch_join_bucket(const uint64_t* recs, const unsigned long long* offB, unsigned n_buckets, int k, int b,
                        int log_slots, const unsigned long long* first, const uint32_t* pass1_bits, EdgeRec* edges,
                        uint64_t cap, unsigned long long* n_edges, unsigned* ticket, int* err, cudaStream_t st) {
    if (!n_buckets) return;
    const int kp = k - 1, a = kmer_bits_a(kp);
    const int r = 2 * kp - a - b;
    if (log_slots < 10) log_slots = 10;
    if (log_slots > 14) log_slots = 14;
    const bool wide = r > 31;
    if (wide && log_slots > 13) log_slots = 13;
    const size_t smem = ((size_t)1 << log_slots) * ((wide ? 8 : 4) + 8);
    const int threads_env = getenv("PEDK_JOIN_THREADS") ? atoi(getenv("PEDK_JOIN_THREADS")) : 0;
    // two CTAs of 512 threads per SM while two tables fit its shared memory, else one of 1024
    const bool small = threads_env ? threads_env <= 512 : smem <= 100 * 1024;
    (void)pass1_bits;  // the kernel evaluates pass 1 on the packed counts; the table serves k_edges
    (void)ticket;
    unsigned grid = kNumSMs * ((small && smem <= 100 * 1024) ? 2 : 1);
    if (grid > n_buckets) grid = n_buckets;
#define PEDK_JB(KT, TH)                                                                                        \
    do {                                                                                                       \
        ensure_dyn_smem(k_join_bucket<KT, TH>, smem);                                                          \
        k_join_bucket<KT, TH><<<grid, TH, smem, st>>>(recs, offB, n_buckets, kp, a, b, log_slots, first, edges, cap, \
                                                      n_edges, err);                                           \
    } while (0)
    if (wide) {
        if (small) PEDK_JB(uint64_t, 512);
        else PEDK_JB(uint64_t, 1024);
    } else {
        if (small) PEDK_JB(uint32_t, 512);
        else PEDK_JB(uint32_t, 1024);
    }
#undef PEDK_JB
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_join_lookup(const uint64_t* recs, const unsigned long long* offB, int k, int b, const uint64_t* p_list,
                        int n_p, GroupRec* out, cudaStream_t st) {
    if (!n_p) return;
    const int kp = k - 1;
    k_join_lookup<<<n_p, 256, 0, st>>>(recs, offB, kp, kmer_bits_a(kp), b, p_list, out);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_rows_digest(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, unsigned long long* out,
                        cudaStream_t st) {
    if (!n) return;
    const uint64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_rows_digest<<<grid, 256, 0, st>>>(row_key, row_val, n, out);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
