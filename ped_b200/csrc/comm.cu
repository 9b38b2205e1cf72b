// Multi-GPU plumbing: one rank per process, NCCL over NVLink 5 / NVSwitch.
//
// ped.pl's own "collective" is a 64x64 matrix of files between countKmerSub and
// countKmerMerge (ped.pl:894-918, 867-886).  Here three record
// streams are exchanged, each by an owner function that every rank evaluates
// identically:
//   reads      (W words)  -> owner = hash of the canonical packed read: exact dedup stays local
//   k-mers     (1 word)   -> owner = hash of the canonical k-mer: all instances of a k-mer meet
//   table rows (2 words)  -> owner = rank that holds the 3-nt bucket of K (contiguous bucket
//                            ranges balanced on the summed histogram of both samples), so no
//                            (k-1)-mer group and no first-of-bucket row straddles two ranks
// An exchange is: count per destination -> all-gather of the count matrix -> scatter into
// per-destination runs -> grouped ncclSend/ncclRecv (all-to-all-v).  NCCL is bound with
// dlopen so a single-GPU host needs no libnccl.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include "comm.h"

namespace pedk {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
#define PEDK_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
            PEDK_SYM(GetUniqueId, "ncclGetUniqueId");
            PEDK_SYM(CommInitRank, "ncclCommInitRank");
            PEDK_SYM(CommDestroy, "ncclCommDestroy");
            PEDK_SYM(AllReduce, "ncclAllReduce");
            PEDK_SYM(AllGather, "ncclAllGather");
            PEDK_SYM(Send, "ncclSend");
            PEDK_SYM(Recv, "ncclRecv");
            PEDK_SYM(GroupStart, "ncclGroupStart");
            PEDK_SYM(GroupEnd, "ncclGroupEnd");
            PEDK_SYM(GetErrorString, "ncclGetErrorString");
#undef PEDK_SYM
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllGather ||
        !api.AllReduce || !api.GroupStart || !api.GroupEnd)
        throw PedkError(PEDK_E_NCCL, "libnccl.so.2 not found (needed only for multi-GPU runs)");
    return &api;
}

void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) {
        NcclApi* a = nccl();
        throw PedkError(PEDK_E_NCCL, std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
    }
}

ncclComm_t comm_of(pedk_ctx* ctx) { return static_cast<ncclComm_t>(ctx->nccl_comm); }

__device__ __forceinline__ int dest_of(const uint64_t* rec, const DestSpec& s) {
    if (s.mode == DEST_ROW) return s.owner[rec[0] >> s.bucket_shift];
    if (s.mode == DEST_READ) return read_owner_rank(rec, s.stride, s.nranks);
    return kmer_owner_rank(rec[0], s.k, s.nranks);
}

__global__ void __launch_bounds__(256) k_dest_count(const uint64_t* __restrict__ rec, uint64_t n, DestSpec s,
                                                   unsigned long long* __restrict__ counts) {
    const unsigned lane = lane_id();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n + 31) & ~31ull;  // whole warps iterate together
    unsigned long long mine = 0;                 // lane r accumulates the count for destination r
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const int d = i < n ? dest_of(rec + i * s.stride, s) : -1;
        for (int r = 0; r < s.nranks; ++r) {
            unsigned m = __ballot_sync(0xffffffffu, d == r);
            if (lane == (unsigned)r) mine += __popc(m);
        }
    }
    if ((int)lane < s.nranks && mine) atomicAdd(&counts[lane], mine);
}

constexpr int XS_ITEMS = 8;
constexpr int XS_TILE = 256 * XS_ITEMS;

// Scatter into per-destination runs.  One global atomic per destination per 2048-record
// tile (a per-warp atomic on R shared addresses would serialise in L2: 150 ms for 1.2 G
// records, measured in round 1); inside the tile positions come from ballots.
__global__ void __launch_bounds__(256) k_dest_scatter(const uint64_t* __restrict__ rec, uint64_t n, DestSpec s,
                                                     unsigned long long* __restrict__ cursors,
                                                     uint64_t* __restrict__ out) {
    __shared__ unsigned w_cnt[8][PEDK_MAX_RANKS];
    __shared__ unsigned long long b_base[PEDK_MAX_RANKS];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned lt = lanemask_lt();
    const uint64_t n_tiles = (n + XS_TILE - 1) / XS_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t warp_base = tile * XS_TILE + (uint64_t)warp * (32 * XS_ITEMS);
        int d[XS_ITEMS];
        unsigned mine = 0;  // lane r counts the records of this warp that go to rank r
#pragma unroll
        for (int j = 0; j < XS_ITEMS; ++j) {
            const uint64_t i = warp_base + 32 * j + lane;
            d[j] = i < n ? dest_of(rec + i * s.stride, s) : -1;
            for (int r = 0; r < s.nranks; ++r) {
                unsigned m = __ballot_sync(0xffffffffu, d[j] == r);
                if (lane == (unsigned)r) mine += __popc(m);
            }
        }
        if ((int)lane < s.nranks) w_cnt[warp][lane] = mine;
        __syncthreads();
        if ((int)tid < s.nranks) {
            unsigned tot = 0;
            for (int w = 0; w < 8; ++w) {
                unsigned c = w_cnt[w][tid];
                w_cnt[w][tid] = tot;
                tot += c;
            }
            b_base[tid] = tot ? atomicAdd(&cursors[tid], (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        unsigned long long running = (int)lane < s.nranks ? b_base[lane] + w_cnt[warp][lane] : 0ull;
#pragma unroll
        for (int j = 0; j < XS_ITEMS; ++j) {
            unsigned mym = 0, add = 0;
            for (int r = 0; r < s.nranks; ++r) {
                unsigned m = __ballot_sync(0xffffffffu, d[j] == r);
                if (d[j] == r) mym = m;
                if (lane == (unsigned)r) add = __popc(m);
            }
            const unsigned long long base = __shfl_sync(0xffffffffu, running, d[j] < 0 ? 0 : d[j]);
            running += add;
            if (d[j] >= 0) {
                const uint64_t i = warp_base + 32 * j + lane;
                const uint64_t* src = rec + i * s.stride;
                uint64_t* dst = out + (base + __popc(mym & lt)) * s.stride;
                for (int w = 0; w < s.stride; ++w) dst[w] = src[w];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_pal_flags(const uint64_t* __restrict__ words, int W, int L, uint64_t n,
                                                  uint8_t* __restrict__ pal) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t fw[PEDK_MAX_WORDS], rc[PEDK_MAX_WORDS];
    for (int w = 0; w < W; ++w) fw[w] = words[i * W + w];
    read_revcomp(fw, W, L, rc);
    pal[i] = words_cmp(fw, rc, W) == 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_rows_pack(const uint64_t* __restrict__ key, const uint32_t* __restrict__ val,
                                                  uint64_t n, uint64_t* __restrict__ rec) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        rec[2 * i] = key[i];
        rec[2 * i + 1] = val[i];
    }
}

__global__ void __launch_bounds__(256) k_rows_unpack(const uint64_t* __restrict__ rec, uint64_t n,
                                                    uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        key[i] = rec[2 * i];
        val[i] = (uint32_t)rec[2 * i + 1];
    }
}

__global__ void __launch_bounds__(256) k_bucket_hist(const uint64_t* __restrict__ key, uint64_t n, int shift,
                                                    unsigned long long* __restrict__ hist) {
    __shared__ unsigned h[64];
    if (threadIdx.x < 64) h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&h[(unsigned)(key[i] >> shift) & 63u], 1u);
    __syncthreads();
    if (threadIdx.x < 64 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// Table rows -> the rank that holds their 3-nt bucket, in ONE kernel: a tile of rows is grouped
// by destination in shared memory and every destination's run is stored straight into that
// rank's (key, value) arrays (peer memory over NVLink for the other ranks), at the slots this
// rank reserved with one atomic per (tile, destination).  Replaces copy + pack + count +
// scatter + ncclSend/Recv + unpack.
constexpr int RR_THREADS = 256;
constexpr int RR_ITEMS = 8;
constexpr int RR_TILE = RR_THREADS * RR_ITEMS;

struct RowPeers {
    uint64_t* key[PEDK_MAX_RANKS];
    uint32_t* val[PEDK_MAX_RANKS];
    uint8_t owner[64];
};

__global__ void __launch_bounds__(RR_THREADS) k_rows_route(const uint64_t* __restrict__ key,
                                                          const uint32_t* __restrict__ val, uint64_t n,
                                                          uint32_t or_bits, int shift, int nranks, RowPeers peers,
                                                          unsigned long long* __restrict__ cursors) {
    __shared__ uint64_t s_key[RR_TILE];
    __shared__ uint32_t s_val[RR_TILE];
    __shared__ uint8_t s_dst[RR_TILE];
    __shared__ unsigned cnt[PEDK_MAX_RANKS], off[PEDK_MAX_RANKS + 1];
    __shared__ unsigned long long base[PEDK_MAX_RANKS];
    const unsigned tid = threadIdx.x;
    const uint64_t n_tiles = (n + RR_TILE - 1) / RR_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (tid < PEDK_MAX_RANKS) cnt[tid] = 0;
        __syncthreads();
        uint64_t k[RR_ITEMS];
        uint32_t v[RR_ITEMS];
        unsigned meta[RR_ITEMS];  // dest << 16 | rank inside (tile, dest); ~0: no row
#pragma unroll
        for (int j = 0; j < RR_ITEMS; ++j) {
            const uint64_t i = tile * RR_TILE + (uint64_t)j * RR_THREADS + tid;
            meta[j] = 0xffffffffu;
            if (i < n) {
                k[j] = key[i];
                v[j] = val[i] | or_bits;
                const unsigned d = peers.owner[(unsigned)(k[j] >> shift) & 63u];
                meta[j] = (d << 16) | atomicAdd(&cnt[d], 1u);
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned acc = 0;
            for (int d = 0; d < nranks; ++d) {
                off[d] = acc;
                acc += cnt[d];
            }
            off[nranks] = acc;
        }
        if ((int)tid < nranks) base[tid] = cnt[tid] ? atomicAdd(&cursors[tid], (unsigned long long)cnt[tid]) : 0ull;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RR_ITEMS; ++j) {
            if (meta[j] != 0xffffffffu) {
                const unsigned d = meta[j] >> 16;
                const unsigned p = off[d] + (meta[j] & 0xffffu);
                s_key[p] = k[j];
                s_val[p] = v[j];
                s_dst[p] = (uint8_t)d;
            }
        }
        __syncthreads();
        const unsigned total = off[nranks];
        for (unsigned i = tid; i < total; i += RR_THREADS) {
            const unsigned d = s_dst[i];
            const unsigned long long o = base[d] + (i - off[d]);
            peers.key[d][o] = s_key[i];
            peers.val[d][o] = s_val[i];
        }
        __syncthreads();
    }
}

unsigned grid_for(uint64_t n) {
    uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    return (unsigned)(blocks < (uint64_t)kNumSMs * 16 ? blocks : (uint64_t)kNumSMs * 16);
}

}  // namespace

void comm_destroy(pedk_ctx* ctx) {
    comm_ipc_destroy(ctx);
    if (ctx->nccl_comm) {
        try {
            nccl()->CommDestroy(comm_of(ctx));
        } catch (...) {
        }
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
}

void comm_allreduce_sum(pedk_ctx* ctx, unsigned long long* dev, size_t n) {
    if (ctx->nranks <= 1 || !n) return;
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 0);
    nccl_check(nccl()->AllReduce(dev, dev, n, ncclUint64, ncclSum, comm_of(ctx), ctx->stream), "ncclAllReduce");
}

void comm_allreduce_sum_u32(pedk_ctx* ctx, uint32_t* dev, size_t n) {
    if (ctx->nranks <= 1 || !n) return;
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 0);
    nccl_check(nccl()->AllReduce(dev, dev, n, ncclUint32, ncclSum, comm_of(ctx), ctx->stream), "ncclAllReduce (u32)");
}

void comm_allreduce_min(pedk_ctx* ctx, unsigned long long* dev, size_t n) {
    if (ctx->nranks <= 1 || !n) return;
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 0);
    nccl_check(nccl()->AllReduce(dev, dev, n, ncclUint64, ncclMin, comm_of(ctx), ctx->stream), "ncclAllReduce (min)");
}

void comm_allreduce_max(pedk_ctx* ctx, unsigned long long* dev, size_t n) {
    if (ctx->nranks <= 1 || !n) return;
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 0);
    nccl_check(nccl()->AllReduce(dev, dev, n, ncclUint64, ncclMax, comm_of(ctx), ctx->stream), "ncclAllReduce (max)");
}

uint64_t comm_allreduce_sum_host(pedk_ctx* ctx, uint64_t v) {
    if (ctx->nranks <= 1) return v;
    unsigned long long x = v;
    PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->scalars + 12, &x, sizeof x, cudaMemcpyHostToDevice, ctx->stream));
    comm_allreduce_sum(ctx, ctx->scalars + 12, 1);
    PEDK_CUDA_TRY(cudaMemcpyAsync(&x, ctx->scalars + 12, sizeof x, cudaMemcpyDeviceToHost, ctx->stream));
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return x;
}

// Contiguous ranges of the 64 buckets, one per rank, as even as the histogram allows.
void plan_bucket_owners(const uint64_t* hist64, int nranks, uint8_t* owner64) {
    uint64_t total = 0;
    for (int b = 0; b < 64; ++b) total += hist64[b];
    uint64_t acc = 0;
    int r = 0;
    for (int b = 0; b < 64; ++b) {
        // move on once this rank holds its share, but keep a bucket for every remaining rank
        if (r < nranks - 1 && total > 0 && acc * (uint64_t)nranks >= total * (uint64_t)(r + 1)) ++r;
        if (64 - b <= nranks - 1 - r) r = nranks - 1 - (63 - b);
        owner64[b] = (uint8_t)r;
        acc += hist64[b];
    }
}

void exchange_records(pedk_ctx* ctx, const uint64_t* rec, uint64_t n, DestSpec spec, DBuf<uint64_t>* recv,
                      uint64_t* n_recv) {
    cudaStream_t st = ctx->stream;
    const int R = ctx->nranks, me = ctx->rank;
    spec.nranks = R;
    NcclApi* a = nccl();
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 2);
    // counts per destination, then the whole R x R matrix on every rank
    DBuf<unsigned long long> counts(ctx, (size_t)PEDK_MAX_RANKS * (PEDK_MAX_RANKS + 2));
    unsigned long long* my_counts = counts.p;
    unsigned long long* all_counts = counts.p + PEDK_MAX_RANKS;
    unsigned long long* cursors = counts.p + PEDK_MAX_RANKS * (PEDK_MAX_RANKS + 1);
    PEDK_CUDA_TRY(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
    if (n) k_dest_count<<<grid_for(n), 256, 0, st>>>(rec, n, spec, my_counts);
    PEDK_CUDA_TRY(cudaGetLastError());
    nccl_check(a->AllGather(my_counts, all_counts, PEDK_MAX_RANKS, ncclUint64, comm_of(ctx), st), "ncclAllGather");
    std::vector<unsigned long long> h((size_t)PEDK_MAX_RANKS * R);
    PEDK_CUDA_TRY(cudaMemcpyAsync(h.data(), all_counts, h.size() * 8, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    trace(ctx, "  exchange: counts");
    // h[src * MAX + dst]
    std::vector<unsigned long long> send_off(R + 1, 0), recv_off(R + 1, 0);
    for (int d = 0; d < R; ++d) send_off[d + 1] = send_off[d] + h[(size_t)me * PEDK_MAX_RANKS + d];
    for (int s = 0; s < R; ++s) recv_off[s + 1] = recv_off[s] + h[(size_t)s * PEDK_MAX_RANKS + me];
    if (send_off[R] != n) throw PedkError(PEDK_E_INTERNAL, "exchange: destination counts do not add up");
    *n_recv = recv_off[R];
    DBuf<uint64_t> sendbuf(ctx, n * spec.stride);
    *recv = DBuf<uint64_t>(ctx, *n_recv * spec.stride);
    PEDK_CUDA_TRY(cudaMemcpyAsync(cursors, send_off.data(), R * 8, cudaMemcpyHostToDevice, st));
    if (n) k_dest_scatter<<<grid_for(n), 256, 0, st>>>(rec, n, spec, cursors, sendbuf.p);
    PEDK_CUDA_TRY(cudaGetLastError());
    if (ctx->trace) {
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        trace(ctx, "  exchange: scatter");
    }
    nccl_check(a->GroupStart(), "ncclGroupStart");
    for (int p = 0; p < R; ++p) {
        if (p == me) continue;
        size_t ns = (size_t)(send_off[p + 1] - send_off[p]) * spec.stride;
        size_t nr = (size_t)(recv_off[p + 1] - recv_off[p]) * spec.stride;
        if (ns) nccl_check(a->Send(sendbuf.p + send_off[p] * spec.stride, ns, ncclUint64, p, comm_of(ctx), st), "ncclSend");
        if (nr) nccl_check(a->Recv(recv->p + recv_off[p] * spec.stride, nr, ncclUint64, p, comm_of(ctx), st), "ncclRecv");
    }
    nccl_check(a->GroupEnd(), "ncclGroupEnd");
    size_t self = (size_t)(send_off[me + 1] - send_off[me]) * spec.stride;
    if (self)
        PEDK_CUDA_TRY(cudaMemcpyAsync(recv->p + recv_off[me] * spec.stride, sendbuf.p + send_off[me] * spec.stride,
                                      self * 8, cudaMemcpyDeviceToDevice, st));
    if (ctx->trace) {
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        char b[96];
        snprintf(b, sizeof b, "  exchange: send/recv (%llu records x %d words out)", (unsigned long long)n, spec.stride);
        trace(ctx, b);
    }
    // the send buffer goes back to the allocator in stream order
    ctx->stats.kernel_launches += 2;
}

}  // namespace pedk (reopened below)

namespace pedk {

// Both samples' strand rows -> the ranks that own their 3-nt buckets (see k_rows_route).
// src_key/src_val/src_n/or_bits: the row stores of the (up to two) samples.  local_hist64 /
// global_hist64: this rank's and all ranks' rows per bucket (host).  Returns this rank's rows
// in *ka / *va (unordered).
void route_rows(pedk_ctx* ctx, const uint64_t* const* src_key, const uint32_t* const* src_val, const uint64_t* src_n,
                const uint32_t* or_bits, int n_src, int k, const uint64_t* local_hist64,
                const uint64_t* global_hist64, DBuf<uint64_t>* ka, DBuf<uint32_t>* va, uint64_t* n_rows) {
    cudaStream_t st = ctx->stream;
    const int R = ctx->nranks, me = ctx->rank;
    RowPeers peers;
    plan_bucket_owners(global_hist64, R, peers.owner);
    // rows this rank sends to every rank, then the whole matrix on every rank
    unsigned long long mine[PEDK_MAX_RANKS] = {0};
    for (int b = 0; b < 64; ++b) mine[peers.owner[b]] += local_hist64[b];
    DBuf<unsigned long long> cbuf(ctx, (size_t)PEDK_MAX_RANKS * (PEDK_MAX_RANKS + 2));
    std::vector<unsigned long long> h;
    PEDK_CUDA_TRY(cudaMemcpyAsync(cbuf.p, mine, sizeof mine, cudaMemcpyHostToDevice, st));
    comm_count_matrix(ctx, cbuf.p, cbuf.p + PEDK_MAX_RANKS, &h);
    uint64_t recv = 0;
    unsigned long long start[PEDK_MAX_RANKS] = {0};
    for (int src = 0; src < R; ++src) recv += h[(size_t)src * PEDK_MAX_RANKS + me];
    for (int d = 0; d < R; ++d)
        for (int src = 0; src < me; ++src) start[d] += h[(size_t)src * PEDK_MAX_RANKS + d];
    *n_rows = recv;
    *ka = DBuf<uint64_t>(ctx, recv);
    *va = DBuf<uint32_t>(ctx, recv);
    void* pk[PEDK_MAX_RANKS];
    void* pv[PEDK_MAX_RANKS];
    comm_peer_pointers(ctx, ka->p, pk);
    comm_peer_pointers(ctx, va->p, pv);
    for (int d = 0; d < R; ++d) {
        peers.key[d] = static_cast<uint64_t*>(pk[d]);
        peers.val[d] = static_cast<uint32_t*>(pv[d]);
    }
    unsigned long long* cursors = cbuf.p + PEDK_MAX_RANKS * (PEDK_MAX_RANKS + 1);
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 2);
    PEDK_CUDA_TRY(cudaMemcpyAsync(cursors, start, sizeof start, cudaMemcpyHostToDevice, st));
    comm_barrier(ctx);  // every rank's receive arrays exist and are free
    for (int i = 0; i < n_src; ++i) {
        if (!src_n[i]) continue;
        const uint64_t tiles = (src_n[i] + RR_TILE - 1) / RR_TILE;
        const unsigned grid = (unsigned)(tiles < (uint64_t)kNumSMs * 8 ? tiles : (uint64_t)kNumSMs * 8);
        k_rows_route<<<grid, RR_THREADS, 0, st>>>(src_key[i], src_val[i], src_n[i], or_bits[i], 2 * k - 6, R, peers,
                                                  cursors);
        PEDK_CUDA_TRY(cudaGetLastError());
    }
    comm_barrier(ctx);  // every rank's stores have landed
    if (ctx->trace) {
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        trace(ctx, "  rows: route + NVLink stores");
    }
}

// ---- NVLink peer memory through CUDA IPC --------------------------------------------------
// Every rank publishes the IPC handle of one of its device blocks; the others map it once
// (the caching allocator hands the same block to the same stage on every step, so the
// mapping is reused) and can then store into it from their own kernels.
// Peer blocks mapped into this process.  The allocators hand the same blocks out step after step, so
// the set of live handles is small; a mapping that has not been asked for during the last kIpcKeep
// requests for its peer is closed (a block its owner no longer uses would otherwise stay mapped --
// and allocated -- for the life of the context).
constexpr size_t kIpcKeep = 12;
struct IpcCache {
    struct Entry {
        cudaIpcMemHandle_t h;
        void* p;
        uint64_t last_use;
    };
    std::vector<Entry> entries[PEDK_MAX_RANKS];
    uint64_t clock = 0;
    std::unordered_map<void*, cudaIpcMemHandle_t> own;  // handles of this rank's exported blocks
};

static IpcCache* ipc_cache(pedk_ctx* ctx) {
    if (!ctx->ipc_cache) ctx->ipc_cache = new IpcCache();
    return static_cast<IpcCache*>(ctx->ipc_cache);
}

void comm_ipc_destroy(pedk_ctx* ctx) {
    if (!ctx->ipc_cache) return;
    IpcCache* c = static_cast<IpcCache*>(ctx->ipc_cache);
    for (auto& v : c->entries)
        for (auto& e : v) cudaIpcCloseMemHandle(e.p);
    delete c;
    ctx->ipc_cache = nullptr;
}

// my_block must be the base of a block of the context's allocator (a whole cudaMalloc)
void comm_peer_pointers(pedk_ctx* ctx, void* my_block, void** peers) {
    const int R = ctx->nranks, me = ctx->rank;
    cudaStream_t st = ctx->stream;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DBuf<unsigned long long> buf(ctx, (size_t)8 * (PEDK_MAX_RANKS + 1));
    // the handle of a block is asked for once (a driver call under the global lock, like cudaMemGetInfo);
    // an exported block stays with the allocator until the context closes, so the handle stays valid
    cudaIpcMemHandle_t mine;
    IpcCache* hc = ipc_cache(ctx);
    auto own = hc->own.find(my_block);
    if (own == hc->own.end()) {
        PEDK_CUDA_TRY(cudaIpcGetMemHandle(&mine, my_block));
        hc->own[my_block] = mine;
        dev_mark_exported(ctx, my_block);
    } else {
        mine = own->second;
    }
    PEDK_CUDA_TRY(cudaMemcpyAsync(buf.p, &mine, 64, cudaMemcpyHostToDevice, st));
    nccl_check(nccl()->AllGather(buf.p, buf.p + 8, 8, ncclUint64, comm_of(ctx), st), "ncclAllGather (ipc handles)");
    std::vector<cudaIpcMemHandle_t> all((size_t)R);
    PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), buf.p + 8, (size_t)R * 64, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    IpcCache* c = ipc_cache(ctx);
    for (int r = 0; r < R; ++r) {
        if (r == me) {
            peers[r] = my_block;
            continue;
        }
        void* p = nullptr;
        ++c->clock;
        for (auto& e : c->entries[r])
            if (memcmp(&e.h, &all[(size_t)r], 64) == 0) {
                p = e.p;
                e.last_use = c->clock;
            }
        if (!p) {
            // every kernel of this rank that stored through an older mapping has finished (the stream was
            // synchronised above): the least recently used ones can go
            while (c->entries[r].size() >= kIpcKeep) {
                size_t old = 0;
                for (size_t i = 1; i < c->entries[r].size(); ++i)
                    if (c->entries[r][i].last_use < c->entries[r][old].last_use) old = i;
                cudaIpcCloseMemHandle(c->entries[r][old].p);
                c->entries[r].erase(c->entries[r].begin() + (long)old);
            }
            cudaError_t err = cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
            if (err != cudaSuccess) {
                cudaGetLastError();
                throw PedkError(PEDK_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(err) +
                                                 " (peer access between the ranks' GPUs is required)");
            }
            c->entries[r].push_back({all[(size_t)r], p, c->clock});
        }
        peers[r] = p;
    }
}

// all ranks' earlier work on their streams (incl. stores into peer memory) is complete
void comm_barrier(pedk_ctx* ctx) {
    if (ctx->nranks <= 1) return;
    PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 13, 0, sizeof(unsigned long long), ctx->stream));
    comm_allreduce_sum(ctx, ctx->scalars + 13, 1);
}

void comm_allgather_u64(pedk_ctx* ctx, const unsigned long long* send, unsigned long long* recv, size_t n) {
    if (ctx->nranks <= 1) return;
    StageScope sc(ctx, PEDK_ST_EXCHANGE, 0);
    nccl_check(nccl()->AllGather(send, recv, n, ncclUint64, comm_of(ctx), ctx->stream), "ncclAllGather");
}

// R x R matrix of per-destination counts: counts_dev[PEDK_MAX_RANKS] of this rank in,
// h[src * PEDK_MAX_RANKS + dst] out (host), all ranks
void comm_count_matrix(pedk_ctx* ctx, unsigned long long* counts_dev, unsigned long long* all_dev,
                       std::vector<unsigned long long>* h) {
    cudaStream_t st = ctx->stream;
    nccl_check(nccl()->AllGather(counts_dev, all_dev, PEDK_MAX_RANKS, ncclUint64, comm_of(ctx), st), "ncclAllGather");
    h->resize((size_t)PEDK_MAX_RANKS * ctx->nranks);
    PEDK_CUDA_TRY(cudaMemcpyAsync(h->data(), all_dev, h->size() * 8, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
}

void launch_pal_flags(const uint64_t* words, int W, int L, uint64_t n, uint8_t* pal, cudaStream_t st) {
    if (!n) return;
    k_pal_flags<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(words, W, L, n, pal);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_rows_pack(const uint64_t* key, const uint32_t* val, uint64_t n, uint64_t* rec, cudaStream_t st) {
    if (!n) return;
    k_rows_pack<<<grid_for(n), 256, 0, st>>>(key, val, n, rec);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_rows_unpack(const uint64_t* rec, uint64_t n, uint64_t* key, uint32_t* val, cudaStream_t st) {
    if (!n) return;
    k_rows_unpack<<<grid_for(n), 256, 0, st>>>(rec, n, key, val);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_bucket_hist(const uint64_t* key, uint64_t n, int k, unsigned long long* hist, cudaStream_t st) {
    if (!n) return;
    k_bucket_hist<<<grid_for(n), 256, 0, st>>>(key, n, 2 * k - 6, hist);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk

using namespace pedk;

extern "C" int pedk_comm_unique_id(pedk_ctx* ctx, void* id_bytes) {
    if (!ctx || !id_bytes) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    static_assert(sizeof(ncclUniqueId) <= PEDK_COMM_ID_BYTES, "id size");
    ncclUniqueId id;
    nccl_check(nccl()->GetUniqueId(&id), "ncclGetUniqueId");
    memset(id_bytes, 0, PEDK_COMM_ID_BYTES);
    memcpy(id_bytes, &id, sizeof id);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_comm_init(pedk_ctx* ctx, const void* id_bytes, int rank, int nranks) {
    if (!ctx || !id_bytes || nranks < 1 || nranks > PEDK_MAX_RANKS || rank < 0 || rank >= nranks) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    comm_destroy(ctx);
    if (nranks == 1) return PEDK_OK;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm;
    nccl_check(nccl()->CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return PEDK_OK;
    PEDK_API_END(ctx)
}

// host-side owner functions, identical to what the kernels evaluate (tests, planning)
extern "C" int pedk_owner_of_read(const uint64_t* words, int words_per_read, int nranks) {
    return read_owner_rank(words, words_per_read, nranks);
}

extern "C" int pedk_owner_of_kmer(uint64_t canonical_kmer, int k, int nranks) {
    if (k < 4 || k > 32 || nranks < 1) return PEDK_E_INVAL;
    return kmer_owner_rank(canonical_kmer, k, nranks);
}

extern "C" uint64_t pedk_kmer_hash(uint64_t kmer, int k) { return kmer_hash(kmer, 2 * k); }
extern "C" uint64_t pedk_kmer_unhash(uint64_t hash, int k) { return kmer_unhash(hash, 2 * k); }

extern "C" int pedk_plan_bucket_owners(const uint64_t* hist64, int nranks, uint8_t* owner64) {
    if (!hist64 || !owner64 || nranks < 1 || nranks > PEDK_MAX_RANKS) return PEDK_E_INVAL;
    plan_bucket_owners(hist64, nranks, owner64);
    return PEDK_OK;
}
