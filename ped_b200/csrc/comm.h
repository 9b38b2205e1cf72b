// Multi-GPU exchange primitives (internal; comm.cu).
#pragma once
#include <vector>

#include "ctx.h"

#define PEDK_MAX_RANKS 8

namespace pedk {

enum { DEST_READ = 0, DEST_KMER = 1, DEST_ROW = 2 };

struct DestSpec {
    int mode = DEST_KMER;
    int stride = 1;        // u64 words per record (reads: W, k-mers: 1, rows: 2)
    int nranks = 1;
    int bucket_shift = 0;  // rows: 2k-6, the key's first three bases
    int k = 20;            // k-mers: the owner is a function of kmer_hash(K, 2k)
    uint8_t owner[64] = {0};
};

void comm_destroy(pedk_ctx* ctx);
void comm_ipc_destroy(pedk_ctx* ctx);
// in-place sum over ranks of n u64 words in device memory (no-op on one rank)
void comm_allreduce_sum(pedk_ctx* ctx, unsigned long long* dev, size_t n);
void comm_allreduce_sum_u32(pedk_ctx* ctx, uint32_t* dev, size_t n);
void comm_allreduce_min(pedk_ctx* ctx, unsigned long long* dev, size_t n);
void comm_allreduce_max(pedk_ctx* ctx, unsigned long long* dev, size_t n);
// Elements to allocate for a block of `elems` 8-byte words that peers map through CUDA IPC: IPC
// shares whole driver allocations, so such a block is its own cudaMalloc of a multiple of 2 MiB
// (blocks under 1 MiB would otherwise be carved out of a shared slab).
inline size_t peer_block_elems(size_t elems, bool shared, size_t elem_bytes = 8) {
    if (!shared) return elems;
    const size_t g = (size_t(2) << 20) / elem_bytes;
    return elems < g ? g : (elems + g - 1) / g * g;
}
uint64_t comm_allreduce_sum_host(pedk_ctx* ctx, uint64_t v);
void plan_bucket_owners(const uint64_t* hist64, int nranks, uint8_t* owner64);
// all-to-all-v of n records; *recv holds the *n_recv records this rank owns, grouped by source rank
void exchange_records(pedk_ctx* ctx, const uint64_t* rec, uint64_t n, DestSpec spec, DBuf<uint64_t>* recv,
                      uint64_t* n_recv);
// NVLink peer memory: peers[r] = rank r's block mapped into this process (peers[me] = my_block)
void comm_peer_pointers(pedk_ctx* ctx, void* my_block, void** peers);
void comm_barrier(pedk_ctx* ctx);
// all-gather of n u64 words per rank (device memory), rank r's words land at recv + r*n
void comm_allgather_u64(pedk_ctx* ctx, const unsigned long long* send, unsigned long long* recv, size_t n);
void comm_count_matrix(pedk_ctx* ctx, unsigned long long* counts_dev, unsigned long long* all_dev,
                       std::vector<unsigned long long>* h);
// both samples' strand rows -> the ranks owning their 3-nt buckets, one fused kernel per sample
// storing into the owners' arrays over NVLink (comm.cu)
void route_rows(pedk_ctx* ctx, const uint64_t* const* src_key, const uint32_t* const* src_val, const uint64_t* src_n,
                const uint32_t* or_bits, int n_src, int k, const uint64_t* local_hist64,
                const uint64_t* global_hist64, DBuf<uint64_t>* ka, DBuf<uint32_t>* va, uint64_t* n_rows);
void launch_pal_flags(const uint64_t* words, int W, int L, uint64_t n, uint8_t* pal, cudaStream_t st);
void launch_rows_pack(const uint64_t* key, const uint32_t* val, uint64_t n, uint64_t* rec, cudaStream_t st);
void launch_rows_unpack(const uint64_t* rec, uint64_t n, uint64_t* key, uint32_t* val, cudaStream_t st);
void launch_bucket_hist(const uint64_t* key, uint64_t n, int k, unsigned long long* hist, cudaStream_t st);

}  // namespace pedk
