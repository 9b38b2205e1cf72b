// Context, device memory and stage timing for libpedk.so (internal).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../../include/pedk.h"
#include "kernels.h"
#include "util.cuh"

struct pedk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    struct Block {
        void* p;
        size_t size;
    };
    std::vector<Block> free_blocks;             // cached device blocks (see dev_alloc)
    std::unordered_map<void*, size_t> live_blocks;
    std::unordered_set<void*> exported_blocks;   // blocks peers may have mapped through CUDA IPC (comm.cu)
    std::unordered_map<void*, size_t> guard_at;  // PEDK_GUARD: block -> offset of its guard zone
    uint64_t guard_hits = 0;                     // guard zones found overwritten
    size_t reserved_bytes = 0;
    size_t mem_free_seen = 0, mem_reserved_seen = 0;  // last cudaMemGetInfo and reserved_bytes at that moment
    cudaStream_t copy_stream = nullptr;  // H2D of FASTQ text runs here, under the kernels of the other sample
    cudaEvent_t ev_alloc = nullptr;
    std::string err;
    // stage timing: event pairs resolved lazily in pedk_get_stats
    struct Pending {
        cudaEvent_t a, b;
        int stage;
        uint64_t sort_keys;  // != 0: an instance-stream radix pass
        uint64_t sort_launches;
        int kclass;          // PEDK_KC_* or -1
        uint64_t kbytes;     // algorithmic bytes of the launches in the scope
    };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    pedk_stats stats{};
    size_t cur_bytes = 0;
    // radix / rle workspace (grown on demand)
    uint32_t* lookback = nullptr;
    size_t lookback_words = 0;  // u32 words
    unsigned int* tile_ctr = nullptr;
    int* err_flag = nullptr;
    unsigned long long* scalars = nullptr;  // 16 device words for counters
    unsigned long long* h_scalars = nullptr;  // pinned mirror
    uint32_t* edge_pass1 = nullptr;           // device: truth table of joinKmerSub's pass 1 (edges.cu)
    void* h_stage = nullptr;                  // pinned staging buffer for results (grown on demand)
    size_t h_stage_bytes = 0;
    // multi-GPU
    void* nccl_comm = nullptr;
    void* ipc_cache = nullptr;  // comm.cu: peer blocks mapped through CUDA IPC
    int rank = 0, nranks = 1;
    int trace = 0;             // PEDK_TRACE=1: host wall-clock between sync points on stderr
    double trace_t0 = 0, trace_last = 0;
    void* file_cache = nullptr;  // files.cu: samples kept resident between the file-level calls
};

#define PEDK_API_BEGIN try {
#define PEDK_API_END(ctx)                                        \
    }                                                            \
    catch (const CudaError& e) { return ctx_fail(ctx, PEDK_E_CUDA, e.msg); } \
    catch (const PedkError& e) { return ctx_fail(ctx, e.code, e.msg); }      \
    catch (const std::bad_alloc&) { return ctx_fail(ctx, PEDK_E_NOMEM, "host allocation failed"); } \
    catch (...) { return ctx_fail(ctx, PEDK_E_INTERNAL, "unknown exception"); }

namespace pedk {

cudaStream_t ctx_stream(pedk_ctx* ctx);
int ctx_fail(pedk_ctx* ctx, int code, const std::string& msg);

void* dev_alloc(pedk_ctx* ctx, size_t bytes);
void dev_free(pedk_ctx* ctx, void* p, size_t bytes);
void dev_release_all(pedk_ctx* ctx);
void dev_mark_exported(pedk_ctx* ctx, void* p);  // the block's IPC handle was handed to the peers
// pinned host buffer of at least `bytes`, owned by the context, valid until the next call
void* host_stage(pedk_ctx* ctx, size_t bytes);

// RAII device array on the context's caching allocator
template <typename T>
struct DBuf {
    pedk_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(pedk_ctx* c, size_t count) : ctx(c), n(count) { p = static_cast<T*>(dev_alloc(c, bytes())); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) {
            reset();
            ctx = o.ctx; p = o.p; n = o.n;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    ~DBuf() { reset(); }
    size_t bytes() const { return (n ? n : 1) * sizeof(T); }
    void reset() {
        if (p) dev_free(ctx, p, bytes());
        p = nullptr;
        n = 0;
    }
    T* release() { T* r = p; p = nullptr; n = 0; return r; }
};

// Records a CUDA-event pair around a stage on the context stream.
struct StageScope {
    pedk_ctx* ctx;
    size_t idx;
    StageScope(pedk_ctx* c, int stage, uint64_t launches, uint64_t sort_keys = 0, uint64_t sort_launches = 0,
               int kclass = -1, uint64_t kbytes = 0);
    ~StageScope();
};

RadixWorkspace radix_workspace(pedk_ctx* ctx, size_t n_keys);
// sorts keys (and vals if not null) on bits [0, 8*n_pass); digit_hist holds the raw
// histograms (n_pass*256) and is consumed.  Returns the buffer that holds the result.
int radix_sort(pedk_ctx* ctx, uint64_t* a, uint64_t* b, uint32_t* va, uint32_t* vb, size_t n, int n_pass,
               unsigned long long* digit_hist, int stage);
void check_err_flag(pedk_ctx* ctx, const char* what);
void trace(pedk_ctx* ctx, const char* label);

}  // namespace pedk
