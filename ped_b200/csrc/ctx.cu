// Context life cycle, stream-ordered device memory, stage timing, the radix
// sort driver and the raw-kernel test entry points of libpedk.so.
#include <chrono>

#include "ctx.h"

#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <mutex>

using namespace pedk;

namespace pedk {

void set_max_dyn_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;  // (kernel, device) -> bytes granted
    int dev = 0;
    PEDK_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return;
    PEDK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done[key] = bytes;
}

cudaStream_t ctx_stream(pedk_ctx* ctx) { return ctx->stream; }

int ctx_fail(pedk_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

// Device memory: a caching allocator private to the context.  Every stage of a step asks
// for the same sizes as the step before, so after the first step no allocation reaches
// the driver (cudaMalloc of tens of GB costs tens of ms and synchronises the device).
// Blocks are handed out again in host order; all kernels of a context run on one stream
// (the copy stream is fenced with events), so stream order makes the reuse safe.
static size_t block_size_for(size_t bytes) {
    const size_t g = bytes < (size_t(1) << 20) ? 512 : (size_t(2) << 20);
    return (bytes + g - 1) / g * g;
}

// Blocks whose IPC handle went to the peers (comm_peer_pointers) stay with the allocator while the
// communicator lives: a peer may still have them mapped, and the driver does not return the memory of
// such a block before every mapping is closed -- freeing it would only lose the block for reuse.
static void release_cached(pedk_ctx* ctx, bool all = false) {
    cudaStreamSynchronize(ctx->stream);
    std::vector<pedk_ctx::Block> kept;
    for (auto& b : ctx->free_blocks) {
        if (!all && ctx->exported_blocks.count(b.p)) {
            kept.push_back(b);
            continue;
        }
        cudaFree(b.p);
        ctx->exported_blocks.erase(b.p);
        ctx->reserved_bytes -= b.size;
    }
    ctx->free_blocks.swap(kept);
}

void dev_mark_exported(pedk_ctx* ctx, void* p) { ctx->exported_blocks.insert(p); }

// PEDK_GUARD=1 (tests; compute-sanitizer is not available on every pool): every block gets a guard
// zone behind the bytes that were asked for, filled with a pattern when the block is handed out and
// checked when it comes back -- a kernel that wrote past the end of any buffer fails the call loudly.
constexpr size_t GUARD_BYTES = 512;
static bool guard_on() {
    static const bool on = getenv("PEDK_GUARD") && atoi(getenv("PEDK_GUARD")) != 0;
    return on;
}

void* dev_alloc(pedk_ctx* ctx, size_t bytes) {
    const size_t asked = bytes ? bytes : 1;
    if (guard_on()) bytes = asked + GUARD_BYTES;
    const size_t want = block_size_for(bytes ? bytes : 1);
    // best fit among cached blocks, but never waste more than 1/8 (+2 MiB) of a block
    size_t best = (size_t)-1, best_size = (size_t)-1;
    for (size_t i = 0; i < ctx->free_blocks.size(); ++i) {
        size_t sz = ctx->free_blocks[i].size;
        if (sz >= want && sz <= want + want / 8 + (size_t(2) << 20) && sz < best_size) {
            best = i;
            best_size = sz;
        }
    }
    void* p = nullptr;
    size_t got = want;
    if (best != (size_t)-1) {
        p = ctx->free_blocks[best].p;
        got = best_size;
        ctx->free_blocks.erase(ctx->free_blocks.begin() + (long)best);
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaMalloc(&p, want);
        if (ctx->trace)
            fprintf(stderr, "[pedk] cudaMalloc of %zu bytes: %.2f ms (%zu cached blocks)\n", want,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
                    ctx->free_blocks.size());
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            // the device is full.  Before giving the cached blocks back to the driver (a free + malloc
            // of tens of GB costs tens of ms, every step again): is any idle block large enough?
            size_t any = (size_t)-1, any_size = (size_t)-1;
            for (size_t i = 0; i < ctx->free_blocks.size(); ++i) {
                const size_t sz = ctx->free_blocks[i].size;
                if (sz >= want && sz < any_size) {
                    any = i;
                    any_size = sz;
                }
            }
            if (any != (size_t)-1) {
                p = ctx->free_blocks[any].p;
                got = any_size;
                ctx->free_blocks.erase(ctx->free_blocks.begin() + (long)any);
                e = cudaSuccess;
            } else {
                release_cached(ctx);  // give the cached blocks back and try once more
                e = cudaMalloc(&p, want);
                if (e == cudaSuccess) ctx->reserved_bytes += want;
            }
        } else if (e == cudaSuccess) {
            ctx->reserved_bytes += want;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            char b[256];
            snprintf(b, sizeof b, "device allocation of %zu bytes failed (%zu in use, %zu reserved): %s", want,
                     ctx->cur_bytes, ctx->reserved_bytes, cudaGetErrorString(e));
            throw PedkError(PEDK_E_NOMEM, b);
        }
    }
    ctx->live_blocks[p] = got;
    ctx->cur_bytes += got;
    if (ctx->cur_bytes > ctx->stats.peak_device_bytes) ctx->stats.peak_device_bytes = ctx->cur_bytes;
    if (guard_on()) {
        ctx->guard_at[p] = asked;
        cudaMemsetAsync(static_cast<char*>(p) + asked, 0xA5, GUARD_BYTES, ctx->stream);
    }
    return p;
}

void dev_free(pedk_ctx* ctx, void* p, size_t /*bytes*/) {
    if (!p) return;
    auto it = ctx->live_blocks.find(p);
    if (it == ctx->live_blocks.end()) return;
    if (guard_on()) {
        auto g = ctx->guard_at.find(p);
        if (g != ctx->guard_at.end()) {
            unsigned char h[GUARD_BYTES];
            memset(h, 0, sizeof h);
            cudaMemcpyAsync(h, static_cast<char*>(p) + g->second, GUARD_BYTES, cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            for (size_t i = 0; i < GUARD_BYTES; ++i)
                if (h[i] != 0xA5) {
                    ++ctx->guard_hits;
                    fprintf(stderr, "[pedk] PEDK_GUARD: a kernel wrote past the end of a %zu-byte device block (guard byte %zu)\n",
                            g->second, i);
                    break;
                }
            ctx->guard_at.erase(g);
        }
    }
    ctx->cur_bytes -= it->second;
    ctx->free_blocks.push_back({p, it->second});
    ctx->live_blocks.erase(it);
}

void dev_release_all(pedk_ctx* ctx) {
    release_cached(ctx, true);
    for (auto& kv : ctx->live_blocks) cudaFree(kv.first);
    ctx->live_blocks.clear();
}

void* host_stage(pedk_ctx* ctx, size_t bytes) {
    if (bytes > ctx->h_stage_bytes) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
        size_t want = bytes + bytes / 4 + (1u << 20);
        PEDK_CUDA_TRY(cudaMallocHost(&ctx->h_stage, want));
        ctx->h_stage_bytes = want;
    }
    return ctx->h_stage;
}

static cudaEvent_t get_event(pedk_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    PEDK_CUDA_TRY(cudaEventCreate(&e));
    return e;
}

StageScope::StageScope(pedk_ctx* c, int stage, uint64_t launches, uint64_t sort_keys, uint64_t sort_launches,
                       int kclass, uint64_t kbytes)
    : ctx(c) {
    pedk_ctx::Pending p;
    p.a = get_event(c);
    p.b = get_event(c);
    p.stage = stage;
    p.sort_keys = sort_keys;
    p.sort_launches = sort_launches;
    p.kclass = kclass;
    p.kbytes = kbytes;
    PEDK_CUDA_TRY(cudaEventRecord(p.a, c->stream));
    idx = c->pending.size();
    c->pending.push_back(p);
    c->stats.stage_launches[stage] += launches;
    c->stats.kernel_launches += launches;
}

StageScope::~StageScope() { cudaEventRecord(ctx->pending[idx].b, ctx->stream); }

static void resolve_pending(pedk_ctx* ctx) {
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto& p : ctx->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            ctx->stats.stage_ms[p.stage] += ms;
            if (p.kclass >= 0 && p.kclass < PEDK_N_KC) {
                ctx->stats.kc_ms[p.kclass] += ms;
                ctx->stats.kc_bytes[p.kclass] += p.kbytes;
                ctx->stats.kc_launches[p.kclass] += 1;
            }
            if (p.sort_keys) {
                ctx->stats.sort_pass_ms += ms;
                ctx->stats.sort_pass_keys += p.sort_keys;
                ctx->stats.sort_pass_launches += p.sort_launches;
            }
        } else {
            cudaGetLastError();
        }
        ctx->event_pool.push_back(p.a);
        ctx->event_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

RadixWorkspace radix_workspace(pedk_ctx* ctx, size_t n_keys) {
    size_t tiles = radix_tiles(n_keys);
    size_t words = tiles * 256;
    if (words > ctx->lookback_words) {
        if (ctx->lookback) dev_free(ctx, ctx->lookback, ctx->lookback_words * sizeof(uint32_t));
        ctx->lookback = nullptr;
        ctx->lookback_words = 0;
        ctx->lookback = static_cast<uint32_t*>(dev_alloc(ctx, words * sizeof(uint32_t)));
        ctx->lookback_words = words;
    }
    RadixWorkspace ws;
    ws.lookback = reinterpret_cast<uint64_t*>(ctx->lookback);
    ws.tile_ctr = ctx->tile_ctr;
    ws.err = ctx->err_flag;
    ws.tiles_cap = ctx->lookback_words / 256;
    return ws;
}

int radix_sort(pedk_ctx* ctx, uint64_t* a, uint64_t* b, uint32_t* va, uint32_t* vb, size_t n, int n_pass,
               unsigned long long* digit_hist, int stage) {
    if (!n) return 0;
    RadixWorkspace ws = radix_workspace(ctx, n);
    {
        StageScope sc(ctx, stage, 1);
        launch_hist_to_base(digit_hist, n_pass, ctx->stream);
    }
    int cur = 0;
    size_t portions = (n + ((size_t(1) << 30) - 1)) >> 30;  // upper bound, only for launch accounting
    if (portions == 0) portions = 1;
    for (int p = 0; p < n_pass; ++p) {
        StageScope sc(ctx, stage, portions, stage == PEDK_ST_SORT ? n : 0, stage == PEDK_ST_SORT ? portions : 0,
                      va ? PEDK_KC_RADIX_PAIRS : PEDK_KC_RADIX_KEYS, (uint64_t)n * (va ? 24 : 16));
        if (cur == 0)
            launch_radix_pass(a, b, va, vb, n, p, digit_hist, ws, ctx->stream);
        else
            launch_radix_pass(b, a, vb, va, n, p, digit_hist, ws, ctx->stream);
        cur ^= 1;
    }
    return cur;
}

void trace(pedk_ctx* ctx, const char* label) {
    if (!ctx->trace) return;
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    if (ctx->trace_t0 == 0) ctx->trace_t0 = ctx->trace_last = now;
    fprintf(stderr, "[pedk %9.2f ms +%8.2f] %s\n", now - ctx->trace_t0, now - ctx->trace_last, label);
    ctx->trace_last = now;
}

void check_err_flag(pedk_ctx* ctx, const char* what) {
    int flag = 0;
    PEDK_CUDA_TRY(cudaMemcpyAsync(&flag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        PEDK_CUDA_TRY(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), ctx->stream));
        throw PedkError(PEDK_E_INTERNAL, std::string("chained scan timed out in ") + what);
    }
}

}  // namespace pedk

extern "C" int pedk_abi_version(void) { return 1; }

extern "C" void* pedk_stream(pedk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" const char* pedk_strerror(int code) {
    switch (code) {
    case PEDK_OK: return "ok";
    case PEDK_E_INVAL: return "invalid argument";
    case PEDK_E_CUDA: return "CUDA error";
    case PEDK_E_NOMEM: return "out of memory";
    case PEDK_E_IO: return "I/O error";
    case PEDK_E_ALPHABET: return "read with a byte outside ACGTN";
    case PEDK_E_UNSUPPORTED: return "unsupported input";
    case PEDK_E_NCCL: return "NCCL error";
    case PEDK_E_INTERNAL: return "internal error";
    case PEDK_E_NODEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
    default: return "unknown error";
    }
}

static thread_local std::string g_open_error;

extern "C" const char* pedk_last_error(pedk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_open_error.c_str(); }

extern "C" int pedk_open(pedk_ctx** out, const int* devices, int ndev) {
    if (!out) return PEDK_E_INVAL;
    *out = nullptr;
    if (ndev > 1) {
        g_open_error = "pedk_open: one device per context (run one rank per GPU and connect with pedk_comm_init)";
        return PEDK_E_INVAL;
    }
    int dev = (devices && ndev == 1) ? devices[0] : 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_open_error = std::string("pedk_open: no CUDA device (") + cudaGetErrorString(e) +
                       "); libpedk has no CPU fallback";
        return PEDK_E_NODEVICE;
    }
    if (dev < 0 || dev >= count) {
        g_open_error = "pedk_open: device ordinal out of range";
        return PEDK_E_INVAL;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        cudaGetLastError();
        g_open_error = "pedk_open: device is not sm_100 (kernels are built for sm_100a only)";
        return PEDK_E_NODEVICE;
    }
    pedk_ctx* ctx = new pedk_ctx();
    ctx->device = dev;
    ctx->trace = getenv("PEDK_TRACE") ? atoi(getenv("PEDK_TRACE")) : 0;
    try {
        PEDK_CUDA_TRY(cudaSetDevice(dev));
        PEDK_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        PEDK_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        PEDK_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_alloc, cudaEventDisableTiming));
        PEDK_CUDA_TRY(cudaMalloc(&ctx->tile_ctr, sizeof(unsigned)));
        PEDK_CUDA_TRY(cudaMalloc(&ctx->err_flag, sizeof(int)));
        PEDK_CUDA_TRY(cudaMemset(ctx->err_flag, 0, sizeof(int)));
        PEDK_CUDA_TRY(cudaMalloc(&ctx->scalars, 16 * sizeof(unsigned long long)));
        PEDK_CUDA_TRY(cudaMallocHost(&ctx->h_scalars, 16 * sizeof(unsigned long long)));
        {
            std::vector<uint32_t> bits(2048);
            build_edge_pass1_table(bits.data());
            PEDK_CUDA_TRY(cudaMalloc(&ctx->edge_pass1, 2048 * sizeof(uint32_t)));
            PEDK_CUDA_TRY(cudaMemcpy(ctx->edge_pass1, bits.data(), 2048 * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    } catch (const CudaError& err) {
        g_open_error = err.msg;
        delete ctx;
        return PEDK_E_CUDA;
    }
    *out = ctx;
    return PEDK_OK;
}

namespace pedk {
void comm_destroy(pedk_ctx* ctx);
void file_cache_destroy(pedk_ctx* ctx);
}

extern "C" int pedk_close(pedk_ctx* ctx) {
    if (!ctx) return PEDK_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    file_cache_destroy(ctx);
    comm_destroy(ctx);
    for (auto& p : ctx->pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->stream);
    dev_release_all(ctx);
    cudaFree(ctx->tile_ctr);
    cudaFree(ctx->err_flag);
    cudaFree(ctx->scalars);
    cudaFree(ctx->edge_pass1);
    cudaFreeHost(ctx->h_scalars);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->ev_alloc);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PEDK_OK;
}

extern "C" uint64_t pedk_guard_hits(pedk_ctx* ctx) { return ctx ? ctx->guard_hits : 0; }

extern "C" int pedk_get_stats(pedk_ctx* ctx, pedk_stats* out) {
    if (!ctx || !out) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    resolve_pending(ctx);
    *out = ctx->stats;
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_reset_stats(pedk_ctx* ctx) {
    if (!ctx) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    resolve_pending(ctx);
    ctx->stats = pedk_stats{};
    ctx->stats.peak_device_bytes = ctx->cur_bytes;
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" void pedk_free(void* p) { free(p); }

// ---------------- raw kernels for unit tests ----------------

extern "C" int pedk_test_sort_u64(pedk_ctx* ctx, uint64_t* keys_dev, uint64_t n, int key_bits) {
    if (!ctx || (!keys_dev && n) || key_bits < 1 || key_bits > 64) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    int n_pass = (key_bits + 7) / 8;
    DBuf<uint64_t> tmp(ctx, n);
    DBuf<unsigned long long> hist(ctx, 8 * 256);
    PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), ctx->stream));
    launch_digit_hist(keys_dev, n, hist.p, n_pass, ctx->stream);
    int cur = radix_sort(ctx, keys_dev, tmp.p, nullptr, nullptr, n, n_pass, hist.p, PEDK_ST_SORT);
    if (cur == 1)
        PEDK_CUDA_TRY(cudaMemcpyAsync(keys_dev, tmp.p, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    check_err_flag(ctx, "radix sort");
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_test_sort_pairs(pedk_ctx* ctx, uint64_t* keys_dev, uint32_t* vals_dev, uint64_t n,
                                    int key_bits) {
    if (!ctx || ((!keys_dev || !vals_dev) && n) || key_bits < 1 || key_bits > 64) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    int n_pass = (key_bits + 7) / 8;
    DBuf<uint64_t> tmp(ctx, n);
    DBuf<uint32_t> vtmp(ctx, n);
    DBuf<unsigned long long> hist(ctx, 8 * 256);
    PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), ctx->stream));
    launch_digit_hist(keys_dev, n, hist.p, n_pass, ctx->stream);
    int cur = radix_sort(ctx, keys_dev, tmp.p, vals_dev, vtmp.p, n, n_pass, hist.p, PEDK_ST_ROWSORT);
    if (cur == 1) {
        PEDK_CUDA_TRY(cudaMemcpyAsync(keys_dev, tmp.p, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
        PEDK_CUDA_TRY(cudaMemcpyAsync(vals_dev, vtmp.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    check_err_flag(ctx, "radix sort (pairs)");
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_test_rle(pedk_ctx* ctx, const uint64_t* sorted_dev, uint64_t n, uint64_t* ukey_dev,
                             uint32_t* count_dev, uint64_t* n_runs) {
    if (!ctx || !n_runs) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 8, 0, 2 * sizeof(unsigned long long), ctx->stream));
    {
        StageScope sc(ctx, PEDK_ST_COUNT, 1);
        launch_count_rows(sorted_dev, n, 20, 0, 0, nullptr, nullptr, 0, ukey_dev, count_dev, n, ctx->scalars + 8,
                          ctx->stream);
    }
    PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars + 8, ctx->scalars + 8, sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *n_runs = ctx->h_scalars[8];
    return PEDK_OK;
    PEDK_API_END(ctx)
}
