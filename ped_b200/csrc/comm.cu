// Multi-GPU plumbing (NCCL), one rank per process.  Filled in by the sharded path.
#include "ctx.h"

namespace pedk {
void comm_destroy(pedk_ctx*) {}
}  // namespace pedk

extern "C" int pedk_comm_unique_id(pedk_ctx* ctx, void*) {
    return pedk::ctx_fail(ctx, PEDK_E_UNSUPPORTED, "pedk_comm_unique_id: not built yet");
}
extern "C" int pedk_comm_init(pedk_ctx* ctx, const void*, int, int) {
    return pedk::ctx_fail(ctx, PEDK_E_UNSUPPORTED, "pedk_comm_init: not built yet");
}
