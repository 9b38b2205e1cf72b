// File-level entry points: what the drop-in ped.pl calls.
#include "sample.h"

extern "C" int pedk_mk_sort_uniq(pedk_ctx* ctx, const char*, const char*, int, int*) {
    return pedk::ctx_fail(ctx, PEDK_E_UNSUPPORTED, "pedk_mk_sort_uniq: not built yet");
}
extern "C" int pedk_count_kmer(pedk_ctx* ctx, const char*, const char*, int) {
    return pedk::ctx_fail(ctx, PEDK_E_UNSUPPORTED, "pedk_count_kmer: not built yet");
}
extern "C" int pedk_lbc_kmer(pedk_ctx* ctx, const char*, const char*, int) {
    return pedk::ctx_fail(ctx, PEDK_E_UNSUPPORTED, "pedk_lbc_kmer: not built yet");
}
extern "C" int pedk_join_kmer(pedk_ctx* ctx, const char*, const char*, const char*, int, const char*, int) {
    return pedk::ctx_fail(ctx, PEDK_E_UNSUPPORTED, "pedk_join_kmer: not built yet");
}
