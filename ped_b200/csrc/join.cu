// Target x control: merged strand rows -> edge list (S5+S6), plus the text
// forms of the per-sample tables that ped.pl keeps on disk.
//
// The device evaluates joinKmerSub (ped.pl:703-782) for every 19-mer group
// except the first group of each of the 64 buckets of either sample: those
// <=128 rows went through lbcKmerSub's uninitialised %count (ped.pl:845-856,
// SURVEY.md F7) and have to be evaluated on their literal text, which is done
// here on the host from the same device table.
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "comm.h"
#include "plan.h"
#include "sample.h"

using namespace pedk;

// The edge records of a result live in pinned host memory: the device-to-host copy lands where the
// caller reads them (no second copy, no first-touch page faults per join).  A result outlives joins and
// may outlive its context, so the buffers are recycled through a small process-wide pool.
namespace {
struct PinnedPool {
    std::mutex mu;
    std::vector<std::pair<void*, size_t>> idle;
    void* take(size_t bytes, size_t* cap) {
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = idle.size();
            for (size_t i = 0; i < idle.size(); ++i)
                if (idle[i].second >= bytes && (best == idle.size() || idle[i].second < idle[best].second)) best = i;
            if (best != idle.size()) {
                void* p = idle[best].first;
                *cap = idle[best].second;
                idle.erase(idle.begin() + (long)best);
                return p;
            }
        }
        size_t want = bytes + bytes / 4;
        want = (want + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
        void* p = nullptr;
        PEDK_CUDA_TRY(cudaHostAlloc(&p, want, cudaHostAllocPortable));
        *cap = want;
        return p;
    }
    void give(void* p, size_t cap) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(mu);
            if (idle.size() < 4) {
                idle.emplace_back(p, cap);
                return;
            }
        }
        cudaFreeHost(p);
    }
};
PinnedPool& pinned_pool() {
    static PinnedPool* pool = new PinnedPool();  // never destroyed: results may be freed during process exit
    return *pool;
}

struct EdgeArray {
    pedk_edge* p = nullptr;
    size_t n = 0, cap_bytes = 0;
    EdgeArray() = default;
    EdgeArray(const EdgeArray&) = delete;
    EdgeArray& operator=(const EdgeArray&) = delete;
    ~EdgeArray() { pinned_pool().give(p, cap_bytes); }
    void reserve(size_t count) {
        if (count * sizeof(pedk_edge) <= cap_bytes) return;
        size_t cap = 0;
        pedk_edge* q = static_cast<pedk_edge*>(pinned_pool().take(count * sizeof(pedk_edge), &cap));
        if (n) memcpy(static_cast<void*>(q), p, n * sizeof(pedk_edge));
        pinned_pool().give(p, cap_bytes);
        p = q;
        cap_bytes = cap;
    }
    void push_back(const pedk_edge& e) {
        if ((n + 1) * sizeof(pedk_edge) > cap_bytes) reserve(n + n / 2 + 64);
        p[n++] = e;
    }
    size_t size() const { return n; }
    pedk_edge* data() { return p; }
    const pedk_edge* data() const { return p; }
    pedk_edge* begin() { return p; }
    pedk_edge* end() { return p + n; }
    const pedk_edge& operator[](size_t i) const { return p[i]; }
};
}  // namespace

struct pedk_result {
    EdgeArray edges;                                        // unordered until finalised
    std::vector<std::pair<uint64_t, std::string>> quirks;   // literal lines of the F7 rows
    std::string text;
    int k = 0;
    bool finalised = false;
};

namespace {

void append_u(std::string& s, uint64_t v) {
    char b[24];
    int n = 0;
    do {
        b[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) s.push_back(b[--n]);
}

void append_kmer(std::string& s, uint64_t v, int k) {
    char b[32];
    for (int j = 0; j < k; ++j) b[j] = "ACGT"[(v >> (2 * (k - 1 - j))) & 3];
    s.append(b, (size_t)k);
}

std::string mask_string(unsigned m) {
    std::string r;
    for (int x = 0; x < 4; ++x)
        if (m & (1u << x)) r.push_back("ACGT"[x]);
    return r;
}

// joinKmerSub on the literal text of a row whose lbc line(s) came out of the
// first-row path: `join` + Perl `split` drop the empty columns, so the counts
// that are present shift left.
bool quirk_row(const GroupRec& g, bool first_c, bool first_t, int k, pedk_edge* e, std::string* line) {
    uint32_t tok[8];
    int m = 0;
    for (int x = 0; x < 4; ++x)
        if (!first_c || g.hasc[x]) tok[m++] = g.c[x];
    for (int x = 0; x < 4; ++x)
        if (!first_t || g.hast[x]) tok[m++] = g.t[x];
    uint32_t row[8];
    bool def[8];
    for (int i = 0; i < 8; ++i) {
        def[i] = i < m;
        row[i] = i < m ? tok[i] : 0;
    }
    int rep = 0, nohita = 0, nohitb = 0, pola = 0, polb = 0;
    unsigned a = 0, b = 0;
    for (int x = 0; x < 4; ++x) {
        uint32_t cv = row[x], tv = row[x + 4];
        if (cv > 100) ++rep;
        else if (cv >= 3) { a |= 1u << x; if (tv <= 1) ++pola; }
        else if (cv <= 1) ++nohita;
        if (tv > 100) ++rep;
        else if (tv >= 3) { b |= 1u << x; if (cv <= 1) ++polb; }
        else if (tv <= 1) ++nohitb;
    }
    if (!(a != b && a && b && rep == 0 && nohita >= 2 && nohitb >= 2 && (pola == 1 || polb == 1))) return false;
    uint64_t ct = 0, tt = 0;
    for (int x = 0; x < 4; ++x) { ct += row[x]; tt += row[x + 4]; }
    if (!ct || !tt) return false;
    unsigned co = 0, al = 0;
    for (int x = 0; x < 4; ++x) {
        if (4ull * row[x] > ct) co |= 1u << x;
        if (4ull * row[x + 4] > tt) al |= 1u << x;
    }
    if (co == al) return false;
    memset(e, 0, sizeof *e);
    e->prefix = g.p;
    for (int x = 0; x < 4; ++x) { e->control[x] = row[x]; e->target[x] = row[x + 4]; }
    e->cont_mask = (uint8_t)co;
    e->alt_mask = (uint8_t)al;
    e->quirk = 1;
    line->clear();
    *line += kmer_string(g.p, k - 1);
    *line += '\t';
    *line += mask_string(co);
    *line += '\t';
    *line += mask_string(al);
    for (int i = 0; i < 8; ++i) {
        *line += '\t';
        if (def[i]) append_u(*line, row[i]);
    }
    *line += '\n';
    return true;
}

// joinKmer without a sort (edges.cu "Hash join"): the strand rows of both samples become 8-byte
// join records binned twice by the hash of their (k-1)-mer -- on several ranks the first level
// stores straight into the owner's buffer over NVLink, like the k-mer instances -- and one CTA
// per bucket applies joinKmerSub to the groups that met in its shared-memory table.
void do_join(pedk_ctx* ctx, pedk_sample* control, pedk_sample* target, pedk_result* res) {
    cudaStream_t st = ctx->stream;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!control->k || control->k != target->k || !(control->counted || control->host_only_table) ||
        !(target->counted || target->host_only_table))
        throw PedkError(PEDK_E_INVAL, "pedk_join: samples not counted with the same k");
    const int k = control->k;
    res->k = k;
    const bool multi = ctx->nranks > 1;
    const int R = ctx->nranks, me = ctx->rank;
    const int kp = k - 1, a = kmer_bits_a(kp);
    // PEDK_JOIN_B (tests): fewer second-level bits -> larger buckets, so that the tables overflow
    const int b = [&]() {
        const int full = kmer_bits_b(kp);
        const char* v = getenv("PEDK_JOIN_B");
        const int x = v && *v ? atoi(v) : full;
        return x < 1 ? 1 : x > full ? full : x;
    }();
    constexpr int NB = 256;

    // ---- the rows of both samples (tables read back from count files are uploaded first) ----
    struct Src {
        const uint64_t* key = nullptr;
        const uint32_t* val = nullptr;
        uint64_t n = 0;
        DBuf<uint64_t> up_key;
        DBuf<uint32_t> up_val;
    } src[2];
    pedk_sample* ss[2] = {control, target};
    uint64_t rows_local = 0;
    for (int i = 0; i < 2; ++i) {
        pedk_sample* s = ss[i];
        if (s->host_only_table) {
            const uint64_t m = s->h_kmer.size();
            src[i].up_key = DBuf<uint64_t>(ctx, m);
            src[i].up_val = DBuf<uint32_t>(ctx, m);
            if (m) {
                StageScope sc(ctx, PEDK_ST_COPY, 0);
                PEDK_CUDA_TRY(cudaMemcpyAsync(src[i].up_key.p, s->h_kmer.data(), m * 8, cudaMemcpyHostToDevice, st));
                PEDK_CUDA_TRY(cudaMemcpyAsync(src[i].up_val.p, s->h_cnt.data(), m * 4, cudaMemcpyHostToDevice, st));
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                ctx->stats.h2d_bytes += m * 12;
            }
            src[i].key = src[i].up_key.p;
            src[i].val = src[i].up_val.p;
            src[i].n = m;
        } else {
            src[i].key = s->row_key_ptr();
            src[i].val = s->row_val_ptr();
            src[i].n = s->n_rows;
        }
        rows_local += src[i].n;
    }
    if (rows_local == 0 && !multi) return;

    // ---- (1) level-A' bin sizes and the first k-mer of every (sample, 3-nt bucket) ----
    DBuf<unsigned long long> counts(ctx, (size_t)(NB + 1) * (PEDK_MAX_RANKS + 1) + 128);
    unsigned long long* d_hist = counts.p;                             // [257]
    unsigned long long* d_all = counts.p + (NB + 1);                   // [R][257]
    unsigned long long* d_first = counts.p + (size_t)(NB + 1) * (PEDK_MAX_RANKS + 1);  // [128]
    PEDK_CUDA_TRY(cudaMemsetAsync(d_hist, 0, (size_t)(NB + 1) * 8, st));
    PEDK_CUDA_TRY(cudaMemsetAsync(d_first, 0xFF, 128 * 8, st));
    for (int i = 0; i < 2; ++i) {
        if (!src[i].n) continue;
        StageScope sc(ctx, PEDK_ST_ROWSORT, 1, 0, 0, PEDK_KC_JOIN_COUNT, src[i].n * 8);
        launch_rows_count(src[i].key, src[i].n, i, k, d_hist, d_first, st);
    }
    std::vector<unsigned long long> all((size_t)(NB + 1) * R, 0);
    unsigned char* hstage = static_cast<unsigned char*>(host_stage(ctx, 1 << 16));
    unsigned long long* h_first = reinterpret_cast<unsigned long long*>(hstage);
    if (multi) {
        comm_allgather_u64(ctx, d_hist, d_all, NB + 1);
        comm_allreduce_min(ctx, d_first, 128);
        PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, all.size() * 8, cudaMemcpyDeviceToHost, st));
    } else {
        PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), d_hist, (size_t)(NB + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    PEDK_CUDA_TRY(cudaMemcpyAsync(h_first, d_first, 128 * 8, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->stats.d2h_bytes += all.size() * 8 + 1024;
    trace(ctx, "  join: row counts");
    LevelAPlan lp;
    plan_level_a_exact(a, R, me, all.data(), &lp);
    const uint64_t m = lp.buffer_elems;  // join records this rank owns
    uint64_t rows_all = 0;
    for (size_t i = 0; i < all.size(); ++i) rows_all += all[i];
    if (rows_all == 0) return;

    // ---- (2) rows -> join records in the level-A' bins (of their owner rank) ----
    DBuf<uint64_t> recA(ctx, peer_block_elems(m, multi));
    DBuf<unsigned long long> d_cur(ctx, NB);
    PEDK_CUDA_TRY(cudaMemcpyAsync(d_cur.p, lp.cur.data(), (size_t)NB * 8, cudaMemcpyHostToDevice, st));
    {
        PeerPtrs dst;
        if (multi) {
            void* peers[PEDK_MAX_RANKS];
            comm_peer_pointers(ctx, recA.p, peers);
            for (int d = 0; d < R; ++d) dst.p[d] = static_cast<uint64_t*>(peers[d]);
        } else {
            dst.p[0] = recA.p;
        }
        if (multi) comm_barrier(ctx);  // nobody stores into a buffer its owner still reads
        for (int i = 0; i < 2; ++i) {
            if (!src[i].n) continue;
            StageScope sc(ctx, multi ? PEDK_ST_EXCHANGE : PEDK_ST_ROWSORT, 1, 0, 0, PEDK_KC_JOIN_SCATTER, src[i].n * 20);
            launch_rows_scatter(src[i].key, src[i].val, src[i].n, i, k, R, d_cur.p, dst, st);
        }
        if (multi) comm_barrier(ctx);  // every rank's stores have landed
    }
    std::vector<GroupRec> groups;
    std::vector<uint64_t> plist;
    for (int i = 0; i < 128; ++i)
        if (h_first[i] != ~0ull) plist.push_back(h_first[i] >> 2);
    std::sort(plist.begin(), plist.end());
    plist.erase(std::unique(plist.begin(), plist.end()), plist.end());
    const std::vector<unsigned long long> first(h_first, h_first + 128);
    uint64_t E = 0;
    if (m) {
        // ---- (3) second binning level: exact bucket sizes, offsets, scatter ----
        const unsigned TILE = sub_tile_keys();
        const int n_seg = (int)lp.seg_start.size();
        std::vector<unsigned> tile_prefix((size_t)n_seg + 1, 0);
        for (int sg = 0; sg < n_seg; ++sg) {
            const unsigned long long len = lp.seg_end[(size_t)sg] - lp.seg_start[(size_t)sg];
            const unsigned long long tiles = (len + TILE - 1) / TILE;
            if ((unsigned long long)tile_prefix[(size_t)sg] + tiles > 0xfffffff0ull)
                throw PedkError(PEDK_E_UNSUPPORTED, "more than 2^45 table rows on one GPU");
            tile_prefix[(size_t)sg + 1] = tile_prefix[(size_t)sg] + (unsigned)tiles;
        }
        const unsigned n_tiles = tile_prefix[(size_t)n_seg];
        DBuf<unsigned long long> d_seg(ctx, (size_t)3 * n_seg);
        DBuf<unsigned> tprefix(ctx, (size_t)n_seg + 1);
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_seg.p, lp.seg_start.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_seg.p + n_seg, lp.seg_end.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_seg.p + 2 * n_seg, lp.seg_bin.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(tprefix.p, tile_prefix.data(), ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, st));
        const size_t n_buckets = ((size_t)1 << a) * NB;
        const int shiftB = join_record_shift_b(k, b);
        const unsigned maskB = (1u << b) - 1;
        DBuf<unsigned> histB(ctx, n_buckets);
        DBuf<unsigned long long> offB(ctx, n_buckets + 1), curB(ctx, n_buckets);
        DBuf<uint64_t> ws(ctx, scan_workspace_words(n_buckets));
        DBuf<uint64_t> recB(ctx, m);
        {
            StageScope sc(ctx, PEDK_ST_ROWSORT, 6, 0, 0, PEDK_KC_JOIN_SUB, m * 24);
            PEDK_CUDA_TRY(cudaMemsetAsync(histB.p, 0, histB.bytes(), st));
            launch_sub_hist_bits(recA.p, 1, d_seg.p, tprefix.p, n_seg, n_tiles, shiftB, maskB, histB.p, st);
            exclusive_scan_u32(histB.p, reinterpret_cast<uint64_t*>(offB.p), n_buckets, ctx->scalars + 3, ws.p, st);
            const unsigned long long total = m;
            PEDK_CUDA_TRY(cudaMemcpyAsync(offB.p + n_buckets, &total, 8, cudaMemcpyHostToDevice, st));
            PEDK_CUDA_TRY(cudaMemcpyAsync(curB.p, offB.p, n_buckets * 8, cudaMemcpyDeviceToDevice, st));
            launch_sub_scatter_bits(recA.p, recB.p, 1, d_seg.p, tprefix.p, n_seg, n_tiles, shiftB, maskB, curB.p, nullptr,
                                    nullptr, st);
        }
        // ---- (4) one CTA per bucket: groups meet in a shared-memory table, joinKmerSub on each ----
        const int log_slots = []() {
            const char* v = getenv("PEDK_JOIN_SLOTS_LOG2");
            const int x = v && *v ? atoi(v) : 14;
            return x < 10 ? 10 : x > 14 ? 14 : x;
        }();
        uint64_t cap = m / 16;
        if (cap < (1u << 16)) cap = 1u << 16;
        DBuf<EdgeRec> edges(ctx, cap);
        DBuf<uint64_t> d_p(ctx, plist.size() ? plist.size() : 1);
        DBuf<GroupRec> d_g(ctx, plist.size() ? plist.size() : 1);
        for (int attempt = 0; attempt < 2; ++attempt) {
            {
                StageScope sc(ctx, PEDK_ST_EDGES, 1, 0, 0, PEDK_KC_JOIN_TABLE, m * 8);
                PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 4, 0, sizeof(unsigned long long), st));
                launch_join_bucket(recB.p, offB.p, (unsigned)n_buckets, k, b, log_slots, d_first, ctx->edge_pass1, edges.p, cap,
                                   ctx->scalars + 4, ctx->tile_ctr, ctx->err_flag, st);
            }
            PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars, ctx->scalars + 4, 8, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            E = ctx->h_scalars[0];
            if (E <= cap) break;
            cap = E;  // rare: more edges than one row in sixteen
            edges = DBuf<EdgeRec>(ctx, cap);
        }
        check_err_flag(ctx, "join buckets");
        // EdgeRec and pedk_edge share one layout (quirk = first pad byte = 0): the records go through
        // a pinned staging buffer of the context straight into the result
        static_assert(sizeof(EdgeRec) == sizeof(pedk_edge), "edge record layout");
        res->edges.reserve(E + plist.size() + 64);  // + the first-of-bucket rows the host adds below
        res->edges.n = E;
        groups.resize(plist.size());
        {
            StageScope sc(ctx, PEDK_ST_COPY, 0);
            if (E)
                PEDK_CUDA_TRY(cudaMemcpyAsync(static_cast<void*>(res->edges.data()), edges.p, E * sizeof(EdgeRec),
                                              cudaMemcpyDeviceToHost, st));
        }
        // ---- first-of-bucket rows: look the groups up and evaluate their text on the host ----
        if (!plist.empty()) {
            StageScope sc(ctx, PEDK_ST_EDGES, 1);
            PEDK_CUDA_TRY(cudaMemcpyAsync(d_p.p, plist.data(), plist.size() * 8, cudaMemcpyHostToDevice, st));
            launch_join_lookup(recB.p, offB.p, k, b, d_p.p, (int)plist.size(), d_g.p, st);
            PEDK_CUDA_TRY(cudaMemcpyAsync(groups.data(), d_g.p, plist.size() * sizeof(GroupRec), cudaMemcpyDeviceToHost, st));
        }
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += E * sizeof(EdgeRec) + plist.size() * sizeof(GroupRec);
        trace(ctx, "join: bins + buckets + edges + D2H");
    }
    auto& quirks = res->quirks;
    for (const GroupRec& g : groups) {
        bool pc = g.hasc[0] | g.hasc[1] | g.hasc[2] | g.hasc[3];
        bool pt = g.hast[0] | g.hast[1] | g.hast[2] | g.hast[3];
        if (!(pc && pt)) continue;  // inner join (on several ranks: only the owner of the bucket sees the group)
        unsigned tag = (unsigned)(g.p >> (2 * (k - 1) - 6));
        bool first_c = first[tag] != ~0ull && (first[tag] >> 2) == g.p;
        bool first_t = first[64 + tag] != ~0ull && (first[64 + tag] >> 2) == g.p;
        pedk_edge e;
        std::string line;
        if (quirk_row(g, first_c, first_t, k, &e, &line)) {
            res->edges.push_back(e);
            quirks.emplace_back(g.p, line);
        }
    }
    trace(ctx, "host: edges collected");
}

// Sort the edges by their (k-1)-mer and lay out <target>.kmer.snp: done on first access, so a
// caller that only wants the count does not pay for the text.  The lines are formatted by a few
// host threads, each over a contiguous range of the sorted edges (the edge list of configs[1] is
// 221 k lines; one thread took longer over it than the join kernels).
void format_edges(const pedk_result* res, size_t lo, size_t hi, std::string* out) {
    const int k = res->k;
    const auto& quirks = res->quirks;
    out->reserve((hi - lo) * 72);
    char buf[128];
    for (size_t i = lo; i < hi; ++i) {
        const pedk_edge& e = res->edges[i];
        if (e.quirk) {
            auto it = std::lower_bound(quirks.begin(), quirks.end(), e.prefix,
                                       [](const std::pair<uint64_t, std::string>& q, uint64_t p) { return q.first < p; });
            if (it != quirks.end() && it->first == e.prefix) *out += it->second;
            continue;
        }
        int n = 0;
        for (int j = 0; j < k - 1; ++j) buf[n++] = "ACGT"[(e.prefix >> (2 * (k - 2 - j))) & 3];
        buf[n++] = '\t';
        for (int x = 0; x < 4; ++x)
            if (e.cont_mask & (1u << x)) buf[n++] = "ACGT"[x];
        buf[n++] = '\t';
        for (int x = 0; x < 4; ++x)
            if (e.alt_mask & (1u << x)) buf[n++] = "ACGT"[x];
        for (int x = 0; x < 8; ++x) {
            buf[n++] = '\t';
            uint32_t v = x < 4 ? e.control[x] : e.target[x - 4];
            char d[12];
            int m = 0;
            do {
                d[m++] = (char)('0' + v % 10);
                v /= 10;
            } while (v);
            while (m) buf[n++] = d[--m];
        }
        buf[n++] = '\n';
        out->append(buf, (size_t)n);
    }
}

void finalise(pedk_result* res) {
    if (res->finalised) return;
    auto& quirks = res->quirks;
    std::sort(res->edges.begin(), res->edges.end(),
              [](const pedk_edge& x, const pedk_edge& y) { return x.prefix < y.prefix; });
    std::sort(quirks.begin(), quirks.end());
    // ---- <target>.kmer.snp ----
    const size_t E = res->edges.size();
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : nt > 16 ? 16 : nt;
    if (E < 4096) nt = 1;
    std::vector<std::string> part(nt);
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(format_edges, res, E * t / nt, E * (t + 1) / nt, &part[t]);
    format_edges(res, 0, E / nt, &part[0]);
    for (auto& x : th) x.join();
    size_t bytes = 0;
    for (const std::string& p : part) bytes += p.size();
    res->text.reserve(bytes);
    for (const std::string& p : part) res->text += p;
    res->finalised = true;
}

char* dup_text(const std::string& s, uint64_t* len) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) throw std::bad_alloc();
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
}

}  // namespace

namespace pedk {
// The sort_uniq/ lines of a sample on the host: every unique read and (unless it is a palindrome) its
// reverse complement, per 3-nt bucket, sorted (ped.pl:1416).  Decoding and the 64 sorts run on host
// threads: this is the text side of the drop-in path (scope B), not of the kernel path.
void ensure_host_reads(pedk_sample* s) {
    if (s->have_reads) return;
    pedk_ctx* ctx = s->ctx;
    for (auto& v : s->h_reads) v.clear();
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : nt > 32 ? 32 : nt;
    for (const ReadGroup& g : s->groups) {
        if (!g.n) continue;
        if (!g.words.p) throw PedkError(PEDK_E_INVAL, "packed reads were released");
        std::vector<uint64_t> w(g.n * g.W);
        std::vector<uint8_t> pal(g.n);
        PEDK_CUDA_TRY(cudaMemcpyAsync(w.data(), g.words.p, w.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PEDK_CUDA_TRY(cudaMemcpyAsync(pal.data(), g.pal.p, g.n, cudaMemcpyDeviceToHost, ctx->stream));
        PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += w.size() * 8 + g.n;
        // every thread decodes a slice of the reads into its own 64 buckets
        std::vector<std::vector<std::string>> part((size_t)nt * 64);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t]() {
                const uint64_t lo = g.n * t / nt, hi = g.n * (t + 1) / nt;
                std::string fwd((size_t)g.L, 'A'), rev((size_t)g.L, 'A');
                for (uint64_t r = lo; r < hi; ++r) {
                    const uint64_t* rw = w.data() + r * g.W;
                    for (int i = 0; i < g.L; ++i) {
                        unsigned c = (unsigned)(rw[i >> 5] >> (62 - 2 * (i & 31))) & 3u;
                        fwd[(size_t)i] = "ACGT"[c];
                        rev[(size_t)(g.L - 1 - i)] = "ACGT"[c ^ 3u];
                    }
                    unsigned tf = (unsigned)(rw[0] >> 58);
                    part[(size_t)t * 64 + tf].push_back(fwd);
                    if (!pal[r]) {
                        unsigned tr = (unsigned)((base_code((unsigned char)rev[0]) << 4) |
                                                 (base_code((unsigned char)rev[1]) << 2) | base_code((unsigned char)rev[2]));
                        part[(size_t)t * 64 + tr].push_back(rev);
                    }
                }
            });
        for (auto& x : th) x.join();
        th.clear();
        std::atomic<int> next(0);
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&]() {
                for (int tag = next.fetch_add(1); tag < 64; tag = next.fetch_add(1)) {
                    auto& dst = s->h_reads[tag];
                    size_t add = 0;
                    for (unsigned p = 0; p < nt; ++p) add += part[(size_t)p * 64 + tag].size();
                    dst.reserve(dst.size() + add);
                    for (unsigned p = 0; p < nt; ++p) {
                        auto& src = part[(size_t)p * 64 + tag];
                        for (auto& x : src) dst.push_back(std::move(x));
                        std::vector<std::string>().swap(src);
                    }
                }
            });
        for (auto& x : th) x.join();
    }
    {
        std::atomic<int> next(0);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&]() {
                for (int tag = next.fetch_add(1); tag < 64; tag = next.fetch_add(1))
                    std::sort(s->h_reads[tag].begin(), s->h_reads[tag].end());
            });
        for (auto& x : th) x.join();
    }
    s->have_reads = true;
}

}  // namespace pedk

extern "C" int pedk_join(pedk_ctx* ctx, pedk_sample* control, pedk_sample* target, pedk_result** out) {
    if (!ctx || !control || !target || !out) return PEDK_E_INVAL;
    *out = nullptr;
    PEDK_API_BEGIN
    pedk_result* r = new pedk_result();
    try {
        do_join(ctx, control, target, r);
    } catch (...) {
        delete r;
        throw;
    }
    *out = r;
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" uint64_t pedk_result_num_edges(const pedk_result* r) { return r ? r->edges.size() : 0; }

extern "C" int pedk_result_text(const pedk_result* r, const char** text, uint64_t* len) {
    if (!r || !text || !len) return PEDK_E_INVAL;
    finalise(const_cast<pedk_result*>(r));
    *text = r->text.c_str();
    *len = r->text.size();
    return PEDK_OK;
}

extern "C" int pedk_result_edges(const pedk_result* r, pedk_edge* out, uint64_t cap, uint64_t* n) {
    if (!r || !n) return PEDK_E_INVAL;
    finalise(const_cast<pedk_result*>(r));
    *n = r->edges.size();
    if (out) {
        if (cap < *n) return PEDK_E_INVAL;
        memcpy(out, r->edges.data(), *n * sizeof(pedk_edge));
    }
    return PEDK_OK;
}

extern "C" int pedk_result_destroy(pedk_result* r) {
    delete r;
    return PEDK_OK;
}

extern "C" int pedk_sample_count_text(pedk_sample* s, int tag, char** text, uint64_t* len) {
    if (!s || !text || !len || tag < 0 || tag >= 64) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    ensure_host_table(s);
    std::string t;
    t.reserve((s->tag_start[tag + 1] - s->tag_start[tag]) * (size_t)(s->k + 8));
    for (uint64_t i = s->tag_start[tag]; i < s->tag_start[tag + 1]; ++i) {
        append_kmer(t, s->h_kmer[i], s->k);
        t += '\t';
        append_u(t, s->h_cnt[i]);
        t += '\n';
    }
    *text = dup_text(t, len);
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

// lbcKmerSub, ped.pl:833-860: the first group of the file prints "" for the last
// bases it never saw (%count is not initialised before the first print).
extern "C" int pedk_sample_lbc_text(pedk_sample* s, int tag, char** text, uint64_t* len) {
    if (!s || !text || !len || tag < 0 || tag >= 64) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    ensure_host_table(s);
    std::string t;
    uint64_t lo = s->tag_start[tag], hi = s->tag_start[tag + 1];
    t.reserve((hi - lo) * (size_t)(s->k + 12));
    if (lo == hi) t = "\t\t\t\t\n";  // ped.pl:856 with nothing read
    bool first = true;
    for (uint64_t i = lo; i < hi;) {
        uint64_t p = s->h_kmer[i] >> 2;
        uint32_t c[4] = {0, 0, 0, 0};
        bool has[4] = {false, false, false, false};
        uint64_t j = i;
        for (; j < hi && (s->h_kmer[j] >> 2) == p; ++j) {
            c[s->h_kmer[j] & 3] = s->h_cnt[j];
            has[s->h_kmer[j] & 3] = true;
        }
        append_kmer(t, p, s->k - 1);
        for (int x = 0; x < 4; ++x) {
            t += '\t';
            if (!first || has[x]) append_u(t, c[x]);
        }
        t += '\n';
        first = false;
        i = j;
    }
    *text = dup_text(t, len);
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_sort_uniq_text(pedk_sample* s, int tag, char** text, uint64_t* len) {
    if (!s || !text || !len || tag < 0 || tag >= 64) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    ensure_host_reads(s);
    std::string t;
    for (const std::string& r : s->h_reads[tag]) {
        t += r;
        t += '\n';
    }
    *text = dup_text(t, len);
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

// host-only (no CUDA call): the 65 536-entry truth table k_edges filters (k-1)-mer groups with
extern "C" int pedk_edge_pass1_table(uint32_t* bits2048) {
    if (!bits2048) return PEDK_E_INVAL;
    pedk::build_edge_pass1_table(bits2048);
    return PEDK_OK;
}

static int kmer_edges_impl(pedk_ctx* ctx, const void* control_text, uint64_t control_bytes, const void* target_fastq,
                           uint64_t target_bytes, int where, int k, int clipping, bool ref_control,
                           pedk_result** out) {
    if (!ctx || !out || where < 0 || where > 2) return PEDK_E_INVAL;
    *out = nullptr;
    pedk_sample *c = nullptr, *t = nullptr;
    int rc = pedk_sample_create(ctx, &c);
    if (rc == PEDK_OK) rc = pedk_sample_create(ctx, &t);
    // both copies are queued before the first kernel so the target's transfer
    // runs under the control's ingest on the copy engine
    if (where == 2) {
        // padded device buffers: read in place, no copy
        const uint64_t T = 16384;
        if (rc == PEDK_OK)
            rc = pedk_sample_borrow_fastq_device(t, const_cast<void*>(target_fastq), target_bytes,
                                                 (target_bytes + T - 1) / T * T + T);
        if (rc == PEDK_OK)
            rc = pedk_sample_borrow_fastq_device(c, const_cast<void*>(control_text), control_bytes,
                                                 (control_bytes + T - 1) / T * T + T);
    } else {
        if (rc == PEDK_OK) rc = pedk_sample_reserve(c, control_bytes);
        if (rc == PEDK_OK) rc = pedk_sample_reserve(t, target_bytes);
        if (rc == PEDK_OK)
            rc = where ? pedk_sample_add_fastq_device(t, target_fastq, target_bytes)
                       : pedk_sample_add_fastq(t, target_fastq, target_bytes);
        if (rc == PEDK_OK)
            rc = where ? pedk_sample_add_fastq_device(c, control_text, control_bytes)
                       : pedk_sample_add_fastq(c, control_text, control_bytes);
    }
    // ped.pl:291-297 ingests the target first, and `$clipping` is a global: a value chosen
    // automatically for the target is an explicit clipping= for the control
    pedk_ingest_info ti;
    if (rc == PEDK_OK) rc = pedk_sample_ingest(t, clipping, &ti);
    if (rc == PEDK_OK && ti.auto_clipped) clipping = ti.clipping_used;
    if (rc == PEDK_OK) rc = pedk_sample_count(t, k, nullptr);
    if (rc == PEDK_OK) rc = pedk_sample_release_reads(t);
    if (rc == PEDK_OK) rc = ref_control ? pedk_sample_ingest_reference(c, 0, 0, nullptr) : pedk_sample_ingest(c, clipping, nullptr);
    if (rc == PEDK_OK) rc = pedk_sample_count(c, k, nullptr);
    if (rc == PEDK_OK) rc = pedk_sample_release_reads(c);
    if (rc == PEDK_OK) rc = pedk_join(ctx, c, t, out);
    pedk_sample_destroy(c);
    pedk_sample_destroy(t);
    pedk::trace(ctx, "samples destroyed (step end)");
    return rc;
}

extern "C" int pedk_kmer_edges(pedk_ctx* ctx, const void* control_fastq, uint64_t control_bytes,
                               const void* target_fastq, uint64_t target_bytes, int where, int k, int clipping,
                               pedk_result** out) {
    return kmer_edges_impl(ctx, control_fastq, control_bytes, target_fastq, target_bytes, where, k, clipping, false, out);
}

extern "C" int pedk_kmer_edges_ref(pedk_ctx* ctx, const void* ref_fasta, uint64_t ref_bytes, const void* target_fastq,
                                   uint64_t target_bytes, int where, int k, int clipping, pedk_result** out) {
    return kmer_edges_impl(ctx, ref_fasta, ref_bytes, target_fastq, target_bytes, where, k, clipping, true, out);
}
