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
#include <string>

#include "sample.h"

using namespace pedk;

struct pedk_result {
    std::vector<pedk_edge> edges;                           // unordered until finalised
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

void do_join(pedk_ctx* ctx, pedk_sample* control, pedk_sample* target, pedk_result* res) {
    cudaStream_t st = ctx->stream;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!control->k || control->k != target->k || !(control->counted || control->host_only_table) ||
        !(target->counted || target->host_only_table))
        throw PedkError(PEDK_E_INVAL, "pedk_join: samples not counted with the same k");
    const int k = control->k;
    res->k = k;
    RowTable rows;
    build_rows(ctx, control, target, &rows);
    if (!rows.n) return;

    DBuf<unsigned long long> first(ctx, 128);
    uint64_t cap = rows.n / 16;
    if (cap < (1u << 16)) cap = 1u << 16;
    if (cap > rows.n) cap = rows.n;
    DBuf<EdgeRec> edges(ctx, cap);
    uint64_t E = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        {
            StageScope sc(ctx, PEDK_ST_EDGES, 2);
            PEDK_CUDA_TRY(cudaMemsetAsync(first.p, 0xFF, first.bytes(), st));
            PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 4, 0, sizeof(unsigned long long), st));
            launch_first_of_bucket(rows.key.p, rows.val.p, rows.n, k, first.p, st);
            launch_edges(rows.key.p, rows.val.p, rows.n, k, first.p, ctx->edge_pass1, edges.p, cap, ctx->scalars + 4, st);
        }
        PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars, ctx->scalars + 4, 8, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        E = ctx->h_scalars[0];
        if (E <= cap) break;
        cap = E;  // rare: more edges than one row in sixteen
        edges = DBuf<EdgeRec>(ctx, cap);
    }
    // EdgeRec and pedk_edge share one layout (quirk = first pad byte = 0): the records go through
    // a pinned staging buffer of the context straight into the result
    static_assert(sizeof(EdgeRec) == sizeof(pedk_edge), "edge record layout");
    unsigned char* stage = static_cast<unsigned char*>(host_stage(ctx, E * sizeof(EdgeRec) + 1024));
    unsigned long long* h_first = reinterpret_cast<unsigned long long*>(stage + E * sizeof(EdgeRec));
    {
        StageScope sc(ctx, PEDK_ST_COPY, 0);
        if (E) PEDK_CUDA_TRY(cudaMemcpyAsync(stage, edges.p, E * sizeof(EdgeRec), cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(h_first, first.p, 128 * 8, cudaMemcpyDeviceToHost, st));
    }
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->stats.d2h_bytes += E * sizeof(EdgeRec) + 1024;
    trace(ctx, "edges kernel + D2H");

    // ---- first-of-bucket rows: look the groups up and evaluate their text on the host ----
    std::vector<uint64_t> plist;
    for (int i = 0; i < 128; ++i)
        if (h_first[i] != ~0ull) plist.push_back(h_first[i] >> 2);
    std::sort(plist.begin(), plist.end());
    plist.erase(std::unique(plist.begin(), plist.end()), plist.end());
    std::vector<GroupRec> groups(plist.size());
    if (!plist.empty()) {
        DBuf<uint64_t> d_p(ctx, plist.size());
        DBuf<GroupRec> d_g(ctx, plist.size());
        StageScope sc(ctx, PEDK_ST_EDGES, 1);
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_p.p, plist.data(), plist.size() * 8, cudaMemcpyHostToDevice, st));
        launch_lookup_groups(rows.key.p, rows.val.p, rows.n, d_p.p, (int)plist.size(), d_g.p, st);
        PEDK_CUDA_TRY(cudaMemcpyAsync(groups.data(), d_g.p, plist.size() * sizeof(GroupRec), cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    }

    res->edges.resize(E);
    if (E) memcpy(static_cast<void*>(res->edges.data()), stage, E * sizeof(pedk_edge));
    auto& quirks = res->quirks;
    for (const GroupRec& g : groups) {
        bool pc = g.hasc[0] | g.hasc[1] | g.hasc[2] | g.hasc[3];
        bool pt = g.hast[0] | g.hast[1] | g.hast[2] | g.hast[3];
        if (!(pc && pt)) continue;  // inner join
        unsigned tag = (unsigned)(g.p >> (2 * (k - 1) - 6));
        bool first_c = h_first[tag] != ~0ull && (h_first[tag] >> 2) == g.p;
        bool first_t = h_first[64 + tag] != ~0ull && (h_first[64 + tag] >> 2) == g.p;
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
// caller that only wants the count does not pay for the text.
void finalise(pedk_result* res) {
    if (res->finalised) return;
    const int k = res->k;
    auto& quirks = res->quirks;
    std::sort(res->edges.begin(), res->edges.end(),
              [](const pedk_edge& x, const pedk_edge& y) { return x.prefix < y.prefix; });
    std::sort(quirks.begin(), quirks.end());
    // ---- <target>.kmer.snp ----
    std::string& t = res->text;
    t.reserve(res->edges.size() * 64);
    size_t qi = 0;
    for (const pedk_edge& e : res->edges) {
        if (e.quirk) {
            while (qi < quirks.size() && quirks[qi].first != e.prefix) ++qi;
            t += quirks[qi].second;
            continue;
        }
        for (int j = 0; j < k - 1; ++j) t.push_back("ACGT"[(e.prefix >> (2 * (k - 2 - j))) & 3]);
        t += '\t';
        for (int x = 0; x < 4; ++x)
            if (e.cont_mask & (1u << x)) t.push_back("ACGT"[x]);
        t += '\t';
        for (int x = 0; x < 4; ++x)
            if (e.alt_mask & (1u << x)) t.push_back("ACGT"[x]);
        for (int x = 0; x < 4; ++x) { t += '\t'; append_u(t, e.control[x]); }
        for (int x = 0; x < 4; ++x) { t += '\t'; append_u(t, e.target[x]); }
        t += '\n';
    }
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
void ensure_host_reads(pedk_sample* s) {
    if (s->have_reads) return;
    pedk_ctx* ctx = s->ctx;
    for (auto& v : s->h_reads) v.clear();
    for (const ReadGroup& g : s->groups) {
        if (!g.n) continue;
        if (!g.words.p) throw PedkError(PEDK_E_INVAL, "packed reads were released");
        std::vector<uint64_t> w(g.n * g.W);
        std::vector<uint8_t> pal(g.n);
        PEDK_CUDA_TRY(cudaMemcpyAsync(w.data(), g.words.p, w.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PEDK_CUDA_TRY(cudaMemcpyAsync(pal.data(), g.pal.p, g.n, cudaMemcpyDeviceToHost, ctx->stream));
        PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->stats.d2h_bytes += w.size() * 8 + g.n;
        std::string fwd((size_t)g.L, 'A'), rev((size_t)g.L, 'A');
        for (uint64_t r = 0; r < g.n; ++r) {
            const uint64_t* rw = w.data() + r * g.W;
            for (int i = 0; i < g.L; ++i) {
                unsigned c = (unsigned)(rw[i >> 5] >> (62 - 2 * (i & 31))) & 3u;
                fwd[(size_t)i] = "ACGT"[c];
                rev[(size_t)(g.L - 1 - i)] = "ACGT"[c ^ 3u];
            }
            unsigned tf = (unsigned)(rw[0] >> 58);
            s->h_reads[tf].push_back(fwd);
            if (!pal[r]) {
                unsigned tr = (unsigned)((base_code((unsigned char)rev[0]) << 4) | (base_code((unsigned char)rev[1]) << 2) |
                                         base_code((unsigned char)rev[2]));
                s->h_reads[tr].push_back(rev);
            }
        }
    }
    for (auto& v : s->h_reads) std::sort(v.begin(), v.end());
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
    for (uint64_t i = s->tag_start[tag]; i < s->tag_start[tag + 1]; ++i) {
        t += kmer_string(s->h_kmer[i], s->k);
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
        t += kmer_string(p, s->k - 1);
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

extern "C" int pedk_kmer_edges(pedk_ctx* ctx, const void* control_fastq, uint64_t control_bytes,
                               const void* target_fastq, uint64_t target_bytes, int where, int k, int clipping,
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
            rc = pedk_sample_borrow_fastq_device(c, const_cast<void*>(control_fastq), control_bytes,
                                                 (control_bytes + T - 1) / T * T + T);
    } else {
        if (rc == PEDK_OK) rc = pedk_sample_reserve(c, control_bytes);
        if (rc == PEDK_OK) rc = pedk_sample_reserve(t, target_bytes);
        if (rc == PEDK_OK)
            rc = where ? pedk_sample_add_fastq_device(t, target_fastq, target_bytes)
                       : pedk_sample_add_fastq(t, target_fastq, target_bytes);
        if (rc == PEDK_OK)
            rc = where ? pedk_sample_add_fastq_device(c, control_fastq, control_bytes)
                       : pedk_sample_add_fastq(c, control_fastq, control_bytes);
    }
    // ped.pl:291-297 ingests the target first, and `$clipping` is a global: a value chosen
    // automatically for the target is an explicit clipping= for the control
    pedk_ingest_info ti;
    if (rc == PEDK_OK) rc = pedk_sample_ingest(t, clipping, &ti);
    if (rc == PEDK_OK && ti.auto_clipped) clipping = ti.clipping_used;
    if (rc == PEDK_OK) rc = pedk_sample_count(t, k, nullptr);
    if (rc == PEDK_OK) rc = pedk_sample_release_reads(t);
    if (rc == PEDK_OK) rc = pedk_sample_ingest(c, clipping, nullptr);
    if (rc == PEDK_OK) rc = pedk_sample_count(c, k, nullptr);
    if (rc == PEDK_OK) rc = pedk_sample_release_reads(c);
    if (rc == PEDK_OK) rc = pedk_join(ctx, c, t, out);
    pedk_sample_destroy(c);
    pedk_sample_destroy(t);
    pedk::trace(ctx, "samples destroyed (step end)");
    return rc;
}
