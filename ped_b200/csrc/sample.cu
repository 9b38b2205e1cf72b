// One sample: FASTQ text in HBM -> unique packed reads (S0+S1) -> canonical
// k-mer count table (S2-S4).  Host code here only sequences kernels and makes
// the few decisions ped.pl makes from whole-file statistics (read-length mode).
#include "sample.h"

#include "comm.h"
#include "plan.h"

#include <string.h>

#include <algorithm>

using namespace pedk;

namespace {

constexpr size_t FQ_ALIGN = INGEST_TILE_BYTES;
constexpr int NBINS_MAX = 256;  // bins per binning level (bucket.cu)

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void fq_reserve(pedk_sample* s, size_t want) {
    size_t cap = round_up(want ? want : 1, FQ_ALIGN) + FQ_ALIGN;
    if (s->fq.p && s->fq.n >= cap) return;
    size_t grow = s->fq.n + s->fq.n / 2;
    if (cap < grow) cap = round_up(grow, FQ_ALIGN);
    DBuf<uint8_t> nb(s->ctx, cap);
    if (s->fq_len) {
        StageScope sc(s->ctx, PEDK_ST_COPY, 0);
        PEDK_CUDA_TRY(cudaMemcpyAsync(nb.p, s->fq.p, s->fq_len, cudaMemcpyDeviceToDevice, s->ctx->stream));
    }
    s->fq = std::move(nb);
}

uint64_t read_scalar(pedk_ctx* ctx, const unsigned long long* dev, const char* label) {
    trace(ctx, "  (host enqueued)");
    PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars, dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += 8;
    trace(ctx, label);
    return ctx->h_scalars[0];
}

struct GroupPlan {
    uint32_t clip;
    int whole;
    uint64_t r_limit;
};

}  // namespace

namespace pedk {
// Exact dedup of n packed canonical reads -> a ReadGroup of the sample.
void dedup_group(pedk_sample* s, int L, int W, DBuf<uint64_t>& words, DBuf<uint8_t>& pal, uint64_t n_c,
                 const uint8_t* valid) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    // the table holds 32-bit read indices and its capacity (a power of two >= 2 n) must fit 32 bits
    if (n_c > (1ull << 30))
        throw PedkError(PEDK_E_UNSUPPORTED, "more than 2^30 packed reads in one length group on one GPU");
    uint64_t cap64 = 1024;
    while (cap64 < 2 * n_c) cap64 <<= 1;
    const uint32_t cap = (uint32_t)cap64;
    DBuf<uint32_t> table(ctx, cap);
    DBuf<uint8_t> uniq(ctx, n_c);
    DBuf<uint64_t> pos(ctx, n_c);
    DBuf<uint64_t> ws2(ctx, scan_workspace_words(n_c));
    {
        StageScope sc(ctx, PEDK_ST_DEDUP, 4);
        PEDK_CUDA_TRY(cudaMemsetAsync(table.p, 0xFF, table.bytes(), st));
        launch_dedup(words.p, W, (uint32_t)n_c, table.p, cap, valid, uniq.p, st);
        exclusive_scan_u8(uniq.p, pos.p, n_c, ctx->scalars, ws2.p, st);
    }
    uint64_t U = read_scalar(ctx, ctx->scalars, "dedup");
    table.reset();
    ReadGroup g;
    g.L = L;
    g.W = W;
    g.n = U;
    g.words = DBuf<uint64_t>(ctx, U * W);
    g.pal = DBuf<uint8_t>(ctx, U);
    {
        StageScope sc(ctx, PEDK_ST_DEDUP, 2);
        launch_compact_reads(words.p, pal.p, uniq.p, pos.p, W, (uint32_t)n_c, g.words.p, g.pal.p, st);
        PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 1, 0, sizeof(unsigned long long), st));
        launch_count_flags(g.pal.p, U, ctx->scalars + 1, st);
    }
    g.n_pal = read_scalar(ctx, ctx->scalars + 1, "compact reads");
    s->groups.push_back(std::move(g));
}

}  // namespace pedk

namespace {

// Build one read-length group: chunk -> pack -> (multi-GPU: route every read to the rank
// that owns its hash) -> dedup -> compact.  R may be 0 on a rank of a multi-GPU job.
void build_group(pedk_sample* s, const GroupPlan& plan, const uint64_t* seq_start, const uint32_t* rlen,
                 const uint8_t* rflag, uint64_t R, uint64_t* strand_reads) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int L = (int)plan.clip;
    const int W = (L + 31) / 32;
    uint64_t n_c = 0;
    DBuf<uint64_t> words;
    DBuf<uint8_t> pal;
    if (R) {
        DBuf<uint32_t> chunks(ctx, R);
        DBuf<uint64_t> base(ctx, R);
        DBuf<uint64_t> ws(ctx, scan_workspace_words(R));
        {
            StageScope sc(ctx, PEDK_ST_INGEST, 4);
            launch_chunk_count(rlen, rflag, R, plan.r_limit, plan.clip, plan.whole, chunks.p, st);
            exclusive_scan_u32(chunks.p, base.p, R, ctx->scalars, ws.p, st);
        }
        n_c = read_scalar(ctx, ctx->scalars, "chunk scan");
        words = DBuf<uint64_t>(ctx, n_c * W);
        pal = DBuf<uint8_t>(ctx, n_c);
        if (n_c) {
            StageScope sc(ctx, PEDK_ST_INGEST, 1);
            launch_pack(s->fq.p, seq_start, chunks.p, base.p, R, plan.clip, W, words.p, pal.p, st);
        }
    }
    *strand_reads += 2 * n_c;
    if (ctx->nranks > 1) {
        DestSpec spec;
        spec.mode = DEST_READ;
        spec.stride = W;
        DBuf<uint64_t> mine;
        uint64_t n_mine = 0;
        exchange_records(ctx, words.p, n_c, spec, &mine, &n_mine);
        words = std::move(mine);
        n_c = n_mine;
        pal = DBuf<uint8_t>(ctx, n_c);
        StageScope sc(ctx, PEDK_ST_EXCHANGE, 1);
        launch_pal_flags(words.p, W, L, n_c, pal.p, st);
    }
    if (n_c == 0) return;
    dedup_group(s, L, W, words, pal, n_c);
}

void do_ingest(pedk_sample* s, int clipping) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    s->groups.clear();
    s->reads_released = false;
    s->info = pedk_ingest_info{};
    s->info.fastq_bytes = s->fq_len;
    s->have_table = s->have_reads = false;
    const size_t n = s->fq_len;
    if (clipping < 0) throw PedkError(PEDK_E_INVAL, "clipping must be >= 0");
    const bool multi = ctx->nranks > 1;  // collective: a rank without input still takes part
    if (n == 0 && !multi) {
        s->ingested = true;
        return;
    }
    if (s->copy_pending) {
        StageScope sc(ctx, PEDK_ST_COPY, 0);  // time the kernels had to wait for the text
        PEDK_CUDA_TRY(cudaStreamWaitEvent(st, s->ev_copied, 0));
        s->copy_pending = false;
    }
    uint64_t R = 0, TL = 0;
    uint8_t last_byte = '\n';
    DBuf<uint32_t> tile_counts;
    DBuf<uint64_t> tile_prefix, ws;
    if (n) {
        fq_reserve(s, n);
        const size_t padded = round_up(n, FQ_ALIGN);
        PEDK_CUDA_TRY(cudaMemsetAsync(s->fq.p + n, 0, padded - n, st));
        const size_t tiles = padded / FQ_ALIGN;
        // ---- record structure: newline count, scan, line index ----
        tile_counts = DBuf<uint32_t>(ctx, tiles);
        tile_prefix = DBuf<uint64_t>(ctx, tiles);
        ws = DBuf<uint64_t>(ctx, scan_workspace_words(tiles));
        {
            StageScope sc(ctx, PEDK_ST_INGEST, 4);
            launch_newline_count(s->fq.p, n, tile_counts.p, st);
            exclusive_scan_u32(tile_counts.p, tile_prefix.p, tiles, ctx->scalars, ws.p, st);
        }
        PEDK_CUDA_TRY(cudaMemcpyAsync(&last_byte, s->fq.p + n - 1, 1, cudaMemcpyDeviceToHost, st));
        uint64_t NL = read_scalar(ctx, ctx->scalars, "newline scan");
        TL = NL + (last_byte != '\n' ? 1 : 0);
        R = TL >= 2 ? (TL - 2) / 4 + 1 : 0;  // records that have a sequence line (line 4r+1)
        if (s->line_format) R = TL;
    }
    s->info.records = R;
    if (R == 0 && !multi) {
        s->fq.reset();
        s->fq_len = 0;
        s->ingested = true;
        return;
    }
    DBuf<uint64_t> seq_start(ctx, R), seq_end(ctx, R);
    DBuf<uint32_t> rlen(ctx, R);
    DBuf<uint8_t> rflag(ctx, R);
    DBuf<unsigned long long> hist(ctx, LEN_HIST_BINS + 4);
    unsigned* minmax = reinterpret_cast<unsigned*>(ctx->scalars + 6);
    {
        StageScope sc(ctx, PEDK_ST_INGEST, 2);
        PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
        const unsigned preset[2] = {0xffffffffu, 0u};
        PEDK_CUDA_TRY(cudaMemcpyAsync(minmax, preset, sizeof preset, cudaMemcpyHostToDevice, st));
        if (R) {
            if (s->line_format) PEDK_CUDA_TRY(cudaMemsetAsync(seq_start.p, 0, sizeof(uint64_t), st));
            launch_line_index(s->fq.p, n, tile_prefix.p, seq_start.p, seq_end.p, R, s->line_format ? 1 : 0, st);
            if (last_byte != '\n' && (s->line_format || ((TL - 1) & 3) == 1)) {
                // the file ends inside a sequence line: it has no terminating newline
                uint64_t end = n;
                PEDK_CUDA_TRY(cudaMemcpyAsync(seq_end.p + (R - 1), &end, sizeof end, cudaMemcpyHostToDevice, st));
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            }
            launch_len_stats(seq_start.p, seq_end.p, R, minmax, st);
        }
    }
    // ---- fast path: every record of the sample has the same length and clipping= is not
    //      given -> ped.pl's fixed-length mode; test, pack and canonicalise in one pass ----
    if (clipping == 0) {
        unsigned mm[2];
        PEDK_CUDA_TRY(cudaMemcpyAsync(mm, minmax, sizeof mm, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        uint64_t has = R ? 1 : 0, uni = (R && mm[0] == mm[1]) ? 1 : 0, Lf = uni ? mm[0] : 0;
        uint64_t sum_has = has, sum_uni = uni, sum_l = Lf, sum_ll = Lf * Lf;
        if (multi) {
            unsigned long long v[4] = {has, uni, Lf, Lf * Lf};
            PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->scalars + 12, v, sizeof v, cudaMemcpyHostToDevice, st));
            comm_allreduce_sum(ctx, ctx->scalars + 12, 4);
            PEDK_CUDA_TRY(cudaMemcpyAsync(v, ctx->scalars + 12, sizeof v, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            sum_has = v[0]; sum_uni = v[1]; sum_l = v[2]; sum_ll = v[3];
        }
        // all ranks that hold records agree on one length  <=>  sum(L)^2 == n * sum(L^2)
        const bool uniform = sum_has && sum_uni == sum_has && sum_l * sum_l == sum_ll * sum_has;
        const uint64_t Lall = uniform ? sum_l / sum_has : 0;
        if (uniform && Lall >= 3 && Lall <= 32 * PEDK_MAX_WORDS) {
            const int L = (int)Lall, W = (L + 31) / 32;
            tile_counts.reset();
            tile_prefix.reset();
            ws.reset();
            DBuf<uint64_t> words(ctx, R * W);
            DBuf<uint8_t> pal(ctx, R), valid(ctx, R);
            {
                StageScope sc(ctx, PEDK_ST_INGEST, 1, 0, 0, PEDK_KC_PACK, R * (uint64_t)(L + 8 + 8 * W + 2));
                launch_pack_fixed(s->fq.p, seq_start.p, R, L, W, words.p, pal.p, valid.p, hist.p + LEN_HIST_BINS, st);
            }
            if (multi) comm_allreduce_sum(ctx, hist.p + LEN_HIST_BINS, 2);
            unsigned long long cnt[2];
            PEDK_CUDA_TRY(cudaMemcpyAsync(cnt, hist.p + LEN_HIST_BINS, sizeof cnt, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            trace(ctx, "pack (fixed length)");
            s->info.reads_kept = cnt[0];
            if (cnt[1]) {
                char b[200];
                snprintf(b, sizeof b,
                         "%llu reads hold a byte outside ACGTN (ped.pl would carry them as text; not supported)", cnt[1]);
                throw PedkError(PEDK_E_ALPHABET, b);
            }
            s->info.clipping_used = cnt[0] ? L : 0;
            s->fq.reset();
            s->fq_len = 0;
            uint64_t n_c = R;
            if (multi) {
                // only the kept reads travel: compact, route by hash, recompute the palindrome flags
                DBuf<uint64_t> pos(ctx, R), ws2(ctx, scan_workspace_words(R));
                {
                    StageScope sc(ctx, PEDK_ST_INGEST, 3);
                    exclusive_scan_u8(valid.p, pos.p, R, ctx->scalars, ws2.p, st);
                }
                uint64_t kept_local = read_scalar(ctx, ctx->scalars, "kept scan");
                DBuf<uint64_t> cw(ctx, kept_local * W);
                DBuf<uint8_t> cp(ctx, kept_local);
                if (R) {
                    StageScope sc(ctx, PEDK_ST_INGEST, 1);
                    launch_compact_reads(words.p, pal.p, valid.p, pos.p, W, (uint32_t)R, cw.p, cp.p, st);
                }
                DestSpec spec;
                spec.mode = DEST_READ;
                spec.stride = W;
                DBuf<uint64_t> mine;
                exchange_records(ctx, cw.p, kept_local, spec, &mine, &n_c);
                words = std::move(mine);
                pal = DBuf<uint8_t>(ctx, n_c);
                valid.reset();
                StageScope sc(ctx, PEDK_ST_EXCHANGE, 1);
                launch_pal_flags(words.p, W, L, n_c, pal.p, st);
            }
            if (n_c) dedup_group(s, L, W, words, pal, n_c, multi ? nullptr : valid.p);
            // strand_reads: what ped.pl writes to the TAG.tmp files (kept reads and their complements)
            s->info.strand_reads = 2 * (multi ? n_c : cnt[0]);
            s->info.n_groups = (int32_t)s->groups.size();
            for (const ReadGroup& g : s->groups) {
                s->info.unique_reads += 2 * g.n - g.n_pal;
                s->info.palindromes += g.n_pal;
            }
            s->ingested = true;
            return;
        }
    }
    if (R) {
        StageScope sc(ctx, PEDK_ST_INGEST, 1);
        launch_read_scan(s->fq.p, seq_start.p, seq_end.p, R, rlen.p, rflag.p, hist.p, hist.p + LEN_HIST_BINS, st);
    }
    // every rank decides the length mode from the statistics of the whole sample
    if (multi) comm_allreduce_sum(ctx, hist.p, LEN_HIST_BINS + 4);
    tile_counts.reset();
    tile_prefix.reset();
    ws.reset();
    std::vector<unsigned long long> h_hist(LEN_HIST_BINS + 4);
    PEDK_CUDA_TRY(cudaMemcpyAsync(h_hist.data(), hist.p, hist.bytes(), cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    trace(ctx, "line index + read scan");
    ctx->stats.d2h_bytes += hist.bytes();
    const uint64_t kept = h_hist[LEN_HIST_BINS];
    const uint64_t bad = h_hist[LEN_HIST_BINS + 1];
    s->info.reads_kept = kept;
    if (bad) {
        char b[200];
        snprintf(b, sizeof b, "%llu reads hold a byte outside ACGTN (ped.pl would carry them as text; not supported)",
                 (unsigned long long)bad);
        throw PedkError(PEDK_E_ALPHABET, b);
    }
    if (h_hist[LEN_HIST_BINS - 1]) throw PedkError(PEDK_E_UNSUPPORTED, "read longer than 1022 bases");
    int n_lengths = 0;
    uint32_t only_len = 0;
    for (uint32_t l = 0; l < LEN_HIST_BINS - 1; ++l)
        if (h_hist[l]) { ++n_lengths; only_len = l; }

    // ---- the length mode ped.pl would end up in (ped.pl:1437-1464, 1478-1527) ----
    std::vector<GroupPlan> plans;
    if (s->line_format) {
        // already strand-closed unique reads: one group per length, nothing is clipped
        for (uint32_t l = 0; l < LEN_HIST_BINS - 1; ++l)
            if (h_hist[l]) plans.push_back({l, 1, R});
        s->info.clipping_used = (int32_t)only_len;
    } else if (clipping > 0 && multi) {
        plans.push_back({(uint32_t)clipping, 0, R});
        s->info.clipping_used = n_lengths == 1 ? (int32_t)only_len : clipping;
    } else if (clipping > 0) {
        plans.push_back({(uint32_t)clipping, 0, R});
        // ped.pl:1526 logs the length of the last kept read even then
        std::vector<uint32_t> hl(R);
        std::vector<uint8_t> hf(R);
        size_t tail = R < 65536 ? R : 65536;
        for (;;) {
            PEDK_CUDA_TRY(cudaMemcpyAsync(hl.data(), rlen.p + (R - tail), tail * 4, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaMemcpyAsync(hf.data(), rflag.p + (R - tail), tail, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            bool found = false;
            for (size_t i = tail; i-- > 0;)
                if ((hf[i] & 1) == 0) { s->info.clipping_used = (int32_t)hl[i]; found = true; break; }
            if (found || tail == R) break;
            tail = R;
        }
    } else if (n_lengths <= 1) {
        s->info.clipping_used = (int32_t)only_len;
        if (kept) plans.push_back({only_len, 1, R});
    } else if (multi) {
        throw PedkError(PEDK_E_UNSUPPORTED,
                        "mixed read lengths without clipping= need the file order of the whole sample: "
                        "run on one GPU or pass clipping=");
    } else {
        // mixed lengths, no clipping=: reads before the first length change were already
        // written unclipped (ped.pl:1440-1464), then everything is re-read and clipped
        std::vector<uint32_t> hl(R);
        std::vector<uint8_t> hf(R);
        PEDK_CUDA_TRY(cudaMemcpyAsync(hl.data(), rlen.p, R * 4, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(hf.data(), rflag.p, R, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += R * 5;
        uint32_t L0 = 0;
        uint64_t r_change = R;
        bool have_first = false;
        for (uint64_t r = 0; r < R; ++r) {
            if (hf[r] & 1) continue;
            // ped.pl:1440 `$readLength != $prev and $prev != 0`
            if (have_first && L0 != 0 && hl[r] != L0) { r_change = r; break; }
            if (!have_first || L0 == 0) { L0 = hl[r]; have_first = true; }
        }
        uint64_t sum = 0;
        uint32_t clip = 0;
        for (uint32_t l = LEN_HIST_BINS - 1; l-- > 0;) {
            if (!h_hist[l]) continue;
            sum += h_hist[l];
            if (sum * 100 / kept >= 90) { clip = l; break; }
        }
        if (clip > 200) clip = clip / (clip / 100);
        else if (clip < 75) clip = 75;
        if (r_change == R) {
            // the only other length is 0 and it came first: ped.pl never raises its flag
            s->info.clipping_used = (int32_t)L0;
            plans.push_back({L0, 1, R});
        } else {
            s->info.clipping_used = (int32_t)clip;
            s->info.auto_clipped = 1;
            if (L0 != clip && L0 != 0) plans.push_back({L0, 1, r_change});
            plans.push_back({clip, 0, R});
        }
    }
    hist.reset();

    uint64_t strand_reads = 0;
    for (const GroupPlan& p : plans) {
        if (p.clip < 3) continue;  // ped.pl: no TAG handle is open for such lines, they go nowhere
        if (p.clip > 32 * PEDK_MAX_WORDS) throw PedkError(PEDK_E_UNSUPPORTED, "reads longer than 512 bases");
        build_group(s, p, seq_start.p, rlen.p, rflag.p, R, &strand_reads);
    }
    s->info.strand_reads = strand_reads;
    s->info.n_groups = (int32_t)s->groups.size();
    for (const ReadGroup& g : s->groups) {
        s->info.unique_reads += 2 * g.n - g.n_pal;
        s->info.palindromes += g.n_pal;
    }
    s->fq.reset();
    s->fq_len = 0;
    s->ingested = true;
}

// Sorted instances -> table rows.  The rows first land in the idle half of the sort's
// ping-pong buffer (room for 2/3 of n rows: enough unless more than a third of the
// instances are distinct), then move to exact-size arrays.
void count_rows_into(pedk_ctx* ctx, const uint64_t* sorted, size_t n, uint64_t* spare, int k, int strand,
                     const uint64_t* corr_key, const uint32_t* corr_cnt, uint64_t n_corr, DBuf<uint64_t>* rkey,
                     DBuf<uint32_t>* rval, uint64_t* n_rows, uint64_t* n_heads, int stage) {
    cudaStream_t st = ctx->stream;
    uint64_t cap = (2 * n) / 3;
    uint64_t* tmp_key = spare;
    uint32_t* tmp_val = reinterpret_cast<uint32_t*>(spare + cap);
    DBuf<uint64_t> big_key;
    DBuf<uint32_t> big_val;
    for (int attempt = 0;; ++attempt) {
        {
            StageScope sc(ctx, stage, 1);
            PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 8, 0, 2 * sizeof(unsigned long long), st));
            launch_count_rows(sorted, n, k, strand, 0, corr_key, corr_cnt, n_corr, tmp_key, tmp_val, cap,
                              ctx->scalars + 8, st);
        }
        PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars + 8, ctx->scalars + 8, 2 * sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += 16;
        trace(ctx, "count rows");
        *n_rows = ctx->h_scalars[8];
        *n_heads = ctx->h_scalars[9];
        if (*n_rows <= cap) break;
        if (attempt) throw PedkError(PEDK_E_INTERNAL, "row count changed between two passes");
        // mostly distinct k-mers (very low coverage): run again into arrays of the exact size
        cap = *n_rows;
        big_key = DBuf<uint64_t>(ctx, cap);
        big_val = DBuf<uint32_t>(ctx, cap);
        tmp_key = big_key.p;
        tmp_val = big_val.p;
    }
    *rkey = DBuf<uint64_t>(ctx, *n_rows);
    *rval = DBuf<uint32_t>(ctx, *n_rows);
    if (*n_rows) {
        StageScope sc(ctx, stage, 0);
        PEDK_CUDA_TRY(cudaMemcpyAsync(rkey->p, tmp_key, *n_rows * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(rval->p, tmp_val, *n_rows * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
}

// ---- bucketed counting (bucket.cu): reads -> level-A bins (-> their owner rank) -> buckets ->
//      shared-memory hash tables -> strand rows.  Collective on several ranks. ----
int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    int x = atoi(v);
    return x < lo ? lo : x > hi ? hi : x;
}

// How many passes the instance stream of this sample is counted in (SURVEY.md 7.4 / 5 "prefix-range
// chunking"; the reference's equivalent is `sort -S 1M -T`, ped.pl:89, behind the 64-way fan-out of
// ped.pl:790-812).  One pass needs the level-A bins (+1/8) and the buckets at once; when that exceeds
// what the device has free, the bins are split into 2, 4, ... groups and the (small) packed reads are
// sliced once per group.  Collective on several ranks (the largest need decides).
int choose_count_passes(pedk_ctx* ctx, uint64_t m_rank, size_t ksz, int nbA) {
    const int R = ctx->nranks;
    int max_pass = nbA / R;  // every rank keeps at least one bin per pass
    if (max_pass < 1) max_pass = 1;
    int forced = env_int("PEDK_COUNT_PASSES", 0, 0, 256);
    int P = 1;
    if (forced > 0) {
        while (P < forced) P <<= 1;
    } else {
        // per pass: level-A segments (expected + 1/8 + slack per (bin, source)) + buckets (expected + 1/8 +
        // 16 sqrt); over all passes: the rows that come out (at most one per three instances at the
        // coverages a comparison makes sense at).  A rank receives about what it sends (the hash
        // spreads the k-mers evenly).
        auto need = [&](int passes) {
            const double inst = (double)m_rank * 1.05;
            return inst / passes * ksz * (1.125 + 1.2) + (double)nbA / passes * 4096 * ksz + inst * 0.34 * 12;
        };
        // blocks the context's allocator holds but nobody uses count as free
        const size_t cached = ctx->reserved_bytes > ctx->cur_bytes ? ctx->reserved_bytes - ctx->cur_bytes : 0;
        const char* cap_env = getenv("PEDK_DEVICE_BUDGET_MB");  // tests: pretend the device is small
        // cudaMemGetInfo takes the driver's global lock: next to a monitoring tool that polls the device
        // (nvidia-smi every 100 ms in bench.py) it was seen to take anything from 0.1 to 70 ms, per sample,
        // inside the step.  Ask once, then follow what this context's own allocator took since; ask again
        // only when one pass would need more than half of what that estimate leaves.
        double free_est = (double)ctx->mem_free_seen + ((double)ctx->mem_reserved_seen - (double)ctx->reserved_bytes);
        const bool ask = !ctx->mem_free_seen || (cap_env && *cap_env) || free_est <= 0 ||
                         need(1) * 2.0 > (free_est + (double)cached) * 0.90;
        if (ask) {
            size_t free_b = 0, total_b = 0;
            PEDK_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
            ctx->mem_free_seen = free_b;
            ctx->mem_reserved_seen = ctx->reserved_bytes;
            free_est = (double)free_b;
        }
        double budget = (free_est + (double)cached) * 0.90;
        if (cap_env && *cap_env) budget = std::min(budget, atof(cap_env) * 1048576.0);
        while (P < max_pass && need(P) > budget) P <<= 1;
    }
    return P > max_pass ? max_pass : P;
}

void count_bucketed(pedk_sample* s, int k, uint64_t* m_out) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const bool multi = ctx->nranks > 1;
    const int R = ctx->nranks, me = ctx->rank;
    const int nbits = 2 * k;
    const int a = kmer_bits_a(k);
    const int b = env_int("PEDK_BUCKET_B", kmer_bits_b(k), nbits - a < 1 ? 0 : 1, kmer_bits_b(k));
    const int log_slots = env_int("PEDK_BUCKET_SLOTS_LOG2", 13, 10, 13);
    const int nbA = 1 << a;
    const int wide = (nbits - a) > 32 ? 1 : 0;
    const size_t ksz = wide ? 8 : 4;
    const int exch_stage = multi ? PEDK_ST_EXCHANGE : PEDK_ST_EXTRACT;

    uint64_t packed_bytes = 0, m_local = 0;
    for (const ReadGroup& g : s->groups)
        if (g.L >= k && g.n) {
            packed_bytes += g.n * (uint64_t)g.W * 8;
            m_local += g.n * (uint64_t)(g.L - k + 1);
        }
    const unsigned TILE = sub_tile_keys();
    DBuf<unsigned long long> histA(ctx, (size_t)(NBINS_MAX + 1) * (PEDK_MAX_RANKS + 1));
    std::vector<unsigned long long> all((size_t)(NBINS_MAX + 1) * R, 0);
    uint64_t m_max = m_local;
    int P = choose_count_passes(ctx, m_local, ksz, nbA);
    if (multi) {
        // one all-gather: what every rank slices (sizes the segments) and the passes it needs (the largest decides)
        const unsigned long long mine[2] = {m_local, (unsigned long long)P};
        PEDK_CUDA_TRY(cudaMemcpyAsync(histA.p, mine, 16, cudaMemcpyHostToDevice, st));
        comm_allgather_u64(ctx, histA.p, histA.p + 8, 2);
        std::vector<unsigned long long> mm((size_t)2 * R);
        PEDK_CUDA_TRY(cudaMemcpyAsync(mm.data(), histA.p + 8, mm.size() * 8, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        for (int r = 0; r < R; ++r) {
            m_max = std::max<uint64_t>(m_max, mm[(size_t)2 * r]);
            P = std::max(P, (int)mm[(size_t)2 * r + 1]);
        }
    }
    s->cinfo.count_passes = P;
    const bool force_exact = getenv("PEDK_BUCKET_EXACT") && atoi(getenv("PEDK_BUCKET_EXACT")) != 0;
    uint64_t m_total = 0;
    s->n_rows = 0;
    s->D = 0;
    uint64_t row_cap = 0;  // several passes: the rows of all passes are collected in s->rkey / s->rval

    for (int pass = 0; pass < P; ++pass) {
        // what the level-A pass leaves for the level-B passes: the segments of this rank's bins
        std::vector<unsigned long long> seg_start, seg_end, seg_bin;
        DBuf<uint8_t> keysA;
        auto scatter_all = [&](unsigned long long* d_cur, const unsigned long long* d_lim, int* d_ovf) {
            PeerPtrs dst;
            if (multi) {
                void* peers[PEDK_MAX_RANKS];
                comm_peer_pointers(ctx, keysA.p, peers);
                for (int d = 0; d < R; ++d) dst.p[d] = static_cast<uint64_t*>(peers[d]);
            } else {
                dst.p[0] = reinterpret_cast<uint64_t*>(keysA.p);
            }
            StageScope sc(ctx, exch_stage, 1, 0, 0, PEDK_KC_PART, packed_bytes + m_local / P * ksz);
            // nobody may store into a buffer before its owner is done with the previous contents
            if (multi) comm_barrier(ctx);
            for (const ReadGroup& g : s->groups)
                if (g.L >= k && g.n)
                    launch_part_scatter(g.words.p, g.W, g.L, k, g.n, a, R, wide, d_cur, d_lim, d_ovf, dst, P, pass, st);
        };
        auto sent_by_me = [&]() {
            uint64_t sent = 0;
            for (int bin = 0; bin < nbA; ++bin)
                if (bin % P == pass) sent += all[(size_t)me * (NBINS_MAX + 1) + bin];
            return sent;
        };

        // (1) OPTIMISTIC: the hash spreads the k-mers evenly over the bins, so every (source rank,
        // bin) run gets a segment of the expected size plus 1/8 (+4096) and the reads are sliced
        // ONCE.  If a run does not fit (a k-mer with millions of instances), fall through to (2).
        bool done = false;
        uint64_t sent_expected = 0;  // P > 1: known only after the pass (the filter drops the other bins)
        if (!force_exact) {
            LevelAPlan lp;
            plan_level_a_optimistic(a, R, me, m_max, &lp, P, pass);
            const uint64_t C = lp.C;
            keysA = DBuf<uint8_t>(ctx, peer_block_elems(lp.buffer_elems * ksz, multi, 1));
            DBuf<unsigned long long> plan(ctx, (size_t)3 * NBINS_MAX + 8);
            unsigned long long* d_cur = plan.p;
            unsigned long long* d_lim = plan.p + NBINS_MAX;
            unsigned long long* d_cnt = plan.p + 2 * NBINS_MAX;  // [NBINS_MAX] counts, then the overflow flag
            PEDK_CUDA_TRY(cudaMemcpyAsync(d_cur, lp.cur.data(), (size_t)NBINS_MAX * 8, cudaMemcpyHostToDevice, st));
            PEDK_CUDA_TRY(cudaMemcpyAsync(d_lim, lp.lim.data(), (size_t)NBINS_MAX * 8, cudaMemcpyHostToDevice, st));
            PEDK_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, ((size_t)NBINS_MAX + 8) * 8, st));
            int* d_ovf = reinterpret_cast<int*>(d_cnt + NBINS_MAX);
            scatter_all(d_cur, d_lim, d_ovf);
            // what every rank really stored: cursor - start per bin (+ the flag), to all ranks.  On
            // several ranks this all-gather is also the fence "every rank's stores have landed".
            launch_cursor_counts(d_cur, d_lim, C, NBINS_MAX, d_cnt, st);
            if (multi) {
                comm_allgather_u64(ctx, d_cnt, histA.p, NBINS_MAX + 1);
                PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), histA.p, all.size() * 8, cudaMemcpyDeviceToHost, st));
            } else {
                PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), d_cnt, ((size_t)NBINS_MAX + 1) * 8, cudaMemcpyDeviceToHost, st));
            }
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            ctx->stats.d2h_bytes += all.size() * 8;
            bool overflow = false;
            for (int r = 0; r < R; ++r) overflow |= all[(size_t)r * (NBINS_MAX + 1) + NBINS_MAX] != 0;
            trace(ctx, overflow ? "  bins: extract + hash + bin (optimistic sizes overflowed)"
                                : multi ? "  bins: extract + hash + bin + NVLink stores" : "  bins: extract + hash + bin");
            if (!overflow) {
                level_a_segments_optimistic(a, R, me, all.data(), &lp);
                seg_start = lp.seg_start;
                seg_end = lp.seg_end;
                seg_bin = lp.seg_bin;
                sent_expected = sent_by_me();
                if (P == 1 && sent_expected != m_local) throw PedkError(PEDK_E_INTERNAL, "level-A bin counts do not add up");
                done = true;
            }
        }
        if (!done) {
            // (2) EXACT: size the bins with a pass over the packed reads (no stores), then slice again
            {
                StageScope sc(ctx, exch_stage, 1, 0, 0, PEDK_KC_PART_COUNT, packed_bytes);
                PEDK_CUDA_TRY(cudaMemsetAsync(histA.p, 0, histA.bytes(), st));
                for (const ReadGroup& g : s->groups)
                    if (g.L >= k && g.n) launch_part_count(g.words.p, g.W, g.L, k, g.n, a, histA.p, st);
            }
            if (multi) {
                comm_allgather_u64(ctx, histA.p, histA.p + NBINS_MAX + 1, NBINS_MAX + 1);
                PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), histA.p + NBINS_MAX + 1, all.size() * 8, cudaMemcpyDeviceToHost, st));
            } else {
                PEDK_CUDA_TRY(cudaMemcpyAsync(all.data(), histA.p, ((size_t)NBINS_MAX + 1) * 8, cudaMemcpyDeviceToHost, st));
            }
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            ctx->stats.d2h_bytes += all.size() * 8;
            trace(ctx, "  bins: counts");
            uint64_t counted = 0;
            for (int bin = 0; bin < nbA; ++bin) counted += all[(size_t)me * (NBINS_MAX + 1) + bin];
            if (counted != m_local) throw PedkError(PEDK_E_INTERNAL, "level-A bin counts do not add up");
            LevelAPlan lp;
            plan_level_a_exact(a, R, me, all.data(), &lp, P, pass);
            seg_start = lp.seg_start;
            seg_end = lp.seg_end;
            seg_bin = lp.seg_bin;
            keysA.reset();
            keysA = DBuf<uint8_t>(ctx, peer_block_elems(lp.buffer_elems * ksz, multi, 1));
            DBuf<unsigned long long> plan(ctx, NBINS_MAX);
            PEDK_CUDA_TRY(cudaMemcpyAsync(plan.p, lp.cur.data(), (size_t)NBINS_MAX * 8, cudaMemcpyHostToDevice, st));
            scatter_all(plan.p, nullptr, nullptr);
            if (multi) comm_barrier(ctx);  // every rank's stores have landed
            if (ctx->trace) {
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                trace(ctx, multi ? "  bins: extract + hash + bin + NVLink stores" : "  bins: extract + hash + bin");
            }
        }
        // the segments of this rank's bins, their tiles
        const int n_seg = (int)seg_start.size();
        std::vector<unsigned> tile_prefix((size_t)n_seg + 1, 0);
        uint64_t m = 0;
        for (int sg = 0; sg < n_seg; ++sg) {
            const unsigned long long len = seg_end[(size_t)sg] - seg_start[(size_t)sg];
            const unsigned long long tiles = (len + TILE - 1) / TILE;
            if ((unsigned long long)tile_prefix[(size_t)sg] + tiles > 0xfffffff0ull)
                throw PedkError(PEDK_E_UNSUPPORTED, "more than 2^45 k-mer instances on one GPU");
            tile_prefix[(size_t)sg + 1] = tile_prefix[(size_t)sg] + (unsigned)tiles;
            m += len;
        }
        const unsigned n_tiles = tile_prefix[(size_t)n_seg];
        m_total += m;
        if (m == 0) continue;
        DBuf<unsigned long long> d_seg(ctx, (size_t)3 * n_seg);
        DBuf<unsigned> tprefix(ctx, (size_t)n_seg + 1);
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_seg.p, seg_start.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_seg.p + n_seg, seg_end.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_seg.p + 2 * n_seg, seg_bin.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(tprefix.p, tile_prefix.data(), ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, st));

        // (4)-(6) second binning level.  OPTIMISTIC first: the hash spreads a bin evenly over its 256
        // buckets, so every bucket gets room for the expected number of keys plus a margin and the
        // keys are scattered at once (no counting pass: k_sub_hist read the whole stream for it).  A
        // bucket that outgrows its room (one k-mer with very many instances) raises a flag and the
        // level is done again with exact sizes.  (7) one CTA per bucket: count, rows.  The rows first
        // land in the level-A buffer (free by then: room for a row per three 4-byte instances); with
        // one pass they stay there, with several they are collected in the sample's own arrays.
        const size_t n_buckets = (size_t)nbA * NBINS_MAX;
        DBuf<unsigned long long> offB(ctx, n_buckets + 1), curB(ctx, n_buckets), limB;
        DBuf<uint8_t> keysB;
        uint64_t cap = keysA.n / 12;
        uint64_t* tmp_key = reinterpret_cast<uint64_t*>(keysA.p);
        uint32_t* tmp_val = reinterpret_cast<uint32_t*>(tmp_key + cap);
        DBuf<uint64_t> big_key;
        DBuf<uint32_t> big_val;
        uint64_t rows_pass = 0, heads_pass = 0;
        int* d_ovfB = reinterpret_cast<int*>(ctx->scalars + 10);
        bool exact_b = force_exact || env_int("PEDK_BUCKET_B_EXACT", 0, 0, 1) != 0;
        for (;;) {
            const unsigned long long* endB = nullptr;
            PEDK_CUDA_TRY(cudaMemsetAsync(d_ovfB, 0, 8, st));
            if (!exact_b) {
                std::vector<int> bin_index((size_t)NBINS_MAX, -1);
                int nb_local = 0;
                for (int sg = 0; sg < n_seg; ++sg)
                    if (bin_index[(size_t)seg_bin[(size_t)sg]] < 0) bin_index[(size_t)seg_bin[(size_t)sg]] = nb_local++;
                const uint64_t mean = m / ((uint64_t)nb_local * NBINS_MAX) + 1;
                uint64_t root = 1;
                while (root * root < mean) ++root;
                // room per bucket: the instances of one k-mer all land in one bucket, so a bucket's size
                // varies like sqrt(mean x instances per k-mer), not like sqrt(mean): 1/8 + 16 sqrt(mean)
                // is 8 sigma at 30x coverage.  Deeper coverage overflows now and then and takes the
                // exact path below.
                const uint64_t capB = (mean + mean / 8 + 16 * root + 64 + 3) & ~3ull;
                keysB = DBuf<uint8_t>(ctx, (uint64_t)nb_local * NBINS_MAX * capB * ksz + 64);
                limB = DBuf<unsigned long long>(ctx, n_buckets);
                DBuf<int> d_bin_index(ctx, NBINS_MAX);
                PEDK_CUDA_TRY(cudaMemcpyAsync(d_bin_index.p, bin_index.data(), (size_t)NBINS_MAX * 4, cudaMemcpyHostToDevice, st));
                {
                    StageScope sc(ctx, PEDK_ST_SORT, 1);
                    launch_bucket_layout(d_bin_index.p, (unsigned)n_buckets, capB, offB.p, limB.p, st);
                    PEDK_CUDA_TRY(cudaMemcpyAsync(curB.p, offB.p, n_buckets * 8, cudaMemcpyDeviceToDevice, st));
                }
                {
                    StageScope sc(ctx, PEDK_ST_SORT, 1, 0, 0, PEDK_KC_SUB_SCATTER, 2 * m * ksz);
                    launch_sub_scatter(keysA.p, keysB.p, wide, d_seg.p, tprefix.p, n_seg, n_tiles, k, a, b, curB.p, limB.p,
                                       d_ovfB, st);
                }
                endB = curB.p;  // where every bucket's keys end
                // the verdict has to be in before the count kernel runs: its rows go into the level-A
                // buffer, which the exact path would have to read again
                PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars + 10, d_ovfB, 8, cudaMemcpyDeviceToHost, st));
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                if ((ctx->h_scalars[10] & 0xffffffffull) != 0) {
                    trace(ctx, "  buckets: optimistic sizes overflowed");
                    exact_b = true;
                    continue;
                }
                if (ctx->trace > 1) {
                    // debug: which buckets outgrew their room
                    std::vector<unsigned long long> hc(n_buckets), hl(n_buckets), ho(n_buckets);
                    int hf = 0;
                    PEDK_CUDA_TRY(cudaMemcpyAsync(hc.data(), curB.p, n_buckets * 8, cudaMemcpyDeviceToHost, st));
                    PEDK_CUDA_TRY(cudaMemcpyAsync(hl.data(), limB.p, n_buckets * 8, cudaMemcpyDeviceToHost, st));
                    PEDK_CUDA_TRY(cudaMemcpyAsync(ho.data(), offB.p, n_buckets * 8, cudaMemcpyDeviceToHost, st));
                    PEDK_CUDA_TRY(cudaMemcpyAsync(&hf, d_ovfB, 4, cudaMemcpyDeviceToHost, st));
                    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                    unsigned long long over = 0, tot = 0, mx = 0;
                    for (size_t i = 0; i < n_buckets; ++i) {
                        if (hc[i] > hl[i]) ++over;
                        tot += hc[i] - ho[i];
                        mx = std::max(mx, hc[i] - ho[i]);
                    }
                    fprintf(stderr, "[pedk debug] optimistic level B: m=%llu nb_local=%d capB=%llu flag=%d over=%llu total=%llu max=%llu\n",
                            (unsigned long long)m, nb_local, (unsigned long long)capB, hf, over, tot, mx);
                }
            } else {
                DBuf<unsigned> histB(ctx, n_buckets);
                DBuf<uint64_t> ws(ctx, scan_workspace_words(n_buckets));
                {
                    StageScope sc(ctx, PEDK_ST_SORT, 1, 0, 0, PEDK_KC_SUB_HIST, m * ksz);
                    PEDK_CUDA_TRY(cudaMemsetAsync(histB.p, 0, histB.bytes(), st));
                    launch_sub_hist(keysA.p, wide, d_seg.p, tprefix.p, n_seg, n_tiles, k, a, b, histB.p, st);
                }
                {
                    StageScope sc(ctx, PEDK_ST_SORT, 3);
                    exclusive_scan_u32(histB.p, reinterpret_cast<uint64_t*>(offB.p), n_buckets, ctx->scalars + 3, ws.p, st);
                    const unsigned long long total = m;
                    PEDK_CUDA_TRY(cudaMemcpyAsync(offB.p + n_buckets, &total, 8, cudaMemcpyHostToDevice, st));
                    PEDK_CUDA_TRY(cudaMemcpyAsync(curB.p, offB.p, n_buckets * 8, cudaMemcpyDeviceToDevice, st));
                }
                keysB.reset();
                keysB = DBuf<uint8_t>(ctx, m * ksz + 64);  // the count kernel reads it with aligned 16-byte loads
                {
                    StageScope sc(ctx, PEDK_ST_SORT, 1, 0, 0, PEDK_KC_SUB_SCATTER, 2 * m * ksz);
                    launch_sub_scatter(keysA.p, keysB.p, wide, d_seg.p, tprefix.p, n_seg, n_tiles, k, a, b, curB.p, nullptr,
                                       nullptr, st);
                }
                endB = offB.p + 1;
            }
            if (ctx->trace) {
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                trace(ctx, exact_b ? "  buckets: histogram + scan + scatter" : "  buckets: scatter (optimistic sizes)");
            }
            for (int attempt = 0;; ++attempt) {
                {
                    StageScope sc(ctx, PEDK_ST_COUNT, 1, 0, 0, PEDK_KC_BUCKET_COUNT, m * ksz);
                    PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 8, 0, 2 * sizeof(unsigned long long), st));
                    launch_bucket_count(keysB.p, wide, offB.p, endB, (unsigned)n_buckets, k, a, b, log_slots, 1, s->ckey.p,
                                        s->ccnt.p, s->n_corr, tmp_key, tmp_val, cap, ctx->scalars + 8, ctx->tile_ctr,
                                        ctx->err_flag, st);
                }
                PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars + 8, ctx->scalars + 8, 2 * sizeof(unsigned long long),
                                              cudaMemcpyDeviceToHost, st));
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                ctx->stats.d2h_bytes += 16;
                trace(ctx, "  buckets: count -> rows");
                rows_pass = ctx->h_scalars[8];
                heads_pass = ctx->h_scalars[9];
                if (rows_pass <= cap) break;
                if (attempt) throw PedkError(PEDK_E_INTERNAL, "row count changed between two passes");
                // mostly distinct k-mers (very low coverage): run again into arrays of the exact size
                cap = rows_pass;
                big_key = DBuf<uint64_t>(ctx, cap);
                big_val = DBuf<uint32_t>(ctx, cap);
                tmp_key = big_key.p;
                tmp_val = big_val.p;
            }
            break;
        }
        check_err_flag(ctx, "bucket count");
        ctx->stats.kc_bytes[PEDK_KC_BUCKET_COUNT] += rows_pass * 12;
        keysB.reset();
        if (P == 1) {
            s->n_rows = rows_pass;
            s->D = heads_pass;
            if (big_key.p) {
                s->rkey = std::move(big_key);
                s->rval = std::move(big_val);
            } else {
                // the rows stay in the level-A buffer: no copy to exact-size arrays
                s->row_key_in_store = tmp_key;
                s->row_val_in_store = tmp_val;
                s->row_store = std::move(keysA);
            }
        } else {
            if (s->n_rows + rows_pass > row_cap) {
                // grow: what the passes so far produced, extrapolated to all passes, plus a margin
                uint64_t want = (s->n_rows + rows_pass) * (uint64_t)P / (uint64_t)(pass + 1);
                want += want / 16 + 1024;
                DBuf<uint64_t> nk(ctx, want);
                DBuf<uint32_t> nv(ctx, want);
                if (s->n_rows) {
                    StageScope sc(ctx, PEDK_ST_COUNT, 0);
                    PEDK_CUDA_TRY(cudaMemcpyAsync(nk.p, s->rkey.p, s->n_rows * 8, cudaMemcpyDeviceToDevice, st));
                    PEDK_CUDA_TRY(cudaMemcpyAsync(nv.p, s->rval.p, s->n_rows * 4, cudaMemcpyDeviceToDevice, st));
                }
                s->rkey = std::move(nk);
                s->rval = std::move(nv);
                row_cap = want;
            }
            if (rows_pass) {
                StageScope sc(ctx, PEDK_ST_COUNT, 0);
                PEDK_CUDA_TRY(cudaMemcpyAsync(s->rkey.p + s->n_rows, tmp_key, rows_pass * 8, cudaMemcpyDeviceToDevice, st));
                PEDK_CUDA_TRY(cudaMemcpyAsync(s->rval.p + s->n_rows, tmp_val, rows_pass * 4, cudaMemcpyDeviceToDevice, st));
            }
            s->n_rows += rows_pass;
            s->D += heads_pass;
        }
    }
    // every instance of every rank was stored in exactly one pass
    if (!multi && m_total != m_local) throw PedkError(PEDK_E_INTERNAL, "passes do not add up to the sample's instances");
    *m_out = m_total;
}

void do_count(pedk_sample* s, int k) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!s->ingested) throw PedkError(PEDK_E_INVAL, "pedk_sample_count before pedk_sample_ingest");
    if (k < 4 || k > 32) throw PedkError(PEDK_E_INVAL, "k must be in [4, 32]");
    s->k = k;
    s->D = 0;
    s->n_rows = 0;
    s->n_corr = 0;
    s->rkey.reset();
    s->rval.reset();
    s->row_store.reset();
    s->ckey.reset();
    s->ccnt.reset();
    s->counted = false;
    s->have_table = false;
    s->cinfo = pedk_count_info{};
    s->cinfo.k = k;
    const int n_pass = (2 * k + 7) / 8;
    s->cinfo.sort_passes = n_pass;
    uint64_t M = 0, Mpal = 0;
    for (const ReadGroup& g : s->groups)
        if (g.L >= k) {
            M += g.n * (uint64_t)(g.L - k + 1);
            Mpal += g.n_pal * (uint64_t)(g.L - k + 1);
        }
    s->cinfo.instances = M;
    s->cinfo.correction_rows = Mpal;
    s->counted = true;
    const bool multi = ctx->nranks > 1;
    // multi-GPU: the decisions below are collective, so they use the totals over all ranks
    uint64_t M_all = M, Mpal_all = Mpal;
    if (multi) {
        unsigned long long v[2] = {M, Mpal};
        PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->scalars + 12, v, sizeof v, cudaMemcpyHostToDevice, st));
        comm_allreduce_sum(ctx, ctx->scalars + 12, 2);
        PEDK_CUDA_TRY(cudaMemcpyAsync(v, ctx->scalars + 12, sizeof v, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        M_all = v[0];
        Mpal_all = v[1];
    }
    if (M_all == 0) return;

    DBuf<unsigned long long> hist(ctx, 8 * 256);
    // extract (all windows, or only those of palindromic reads) -> [route every k-mer to the
    // rank that owns its hash] -> sort.  Returns which of a/b holds the sorted instances.
    auto extract_sort = [&](bool pal_only, uint64_t m_local, DBuf<uint64_t>& a, DBuf<uint64_t>& b, uint64_t& m_final,
                            int stage) -> int {
        PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
        a = DBuf<uint64_t>(ctx, m_local);
        if (pal_only) PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 2, 0, sizeof(unsigned long long), st));
        uint64_t off = 0;
        for (const ReadGroup& g : s->groups) {
            if (g.L < k || !g.n) continue;
            StageScope sc(ctx, PEDK_ST_EXTRACT, 1);
            if (pal_only) {
                if (g.n_pal) launch_extract_pal(g.words.p, g.pal.p, g.W, g.L, k, g.n, a.p, ctx->scalars + 2, st);
            } else {
                // the histograms come for free here unless the keys are about to change hands
                launch_extract(g.words.p, g.W, g.L, k, g.n, a.p + off, hist.p, multi ? 0 : n_pass, st);
                off += g.n * (uint64_t)(g.L - k + 1);
            }
        }
        m_final = m_local;
        if (multi) {
            DestSpec spec;
            spec.mode = DEST_KMER;
            spec.stride = 1;
            spec.k = k;
            DBuf<uint64_t> mine;
            exchange_records(ctx, a.p, m_local, spec, &mine, &m_final);
            a = std::move(mine);
        }
        if (multi || pal_only) {
            StageScope sc(ctx, PEDK_ST_EXTRACT, 1);
            launch_digit_hist(a.p, m_final, hist.p, n_pass, st);
        }
        b = DBuf<uint64_t>(ctx, m_final);
        trace(ctx, "  (extract enqueued)");
        int cur = radix_sort(ctx, a.p, b.p, nullptr, nullptr, m_final, n_pass, hist.p, stage);
        check_err_flag(ctx, "radix sort");
        trace(ctx, "extract + radix sort");
        return cur;
    };

    if (Mpal_all) {
        // windows of palindromic reads: each is emitted once below but stands for one strand
        // only; their sorted (k-mer, windows) table corrects the counts (Appendix A.3)
        DBuf<uint64_t> a, b;
        uint64_t m = 0;
        int cur = extract_sort(true, Mpal, a, b, m, PEDK_ST_EXTRACT);
        DBuf<uint64_t> ck;
        DBuf<uint32_t> cc;
        uint64_t heads;
        count_rows_into(ctx, cur ? b.p : a.p, m, cur ? a.p : b.p, k, 0, nullptr, nullptr, 0, &ck, &cc, &s->n_corr,
                        &heads, PEDK_ST_EXTRACT);
        // the rows come out unordered: sort them so the count kernel can bisect
        DBuf<uint64_t> ck2(ctx, s->n_corr);
        DBuf<uint32_t> cc2(ctx, s->n_corr);
        PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
        {
            StageScope sc(ctx, PEDK_ST_EXTRACT, 1);
            launch_digit_hist(ck.p, s->n_corr, hist.p, n_pass, st);
        }
        int c2 = radix_sort(ctx, ck.p, ck2.p, cc.p, cc2.p, s->n_corr, n_pass, hist.p, PEDK_ST_EXTRACT);
        check_err_flag(ctx, "radix sort (correction table)");
        s->ckey = std::move(c2 ? ck2 : ck);
        s->ccnt = std::move(c2 ? cc2 : cc);
    }
    const bool sort_mode = getenv("PEDK_COUNT_MODE") && !strcmp(getenv("PEDK_COUNT_MODE"), "sort");
    if (!sort_mode) {
        uint64_t m = 0;
        count_bucketed(s, k, &m);
        s->cinfo.instances = m;
        s->cinfo.sort_passes = 0;
    } else {
        DBuf<uint64_t> a, b;
        uint64_t m = 0;
        int cur = extract_sort(false, M, a, b, m, PEDK_ST_SORT);
        s->cinfo.instances = m;
        uint64_t* sorted = cur ? b.p : a.p;
        uint64_t* spare = cur ? a.p : b.p;
        count_rows_into(ctx, sorted, m, spare, k, 1, s->ckey.p, s->ccnt.p, s->n_corr, &s->rkey, &s->rval, &s->n_rows,
                        &s->D, PEDK_ST_COUNT);
    }
    s->cinfo.canonical_kmers = s->D;
}

}  // namespace

namespace pedk {

std::string kmer_string(uint64_t v, int k) {
    std::string r((size_t)k, 'A');
    for (int j = 0; j < k; ++j) r[(size_t)j] = "ACGT"[(v >> (2 * (k - 1 - j))) & 3];
    return r;
}

void build_rows(pedk_ctx* ctx, pedk_sample* control, pedk_sample* target, RowTable* out) {
    cudaStream_t st = ctx->stream;
    const int k = control->k;
    const int n_pass = (2 * k + 7) / 8;
    auto rows_of = [](pedk_sample* s) -> uint64_t { return !s ? 0 : s->host_only_table ? s->h_kmer.size() : s->n_rows; };
    uint64_t cap = rows_of(control) + rows_of(target);
    const bool multi = ctx->nranks > 1;
    out->n = 0;
    if (cap == 0 && !multi) return;
    DBuf<uint64_t> ka, kb;
    DBuf<uint32_t> va, vb;
    DBuf<unsigned long long> hist(ctx, 8 * 256);
    PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
    uint64_t rows = 0;
    pedk_sample* ss[2] = {control, target};
    bool hist_done = false;
    const bool device_rows = !(control && control->host_only_table) && !(target && target->host_only_table);
    if (multi && device_rows) {
        // rows move to the rank that holds their 3-nt bucket: contiguous bucket ranges, balanced
        // on the histogram of both samples over all ranks, so every (k-1)-mer group and every
        // first-of-bucket row is local to one rank and rank order is key order
        const uint64_t* sk[2];
        const uint32_t* sv[2];
        uint64_t sn[2];
        uint32_t ob[2];
        int n_src = 0;
        {
            StageScope sc(ctx, PEDK_ST_EXCHANGE, 2);
            for (int i = 0; i < 2; ++i) {
                pedk_sample* s = ss[i];
                if (!s) continue;
                sk[n_src] = s->row_key_ptr();
                sv[n_src] = s->row_val_ptr();
                sn[n_src] = s->n_rows;
                ob[n_src] = i ? 0x80000000u : 0u;
                launch_bucket_hist(sk[n_src], sn[n_src], k, hist.p, st);
                ++n_src;
            }
        }
        uint64_t local64[64], global64[64];
        PEDK_CUDA_TRY(cudaMemcpyAsync(local64, hist.p, sizeof local64, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));  // the local counts are on the host before they are summed in place
        comm_allreduce_sum(ctx, hist.p, 64);
        PEDK_CUDA_TRY(cudaMemcpyAsync(global64, hist.p, sizeof global64, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        route_rows(ctx, sk, sv, sn, ob, n_src, k, local64, global64, &ka, &va, &rows);
    } else {
        ka = DBuf<uint64_t>(ctx, cap);
        va = DBuf<uint32_t>(ctx, cap);
        // one rank: the digit histograms of the row sort are taken while the rows are copied
        hist_done = !multi;
        for (int i = 0; i < 2; ++i) {
            pedk_sample* s = ss[i];
            if (s && s->host_only_table) {
                // strand table read back from count files: upload it as rows
                uint64_t m = s->h_kmer.size();
                if (!m) continue;
                std::vector<uint32_t> v(s->h_cnt);
                if (i) for (auto& x : v) x |= 0x80000000u;
                StageScope sc(ctx, PEDK_ST_COPY, 0);
                PEDK_CUDA_TRY(cudaMemcpyAsync(ka.p + rows, s->h_kmer.data(), m * 8, cudaMemcpyHostToDevice, st));
                PEDK_CUDA_TRY(cudaMemcpyAsync(va.p + rows, v.data(), m * 4, cudaMemcpyHostToDevice, st));
                PEDK_CUDA_TRY(cudaStreamSynchronize(st));
                ctx->stats.h2d_bytes += m * 12;
                rows += m;
                hist_done = false;
                continue;
            }
            if (!s || !s->n_rows) continue;
            {
                StageScope sc(ctx, PEDK_ST_EXPAND, 1);
                launch_copy_rows(s->row_key_ptr(), s->row_val_ptr(), s->n_rows, ka.p + rows, va.p + rows,
                                 i ? 0x80000000u : 0u, hist_done ? hist.p : nullptr, n_pass, st);
            }
            rows += s->n_rows;
        }
        if (multi) {
            // (tables uploaded from count files: the staged all-to-all-v)
            PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, 64 * sizeof(unsigned long long), st));
            {
                StageScope sc(ctx, PEDK_ST_EXCHANGE, 1);
                launch_bucket_hist(ka.p, rows, k, hist.p, st);
            }
            comm_allreduce_sum(ctx, hist.p, 64);
            uint64_t h64[64];
            PEDK_CUDA_TRY(cudaMemcpyAsync(h64, hist.p, sizeof h64, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
            DestSpec spec;
            spec.mode = DEST_ROW;
            spec.stride = 2;
            spec.bucket_shift = 2 * k - 6;
            plan_bucket_owners(h64, ctx->nranks, spec.owner);
            DBuf<uint64_t> rec(ctx, 2 * rows), mine;
            uint64_t n_mine = 0;
            {
                StageScope sc(ctx, PEDK_ST_EXCHANGE, 1);
                launch_rows_pack(ka.p, va.p, rows, rec.p, st);
            }
            exchange_records(ctx, rec.p, rows, spec, &mine, &n_mine);
            rec.reset();
            rows = n_mine;
            ka = DBuf<uint64_t>(ctx, rows);
            va = DBuf<uint32_t>(ctx, rows);
            {
                StageScope sc(ctx, PEDK_ST_EXCHANGE, 1);
                launch_rows_unpack(mine.p, rows, ka.p, va.p, st);
            }
        }
    }
    if (rows == 0) return;
    kb = DBuf<uint64_t>(ctx, rows);
    vb = DBuf<uint32_t>(ctx, rows);
    if (!hist_done) {
        PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
        StageScope sc(ctx, PEDK_ST_ROWSORT, 1);
        launch_digit_hist(ka.p, rows, hist.p, n_pass, st);
    }
    int cur = radix_sort(ctx, ka.p, kb.p, va.p, vb.p, rows, n_pass, hist.p, PEDK_ST_ROWSORT);
    check_err_flag(ctx, "radix sort (rows)");
    trace(ctx, "row sort");
    out->key = std::move(cur ? kb : ka);
    out->val = std::move(cur ? vb : va);
    out->n = rows;
}

void ensure_host_table(pedk_sample* s) {
    if (s->have_table) return;
    pedk_ctx* ctx = s->ctx;
    if (!s->k) throw PedkError(PEDK_E_INVAL, "sample has not been counted");
    RowTable t;
    build_rows(ctx, s, nullptr, &t);
    s->h_kmer.resize(t.n);
    s->h_cnt.resize(t.n);
    if (t.n) {
        StageScope sc(ctx, PEDK_ST_COPY, 0);
        PEDK_CUDA_TRY(cudaMemcpyAsync(s->h_kmer.data(), t.key.p, t.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PEDK_CUDA_TRY(cudaMemcpyAsync(s->h_cnt.data(), t.val.p, t.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += t.n * 12;
    // bucket boundaries by the first three bases
    const int shift = 2 * s->k - 6;
    size_t i = 0;
    for (int tag = 0; tag < 64; ++tag) {
        while (i < s->h_kmer.size() && (int)(s->h_kmer[i] >> shift) < tag) ++i;
        s->tag_start[tag] = i;
    }
    s->tag_start[64] = s->h_kmer.size();
    s->have_table = true;
}

}  // namespace pedk

// ---------------- C ABI ----------------

extern "C" int pedk_sample_create(pedk_ctx* ctx, pedk_sample** out) {
    if (!ctx || !out) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    pedk_sample* s = new pedk_sample();
    s->ctx = ctx;
    *out = s;
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_sample_destroy(pedk_sample* s) {
    if (!s) return PEDK_OK;
    cudaSetDevice(s->ctx->device);
    if (s->copy_pending) cudaStreamSynchronize(s->ctx->copy_stream);
    if (s->ev_copied) cudaEventDestroy(s->ev_copied);
    delete s;
    return PEDK_OK;
}

extern "C" int pedk_sample_reserve(pedk_sample* s, uint64_t fastq_bytes) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    fq_reserve(s, fastq_bytes);
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

static int add_fastq(pedk_sample* s, const void* bytes, uint64_t n, cudaMemcpyKind kind) {
    if (!s || (!bytes && n)) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    if (!n) return PEDK_OK;
    pedk_ctx* ctx = s->ctx;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (s->ingested) {
        s->ingested = false;
        s->groups.clear();
    }
    if (s->copy_pending && !(s->fq.p && s->fq.n >= s->fq_len + n + 2 * INGEST_TILE_BYTES)) {
        // the buffer is about to move: earlier chunks must have landed first
        PEDK_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copied, 0));
    }
    fq_reserve(s, s->fq_len + n);
    if (kind == cudaMemcpyHostToDevice) {
        // host text goes through the copy stream so it overlaps kernels already queued
        PEDK_CUDA_TRY(cudaEventRecord(ctx->ev_alloc, ctx->stream));
        PEDK_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_alloc, 0));
        PEDK_CUDA_TRY(cudaMemcpyAsync(s->fq.p + s->fq_len, bytes, n, kind, ctx->copy_stream));
        if (!s->ev_copied) PEDK_CUDA_TRY(cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming));
        PEDK_CUDA_TRY(cudaEventRecord(s->ev_copied, ctx->copy_stream));
        s->copy_pending = true;
        ctx->stats.h2d_bytes += n;
    } else {
        StageScope sc(ctx, PEDK_ST_COPY, 0);
        PEDK_CUDA_TRY(cudaMemcpyAsync(s->fq.p + s->fq_len, bytes, n, kind, ctx->stream));
    }
    s->fq_len += n;
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_add_fastq(pedk_sample* s, const void* host_bytes, uint64_t n) {
    return add_fastq(s, host_bytes, n, cudaMemcpyHostToDevice);
}

extern "C" int pedk_sample_add_fastq_device(pedk_sample* s, const void* device_bytes, uint64_t n) {
    return add_fastq(s, device_bytes, n, cudaMemcpyDeviceToDevice);
}

extern "C" int pedk_sample_borrow_fastq_device(pedk_sample* s, void* device_bytes, uint64_t n, uint64_t capacity) {
    if (!s || (!device_bytes && n)) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    if (s->fq_len || s->copy_pending) throw PedkError(PEDK_E_INVAL, "pedk_sample_borrow_fastq_device: the sample already holds text");
    const uint64_t need = (n + INGEST_TILE_BYTES - 1) / INGEST_TILE_BYTES * INGEST_TILE_BYTES + INGEST_TILE_BYTES;
    if (capacity < need)
        throw PedkError(PEDK_E_INVAL, "pedk_sample_borrow_fastq_device: capacity must cover the text rounded up to 16 KiB plus 16 KiB");
    if (!n) return PEDK_OK;
    if (s->ingested) {
        s->ingested = false;
        s->groups.clear();
    }
    // a DBuf over memory the context's allocator does not know: releasing it is a no-op
    s->fq.reset();
    s->fq.ctx = s->ctx;
    s->fq.p = static_cast<uint8_t*>(device_bytes);
    s->fq.n = capacity;
    s->fq_len = n;
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_ingest(pedk_sample* s, int clipping, pedk_ingest_info* info) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_ingest(s, clipping);
    if (info) *info = s->info;
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_count(pedk_sample* s, int k, pedk_count_info* info) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_count(s, k);
    if (info) *info = s->cinfo;
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_table(pedk_sample* s, uint64_t* kmers, uint32_t* counts, uint64_t cap, uint64_t* n) {
    if (!s || !n) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    ensure_host_table(s);
    *n = s->h_kmer.size();
    if (kmers && counts) {
        if (cap < *n) throw PedkError(PEDK_E_INVAL, "pedk_sample_table: buffer too small");
        memcpy(kmers, s->h_kmer.data(), *n * 8);
        memcpy(counts, s->h_cnt.data(), *n * 4);
    }
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_table_digest(pedk_sample* s, uint64_t* rows, uint64_t* digest) {
    if (!s || !rows || !digest) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    pedk_ctx* ctx = s->ctx;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (s->host_only_table) {
        uint64_t acc = 0;
        for (size_t i = 0; i < s->h_kmer.size(); ++i)
            acc += mix64(s->h_kmer[i] * 0x9E3779B97F4A7C15ull + (uint64_t)s->h_cnt[i]);
        *rows = s->h_kmer.size();
        *digest = acc;
        return PEDK_OK;
    }
    if (!s->counted) throw PedkError(PEDK_E_INVAL, "sample has not been counted");
    PEDK_CUDA_TRY(cudaMemsetAsync(ctx->scalars + 10, 0, sizeof(unsigned long long), ctx->stream));
    {
        StageScope sc(ctx, PEDK_ST_EDGES, 1);
        launch_rows_digest(s->row_key_ptr(), s->row_val_ptr(), s->n_rows, ctx->scalars + 10, ctx->stream);
    }
    PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars + 10, ctx->scalars + 10, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *rows = s->n_rows;
    *digest = ctx->h_scalars[10];
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_reads(pedk_sample* s, int group, uint64_t* words, uint8_t* pal, uint64_t cap_reads,
                                 uint64_t* n_reads, int32_t* read_len, int32_t* words_per_read) {
    if (!s || !n_reads) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    if (group < 0 || group >= (int)s->groups.size()) throw PedkError(PEDK_E_INVAL, "no such read group");
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const ReadGroup& g = s->groups[(size_t)group];
    *n_reads = g.n;
    if (read_len) *read_len = g.L;
    if (words_per_read) *words_per_read = g.W;
    if (words && pal) {
        if (cap_reads < g.n) throw PedkError(PEDK_E_INVAL, "pedk_sample_reads: buffer too small");
        PEDK_CUDA_TRY(cudaMemcpyAsync(words, g.words.p, g.n * g.W * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
        PEDK_CUDA_TRY(cudaMemcpyAsync(pal, g.pal.p, g.n, cudaMemcpyDeviceToHost, s->ctx->stream));
        PEDK_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
        s->ctx->stats.d2h_bytes += g.n * (g.W * 8 + 1);
    }
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_release_reads(pedk_sample* s) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
    for (ReadGroup& g : s->groups) {
        g.words.reset();
        g.pal.reset();
    }
    s->reads_released = true;
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}
