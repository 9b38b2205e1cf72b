// a10: the reference itself as the control sample (`ref=` runs, BASELINE config 5).
//
// mkChrFromFile (ped.pl:1027-1060) turns a FASTA file into one newline-free,
// upper-case, non-ACGT -> 'N' sequence per chromosome; mkControlReadChr
// (ped.pl:1089-1130) tiles every chromosome into 100-base windows every 2 bases,
// drops the windows that hold an 'N' and writes each window and its reverse
// complement; mkControlReadSub (ped.pl:1132-1136) does `sort | uniq` per bucket.
// The result is a sort_uniq/ read set like any sample's, so from here on the
// reference is ingested like reads:
//
//   FASTA text in HBM -> line spans (newline count + scan + k_line_index, as for FASTQ)
//   k_fasta_lines   per line: header flag, number of sequence bytes
//   scan            offset of every line's bases in the cleaned text
//   k_fasta_clean   a warp per line: upper-case, non-ACGT -> N (a '\r' becomes N too:
//                   chomp only removes the '\n'), written at the line's offset
//   k_window_starts byte offset of every window of every chromosome
//   k_pack_fixed    the fixed-length pack of ingest.cu on those windows: N test, 2-bit
//                   pack, canonical strand, palindrome flag
//   dedup_group     exact de-duplication, as for reads
#include <string.h>

#include <algorithm>
#include <unordered_map>

#include "comm.h"
#include "sample.h"

using namespace pedk;

namespace {

__global__ void __launch_bounds__(256) k_fasta_lines(const uint8_t* __restrict__ text,
                                                    const uint64_t* __restrict__ line_start,
                                                    const uint64_t* __restrict__ line_end, uint64_t n_lines,
                                                    uint8_t* __restrict__ is_header, uint32_t* __restrict__ seq_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    const uint64_t s = line_start[i], e = line_end[i];
    const bool hdr = e > s && text[s] == '>';  // /^>/ (ped.pl:1043)
    is_header[i] = hdr ? 1 : 0;
    seq_len[i] = hdr ? 0u : (uint32_t)(e - s);
}

__global__ void __launch_bounds__(256) k_compact_flagged(const uint8_t* __restrict__ flag,
                                                        const uint64_t* __restrict__ pos, uint64_t n,
                                                        uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = i;
}

// y/a-z/A-Z/; y/ACGT/N/c  (ped.pl:1053-1054)
__device__ __forceinline__ uint8_t clean_base(uint8_t c) {
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (uint8_t)'N';
}

__global__ void __launch_bounds__(256) k_fasta_clean(const uint8_t* __restrict__ text,
                                                    const uint64_t* __restrict__ line_start,
                                                    const uint32_t* __restrict__ seq_len,
                                                    const uint64_t* __restrict__ seq_off, uint64_t first_line,
                                                    uint64_t n_lines, uint64_t base_off, uint8_t* __restrict__ out) {
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t line = first_line + warp; line < n_lines; line += n_warps) {
        const uint32_t len = seq_len[line];
        const uint8_t* src = text + line_start[line];
        uint8_t* dst = out + (seq_off[line] - base_off);
        for (uint32_t i = lane; i < len; i += 32) dst[i] = clean_base(src[i]);
    }
}

// window w of the whole reference -> byte offset in the cleaned text
__global__ void __launch_bounds__(256) k_window_starts(const uint64_t* __restrict__ chr_start,
                                                      const uint64_t* __restrict__ win_base, int n_chr, int step,
                                                      uint64_t n_win, uint64_t* __restrict__ seq_start) {
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    int lo = 0, hi = n_chr;  // win_base[lo] <= w < win_base[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (win_base[mid] <= w) lo = mid; else hi = mid;
    }
    seq_start[w] = chr_start[lo] + (w - win_base[lo]) * (uint64_t)step;
}

// multi-GPU: every rank tiles the whole reference and keeps the windows it owns
__global__ void __launch_bounds__(256) k_keep_owned(const uint64_t* __restrict__ words, int W, uint64_t n, int nranks,
                                                   int me, uint8_t* __restrict__ valid) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !valid[i]) return;
    if (read_owner_rank(words + i * (uint64_t)W, W, nranks) != me) valid[i] = 0;
}

size_t round_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// "chr" + the name ped.pl derives from a header line (ped.pl:1045-1050)
std::string chr_name_of(const std::string& header_line) {
    // (split)[0]: first whitespace-separated field of the line
    size_t b = 0;
    while (b < header_line.size() && (header_line[b] == ' ' || header_line[b] == '\t' || header_line[b] == '\r' ||
                                      header_line[b] == '\f' || header_line[b] == '\v'))
        ++b;
    size_t e = b;
    while (e < header_line.size() && !(header_line[e] == ' ' || header_line[e] == '\t' || header_line[e] == '\r' ||
                                       header_line[e] == '\f' || header_line[e] == '\v'))
        ++e;
    std::string out = header_line.substr(b, e - b);
    size_t gt = out.find('>');  // s/>//
    if (gt != std::string::npos) out.erase(gt, 1);
    if (out.size() >= 3 && (out[0] == 'c' || out[0] == 'C') && (out[1] == 'h' || out[1] == 'H') &&
        (out[2] == 'r' || out[2] == 'R'))
        out.erase(0, 3);  // s/^chr//i
    if (!out.empty() && std::all_of(out.begin(), out.end(), [](char c) { return c >= '0' && c <= '9'; })) {
        // $out += 0: a decimal number loses its leading zeros
        size_t nz = 0;
        while (nz + 1 < out.size() && out[nz] == '0') ++nz;
        out.erase(0, nz);
    }
    return "chr" + out;
}

void do_ingest_reference(pedk_sample* s, int window, int step) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (window == 0) window = 100;  // ped.pl:1108
    if (step == 0) step = 2;        // ped.pl:1119
    if (window < 3 || window > 32 * PEDK_MAX_WORDS || step < 1)
        throw PedkError(PEDK_E_INVAL, "reference tiling: window must be in [3, 512] and step >= 1");
    s->groups.clear();
    s->chromosomes.clear();
    s->reads_released = false;
    s->info = pedk_ingest_info{};
    s->info.fastq_bytes = s->fq_len;
    s->have_table = s->have_reads = false;
    s->ingested = true;
    const size_t n = s->fq_len;
    const bool multi = ctx->nranks > 1;
    if (s->copy_pending) {
        StageScope sc(ctx, PEDK_ST_COPY, 0);
        PEDK_CUDA_TRY(cudaStreamWaitEvent(st, s->ev_copied, 0));
        s->copy_pending = false;
    }
    if (n == 0) return;
    // ---- line spans, exactly as for one-read-per-line text ----
    const size_t padded = round_up_sz(n, INGEST_TILE_BYTES);
    if (!s->fq.p || s->fq.n < padded) throw PedkError(PEDK_E_INTERNAL, "text buffer smaller than its padding");
    PEDK_CUDA_TRY(cudaMemsetAsync(s->fq.p + n, 0, padded - n, st));
    const size_t tiles = padded / INGEST_TILE_BYTES;
    DBuf<uint32_t> tile_counts(ctx, tiles);
    DBuf<uint64_t> tile_prefix(ctx, tiles), ws(ctx, scan_workspace_words(tiles));
    uint8_t last_byte = '\n';
    {
        StageScope sc(ctx, PEDK_ST_INGEST, 4);
        launch_newline_count(s->fq.p, n, tile_counts.p, st);
        exclusive_scan_u32(tile_counts.p, tile_prefix.p, tiles, ctx->scalars, ws.p, st);
    }
    PEDK_CUDA_TRY(cudaMemcpyAsync(&last_byte, s->fq.p + n - 1, 1, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars, ctx->scalars, 8, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t TL = ctx->h_scalars[0] + (last_byte != '\n' ? 1 : 0);
    if (TL == 0) return;
    DBuf<uint64_t> line_start(ctx, TL), line_end(ctx, TL);
    DBuf<uint8_t> is_header(ctx, TL);
    DBuf<uint32_t> seq_len(ctx, TL);
    DBuf<uint64_t> seq_off(ctx, TL + 1), hdr_pos(ctx, TL), ws2(ctx, scan_workspace_words(TL));
    {
        StageScope sc(ctx, PEDK_ST_INGEST, 8);
        PEDK_CUDA_TRY(cudaMemsetAsync(line_start.p, 0, sizeof(uint64_t), st));
        launch_line_index(s->fq.p, n, tile_prefix.p, line_start.p, line_end.p, TL, 1, st);
        if (last_byte != '\n') {
            const uint64_t end = n;
            PEDK_CUDA_TRY(cudaMemcpyAsync(line_end.p + (TL - 1), &end, sizeof end, cudaMemcpyHostToDevice, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        }
        k_fasta_lines<<<(unsigned)((TL + 255) / 256), 256, 0, st>>>(s->fq.p, line_start.p, line_end.p, TL, is_header.p,
                                                                    seq_len.p);
        PEDK_CUDA_TRY(cudaGetLastError());
        exclusive_scan_u32(seq_len.p, seq_off.p, TL, ctx->scalars + 1, ws2.p, st);
        exclusive_scan_u8(is_header.p, hdr_pos.p, TL, ctx->scalars + 2, ws2.p, st);
    }
    PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars, ctx->scalars + 1, 16, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t total_seq = ctx->h_scalars[0];
    const uint64_t H = ctx->h_scalars[1];
    if (H == 0) {
        // no header at all: every line goes to an unopened handle (ped.pl:1055) -- nothing is kept
        s->fq.reset();
        s->fq_len = 0;
        return;
    }
    if (H > (1u << 24)) throw PedkError(PEDK_E_UNSUPPORTED, "more than 2^24 sequences in the reference FASTA");
    // ---- the header lines: where each chromosome starts, and its name ----
    DBuf<uint64_t> hdr_line(ctx, H);
    {
        StageScope sc(ctx, PEDK_ST_INGEST, 1);
        k_compact_flagged<<<(unsigned)((TL + 255) / 256), 256, 0, st>>>(is_header.p, hdr_pos.p, TL, hdr_line.p);
        PEDK_CUDA_TRY(cudaGetLastError());
        // seq_off[TL] = total, so that "the line after the last header" always has an offset
        PEDK_CUDA_TRY(cudaMemcpyAsync(seq_off.p + TL, ctx->scalars + 1, 8, cudaMemcpyDeviceToDevice, st));
    }
    std::vector<uint64_t> h_line(H), h_ls(H), h_le(H), h_off(H);
    PEDK_CUDA_TRY(cudaMemcpyAsync(h_line.data(), hdr_line.p, H * 8, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    for (uint64_t c = 0; c < H; ++c) {
        PEDK_CUDA_TRY(cudaMemcpyAsync(&h_ls[c], line_start.p + h_line[c], 8, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(&h_le[c], line_end.p + h_line[c], 8, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(&h_off[c], seq_off.p + h_line[c] + 1, 8, cudaMemcpyDeviceToHost, st));
    }
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t base_off = h_off[0];  // bases before the first header are dropped
    const uint64_t clean_len = total_seq - base_off;
    std::vector<uint64_t> chr_start(H), chr_len(H), win_base(H + 1, 0);
    for (uint64_t c = 0; c < H; ++c) {
        chr_start[c] = h_off[c] - base_off;
        const uint64_t next = c + 1 < H ? h_off[c + 1] - base_off : clean_len;
        chr_len[c] = next - chr_start[c];
        std::string line(h_le[c] - h_ls[c], '\0');
        if (!line.empty())
            PEDK_CUDA_TRY(cudaMemcpyAsync(&line[0], s->fq.p + h_ls[c], line.size(), cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        pedk_sample::Chromosome ch;
        ch.name = chr_name_of(line);
        ch.start = chr_start[c];
        ch.length = chr_len[c];
        s->chromosomes.push_back(ch);
    }
    // a later sequence of the same name overwrites the earlier file (open ">", ped.pl:1051):
    // only the last one of each name is tiled (the duplicates would vanish in `uniq` anyway
    // if they were equal, but they need not be)
    std::vector<char> live(H, 1);
    {
        std::unordered_map<std::string, uint64_t> last;
        for (uint64_t c = 0; c < H; ++c) {
            auto it = last.find(s->chromosomes[c].name);
            if (it != last.end()) live[it->second] = 0;
            last[s->chromosomes[c].name] = c;
        }
    }
    for (uint64_t c = 0; c < H; ++c) {
        // mkControlRead skips a chromosome called NOP (ped.pl:1069); its chrNOP file is still written
        const bool nop = s->chromosomes[c].name == "chrNOP";
        const uint64_t nw =
            (live[c] && !nop && chr_len[c] >= (uint64_t)window) ? (chr_len[c] - window) / step + 1 : 0;
        win_base[c + 1] = win_base[c] + nw;
        s->chromosomes[c].live = live[c] != 0;
    }
    const uint64_t R = win_base[H];
    // ---- cleaned text (kept: the chromosome files are part of what mkChrFromFile leaves) ----
    s->ref_clean = DBuf<uint8_t>(ctx, clean_len + 64);
    {
        StageScope sc(ctx, PEDK_ST_INGEST, 1);
        PEDK_CUDA_TRY(cudaMemsetAsync(s->ref_clean.p + clean_len, 'N', 64, st));
        const uint64_t lines = TL - (h_line[0] + 1);
        if (lines) {
            const uint64_t blocks = (lines + 7) / 8;
            const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 16 ? blocks : (uint64_t)kNumSMs * 16);
            k_fasta_clean<<<grid, 256, 0, st>>>(s->fq.p, line_start.p, seq_len.p, seq_off.p, h_line[0] + 1, TL, base_off,
                                                s->ref_clean.p);
            PEDK_CUDA_TRY(cudaGetLastError());
        }
    }
    s->ref_clean_len = clean_len;
    s->info.records = R;
    s->info.clipping_used = window;
    if (R == 0) {
        s->fq.reset();
        s->fq_len = 0;
        return;
    }
    // ---- windows -> packed canonical reads -> dedup ----
    const int L = window, W = (L + 31) / 32;
    DBuf<uint64_t> d_chr(ctx, 2 * H + 1);
    DBuf<uint64_t> seq_start(ctx, R);
    DBuf<uint64_t> words(ctx, R * W);
    DBuf<uint8_t> pal(ctx, R), valid(ctx, R);
    DBuf<unsigned long long> counters(ctx, 2);
    PEDK_CUDA_TRY(cudaMemcpyAsync(d_chr.p, chr_start.data(), H * 8, cudaMemcpyHostToDevice, st));
    PEDK_CUDA_TRY(cudaMemcpyAsync(d_chr.p + H, win_base.data(), (H + 1) * 8, cudaMemcpyHostToDevice, st));
    {
        StageScope sc(ctx, PEDK_ST_INGEST, 2);
        PEDK_CUDA_TRY(cudaMemsetAsync(counters.p, 0, counters.bytes(), st));
        k_window_starts<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(d_chr.p, d_chr.p + H, (int)H, step, R,
                                                                     seq_start.p);
        PEDK_CUDA_TRY(cudaGetLastError());
        launch_pack_fixed(s->ref_clean.p, seq_start.p, R, L, W, words.p, pal.p, valid.p, counters.p, st);
        if (multi) {
            k_keep_owned<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(words.p, W, R, ctx->nranks, ctx->rank, valid.p);
            PEDK_CUDA_TRY(cudaGetLastError());
        }
    }
    unsigned long long cnt[2];
    PEDK_CUDA_TRY(cudaMemcpyAsync(cnt, counters.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    trace(ctx, "reference: clean + tile + pack");
    s->info.reads_kept = cnt[0];       // windows without N
    s->info.strand_reads = 2 * cnt[0]; // each is written with its reverse complement (ped.pl:1111-1116)
    s->fq.reset();
    s->fq_len = 0;
    seq_start.reset();
    dedup_group(s, L, W, words, pal, R, valid.p);
    s->info.n_groups = (int32_t)s->groups.size();
    for (const ReadGroup& g : s->groups) {
        s->info.unique_reads += 2 * g.n - g.n_pal;
        s->info.palindromes += g.n_pal;
    }
}

}  // namespace

extern "C" int pedk_sample_ingest_reference(pedk_sample* s, int window, int step, pedk_ingest_info* info) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_ingest_reference(s, window, step);
    if (info) *info = s->info;
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

// host-only: the file name ped.pl gives the sequence of a FASTA header line (no CUDA call)
extern "C" int pedk_chromosome_file_name(const char* header_line, char* out, uint64_t cap) {
    if (!header_line || !out) return PEDK_E_INVAL;
    const std::string name = chr_name_of(header_line);
    if (cap < name.size() + 1) return PEDK_E_INVAL;
    memcpy(out, name.c_str(), name.size() + 1);
    return PEDK_OK;
}

extern "C" int pedk_sample_num_chromosomes(pedk_sample* s, int* n) {
    if (!s || !n) return PEDK_E_INVAL;
    *n = (int)s->chromosomes.size();
    return PEDK_OK;
}

extern "C" int pedk_sample_chromosome(pedk_sample* s, int i, char* name, uint64_t name_cap, uint64_t* length,
                                      int* overwritten, char* seq, uint64_t seq_cap) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    if (i < 0 || i >= (int)s->chromosomes.size()) throw PedkError(PEDK_E_INVAL, "no such chromosome");
    const pedk_sample::Chromosome& c = s->chromosomes[(size_t)i];
    if (name) {
        if (name_cap < c.name.size() + 1) throw PedkError(PEDK_E_INVAL, "pedk_sample_chromosome: name buffer too small");
        memcpy(name, c.name.c_str(), c.name.size() + 1);
    }
    if (length) *length = c.length;
    if (overwritten) *overwritten = c.live ? 0 : 1;
    if (seq) {
        if (seq_cap < c.length) throw PedkError(PEDK_E_INVAL, "pedk_sample_chromosome: sequence buffer too small");
        if (!s->ref_clean.p) throw PedkError(PEDK_E_INVAL, "the cleaned reference has been released");
        PEDK_CUDA_TRY(cudaSetDevice(s->ctx->device));
        if (c.length)
            PEDK_CUDA_TRY(cudaMemcpyAsync(seq, s->ref_clean.p + c.start, c.length, cudaMemcpyDeviceToHost, s->ctx->stream));
        PEDK_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
        s->ctx->stats.d2h_bytes += c.length;
    }
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}
