// f1: the unique-20-mer index of the reference and the edge -> position map.
//
// mk20mer (ped.pl:1162-1211) writes, for every 20-base window of every chromosome that holds no N,
// the line `FORWARD chr pos f` and the line `REVCOMP chr pos+19 r`; mkUniq (ped.pl:1138-1160)
// sorts them per 3-nt bucket and keeps the 20-mers that occur on exactly ONE line -> ref20_uniq.TAG.gz.
// mapKmerSub (ped.pl:641-685) then looks every edge whose control allele is a single base up in that
// index: the 20-mer (19-mer + control base) gives the edge a chromosome position.
//
// Here: a 20-mer and its reverse complement occur on equally many lines, so "exactly one line" is
// "the canonical form has exactly one window, and the 20-mer is not its own reverse complement".
//   k_ref_kmers     one thread per base of the cleaned reference: canonical k-mer of the window that
//                   starts there (~0 if it holds an N, leaves its chromosome, or the chromosome is
//                   skipped), value = offset << 1 | (canonical form is the reverse strand)
//   radix sort      (key, value) pairs by key (radix.cu): equal k-mers become neighbours
//   k_ref_unique    keys unlike both neighbours (and not palindromic) -> the two strand rows
//                   (forward k-mer, offset << 1), (reverse complement, offset << 1 | 1)
//   radix sort      the strand rows by k-mer: the order of the ref20_uniq files
//   k_lookup_sorted edges: binary search of (19-mer << 2 | control base) in that table
#include <string.h>

#include <algorithm>
#include <string>

#include "comm.h"
#include "hostpar.h"
#include "sample.h"

using namespace pedk;

namespace {

__global__ void __launch_bounds__(256) k_ref_kmers(const uint8_t* __restrict__ clean, uint64_t n,
                                                  const uint64_t* __restrict__ chr_start,
                                                  const uint64_t* __restrict__ chr_end, int n_chr, int k,
                                                  uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    // chromosome of g: the last one that starts at or before g (chr_end == chr_start marks a skipped one)
    int lo = 0, hi = n_chr;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chr_start[mid] <= g) lo = mid; else hi = mid;
    }
    uint64_t K = 0;
    bool ok = g >= chr_start[lo] && g + (uint64_t)k <= chr_end[lo];
    if (ok) {
        for (int j = 0; j < k; ++j) {
            const uint32_t c = clean[g + j];
            ok = ok && c != 'N';
            K = (K << 2) | base_code(c);
        }
    }
    if (!ok) {
        key[g] = ~0ull;
        val[g] = 0;
        return;
    }
    const uint64_t rc = kmer_revcomp(K, k);
    const bool rev = rc < K;
    key[g] = rev ? rc : K;
    val[g] = (uint32_t)(g << 1) | (rev ? 1u : 0u);
}

__global__ void __launch_bounds__(256) k_ref_unique_flags(const uint64_t* __restrict__ key, uint64_t n, int k,
                                                         uint8_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t K = key[i];
    bool u = K != ~0ull;
    if (u && i && key[i - 1] == K) u = false;
    if (u && i + 1 < n && key[i + 1] == K) u = false;
    if (u && kmer_revcomp(K, k) == K) u = false;  // both of its lines carry the same 20-mer
    flag[i] = u ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_ref_strand_rows(const uint64_t* __restrict__ key,
                                                        const uint32_t* __restrict__ val,
                                                        const uint8_t* __restrict__ flag,
                                                        const uint64_t* __restrict__ pos, uint64_t n, int k,
                                                        uint64_t* __restrict__ out_key, uint32_t* __restrict__ out_val) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const uint64_t K = key[i], rc = kmer_revcomp(K, k);
    const uint32_t v = val[i];
    const uint64_t fwd = (v & 1u) ? rc : K;  // the k-mer as the chromosome spells it
    const uint64_t o = 2 * pos[i];
    out_key[o] = fwd;
    out_val[o] = v & ~1u;
    out_key[o + 1] = (v & 1u) ? K : rc;
    out_val[o + 1] = v | 1u;
}

__global__ void __launch_bounds__(128) k_lookup_sorted(const uint64_t* __restrict__ key, uint64_t n,
                                                      const uint64_t* __restrict__ q, int nq,
                                                      long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint64_t want = q[i];
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (key[mid] < want) lo = mid + 1; else hi = mid;
    }
    out[i] = (lo < n && key[lo] == want) ? (long long)lo : -1ll;
}

// f3, cnvJoinSub (ped.pl:949-989): every unique reference k-mer looked up in the two sorted count tables
__global__ void __launch_bounds__(256) k_cnv_lookup(const uint64_t* __restrict__ ukey, uint64_t n,
                                                   const uint64_t* __restrict__ tkey, const uint32_t* __restrict__ tval,
                                                   uint64_t nt, uint32_t* __restrict__ out_t,
                                                   uint32_t* __restrict__ out_c) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t want = ukey[i];
    uint64_t lo = 0, hi = nt;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (tkey[mid] < want) lo = mid + 1; else hi = mid;
    }
    // the merged table of build_rows holds both samples: bit 31 of the value marks the target
    uint32_t t = 0, c = 0;
    for (; lo < nt && tkey[lo] == want; ++lo) {
        const uint32_t v = tval[lo];
        if (v >> 31) t = v & 0x7fffffffu; else c = v;
    }
    out_t[i] = t;
    out_c[i] = c;
}

void append_u(std::string& s, uint64_t v) {
    char b[24];
    int n = 0;
    do {
        b[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) s.push_back(b[--n]);
}

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// Perl `split` without arguments: fields separated by runs of whitespace, leading whitespace skipped
std::vector<std::string> awk_split(const char* p, const char* end) {
    std::vector<std::string> f;
    while (p < end) {
        while (p < end && is_space(*p)) ++p;
        if (p >= end) break;
        const char* b = p;
        while (p < end && !is_space(*p)) ++p;
        f.emplace_back(b, p);
    }
    return f;
}

// ped.pl:3464-3484 on a short allele string: reverse, A<->T, C<->G, anything else -> N
std::string complement_str(const std::string& s) {
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    return r;
}

void build_ref20(pedk_sample* s, int k) {
    pedk_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (k < 4 || k > 32) throw PedkError(PEDK_E_INVAL, "k must be in [4, 32]");
    if (s->chromosomes.empty() && !s->ref_clean_len) {
        // an empty reference: an empty index
        s->ref20 = pedk_sample::Ref20{};
        s->ref20.k = k;
        s->ref20.built = true;
        return;
    }
    if (!s->ref_clean.p) throw PedkError(PEDK_E_INVAL, "pedk_sample_ref20_uniq: the sample is not an ingested reference");
    const uint64_t n = s->ref_clean_len;
    if (n >= (1ull << 31)) throw PedkError(PEDK_E_UNSUPPORTED, "reference longer than 2^31 bases (32-bit offsets in the index)");
    pedk_sample::Ref20 idx;
    idx.k = k;
    const size_t H = s->chromosomes.size();
    // every chromosome file is one sequence (a name that came back replaced the earlier file);
    // mk20 skips a chromosome called NOP (ped.pl:1217)
    std::vector<uint64_t> cs(H), ce(H);
    for (size_t c = 0; c < H; ++c) {
        const pedk_sample::Chromosome& ch = s->chromosomes[c];
        cs[c] = ch.start;
        ce[c] = (ch.live && ch.name != "chrNOP") ? ch.start + ch.length : ch.start;
    }
    uint64_t n_rows = 0;
    DBuf<uint64_t> rk, rk2;
    DBuf<uint32_t> rv, rv2;
    int cur2 = 0;
    if (n && H) {
        DBuf<uint64_t> d_chr(ctx, 2 * H);
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_chr.p, cs.data(), H * 8, cudaMemcpyHostToDevice, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_chr.p + H, ce.data(), H * 8, cudaMemcpyHostToDevice, st));
        const int n_pass = (2 * k + 7) / 8;
        DBuf<uint64_t> ka(ctx, n), kb(ctx, n);
        DBuf<uint32_t> va(ctx, n), vb(ctx, n);
        DBuf<unsigned long long> hist(ctx, 8 * 256);
        {
            StageScope sc(ctx, PEDK_ST_EXTRACT, 2);
            PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
            k_ref_kmers<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->ref_clean.p, n, d_chr.p, d_chr.p + H, (int)H, k, ka.p,
                                                                     va.p);
            PEDK_CUDA_TRY(cudaGetLastError());
            launch_digit_hist(ka.p, n, hist.p, n_pass, st);
        }
        const int cur = radix_sort(ctx, ka.p, kb.p, va.p, vb.p, n, n_pass, hist.p, PEDK_ST_ROWSORT);
        check_err_flag(ctx, "radix sort (reference k-mers)");
        const uint64_t* sk = cur ? kb.p : ka.p;
        const uint32_t* sv = cur ? vb.p : va.p;
        DBuf<uint8_t> flag(ctx, n);
        DBuf<uint64_t> pos(ctx, n), ws(ctx, scan_workspace_words(n));
        {
            StageScope sc(ctx, PEDK_ST_COUNT, 4);
            k_ref_unique_flags<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sk, n, k, flag.p);
            PEDK_CUDA_TRY(cudaGetLastError());
            exclusive_scan_u8(flag.p, pos.p, n, ctx->scalars, ws.p, st);
        }
        PEDK_CUDA_TRY(cudaMemcpyAsync(ctx->h_scalars, ctx->scalars, 8, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        const uint64_t U = ctx->h_scalars[0];
        n_rows = 2 * U;
        if (n_rows) {
            rk = DBuf<uint64_t>(ctx, n_rows);
            rv = DBuf<uint32_t>(ctx, n_rows);
            rk2 = DBuf<uint64_t>(ctx, n_rows);
            rv2 = DBuf<uint32_t>(ctx, n_rows);
            {
                StageScope sc(ctx, PEDK_ST_COUNT, 2);
                k_ref_strand_rows<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sk, sv, flag.p, pos.p, n, k, rk.p, rv.p);
                PEDK_CUDA_TRY(cudaGetLastError());
                PEDK_CUDA_TRY(cudaMemsetAsync(hist.p, 0, hist.bytes(), st));
                launch_digit_hist(rk.p, n_rows, hist.p, n_pass, st);
            }
            cur2 = radix_sort(ctx, rk.p, rk2.p, rv.p, rv2.p, n_rows, n_pass, hist.p, PEDK_ST_ROWSORT);
            check_err_flag(ctx, "radix sort (unique reference k-mers)");
        }
    }
    idx.kmer.resize(n_rows);
    idx.chr.resize(n_rows);
    idx.pos.resize(n_rows);
    idx.strand.resize(n_rows);
    if (n_rows) {
        std::vector<uint32_t> v(n_rows);
        PEDK_CUDA_TRY(cudaMemcpyAsync(idx.kmer.data(), cur2 ? rk2.p : rk.p, n_rows * 8, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(v.data(), cur2 ? rv2.p : rv.p, n_rows * 4, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += n_rows * 12;
        idx.d_kmer = std::move(cur2 ? rk2 : rk);
        for (uint64_t i = 0; i < n_rows; ++i) {
            const uint64_t g = v[i] >> 1;
            size_t c = (size_t)(std::upper_bound(cs.begin(), cs.end(), g) - cs.begin()) - 1;
            while (ce[c] <= g) --c;  // skipped chromosomes have no bases of their own
            const uint64_t off = g - cs[c];
            idx.chr[i] = (uint32_t)c;
            idx.strand[i] = (uint8_t)(v[i] & 1u);
            // $fpos = $i + 1; $rpos = $fpos + 20 - 1 (ped.pl:1191-1192)
            idx.pos[i] = (v[i] & 1u) ? off + (uint64_t)k : off + 1;
        }
    }
    idx.names.resize(H);
    for (size_t c = 0; c < H; ++c) idx.names[c] = s->chromosomes[c].name.substr(3);  // mk20mer prints $chr
    const int shift = 2 * k - 6;
    size_t i = 0;
    for (int tag = 0; tag < 64; ++tag) {
        while (i < idx.kmer.size() && (int)(idx.kmer[i] >> shift) < tag) ++i;
        idx.tag_start[tag] = i;
    }
    idx.tag_start[64] = idx.kmer.size();
    idx.any_chromosome = false;
    for (size_t c = 0; c < H; ++c) idx.any_chromosome |= s->chromosomes[c].live && s->chromosomes[c].name != "chrNOP";
    idx.built = true;
    s->ref20 = std::move(idx);
}

}  // namespace

namespace pedk {

// index read back from ref20_uniq.TAG.gz text (tag by tag, in order): `KMER chr pos f|r`
void ref20_append_text(pedk_sample* s, int tag, const std::string& text) {
    pedk_sample::Ref20& idx = s->ref20;
    idx.tag_start[tag] = idx.kmer.size();
    size_t p = 0;
    while (p < text.size()) {
        size_t nl = text.find('\n', p);
        if (nl == std::string::npos) nl = text.size();
        std::vector<std::string> f = awk_split(text.data() + p, text.data() + nl);
        p = nl + 1;
        if (f.size() < 4) continue;
        if (!idx.k) idx.k = (int)f[0].size();
        if ((int)f[0].size() != idx.k || idx.k > 32) throw PedkError(PEDK_E_IO, "malformed ref20_uniq line");
        uint64_t v = 0;
        for (char c : f[0]) {
            if (!is_acgt((unsigned char)c)) throw PedkError(PEDK_E_IO, "unexpected character in a ref20_uniq k-mer");
            v = (v << 2) | base_code((unsigned char)c);
        }
        uint32_t ci = 0;
        for (; ci < idx.names.size(); ++ci)
            if (idx.names[ci] == f[1]) break;
        if (ci == idx.names.size()) idx.names.push_back(f[1]);
        idx.kmer.push_back(v);
        idx.chr.push_back(ci);
        idx.pos.push_back(strtoull(f[2].c_str(), nullptr, 10));
        idx.strand.push_back(f[3] == "f" ? 0 : 1);
    }
    idx.tag_start[tag + 1] = idx.kmer.size();
}

std::string ref20_text(pedk_sample* s, int tag) {
    const pedk_sample::Ref20& idx = s->ref20;
    std::string t;
    t.reserve((idx.tag_start[tag + 1] - idx.tag_start[tag]) * (size_t)(idx.k + 16));
    for (uint64_t i = idx.tag_start[tag]; i < idx.tag_start[tag + 1]; ++i) {
        t += kmer_string(idx.kmer[i], idx.k);
        t += '\t';
        t += idx.names[idx.chr[i]];
        t += '\t';
        append_u(t, idx.pos[i]);
        t += idx.strand[i] ? "\tr\n" : "\tf\n";
    }
    return t;
}

// mapKmerSub (ped.pl:641-685) over the lines of <tmpdir>/<t>.snp.TAG (any tags, in key order):
// the lines whose control allele is one base are looked up as 19-mer + that base.
std::string map_kmer_text(pedk_sample* ref, const char* snp, uint64_t snp_len) {
    pedk_ctx* ctx = ref->ctx;
    cudaStream_t st = ctx->stream;
    pedk_sample::Ref20& idx = ref->ref20;
    if (!idx.built) throw PedkError(PEDK_E_INVAL, "pedk_sample_map_kmer_text: no unique k-mer index (pedk_sample_ref20_uniq first)");
    const int k = idx.k;
    struct Row {
        std::vector<std::string> f;  // seq, cont, alt, counts...
        uint64_t key;
    };
    std::vector<Row> rows;
    {
        std::vector<std::pair<const char*, const char*>> lines;
        const char* p = snp;
        const char* end = snp + snp_len;
        while (p < end) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            if (!nl) nl = end;
            lines.emplace_back(p, nl);
            p = nl + 1;
        }
        std::vector<Row> parsed(lines.size());
        std::vector<uint8_t> keep(lines.size(), 0);
        parallel_for(lines.size(), [&](size_t lo, size_t hi) {
            for (size_t li = lo; li < hi; ++li) {
                Row& r = parsed[li];
                r.f = awk_split(lines[li].first, lines[li].second);
                // $seq = shift(@row); if (length($row[0]) == 1) { $dat{$seq} = ... }
                if (r.f.size() < 2 || r.f[1].size() != 1 || (int)r.f[0].size() != k - 1) continue;
                const char cb = r.f[1][0];
                if (!is_acgt((unsigned char)cb)) continue;
                uint64_t v = 0;
                bool ok = true;
                for (char c : r.f[0]) {
                    ok = ok && is_acgt((unsigned char)c);
                    v = (v << 2) | base_code((unsigned char)c);
                }
                if (!ok) continue;
                r.key = (v << 2) | base_code((unsigned char)cb);
                r.f.resize(11);  // undefined fields print as nothing
                keep[li] = 1;
            }
        });
        for (size_t li = 0; li < parsed.size(); ++li)
            if (keep[li]) rows.push_back(std::move(parsed[li]));
    }
    std::string out;
    if (rows.empty() || idx.kmer.empty()) return out;
    // the index on the device (kept after the build; uploaded once when it came from files)
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!idx.d_kmer.p) {
        idx.d_kmer = DBuf<uint64_t>(ctx, idx.kmer.size());
        PEDK_CUDA_TRY(cudaMemcpyAsync(idx.d_kmer.p, idx.kmer.data(), idx.kmer.size() * 8, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += idx.kmer.size() * 8;
    }
    const int nq = (int)rows.size();
    std::vector<uint64_t> q((size_t)nq);
    for (int i = 0; i < nq; ++i) q[(size_t)i] = rows[(size_t)i].key;
    std::vector<long long> hit((size_t)nq);
    DBuf<uint64_t> d_q(ctx, (size_t)nq);
    DBuf<long long> d_hit(ctx, (size_t)nq);
    {
        StageScope sc(ctx, PEDK_ST_EDGES, 1);
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_q.p, q.data(), (size_t)nq * 8, cudaMemcpyHostToDevice, st));
        k_lookup_sorted<<<(nq + 127) / 128, 128, 0, st>>>(idx.d_kmer.p, idx.kmer.size(), d_q.p, nq, d_hit.p);
        PEDK_CUDA_TRY(cudaGetLastError());
        PEDK_CUDA_TRY(cudaMemcpyAsync(hit.data(), d_hit.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, st));
    }
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    // ref20_uniq is read in file order = k-mer order; a later edge line with the same 19-mer
    // would overwrite the earlier hash entry (edge lists hold every 19-mer once)
    std::vector<int> order((size_t)nq);
    for (int i = 0; i < nq; ++i) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rows[(size_t)a].key < rows[(size_t)b].key; });
    constexpr size_t PARTS = 16;
    std::vector<std::string> part(PARTS);
    parallel_for(PARTS, 1, [&](size_t plo, size_t phi) {
    for (size_t pi = plo; pi < phi; ++pi) {
    std::string& out = part[pi];
    for (size_t oi = order.size() * pi / PARTS; oi < order.size() * (pi + 1) / PARTS; ++oi) {
        const int i = order[oi];
        if (oi + 1 < order.size() && (rows[(size_t)order[oi + 1]].key >> 2) == (rows[(size_t)i].key >> 2)) continue;
        const long long h = hit[(size_t)i];
        if (h < 0) continue;
        const std::vector<std::string>& f = rows[(size_t)i].f;
        const std::string& chr = idx.names[idx.chr[(size_t)h]];
        const uint64_t pos = idx.pos[(size_t)h];
        if (!idx.strand[(size_t)h]) {
            // "$row[1]\t$pos\t$seq\t$nuc\t$dat[0]\t$dat[1]\tf\t$dat[2..9]"   ($pos = $row[2] + 19)
            out += chr;
            out += '\t';
            append_u(out, pos + (uint64_t)(k - 1));
            out += '\t' + f[0] + '\t' + f[1] + '\t' + f[1] + '\t' + f[2] + "\tf";
            for (int j = 3; j < 11; ++j) out += '\t' + f[(size_t)j];
            out += '\n';
        } else {
            // strand r: position - 19, everything complemented, the counts in T,G,C,A order
            out += chr;
            out += '\t';
            const long long rp = (long long)pos - (k - 1);
            if (rp < 0) {
                out += '-';
                append_u(out, (uint64_t)(-rp));
            } else {
                append_u(out, (uint64_t)rp);
            }
            const std::string c0 = complement_str(f[1]);
            out += '\t' + complement_str(f[0]) + '\t' + c0 + '\t' + c0 + '\t' + complement_str(f[2]) + "\tr";
            static const int perm[8] = {6, 5, 4, 3, 10, 9, 8, 7};
            for (int j = 0; j < 8; ++j) out += '\t' + f[(size_t)perm[j]];
            out += '\n';
        }
    }
    }
    });
    for (const std::string& x : part) out += x;
    return out;
}

// the number GNU `sort -n` reads at the start of a key: blanks, an optional '-', digits, an optional
// fraction; anything else ends it (no exponent, no hex), nothing at all counts as 0
static double sort_n_value(const std::string& s) {
    size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i;
    bool neg = false;
    if (i < s.size() && s[i] == '-') { neg = true; ++i; }
    double v = 0;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
    if (i < s.size() && s[i] == '.') {
        double f = 0.1;
        for (++i; i < s.size() && s[i] >= '0' && s[i] <= '9'; ++i, f *= 0.1) v += (s[i] - '0') * f;
    }
    return neg ? -v : v;
}

// cnvJoin + cnvJoinSub (ped.pl:934-989): the full outer join of the two count tables, restricted to
// the unique reference k-mers: `chr pos strand target_count control_count`, a count a sample does
// not have printed as 0, sorted like `sort -k 1 -n -k 2 -n` under LANG=C.
std::string cnv_text(pedk_sample* ref, pedk_sample* control, pedk_sample* target) {
    pedk_ctx* ctx = ref->ctx;
    cudaStream_t st = ctx->stream;
    pedk_sample::Ref20& idx = ref->ref20;
    if (!idx.built) throw PedkError(PEDK_E_INVAL, "pedk_cnv_text: no unique k-mer index (pedk_sample_ref20_uniq first)");
    if (!control->k || control->k != target->k || control->k != idx.k)
        throw PedkError(PEDK_E_INVAL, "pedk_cnv_text: the samples and the index must use the same k");
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t U = idx.kmer.size();
    std::string out;
    if (!U) return out;
    if (!idx.d_kmer.p) {
        idx.d_kmer = DBuf<uint64_t>(ctx, U);
        PEDK_CUDA_TRY(cudaMemcpyAsync(idx.d_kmer.p, idx.kmer.data(), U * 8, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += U * 8;
    }
    RowTable rows;  // both tables in one array sorted by k-mer
    build_rows(ctx, control, target, &rows);
    // On several ranks every rank holds the rows of a key range (build_rows routes them) and the whole
    // index: it looks all index k-mers up among its rows, finds each in at most one rank's range, and
    // the two count arrays are summed over the ranks (collective; every rank gets the whole text).
    std::vector<uint32_t> ht(U, 0), hc(U, 0);
    const bool multi = ctx->nranks > 1;
    if (rows.n || multi) {
        DBuf<uint32_t> d_t(ctx, U), d_c(ctx, U);
        if (rows.n) {
            StageScope sc(ctx, PEDK_ST_EDGES, 1);
            k_cnv_lookup<<<(unsigned)((U + 255) / 256), 256, 0, st>>>(idx.d_kmer.p, U, rows.key.p, rows.val.p, rows.n,
                                                                      d_t.p, d_c.p);
            PEDK_CUDA_TRY(cudaGetLastError());
        } else {
            PEDK_CUDA_TRY(cudaMemsetAsync(d_t.p, 0, U * 4, st));
            PEDK_CUDA_TRY(cudaMemsetAsync(d_c.p, 0, U * 4, st));
        }
        comm_allreduce_sum_u32(ctx, d_t.p, U);
        comm_allreduce_sum_u32(ctx, d_c.p, U);
        PEDK_CUDA_TRY(cudaMemcpyAsync(ht.data(), d_t.p, U * 4, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaMemcpyAsync(hc.data(), d_c.p, U * 4, cudaMemcpyDeviceToHost, st));
        PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->stats.d2h_bytes += U * 8;
    }
    // `sort -k 1 -n -k 2 -n`: the numeric value of the leading number of the chromosome name (a name
    // that is not a number counts as 0), then of the position, then the whole line bytewise
    struct Line {
        double chr_num;
        uint64_t pos;
        std::string text;
    };
    std::vector<double> name_num(idx.names.size(), 0.0);
    for (size_t c = 0; c < idx.names.size(); ++c) name_num[c] = sort_n_value(idx.names[c]);
    std::vector<Line> lines;
    for (uint64_t i = 0; i < U; ++i) {
        if (!ht[i] && !hc[i]) continue;  // in neither table: the outer join has no such line
        Line l;
        l.chr_num = name_num[idx.chr[i]];
        l.pos = idx.pos[i];
        l.text = idx.names[idx.chr[i]];
        l.text += '\t';
        append_u(l.text, idx.pos[i]);
        l.text += idx.strand[i] ? "\tr\t" : "\tf\t";
        append_u(l.text, ht[i]);
        l.text += '\t';
        append_u(l.text, hc[i]);
        l.text += '\n';
        lines.push_back(std::move(l));
    }
    std::sort(lines.begin(), lines.end(), [](const Line& a, const Line& b) {
        if (a.chr_num != b.chr_num) return a.chr_num < b.chr_num;
        if (a.pos != b.pos) return a.pos < b.pos;
        return a.text < b.text;
    });
    size_t bytes = 0;
    for (const Line& l : lines) bytes += l.text.size();
    out.reserve(bytes);
    for (const Line& l : lines) out += l.text;
    return out;
}

}  // namespace pedk

static char* dup_text(const std::string& s, uint64_t* len) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) throw std::bad_alloc();
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
}

extern "C" int pedk_sample_ref20_uniq(pedk_sample* s, int k, uint64_t* rows) {
    if (!s) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    build_ref20(s, k);
    if (rows) *rows = s->ref20.kmer.size();
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_ref20_uniq_text(pedk_sample* s, int tag, char** text, uint64_t* len) {
    if (!s || !text || !len || tag < 0 || tag >= 64) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    if (!s->ref20.built) throw PedkError(PEDK_E_INVAL, "pedk_sample_ref20_uniq_text: no index (pedk_sample_ref20_uniq first)");
    *text = dup_text(ref20_text(s, tag), len);
    return PEDK_OK;
    PEDK_API_END(s->ctx)
}

extern "C" int pedk_sample_map_kmer_text(pedk_sample* ref, const char* snp_text, uint64_t snp_len, char** text,
                                         uint64_t* len) {
    if (!ref || (!snp_text && snp_len) || !text || !len) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    *text = dup_text(map_kmer_text(ref, snp_text, snp_len), len);
    return PEDK_OK;
    PEDK_API_END(ref->ctx)
}

extern "C" int pedk_cnv_text(pedk_sample* ref, pedk_sample* control, pedk_sample* target, char** text, uint64_t* len) {
    if (!ref || !control || !target || !text || !len) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    *text = dup_text(cnv_text(ref, control, target), len);
    return PEDK_OK;
    PEDK_API_END(ref->ctx)
}
