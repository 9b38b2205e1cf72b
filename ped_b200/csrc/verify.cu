// f2: read-level verification and genotype of the mapped candidates, and toVcf.
//
// What ped.pl does between the position map and <target>.kmer.snp (ped.pl:380-387):
//   snpMkT   (2087-2229)  every candidate `chr pos ref alt` -> the L windows of the reference around it
//                         ("tw") and of the reference with the alternative allele put in ("tm"), L = the
//                         target's read length; up to ten earlier candidates of the chromosome that fall
//                         into the window are put in as well
//   sortSeq, joinTarget (2622-2641, 2601-2620)  the windows that ARE reads of the target (`join` with the
//                         sort_uniq files)
//   snpMkC, sortSeq, joinControl (2231-2305, 2576-2599)  the same against the control, with its read length
//   kmerReadCount (2388-2500)  windows found per candidate and kind -> cw cm tw tm, genotype M/H/R/N,
//                         `join` with the map lines -> <target>.kmer.snp
//
// Here the windows never become text files: the two strings of a candidate (2L-1 reference bases, and the
// mutant) go to HBM once, one warp per window packs it like a read, makes it canonical and looks it up in
// the hash table of the sample's de-duplicated reads (`k_probe_windows`, ingest.cu; the table is k_dedup's).
// sort_uniq holds both strands of every read, so "the window is a line of sort_uniq" is "its canonical form
// is one of the unique canonical reads".  Everything that is text in ped.pl (split on blanks, zero padding,
// the order of `sort` and of `sort keys`, `join` on the first field) is kept as text on the host: it is a few
// hundred bytes per candidate.  Bytewise (C locale) order throughout.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <exception>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "comm.h"
#include "hostpar.h"
#include "kernels.h"
#include "sample.h"

namespace pedk {

namespace {

typedef std::vector<std::string> Row;

// Perl's split(' ', $_): leading blanks ignored, fields separated by runs of blanks, no trailing empty field
Row split_blanks(const char* p, size_t n) {
    Row r;
    size_t i = 0;
    while (i < n) {
        while (i < n && (p[i] == ' ' || p[i] == '\t' || p[i] == '\n' || p[i] == '\r' || p[i] == '\f' || p[i] == '\v')) ++i;
        size_t a = i;
        while (i < n && !(p[i] == ' ' || p[i] == '\t' || p[i] == '\n' || p[i] == '\r' || p[i] == '\f' || p[i] == '\v')) ++i;
        if (i > a) r.emplace_back(p + a, i - a);
    }
    return r;
}
Row split_blanks(const std::string& s) { return split_blanks(s.data(), s.size()); }

const std::string kEmpty;
const std::string& field(const Row& r, size_t i) { return i < r.size() ? r[i] : kEmpty; }  // undef prints as ""

bool all_digits(const std::string& s) {
    if (s.empty()) return false;
    for (char c : s)
        if (c < '0' || c > '9') return false;
    return true;
}

// numeric value of a string the way `$x += 0` / `$x < $y` see it (the fields here are integers or not numbers)
long long perl_int(const std::string& s) {
    size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n')) ++i;
    bool neg = false;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    long long v = 0;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
    return neg ? -v : v;
}

std::string pad_chr(const std::string& c) {  // ped.pl:2106-2109
    if (!all_digits(c)) return c;
    std::string t = "000" + c;
    return t.substr(t.size() - 3);
}

std::string pad_pos(const std::string& p) {  // ped.pl:2110-2111
    std::string t = "000000000000" + p;
    return t.substr(t.size() - 11);
}

std::string strip_zeros(const std::string& c) {  // s/^0+//
    size_t i = 0;
    while (i < c.size() && c[i] == '0') ++i;
    return c.substr(i);
}

std::vector<std::string> lines_of(const char* p, uint64_t n) {
    std::vector<std::string> v;
    uint64_t a = 0;
    while (a < n) {
        const char* e = (const char*)memchr(p + a, '\n', n - a);
        uint64_t b = e ? (uint64_t)(e - p) : n;
        v.emplace_back(p + a, b - a);
        a = b + 1;
    }
    return v;
}

struct Candidate {
    std::string chr, ref, alt;  // chr without leading zeros
    long long pos = 0;
    std::string ialt;           // what `($tpos, $iref, $ialt) = split` of "$pos $ref $alt" gives (ped.pl:2185)
};

// snpMkT up to snp.list (ped.pl:2094-2152)
std::vector<Candidate> candidates_of(const std::vector<std::string>& map_lines) {
    std::vector<std::string> tmp(map_lines.size());
    parallel_for(map_lines.size(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            Row row = split_blanks(map_lines[i]);
            const std::string ref = field(row, 4);
            std::string alt = field(row, 5);
            if (!ref.empty()) {
                size_t at = alt.find(ref);  // $alt =~ s/$ref//: ref is a nucleotide here (mapKmerSub prints it only if it is one)
                if (at != std::string::npos) alt.erase(at, ref.size());
            }
            tmp[i] = pad_chr(field(row, 0)) + " " + pad_pos(field(row, 1)) + " " + ref + " " + alt;
        }
    });
    parallel_sort(tmp, std::less<std::string>());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    // every line re-split and re-joined (snp.tmp -> snp.list, ped.pl:2145-2150), then `uniq -c`
    std::vector<std::string> relined(tmp.size());
    std::vector<Candidate> parsed(tmp.size());
    parallel_for(tmp.size(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            Row row = split_blanks(tmp[i]);
            relined[i] = field(row, 0) + " " + field(row, 1) + " " + field(row, 2) + " " + field(row, 3);
            Row r2 = split_blanks(relined[i]);  // ($count, $chr, $pos, $ref, $alt) = split, the count in front
            Candidate& c = parsed[i];
            c.chr = strip_zeros(field(r2, 0));
            c.pos = perl_int(field(r2, 1));
            c.ref = field(r2, 2);
            c.alt = field(r2, 3);
            // "$pos $ref $alt" split on blanks again: an empty field lets the next one move up
            const std::string* third[2] = {&c.ref, &c.alt};
            int seen = 0;
            for (const std::string* f : third)
                if (!f->empty() && ++seen == 2) c.ialt = *f;
        }
    });
    std::vector<Candidate> out;
    out.reserve(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) {
        if (i && relined[i] == relined[i - 1]) continue;  // `uniq -c`
        out.push_back(std::move(parsed[i]));
    }
    return out;
}

// The two strings of every candidate for read length L (ped.pl:2153-2196 / 2246-2291): index 2c = the
// reference around the candidate, 2c+1 = the mutant; an empty pair where ped.pl says `next`.
struct Strings {
    std::vector<char> text;
    std::vector<uint64_t> start;
    std::vector<uint32_t> length;
};

Strings strings_of(const std::vector<Candidate>& cand, const ChromosomeLookup& chromosome, long long L) {
    const size_t n = cand.size();
    // which chromosome text and which earlier candidates (@dat: the last ten of this chromosome, the
    // candidate itself included, ped.pl:2169-2176) every candidate sees -- sequential, it is the file order
    std::vector<const std::string*> seq_of(n, nullptr);
    std::vector<size_t> dat_from(n, 0);
    {
        const std::string* seq = nullptr;
        size_t chr_first = 0;
        for (size_t i = 0; i < n; ++i) {
            if (i == 0 || cand[i - 1].chr != cand[i].chr) {
                seq = chromosome("chr" + cand[i].chr);  // open(IN, "$refdir/chr$chr"): no such file reads as nothing
                chr_first = i;
            }
            seq_of[i] = seq;
            dat_from[i] = i - chr_first >= 10 ? i - 9 : chr_first;  // $count = int(1 + 25/4) >= 5: always pushed
        }
    }
    std::vector<std::string> str(2 * n);
    parallel_for(n, [&](size_t lo, size_t hi) {
        for (size_t ci = lo; ci < hi; ++ci) {
            const Candidate& c = cand[ci];
            const std::string* seq = seq_of[ci];
            if (c.pos - L < 0 || !seq) continue;
            const unsigned long long at = (unsigned long long)(c.pos - L);
            if (at >= seq->size()) continue;
            std::string ref_seq = seq->substr(at, (size_t)(2 * L - 1));
            if ((long long)ref_seq.size() != 2 * L - 1) continue;
            std::string mut_seq = ref_seq.substr(0, (size_t)(L - 1)) + c.alt + ref_seq.substr((size_t)L);
            for (size_t d = dat_from[ci]; d <= ci; ++d) {
                const long long i = cand[d].pos - (c.pos - L) - 1;
                if (i > 0 && i < 2 * L && i != L - 1) {
                    const size_t ui = (size_t)i;
                    std::string head = mut_seq.substr(0, std::min(ui, mut_seq.size()));
                    std::string tail = ui + 1 <= mut_seq.size() ? mut_seq.substr(ui + 1) : std::string();
                    mut_seq = head + cand[d].ialt + tail;
                }
            }
            str[2 * ci] = std::move(ref_seq);
            str[2 * ci + 1] = std::move(mut_seq);
        }
    });
    Strings s;
    s.start.resize(2 * n);
    s.length.resize(2 * n);
    uint64_t total = 0;
    for (size_t i = 0; i < 2 * n; ++i) {
        s.start[i] = total;
        s.length[i] = (uint32_t)str[i].size();
        total += str[i].size();
    }
    s.text.resize(total);
    parallel_for(2 * n, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i)
            if (!str[i].empty()) memcpy(s.text.data() + s.start[i], str[i].data(), str[i].size());
    });
    return s;
}

// windows of every string that are reads of the sample -> hits per string
std::vector<uint32_t> probe(pedk_sample* smp, const Strings& s, int L) {
    std::vector<uint32_t> hits(s.start.size(), 0);
    if (s.start.empty() || L <= 0) return hits;
    const ReadGroup* g = nullptr;
    for (const ReadGroup& x : smp->groups)
        if (x.L == L && x.n) g = &x;
    pedk_ctx* ctx = smp->ctx;
    cudaStream_t st = ctx->stream;
    if (!g) {
        // no read of this length here: nothing joins.  On several ranks the reads of a sample are spread
        // over the ranks by their hash and every rank looks ALL windows up among its own: a window is a
        // read of the sample iff it is one on exactly one rank, so the counters add up (collective)
        if (ctx->nranks > 1) {
            DBuf<uint32_t> d_zero(ctx, hits.size());
            PEDK_CUDA_TRY(cudaMemsetAsync(d_zero.p, 0, d_zero.bytes(), st));
            comm_allreduce_sum_u32(ctx, d_zero.p, hits.size());
            PEDK_CUDA_TRY(cudaMemcpyAsync(hits.data(), d_zero.p, hits.size() * 4, cudaMemcpyDeviceToHost, st));
            PEDK_CUDA_TRY(cudaStreamSynchronize(st));
        }
        return hits;
    }
    if (g->n > (1ull << 30)) throw PedkError(PEDK_E_UNSUPPORTED, "more than 2^30 unique reads in one length group");
    uint64_t cap64 = 1024;
    while (cap64 < 2 * g->n) cap64 <<= 1;
    // PEDK_PROBE=windows: the first form of the look-up (one warp per window, packed from its bytes, over
    // k_dedup's table); default: one warp per string with rolling windows over its own table (ingest.cu)
    const char* form = getenv("PEDK_PROBE");
    const bool per_window = form && strcmp(form, "windows") == 0;
    DBuf<uint32_t> table(ctx, cap64);
    DBuf<uint8_t> uniq(ctx, per_window ? g->n : 1);
    DBuf<uint8_t> d_text(ctx, s.text.size() + 64);
    DBuf<uint64_t> d_start(ctx, s.start.size());
    DBuf<uint32_t> d_len(ctx, s.length.size());
    DBuf<uint32_t> d_hits(ctx, hits.size());
    StageScope sc(ctx, PEDK_ST_EDGES, 2);
    PEDK_CUDA_TRY(cudaMemsetAsync(table.p, 0xFF, table.bytes(), st));
    if (per_window) launch_dedup(g->words.p, g->W, (uint32_t)g->n, table.p, (uint32_t)cap64, nullptr, uniq.p, st);
    else launch_read_table(g->words.p, g->W, (uint32_t)g->n, table.p, (uint32_t)cap64, st);
    if (!s.text.empty())
        PEDK_CUDA_TRY(cudaMemcpyAsync(d_text.p, s.text.data(), s.text.size(), cudaMemcpyHostToDevice, st));
    PEDK_CUDA_TRY(cudaMemcpyAsync(d_start.p, s.start.data(), s.start.size() * 8, cudaMemcpyHostToDevice, st));
    PEDK_CUDA_TRY(cudaMemcpyAsync(d_len.p, s.length.data(), s.length.size() * 4, cudaMemcpyHostToDevice, st));
    PEDK_CUDA_TRY(cudaMemsetAsync(d_hits.p, 0, d_hits.bytes(), st));
    if (per_window)
        launch_probe_windows(d_text.p, d_start.p, d_len.p, s.start.size(), L, g->W, g->words.p, table.p,
                             (uint32_t)cap64, d_hits.p, st);
    else
        launch_probe_strings(d_text.p, d_start.p, d_len.p, s.start.size(), L, g->W, g->words.p, table.p,
                             (uint32_t)cap64, d_hits.p, st);
    comm_allreduce_sum_u32(ctx, d_hits.p, hits.size());  // several ranks: see above
    PEDK_CUDA_TRY(cudaMemcpyAsync(hits.data(), d_hits.p, hits.size() * 4, cudaMemcpyDeviceToHost, st));
    PEDK_CUDA_TRY(cudaStreamSynchronize(st));
    return hits;
}

char genotype(long long cw, long long cm, long long tw, long long tm) {  // ped.pl:2435-2445
    if (cw >= 5 && cm <= 1) {
        if (tm >= 5 && tw <= 1) return 'M';
        if ((tm + tw) > 0 && (double)tm / (double)(tm + tw) > 0.3 && (double)tm / (double)(tm + tw) < 0.7 && cm <= 1 &&
            tw >= 5)
            return 'H';
        return 'R';
    }
    return 'N';
}

// first blank-separated field of a line, and the non-empty fields behind it
std::string key_of(const std::string& line) {
    size_t i = 0, n = line.size();
    auto blank = [&](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; };
    while (i < n && blank(line[i])) ++i;
    size_t a = i;
    while (i < n && !blank(line[i])) ++i;
    return line.substr(a, i - a);
}

// kmerReadCount (ped.pl:2388-2500)
std::string read_count(const std::vector<std::string>& map_lines, const std::vector<Candidate>& cand,
                       const std::vector<uint32_t>& t_hits, const std::vector<uint32_t>& c_hits, long long length,
                       long long clength) {
    // ---- `$tmpdir/$target.map`: the map lines re-keyed `chr:pos`, sorted (ped.pl:2391-2406) ----
    std::vector<std::string> mp(map_lines.size());
    std::vector<uint8_t> keep(map_lines.size(), 0);
    parallel_for(map_lines.size(), [&](size_t lo, size_t hi) {
        for (size_t li = lo; li < hi; ++li) {
            Row row = split_blanks(map_lines[li]);
            const long long pos = perl_int(field(row, 1));
            if (pos < length || pos < clength) continue;
            std::string l = pad_chr(field(row, 0)) + ":" + pad_pos(field(row, 1));
            for (size_t i = 2; i <= 14; ++i) {
                l += '\t';
                l += field(row, i);
            }
            mp[li] = std::move(l);
            keep[li] = 1;
        }
    });
    {
        size_t o = 0;
        for (size_t i = 0; i < mp.size(); ++i)
            if (keep[i]) {
                if (o != i) mp[o] = std::move(mp[i]);
                ++o;
            }
        mp.resize(o);
    }
    parallel_sort(mp, std::less<std::string>());
    // ---- %count: one entry per `chr pos ref alt kind` that has at least one window among the reads; the key
    // is the line `join | cut` leaves, split on blanks and re-padded (ped.pl:2410-2421) ----
    struct Entry {
        std::string key;
        Row row;
        std::string group;  // the loop over `sort keys` compares this: the fields but the last, tab-joined
        long long n;
    };
    static const char* kinds[4] = {"tw", "tm", "cw", "cm"};
    std::vector<Entry> slot(cand.size() * 4);
    parallel_for(cand.size(), [&](size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; ++c)
            for (int kind = 0; kind < 4; ++kind) {
                Entry& e = slot[c * 4 + (size_t)kind];
                e.n = (kind < 2 ? t_hits : c_hits)[2 * c + (size_t)(kind & 1)];
                if (!e.n) continue;
                Row row;  // = split of "$chr $pos $ref $alt kind": empty fields vanish
                if (!cand[c].chr.empty()) row.push_back(cand[c].chr);
                row.push_back(std::to_string(cand[c].pos));
                if (!cand[c].ref.empty()) row.push_back(cand[c].ref);
                if (!cand[c].alt.empty()) row.push_back(cand[c].alt);
                row.push_back(kinds[kind]);
                while (row.size() < 2) row.emplace_back();
                row[0] = pad_chr(row[0]);
                row[1] = pad_pos(row[1]);
                e.key = row[0];
                for (size_t t = 1; t < row.size(); ++t) {
                    e.key += ' ';
                    e.key += row[t];
                }
                // ... and what the loop over `sort keys` sees when it splits the key again
                e.row = split_blanks(e.key);
                while (e.row.size() < 2) e.row.emplace_back();
                e.group = strip_zeros(e.row[0]);  // $row[0] =~ s/^0+//g; $row[1] += 0; join("\t", @row[0 .. $#row - 1])
                if (e.row.size() > 2) {
                    e.group += '\t';
                    e.group += std::to_string(perl_int(e.row[1]));
                }
                for (size_t t = 2; t + 1 < e.row.size(); ++t) {
                    e.group += '\t';
                    e.group += e.row[t];
                }
            }
    });
    std::vector<Entry> count;
    count.reserve(slot.size());
    for (Entry& e : slot)
        if (e.n) count.push_back(std::move(e));
    std::vector<Entry>().swap(slot);
    parallel_sort(count, [](const Entry& x, const Entry& y) { return x.key < y.key; });  // `sort keys`
    {
        size_t o = 0;  // equal keys are one hash entry
        for (size_t i = 0; i < count.size(); ++i) {
            if (o && count[o - 1].key == count[i].key) count[o - 1].n += count[i].n;
            else {
                if (o != i) count[o] = std::move(count[i]);
                ++o;
            }
        }
        count.resize(o);
    }
    // ---- `$tmpdir/$target.verify` (ped.pl:2423-2480) ----
    struct Group {
        const std::string* key;  // the group's fields, tab-joined (null: the empty $prev of the first flush)
        long long cw, cm, tw, tm;
    };
    std::vector<Group> groups;
    {
        long long cm = 0, cw = 0, tm = 0, tw = 0;
        const std::string* prev = nullptr;
        for (const Entry& e : count) {
            if (!prev || e.group != *prev) {
                if (prev && !prev->empty()) groups.push_back(Group{prev, cw, cm, tw, tm});
                cm = cw = tm = tw = 0;
            }
            const std::string& kind = e.row.back();
            if (kind == "cm") cm = e.n;
            if (kind == "cw") cw = e.n;
            if (kind == "tm") tm = e.n;
            if (kind == "tw") tw = e.n;
            prev = &e.group;
        }
        groups.push_back(Group{prev, cw, cm, tw, tm});  // ped.pl prints the last group without looking at it (2461-2480)
    }
    std::vector<std::string> verify(groups.size());
    parallel_for(groups.size(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const Group& g = groups[i];
            Row pr = g.key ? split_blanks(*g.key) : Row();
            while (pr.size() < 2) pr.emplace_back();
            verify[i] = pad_chr(pr[0]) + ":" + pad_pos(pr[1]) + "\t" + std::to_string(g.cw) + "\t" +
                        std::to_string(g.cm) + "\t" + std::to_string(g.tw) + "\t" + std::to_string(g.tm) + "\t" +
                        genotype(g.cw, g.cm, g.tw, g.tm);
        }
    });
    // ---- `join map verify` on the first blank-separated field: a merge, whatever the order of the inputs
    // (GNU join does not sort), every line of a key with every line of the other side ----
    std::vector<std::string> ka(mp.size()), kb(verify.size());
    parallel_for(mp.size(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) ka[i] = key_of(mp[i]);
    });
    parallel_for(verify.size(), [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) kb[i] = key_of(verify[i]);
    });
    std::vector<std::pair<uint32_t, uint32_t>> pairs;
    {
        size_t i = 0, j = 0;
        while (i < ka.size() && j < kb.size()) {
            if (ka[i] < kb[j]) ++i;
            else if (kb[j] < ka[i]) ++j;
            else {
                size_t i2 = i, j2 = j;
                while (i2 < ka.size() && ka[i2] == kb[j]) ++i2;
                while (j2 < kb.size() && kb[j2] == ka[i]) ++j2;
                for (size_t p = i; p < i2; ++p)
                    for (size_t q = j; q < j2; ++q) pairs.emplace_back((uint32_t)p, (uint32_t)q);
                i = i2;
                j = j2;
            }
        }
    }
    // ---- the joined line split on blanks again, `chr pos` from its key, fields 1..18 (ped.pl:2483-2497) ----
    std::vector<std::string> lines(pairs.size());
    parallel_for(pairs.size(), [&](size_t lo, size_t hi) {
        for (size_t x = lo; x < hi; ++x) {
            Row row = split_blanks(mp[pairs[x].first]);
            const Row rb = split_blanks(verify[pairs[x].second]);
            for (size_t t = 1; t < rb.size(); ++t) row.push_back(rb[t]);
            const std::string& key = field(row, 0);
            const size_t colon = key.find(':');
            const std::string chr = strip_zeros(key.substr(0, colon));
            std::string rest = colon == std::string::npos ? std::string() : key.substr(colon + 1);
            const size_t colon2 = rest.find(':');  // split(':', ...) takes the second piece
            if (colon2 != std::string::npos) rest.resize(colon2);
            std::string l = chr + "\t" + std::to_string(perl_int(rest));
            for (size_t i = 1; i <= 18; ++i) {
                l += '\t';
                l += field(row, i);
            }
            l += '\n';
            lines[x] = std::move(l);
        }
    });
    size_t bytes = 0;
    for (const std::string& l : lines) bytes += l.size();
    std::string out;
    out.reserve(bytes);
    for (const std::string& l : lines) out += l;
    return out;
}

long long first_read_length(pedk_sample* s) {  // getLength (ped.pl:3486-3502): the first line of sort_uniq/*.gz
    uint64_t live = 0;
    int L = 0;
    for (const ReadGroup& g : s->groups)
        if (g.n) {
            ++live;
            L = g.L;
        }
    if (live == 0) return 0;
    if (live == 1) return L;
    ensure_host_reads(s);
    for (int t = 0; t < 64; ++t)
        if (!s->h_reads[t].empty()) return (long long)s->h_reads[t][0].size();
    return 0;
}

// Perl prints a number with %.15g
std::string perl_number(double x) {
    char b[64];
    snprintf(b, sizeof b, "%.15g", x);
    return b;
}

// numeric value of a field for `+` (ped.pl:1560): the fields are counts, or a genotype letter in a shifted row
double perl_double(const std::string& s) {
    const char* p = s.c_str();
    while (*p == ' ' || *p == '\t' || *p == '\n') ++p;
    const char* q = p;
    if (*q == '+' || *q == '-') ++q;
    if (!((*q >= '0' && *q <= '9') || (*q == '.' && q[1] >= '0' && q[1] <= '9'))) return 0.0;  // not "inf"/"nan"/hex
    return strtod(p, nullptr);
}

}  // namespace

std::string verify_text(pedk_sample* target, pedk_sample* control, const ChromosomeLookup& chromosome, const char* map,
                        uint64_t map_len) {
    if (!target->ingested || !control->ingested || target->reads_released || control->reads_released)
        throw PedkError(PEDK_E_INVAL, "verification needs the reads of both samples (ingested, not released)");
    PEDK_CUDA_TRY(cudaSetDevice(target->ctx->device));
    long long length = first_read_length(target), clength = first_read_length(control);
    if (target->ctx->nranks > 1) {
        // a rank may hold no read of a (small) sample; the samples of a multi-rank job have one read length
        // (mixed lengths are refused at ingest), so the largest value over the ranks is everybody's
        pedk_ctx* cx = target->ctx;
        unsigned long long v[2] = {(unsigned long long)length, (unsigned long long)clength};
        PEDK_CUDA_TRY(cudaMemcpyAsync(cx->scalars + 12, v, sizeof v, cudaMemcpyHostToDevice, cx->stream));
        comm_allreduce_max(cx, cx->scalars + 12, 2);
        PEDK_CUDA_TRY(cudaMemcpyAsync(v, cx->scalars + 12, sizeof v, cudaMemcpyDeviceToHost, cx->stream));
        PEDK_CUDA_TRY(cudaStreamSynchronize(cx->stream));
        length = (long long)v[0];
        clength = (long long)v[1];
    }
    if (length <= 0 || clength <= 0)
        throw PedkError(PEDK_E_INVAL, "verification needs reads in both samples (ped.pl dies in read() with a negative length)");
    pedk_ctx* ctx = target->ctx;
    trace(ctx, "verify: start");
    const std::vector<std::string> map_lines = lines_of(map, map_len);
    const std::vector<Candidate> cand = candidates_of(map_lines);
    trace(ctx, "verify: candidates");
    std::vector<uint32_t> t_hits, c_hits;
    {
        const Strings s = strings_of(cand, chromosome, length);
        trace(ctx, "verify: target windows built");
        t_hits = probe(target, s, (int)length);
        trace(ctx, "verify: target windows looked up");
    }
    {
        const Strings s = strings_of(cand, chromosome, clength);
        trace(ctx, "verify: control windows built");
        c_hits = probe(control, s, (int)clength);
        trace(ctx, "verify: control windows looked up");
    }
    std::string out = read_count(map_lines, cand, t_hits, c_hits, length, clength);
    trace(ctx, "verify: counts, genotypes, join");
    return out;
}

// toVcf, method kmer (ped.pl:1532-1574)
std::string vcf_text(const char* snp, uint64_t snp_len, const std::string& target) {
    std::string out =
        "##fileformat=VCFv4.2\n"
        "##INFO=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
        "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Approximate read depth)\">\n"
        "##INFO=<ID=AF,Number=.,Type=Float,Description=\"Allele Frequency\">\n"
        "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
        "##FORMAT=<ID=AD,Number=.,Type=Integer,Description=\"Allelic depths for the reference and alternate alleles in "
        "the order listed\">\n"
        "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read depth\">\n"
        "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + target + "\n";
    // a line with M or H anywhere in it makes a record, any other line with depth repeats the record before
    // it, and a record equal to the one printed last is not printed again (ped.pl:1553-1573)
    const std::vector<std::string> lines = lines_of(snp, snp_len);
    std::vector<std::string> rec(lines.size());
    std::vector<uint8_t> kind(lines.size(), 0);  // 0: `next` (depth 0), 1: own record, 2: carries the previous one
    parallel_for(lines.size(), [&](size_t lo, size_t hi) {
        for (size_t li = lo; li < hi; ++li) {
            const std::string& line = lines[li];
            Row row;  // split('\t', $_): trailing empty fields are dropped
            size_t a = 0;
            for (;;) {
                size_t b = line.find('\t', a);
                row.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
                if (b == std::string::npos) break;
                a = b + 1;
            }
            while (!row.empty() && row.back().empty()) row.pop_back();
            if (row.empty()) row.emplace_back();
            if (row[0].size() <= 3 || all_digits(row[0])) row[0] = "chr" + row[0];
            const double r17 = perl_double(field(row, 17)), r18 = perl_double(field(row, 18));
            const double dp = r17 + r18;
            if (dp == 0) continue;
            const bool m = line.find('M') != std::string::npos, h = line.find('H') != std::string::npos;
            if (!(m || h)) {
                kind[li] = 2;
                continue;
            }
            const double af = (trunc(1000 * r18 / dp) + 0.0) / 1000;  // int() gives an integer: no negative zero
            const double p = pow(0.001 * dp, r18);
            const double qual = p == 0 ? 1000 : (trunc((-10 * log(p) / log(10.0)) * 1000) + 0.0) / 1000;
            rec[li] = row[0] + "\t" + field(row, 1) + "\t.\t" + field(row, 4) + "\t" + field(row, 5) + "\t" +
                      perl_number(qual) + "\t.\tGT=" + (m ? "1/1" : "0/1") + ";AF=" + perl_number(af) + ";DP=" +
                      field(row, 18) + "\tGT:AD:DP\t" + (m ? "1/1" : "0/1") + ":" + field(row, 17) + "," +
                      field(row, 18) + ":" + perl_number(dp) + "\n";
            kind[li] = 1;
        }
    });
    const std::string* output = nullptr;  // $output
    const std::string* prev = nullptr;    // $prev
    static const std::string none;
    for (size_t li = 0; li < lines.size(); ++li) {
        if (kind[li] == 0) continue;
        if (kind[li] == 1) output = &rec[li];
        const std::string& o = output ? *output : none;
        const std::string& pv = prev ? *prev : none;
        if (o != pv) out += o;
        prev = output;
    }
    return out;
}

}  // namespace pedk

using namespace pedk;

static char* dup_text(const std::string& s, uint64_t* len) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) throw std::bad_alloc();
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    *len = s.size();
    return p;
}

extern "C" int pedk_verify_text(pedk_sample* target, pedk_sample* control, pedk_sample* ref, const char* map_text,
                                uint64_t map_len, char** text, uint64_t* len) {
    if (!target || !control || !ref || (!map_text && map_len) || !text || !len) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    // the chromosome files of the reference sample (reference.cu), fetched from HBM when a candidate needs them
    std::map<std::string, std::string> cache;
    ChromosomeLookup lookup = [&](const std::string& name) -> const std::string* {
        auto it = cache.find(name);
        if (it != cache.end()) return &it->second;
        int found = -1;
        for (size_t i = 0; i < ref->chromosomes.size(); ++i)
            if (ref->chromosomes[i].live && ref->chromosomes[i].name == name) found = (int)i;
        if (found < 0) return nullptr;
        const uint64_t n = ref->chromosomes[(size_t)found].length;
        std::string seq((size_t)n, '\0');
        int rc = pedk_sample_chromosome(ref, found, nullptr, 0, nullptr, nullptr, n ? &seq[0] : nullptr, n);
        if (rc) throw PedkError(rc, ref->ctx->err);
        return &(cache[name] = std::move(seq));
    };
    *text = dup_text(verify_text(target, control, lookup, map_text, map_len), len);
    return PEDK_OK;
    PEDK_API_END(target->ctx)
}

// Host-only halves of the verification (no CUDA call), so that the text logic around the look-up kernel can
// be checked without a device: the strings whose windows are looked up, and the text made from the counters.
namespace {
ChromosomeLookup lookup_in(const std::map<std::string, std::string>& files) {
    return [&files](const std::string& name) -> const std::string* {
        auto it = files.find(name);
        return it == files.end() ? nullptr : &it->second;
    };
}
std::map<std::string, std::string> files_of(const char* const* names, const char* const* seqs, const uint64_t* lens,
                                            int n) {
    std::map<std::string, std::string> m;
    for (int i = 0; i < n; ++i) m[names[i]] = std::string(seqs[i], (size_t)lens[i]);
    return m;
}
}  // namespace

extern "C" int pedk_verify_strings_text(const char* map_text, uint64_t map_len, const char* const* chr_names,
                                        const char* const* chr_seqs, const uint64_t* chr_lens, int n_chr,
                                        int read_length, char** text, uint64_t* len) {
    if ((!map_text && map_len) || !text || !len || n_chr < 0 || read_length <= 0) return PEDK_E_INVAL;
    try {
        const std::map<std::string, std::string> files = files_of(chr_names, chr_seqs, chr_lens, n_chr);
        const std::vector<Candidate> cand = candidates_of(lines_of(map_text, map_len));
        const Strings s = strings_of(cand, lookup_in(files), read_length);
        std::string out;
        for (size_t c = 0; c < cand.size(); ++c) {
            out += cand[c].chr + " " + std::to_string(cand[c].pos) + " " + cand[c].ref + " " + cand[c].alt + "\t";
            out.append(s.text.data() + s.start[2 * c], s.length[2 * c]);
            out += '\t';
            out.append(s.text.data() + s.start[2 * c + 1], s.length[2 * c + 1]);
            out += '\n';
        }
        *text = dup_text(out, len);
        return PEDK_OK;
    } catch (...) {
        return PEDK_E_INTERNAL;
    }
}

extern "C" int pedk_verify_counts_text(const char* map_text, uint64_t map_len, const uint32_t* target_hits,
                                       const uint32_t* control_hits, uint64_t n_strings, int length, int clength,
                                       char** text, uint64_t* len) {
    if ((!map_text && map_len) || !text || !len || (n_strings && (!target_hits || !control_hits))) return PEDK_E_INVAL;
    try {
        const std::vector<std::string> map_lines = lines_of(map_text, map_len);
        const std::vector<Candidate> cand = candidates_of(map_lines);
        if (n_strings != 2 * cand.size()) return PEDK_E_INVAL;
        const std::vector<uint32_t> th(target_hits, target_hits + n_strings), ch(control_hits, control_hits + n_strings);
        *text = dup_text(read_count(map_lines, cand, th, ch, length, clength), len);
        return PEDK_OK;
    } catch (...) {
        return PEDK_E_INTERNAL;
    }
}

extern "C" int pedk_vcf_text(const char* kmer_snp, uint64_t snp_len, const char* target_name, char** text,
                             uint64_t* len) {
    if ((!kmer_snp && snp_len) || !target_name || !text || !len) return PEDK_E_INVAL;
    try {
        *text = dup_text(vcf_text(kmer_snp, snp_len, target_name), len);
        return PEDK_OK;
    } catch (...) {
        return PEDK_E_NOMEM;
    }
}
