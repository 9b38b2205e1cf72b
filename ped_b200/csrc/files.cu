// File-level entry points: what the drop-in ped.pl calls.
//
// These keep ped.pl's file contract: the same
// directories, names, line formats and order, `zcat`-identical content, and the
// TTT file of a stage written last because ped.pl gates on it (ped.pl:291,
// 295, 369-372).  A sample that was ingested/counted earlier in the same
// process stays resident in HBM (ctx cache) so count -> lbc -> join do not go
// back through the files; a fresh process resumes from the files.
#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <map>
#include <string>
#include <thread>

#include "sample.h"

using namespace pedk;

namespace {

const char* NUC = "ACGT";

std::string tag_name(int t) {
    std::string s(3, 'A');
    s[0] = NUC[(t >> 4) & 3];
    s[1] = NUC[(t >> 2) & 3];
    s[2] = NUC[t & 3];
    return s;
}

bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

void mkdir_p(const std::string& p) {
    if (exists(p)) return;
    if (mkdir(p.c_str(), 0777) != 0 && errno != EEXIST)
        throw PedkError(PEDK_E_IO, "mkdir " + p + ": " + strerror(errno));
}

bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

int gz_level() {
    const char* e = getenv("PEDK_GZ_LEVEL");
    int l = e ? atoi(e) : 1;
    return l < 1 ? 1 : l > 9 ? 9 : l;
}

// write <path>.tmp then rename: a failed call never leaves a gate file behind
void write_gz_atomic(const std::string& path, const char* data, size_t len) {
    std::string tmp = path + ".tmp";
    char mode[8];
    snprintf(mode, sizeof mode, "wb%d", gz_level());
    gzFile f = gzopen(tmp.c_str(), mode);
    if (!f) throw PedkError(PEDK_E_IO, "cannot write " + tmp);
    gzbuffer(f, 1 << 20);
    size_t off = 0;
    while (off < len) {
        unsigned n = (unsigned)std::min<size_t>(len - off, 1u << 30);
        if (gzwrite(f, data + off, n) != (int)n) {
            gzclose(f);
            unlink(tmp.c_str());
            throw PedkError(PEDK_E_IO, "short write to " + tmp);
        }
        off += n;
    }
    if (gzclose(f) != Z_OK) {
        unlink(tmp.c_str());
        throw PedkError(PEDK_E_IO, "gzclose failed on " + tmp);
    }
    if (rename(tmp.c_str(), path.c_str()) != 0) throw PedkError(PEDK_E_IO, "rename " + tmp + ": " + strerror(errno));
}

void write_plain_atomic(const std::string& path, const char* data, size_t len) {
    std::string tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) throw PedkError(PEDK_E_IO, "cannot write " + tmp);
    if (len && fwrite(data, 1, len, f) != len) {
        fclose(f);
        unlink(tmp.c_str());
        throw PedkError(PEDK_E_IO, "short write to " + tmp);
    }
    fclose(f);
    if (rename(tmp.c_str(), path.c_str()) != 0) throw PedkError(PEDK_E_IO, "rename " + tmp + ": " + strerror(errno));
}

std::string read_gz(const std::string& path) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw PedkError(PEDK_E_IO, "cannot read " + path);
    gzbuffer(f, 1 << 20);
    std::string out;
    std::vector<char> buf(1 << 22);
    for (;;) {
        int n = gzread(f, buf.data(), (unsigned)buf.size());
        if (n < 0) {
            gzclose(f);
            throw PedkError(PEDK_E_IO, "gzread failed on " + path);
        }
        if (n == 0) break;
        out.append(buf.data(), (size_t)n);
    }
    gzclose(f);
    return out;
}

// run fn(tag) for the 64 tags on a few host threads; TTT (the gate) is done last
void for_each_tag(const std::function<void(int)>& fn) {
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : nt > 16 ? 16 : nt;
    std::atomic<int> next(0);
    std::atomic<bool> failed(false);
    std::string err;
    int err_code = 0;
    std::vector<std::thread> th;
    auto work = [&]() {
        for (;;) {
            int t = next.fetch_add(1);
            if (t >= 63 || failed.load()) return;
            try {
                fn(t);
            } catch (const PedkError& e) {
                if (!failed.exchange(true)) { err = e.msg; err_code = e.code; }
            } catch (...) {
                if (!failed.exchange(true)) { err = "unknown error in writer thread"; err_code = PEDK_E_INTERNAL; }
            }
        }
    };
    for (unsigned i = 0; i < nt; ++i) th.emplace_back(work);
    for (auto& t : th) t.join();
    if (failed.load()) throw PedkError(err_code, err);
    fn(63);
}

struct Cache {
    std::map<std::string, pedk_sample*> samples;
};

Cache* cache_of(pedk_ctx* ctx) {
    if (!ctx->file_cache) ctx->file_cache = new Cache();
    return static_cast<Cache*>(ctx->file_cache);
}

pedk_sample* cached(pedk_ctx* ctx, const std::string& key) {
    Cache* c = cache_of(ctx);
    auto it = c->samples.find(key);
    return it == c->samples.end() ? nullptr : it->second;
}

pedk_sample* fresh_sample(pedk_ctx* ctx, const std::string& key) {
    Cache* c = cache_of(ctx);
    auto it = c->samples.find(key);
    if (it != c->samples.end()) {
        pedk_sample_destroy(it->second);
        c->samples.erase(it);
    }
    pedk_sample* s = nullptr;
    int rc = pedk_sample_create(ctx, &s);
    if (rc) throw PedkError(rc, ctx->err);
    c->samples[key] = s;
    return s;
}

void must(pedk_ctx* ctx, int rc) {
    if (rc) throw PedkError(rc, ctx->err);
}

// ---- streaming a group of read files into HBM (ped.pl:1383-1398) ----
struct Feeder {
    pedk_sample* s;
    static constexpr size_t CHUNK = 32u << 20;
    static constexpr int NBUF = 4;
    char* buf[NBUF] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[NBUF];
    bool used[NBUF] = {false, false, false, false};
    int cur = 0;
    size_t fill = 0;
    explicit Feeder(pedk_sample* sample) : s(sample) {
        for (int i = 0; i < NBUF; ++i) {
            PEDK_CUDA_TRY(cudaMallocHost(&buf[i], CHUNK));
            PEDK_CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
    }
    ~Feeder() {
        cudaStreamSynchronize(s->ctx->copy_stream);
        for (int i = 0; i < NBUF; ++i) {
            if (buf[i]) cudaFreeHost(buf[i]);
            cudaEventDestroy(ev[i]);
        }
    }
    char* space(size_t* room) {
        *room = CHUNK - fill;
        return buf[cur] + fill;
    }
    void advance(size_t n) {
        fill += n;
        if (fill == CHUNK) flush();
    }
    void flush() {
        if (!fill) return;
        must(s->ctx, pedk_sample_add_fastq(s, buf[cur], fill));
        PEDK_CUDA_TRY(cudaEventRecord(ev[cur], s->ctx->copy_stream));
        used[cur] = true;
        cur = (cur + 1) % NBUF;
        fill = 0;
        if (used[cur]) PEDK_CUDA_TRY(cudaEventSynchronize(ev[cur]));
    }
    // plain files are read with several preads per chunk (a single fread stream tops out far below
    // what the page cache or an NVMe array delivers)
    void feed_plain(const std::string& path) {
        int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw PedkError(PEDK_E_IO, "cannot read " + path);
        struct stat sb;
        if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
            // not a regular file (a pipe, a device): stream it
            close(fd);
            FILE* f = fopen(path.c_str(), "rb");
            if (!f) throw PedkError(PEDK_E_IO, "cannot read " + path);
            for (;;) {
                size_t room;
                char* p = space(&room);
                size_t n = fread(p, 1, room, f);
                if (n == 0) break;
                advance(n);
            }
            fclose(f);
            return;
        }
        const size_t total = (size_t)sb.st_size;
        size_t off = 0;
        std::atomic<bool> bad(false);
        while (off < total) {
            size_t room;
            char* p = space(&room);
            const size_t n = std::min(room, total - off);
            const unsigned parts = n >= (8u << 20) ? 8 : 1;
            std::vector<std::thread> th;
            for (unsigned i = 0; i < parts; ++i) {
                const size_t a = n * i / parts, b = n * (i + 1) / parts;
                th.emplace_back([&, a, b]() {
                    size_t done = a;
                    while (done < b) {
                        ssize_t r = pread(fd, p + done, b - done, (off_t)(off + done));
                        if (r <= 0) {
                            bad = true;
                            return;
                        }
                        done += (size_t)r;
                    }
                });
            }
            for (auto& t : th) t.join();
            if (bad) {
                close(fd);
                throw PedkError(PEDK_E_IO, "read failed on " + path);
            }
            advance(n);
            off += n;
        }
        close(fd);
    }
    void feed_gz(const std::string& path) {
        gzFile f = gzopen(path.c_str(), "rb");
        if (!f) throw PedkError(PEDK_E_IO, "cannot read " + path);
        gzbuffer(f, 1 << 20);
        for (;;) {
            size_t room;
            char* p = space(&room);
            int n = gzread(f, p, (unsigned)std::min<size_t>(room, 1u << 30));
            if (n < 0) {
                gzclose(f);
                throw PedkError(PEDK_E_IO, "gzread failed on " + path);
            }
            if (n == 0) break;
            advance((size_t)n);
        }
        gzclose(f);
    }
    void feed_pipe(const std::string& cmd) {
        FILE* f = popen(cmd.c_str(), "r");
        if (!f) throw PedkError(PEDK_E_IO, "cannot run " + cmd);
        for (;;) {
            size_t room;
            char* p = space(&room);
            size_t n = fread(p, 1, room, f);
            if (n == 0) break;
            advance(n);
        }
        if (pclose(f) != 0) throw PedkError(PEDK_E_IO, "failed: " + cmd);
    }
};

std::string shell_quote(const std::string& s) {
    std::string r = "'";
    for (char c : s) {
        if (c == '\'') r += "'\\''";
        else r += c;
    }
    return r + "'";
}

void do_mk_sort_uniq(pedk_ctx* ctx, const std::string& wd, const std::string& subject, int clipping, int* auto_clipped,
                     int* clipping_used) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    std::string rd = wd + "/" + subject + "/read";
    DIR* d = opendir(rd.c_str());
    if (!d) throw PedkError(PEDK_E_IO, "cannot open " + rd);
    std::vector<std::string> names;
    while (dirent* e = readdir(d)) names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());  // `sort readdir` (ped.pl:1370)
    std::vector<std::string> cls[4];       // gz, bz2, xz, *q  (ped.pl:1371-1379)
    for (const std::string& n : names) {
        if (ends_with(n, "gz")) cls[0].push_back(n);
        else if (ends_with(n, "bz2")) cls[1].push_back(n);
        else if (ends_with(n, "xz")) cls[2].push_back(n);
        else if (ends_with(n, "q")) cls[3].push_back(n);
    }
    // every class re-opens the TAG.tmp files with ">" (ped.pl:1430), so only the last
    // non-empty class reaches `sort | uniq`
    int last = -1;
    for (int c = 0; c < 4; ++c)
        if (!cls[c].empty()) last = c;
    pedk_sample* s = fresh_sample(ctx, wd + "/" + subject);
    if (last >= 0) {
        Feeder feed(s);
        for (const std::string& n : cls[last]) {
            std::string p = rd + "/" + n;
            if (last == 0) feed.feed_gz(p);
            else if (last == 1) feed.feed_pipe("bzcat " + shell_quote(p));
            else if (last == 2) feed.feed_pipe("xzcat " + shell_quote(p));
            else feed.feed_plain(p);
        }
        feed.flush();
    }
    pedk_ingest_info info;
    must(ctx, pedk_sample_ingest(s, clipping, &info));
    if (clipping_used) *clipping_used = info.clipping_used;
    if (auto_clipped) *auto_clipped = info.auto_clipped;
    std::string od = wd + "/" + subject + "/sort_uniq";
    mkdir_p(od);
    ensure_host_reads(s);
    for_each_tag([&](int t) {
        std::string text;
        size_t bytes = 0;
        for (const std::string& r : s->h_reads[t]) bytes += r.size() + 1;
        text.reserve(bytes);
        for (const std::string& r : s->h_reads[t]) {
            text += r;
            text += '\n';
        }
        write_gz_atomic(od + "/" + subject + ".sort_uniq." + tag_name(t) + ".gz", text.data(), text.size());
    });
    for (auto& v : s->h_reads) std::vector<std::string>().swap(v);
    s->have_reads = false;
}

// mkChrFromFile + mkControlRead: the reference FASTA as the control sample
void do_mk_control_read(pedk_ctx* ctx, const std::string& wd, const std::string& ref, const std::string& file) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    const std::string dir = wd + "/" + ref;
    const std::string path = dir + "/" + file;
    if (!exists(path)) throw PedkError(PEDK_E_IO, file + " is not found in " + dir);  // ped.pl:1031
    pedk_sample* s = fresh_sample(ctx, wd + "/" + ref);
    {
        Feeder feed(s);
        if (ends_with(file, "gz")) feed.feed_gz(path);                                   // ped.pl:1033-1038
        else if (ends_with(file, "bz2")) feed.feed_pipe("bzcat " + shell_quote(path));
        else feed.feed_plain(path);
        feed.flush();
    }
    must(ctx, pedk_sample_ingest_reference(s, 0, 0, nullptr));
    // the chromosome files (later stages of ped.pl read them; a repeated name overwrites)
    for (size_t i = 0; i < s->chromosomes.size(); ++i) {
        const pedk_sample::Chromosome& c = s->chromosomes[i];
        if (!c.live) continue;
        std::string seq((size_t)c.length, '\0');
        must(ctx, pedk_sample_chromosome(s, (int)i, nullptr, 0, nullptr, nullptr, c.length ? &seq[0] : nullptr,
                                         c.length));
        write_plain_atomic(dir + "/" + c.name, seq.data(), seq.size());
    }
    const std::string od = dir + "/sort_uniq";
    mkdir_p(od);
    ensure_host_reads(s);
    for_each_tag([&](int t) {
        std::string text;
        size_t bytes = 0;
        for (const std::string& r : s->h_reads[t]) bytes += r.size() + 1;
        text.reserve(bytes);
        for (const std::string& r : s->h_reads[t]) {
            text += r;
            text += '\n';
        }
        write_gz_atomic(od + "/" + ref + ".sort_uniq." + tag_name(t) + ".gz", text.data(), text.size());
    });
    for (auto& v : s->h_reads) std::vector<std::string>().swap(v);
    s->have_reads = false;
}

// a sample whose reads come from existing sort_uniq files (one read per line)
pedk_sample* sample_from_sort_uniq(pedk_ctx* ctx, const std::string& wd, const std::string& subject) {
    pedk_sample* s = fresh_sample(ctx, wd + "/" + subject);
    s->line_format = true;
    std::string dir = wd + "/" + subject + "/sort_uniq/";
    for (int t = 0; t < 64; ++t) {
        // ped.pl:903-907 also accepts the older <subject>.TAG.gz name
        std::string p = dir + subject + ".sort_uniq." + tag_name(t) + ".gz";
        if (!exists(p)) p = dir + subject + "." + tag_name(t) + ".gz";
        if (!exists(p)) continue;
        std::string text = read_gz(p);
        must(ctx, pedk_sample_add_fastq(s, text.data(), text.size()));
        PEDK_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    }
    must(ctx, pedk_sample_ingest(s, 0, nullptr));
    return s;
}

std::string count_text_of(pedk_sample* s, int tag) {
    char* p = nullptr;
    uint64_t n = 0;
    must(s->ctx, pedk_sample_count_text(s, tag, &p, &n));
    std::string r(p, n);
    free(p);
    return r;
}

std::string lbc_text_of(pedk_sample* s, int tag) {
    char* p = nullptr;
    uint64_t n = 0;
    must(s->ctx, pedk_sample_lbc_text(s, tag, &p, &n));
    std::string r(p, n);
    free(p);
    return r;
}

void do_count_kmer(pedk_ctx* ctx, const std::string& wd, const std::string& subject, int k) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    pedk_sample* s = cached(ctx, wd + "/" + subject);
    if (!s || !s->ingested || s->reads_released || s->host_only_table)
        s = sample_from_sort_uniq(ctx, wd, subject);
    must(ctx, pedk_sample_count(s, k, nullptr));
    must(ctx, pedk_sample_release_reads(s));
    ensure_host_table(s);
    std::string od = wd + "/" + subject + "/count";
    mkdir_p(od);
    for_each_tag([&](int t) {
        std::string text = count_text_of(s, t);
        write_gz_atomic(od + "/" + subject + ".count." + tag_name(t) + ".gz", text.data(), text.size());
    });
}

uint64_t encode_kmer(const char* p, size_t n) {
    uint64_t v = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!is_acgt((unsigned char)p[i])) throw PedkError(PEDK_E_IO, "unexpected character in a k-mer table");
        v = (v << 2) | base_code((unsigned char)p[i]);
    }
    return v;
}

// a sample whose strand table comes from existing count files
pedk_sample* sample_from_count_files(pedk_ctx* ctx, const std::string& wd, const std::string& subject, int k) {
    pedk_sample* s = fresh_sample(ctx, wd + "/" + subject);
    s->k = k;
    s->ingested = true;
    std::string dir = wd + "/" + subject + "/count/";
    for (int t = 0; t < 64; ++t) {
        s->tag_start[t] = s->h_kmer.size();
        std::string text = read_gz(dir + subject + ".count." + tag_name(t) + ".gz");
        size_t pos = 0;
        while (pos < text.size()) {
            size_t nl = text.find('\n', pos);
            if (nl == std::string::npos) nl = text.size();
            size_t tab = text.find('\t', pos);
            if (tab == std::string::npos || tab > nl) throw PedkError(PEDK_E_IO, "malformed count file for " + subject);
            if ((int)(tab - pos) != k) throw PedkError(PEDK_E_INVAL, "count files of " + subject + " hold another k");
            s->h_kmer.push_back(encode_kmer(text.data() + pos, tab - pos));
            s->h_cnt.push_back((uint32_t)strtoul(text.c_str() + tab + 1, nullptr, 10));
            pos = nl + 1;
        }
    }
    s->tag_start[64] = s->h_kmer.size();
    s->have_table = true;
    s->host_only_table = true;
    return s;
}

pedk_sample* counted_sample(pedk_ctx* ctx, const std::string& wd, const std::string& subject, int k) {
    pedk_sample* s = cached(ctx, wd + "/" + subject);
    if (s && s->k == k && (s->have_table || s->counted)) return s;
    return sample_from_count_files(ctx, wd, subject, k);
}

void do_lbc_kmer(pedk_ctx* ctx, const std::string& wd, const std::string& subject, int k) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    pedk_sample* s = counted_sample(ctx, wd, subject, k);
    std::string od = wd + "/" + subject + "/lbc";
    mkdir_p(od);
    for_each_tag([&](int t) {
        std::string text = lbc_text_of(s, t);
        write_gz_atomic(od + "/" + subject + ".lbc." + tag_name(t) + ".gz", text.data(), text.size());
    });
}

void do_join_kmer(pedk_ctx* ctx, const std::string& wd, const std::string& target, const std::string& control, int k,
                  const std::string& tmpdir, bool have_ref) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    pedk_sample* c = counted_sample(ctx, wd, control, k);
    pedk_sample* t = counted_sample(ctx, wd, target, k);
    pedk_result* r = nullptr;
    must(ctx, pedk_join(ctx, c, t, &r));
    const char* text;
    uint64_t len;
    pedk_result_text(r, &text, &len);
    try {
        if (!have_ref) {
            // ped.pl:375 `cat $tmpdir/$target.snp.* > $wd/$target/$target.kmer.snp`
            write_plain_atomic(wd + "/" + target + "/" + target + ".kmer.snp", text, len);
        } else {
            // ped.pl:750: one file per tag for mapKmerSub (ped.pl:647)
            mkdir_p(tmpdir);
            const char* p = text;
            const char* end = text + len;
            for (int tag = 0; tag < 64; ++tag) {
                std::string tn = tag_name(tag);
                const char* q = p;
                while (q < end && memcmp(q, tn.data(), 3) == 0) {
                    const char* nl = (const char*)memchr(q, '\n', (size_t)(end - q));
                    q = nl ? nl + 1 : end;
                }
                write_plain_atomic(tmpdir + "/" + target + ".snp." + tn, p, (size_t)(q - p));
                p = q;
            }
        }
    } catch (...) {
        pedk_result_destroy(r);
        throw;
    }
    pedk_result_destroy(r);
}

std::string read_plain(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return std::string();  // ped.pl reads a missing file as an empty one
    std::string out;
    std::vector<char> buf(1 << 20);
    for (;;) {
        size_t n = fread(buf.data(), 1, buf.size(), f);
        if (!n) break;
        out.append(buf.data(), n);
    }
    fclose(f);
    return out;
}

// the reference as a resident sample: from this process's mkControlRead, or parsed again from the FASTA
pedk_sample* reference_sample(pedk_ctx* ctx, const std::string& wd, const std::string& ref, const std::string& file) {
    pedk_sample* s = cached(ctx, wd + "/" + ref);
    if (s && s->ref_clean.p) return s;
    if (file.empty()) return nullptr;
    const std::string path = wd + "/" + ref + "/" + file;
    if (!exists(path)) throw PedkError(PEDK_E_IO, file + " is not found in " + wd + "/" + ref);
    s = fresh_sample(ctx, wd + "/" + ref);
    {
        Feeder feed(s);
        if (ends_with(file, "gz")) feed.feed_gz(path);
        else if (ends_with(file, "bz2")) feed.feed_pipe("bzcat " + shell_quote(path));
        else feed.feed_plain(path);
        feed.flush();
    }
    must(ctx, pedk_sample_ingest_reference(s, 0, 0, nullptr));
    return s;
}

// mk20 (ped.pl:1213-1249): mk20mer per chromosome + mkUniq per tag -> <ref>/ref20_uniq.TAG.gz
void do_mk_ref20_uniq(pedk_ctx* ctx, const std::string& wd, const std::string& ref, const std::string& file, int k) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    pedk_sample* s = reference_sample(ctx, wd, ref, file);
    if (!s) throw PedkError(PEDK_E_INVAL, "pedk_mk_ref20_uniq: the reference is not resident and no FASTA file was named");
    must(ctx, pedk_sample_ref20_uniq(s, k, nullptr));
    if (!s->ref20.any_chromosome) return;  // no mk20mer ran: no tmp files, no mkUniq (ped.pl:1238)
    const std::string dir = wd + "/" + ref;
    for_each_tag([&](int t) {
        const std::string text = ref20_text(s, t);
        write_gz_atomic(dir + "/ref20_uniq." + tag_name(t) + ".gz", text.data(), text.size());
    });
}

// the unique k-mer index of the reference: resident, or read back from <ref>/ref20_uniq.TAG.gz
pedk_sample* indexed_reference(pedk_ctx* ctx, const std::string& wd, const std::string& ref) {
    pedk_sample* s = cached(ctx, wd + "/" + ref);
    if (s && s->ref20.built) return s;
    if (!s) s = fresh_sample(ctx, wd + "/" + ref);
    s->ref20 = pedk_sample::Ref20{};
    for (int t = 0; t < 64; ++t) {
        const std::string p = wd + "/" + ref + "/ref20_uniq." + tag_name(t) + ".gz";
        ref20_append_text(s, t, exists(p) ? read_gz(p) : std::string());
    }
    s->ref20.built = true;
    return s;
}

// mapKmer (ped.pl:625-685): <tmpdir>/<t>.snp.TAG x <ref>/ref20_uniq.TAG.gz -> <tmpdir>/<t>.map.TAG
void do_map_kmer(pedk_ctx* ctx, const std::string& wd, const std::string& target, const std::string& ref,
                 const std::string& tmpdir) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    pedk_sample* s = indexed_reference(ctx, wd, ref);  // resident, or back from the ref20_uniq files
    for (int t = 0; t < 64; ++t) {
        const std::string snp = read_plain(tmpdir + "/" + target + ".snp." + tag_name(t));
        const std::string m = map_kmer_text(s, snp.data(), snp.size());
        write_plain_atomic(tmpdir + "/" + target + ".map." + tag_name(t), m.data(), m.size());
    }
}

// cnv (ped.pl:362-366) after the count tables: cnvJoin (ped.pl:934-947) -> <wd>/<target>/<target>.<control>.cnv
void do_cnv(pedk_ctx* ctx, const std::string& wd, const std::string& target, const std::string& control,
            const std::string& ref, int k) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    pedk_sample* r = indexed_reference(ctx, wd, ref);
    pedk_sample* c = counted_sample(ctx, wd, control, k);
    pedk_sample* t = counted_sample(ctx, wd, target, k);
    const std::string text = cnv_text(r, c, t);
    write_plain_atomic(wd + "/" + target + "/" + target + "." + control + ".cnv", text.data(), text.size());
}

// snpMkT ... kmerReadCount (ped.pl:380-386): <tmpdir>/<t>.map.TAG + the reads of both samples + <ref>/chrNAME
// -> <wd>/<target>/<target>.kmer.snp
void do_verify_kmer(pedk_ctx* ctx, const std::string& wd, const std::string& target, const std::string& control,
                    const std::string& ref, const std::string& tmpdir) {
    PEDK_CUDA_TRY(cudaSetDevice(ctx->device));
    auto with_reads = [&](const std::string& subject) {
        pedk_sample* s = cached(ctx, wd + "/" + subject);
        if (!s || !s->ingested || s->reads_released || s->host_only_table) s = sample_from_sort_uniq(ctx, wd, subject);
        return s;
    };
    pedk_sample* t = with_reads(target);
    pedk_sample* c = control == target ? t : with_reads(control);
    std::string map;
    for (int tag = 0; tag < 64; ++tag) map += read_plain(tmpdir + "/" + target + ".map." + tag_name(tag));  // `cat *.map.*`
    std::map<std::string, std::string> files;
    std::map<std::string, bool> missing;
    ChromosomeLookup lookup = [&](const std::string& name) -> const std::string* {
        auto it = files.find(name);
        if (it != files.end()) return &it->second;
        if (missing.count(name)) return nullptr;
        const std::string p = wd + "/" + ref + "/" + name;
        struct stat sb;
        if (name.find('/') != std::string::npos || stat(p.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) {
            missing[name] = true;
            return nullptr;
        }
        return &(files[name] = read_plain(p));
    };
    const std::string text = verify_text(t, c, lookup, map.data(), map.size());
    write_plain_atomic(wd + "/" + target + "/" + target + ".kmer.snp", text.data(), text.size());
}

void do_to_vcf(const std::string& wd, const std::string& target) {
    const std::string snp = read_plain(wd + "/" + target + "/" + target + ".kmer.snp");
    const std::string vcf = vcf_text(snp.data(), snp.size(), target);
    write_plain_atomic(wd + "/" + target + "/" + target + ".kmer.vcf", vcf.data(), vcf.size());
}

}  // namespace

namespace pedk {
void file_cache_destroy(pedk_ctx* ctx) {
    if (!ctx->file_cache) return;
    Cache* c = static_cast<Cache*>(ctx->file_cache);
    for (auto& kv : c->samples) pedk_sample_destroy(kv.second);
    delete c;
    ctx->file_cache = nullptr;
}
}  // namespace pedk

extern "C" int pedk_mk_control_read(pedk_ctx* ctx, const char* wd, const char* ref, const char* file) {
    if (!ctx || !wd || !ref || !file) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_mk_control_read(ctx, wd, ref, file);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_mk_ref20_uniq(pedk_ctx* ctx, const char* wd, const char* ref, const char* file, int k) {
    if (!ctx || !wd || !ref) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_mk_ref20_uniq(ctx, wd, ref, file ? file : "", k);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_map_kmer(pedk_ctx* ctx, const char* wd, const char* target, const char* ref, const char* tmpdir) {
    if (!ctx || !wd || !target || !ref || !tmpdir) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_map_kmer(ctx, wd, target, ref, tmpdir);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_cnv(pedk_ctx* ctx, const char* wd, const char* target, const char* control, const char* ref, int k) {
    if (!ctx || !wd || !target || !control || !ref) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_cnv(ctx, wd, target, control, ref, k);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_verify_kmer(pedk_ctx* ctx, const char* wd, const char* target, const char* control, const char* ref,
                                const char* tmpdir) {
    if (!ctx || !wd || !target || !control || !ref || !tmpdir) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_verify_kmer(ctx, wd, target, control, ref, tmpdir);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_to_vcf(pedk_ctx* ctx, const char* wd, const char* target) {
    if (!ctx || !wd || !target) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_to_vcf(wd, target);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_mk_sort_uniq(pedk_ctx* ctx, const char* wd, const char* subject, int clipping,
                                 int* clipping_used, int* auto_clipped) {
    if (!ctx || !wd || !subject) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_mk_sort_uniq(ctx, wd, subject, clipping, auto_clipped, clipping_used);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_count_kmer(pedk_ctx* ctx, const char* wd, const char* subject, int k) {
    if (!ctx || !wd || !subject) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_count_kmer(ctx, wd, subject, k);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_lbc_kmer(pedk_ctx* ctx, const char* wd, const char* subject, int k) {
    if (!ctx || !wd || !subject) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_lbc_kmer(ctx, wd, subject, k);
    return PEDK_OK;
    PEDK_API_END(ctx)
}

extern "C" int pedk_join_kmer(pedk_ctx* ctx, const char* wd, const char* target, const char* control, int k,
                              const char* tmpdir, int have_ref) {
    if (!ctx || !wd || !target || !control || (have_ref && !tmpdir)) return PEDK_E_INVAL;
    PEDK_API_BEGIN
    do_join_kmer(ctx, wd, target, control, k, tmpdir ? tmpdir : "", have_ref != 0);
    return PEDK_OK;
    PEDK_API_END(ctx)
}
