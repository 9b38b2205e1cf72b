// Device-resident state of one sample (internal).
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "ctx.h"

struct ReadGroup {
    int L = 0, W = 0;      // bases per read, 64-bit words per read
    uint64_t n = 0;        // unique canonical representatives
    uint64_t n_pal = 0;    // of which palindromic
    pedk::DBuf<uint64_t> words;
    pedk::DBuf<uint8_t> pal;
};

struct pedk_sample {
    pedk_ctx* ctx = nullptr;
    // FASTQ text in HBM
    pedk::DBuf<uint8_t> fq;
    size_t fq_len = 0;
    cudaEvent_t ev_copied = nullptr;  // last H2D chunk has landed
    bool copy_pending = false;
    bool ingested = false;
    bool line_format = false;      // text is one read per line (sort_uniq files), not FASTQ
    bool reads_released = false;
    bool host_only_table = false;  // strand table was loaded from count files; nothing in HBM
    std::vector<ReadGroup> groups;
    pedk_ingest_info info{};
    // canonical count table
    int k = 0;
    bool counted = false;
    uint64_t D = 0;        // distinct canonical k-mers
    uint64_t n_rows = 0;   // rows of the strand table (K and revcomp K), unordered, in HBM
    pedk::DBuf<uint64_t> rkey;
    pedk::DBuf<uint32_t> rval;
    // bucketed path: the rows stay where the count kernel wrote them (the level-A buffer,
    // free by then) instead of being copied to exact-size arrays
    pedk::DBuf<uint8_t> row_store;
    const uint64_t* row_key_ptr() const { return row_store.p ? row_key_in_store : rkey.p; }
    const uint32_t* row_val_ptr() const { return row_store.p ? row_val_in_store : rval.p; }
    const uint64_t* row_key_in_store = nullptr;
    const uint32_t* row_val_in_store = nullptr;
    uint64_t n_corr = 0;   // sorted (k-mer, windows of palindromic reads) corrections
    pedk::DBuf<uint64_t> ckey;
    pedk::DBuf<uint32_t> ccnt;
    pedk_count_info cinfo{};
    // host caches for the text emitters
    bool have_table = false;
    std::vector<uint64_t> h_kmer;
    std::vector<uint32_t> h_cnt;
    uint64_t tag_start[65] = {0};
    bool have_reads = false;
    std::vector<std::string> h_reads[64];  // sorted strand-closed reads per bucket
    // reference as a sample (reference.cu): the cleaned chromosomes, back to back in HBM
    struct Chromosome {
        std::string name;   // "chr" + the name ped.pl derives from the header (ped.pl:1045-1050)
        uint64_t start = 0, length = 0;
        bool live = true;   // false: a later sequence of the same name overwrote this one
    };
    std::vector<Chromosome> chromosomes;
    pedk::DBuf<uint8_t> ref_clean;
    uint64_t ref_clean_len = 0;
    // f1: the unique-k-mer index of the reference (refindex.cu), in the order of the ref20_uniq files
    struct Ref20 {
        bool built = false;
        bool any_chromosome = false;   // mk20 wrote its 64 files (there was a chromosome to process)
        int k = 0;
        std::vector<uint64_t> kmer;    // ascending
        std::vector<uint32_t> chr;     // index into names
        std::vector<uint64_t> pos;     // 1-based: first base (strand f) or last base (strand r) of the k-mer
        std::vector<uint8_t> strand;   // 0 = f, 1 = r
        std::vector<std::string> names;  // chromosome names as mk20mer prints them (no "chr" prefix)
        uint64_t tag_start[65] = {0};
        pedk::DBuf<uint64_t> d_kmer;   // the keys in HBM, for the edge look-ups
    } ref20;
};

namespace pedk {
// strand table of one sample (or the merged table of two) in HBM
struct RowTable {
    DBuf<uint64_t> key;
    DBuf<uint32_t> val;
    uint64_t n = 0;
};
// expand + sort; target may be null (single-sample table, sample bit 0)
void build_rows(pedk_ctx* ctx, pedk_sample* control, pedk_sample* target, RowTable* out);
void ensure_host_table(pedk_sample* s);
// exact de-duplication of n_c packed canonical reads (valid[i] == 0: skip) -> a ReadGroup of s
void dedup_group(pedk_sample* s, int L, int W, DBuf<uint64_t>& words, DBuf<uint8_t>& pal, uint64_t n_c,
                 const uint8_t* valid = nullptr);
// reads the sequence read_scalar-style helpers of sample.cu rely on

void ensure_host_reads(pedk_sample* s);
// f1 (refindex.cu)
void ref20_append_text(pedk_sample* s, int tag, const std::string& text);
std::string ref20_text(pedk_sample* s, int tag);
std::string map_kmer_text(pedk_sample* ref, const char* snp, uint64_t snp_len);
std::string cnv_text(pedk_sample* ref, pedk_sample* control, pedk_sample* target);
std::string kmer_string(uint64_t v, int k);
// f2 (verify.cu): the text of a chromosome file ("chr" + name) or null when there is none
typedef std::function<const std::string*(const std::string&)> ChromosomeLookup;
std::string verify_text(pedk_sample* target, pedk_sample* control, const ChromosomeLookup& chromosome, const char* map,
                        uint64_t map_len);
std::string vcf_text(const char* snp, uint64_t snp_len, const std::string& target);
}  // namespace pedk
