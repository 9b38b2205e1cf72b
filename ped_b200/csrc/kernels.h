// Host-side launchers of the pedk kernels (one per stage of SURVEY.md 7.2).
// Every launcher enqueues on `st` and returns; none of them synchronises.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace pedk {

// ---------------- S0: ingest (ingest.cu) ----------------
constexpr int INGEST_TILE_BYTES = 16384;
constexpr int LEN_HIST_BINS = 1024;  // lengths >= LEN_HIST_BINS-1 land in the last bin

// tile_counts[t] = number of '\n' in bytes [t*TILE, (t+1)*TILE); fq must be
// readable (and newline-free) up to the next multiple of INGEST_TILE_BYTES.
void launch_newline_count(const uint8_t* fq, size_t n, uint32_t* tile_counts, cudaStream_t st);
// seq_start[r] / seq_end[r]: byte span of line 4r+1 (ped.pl:1435-1436 state machine)
// plain_lines != 0: every line is a read (sort_uniq text); seq_start[0] must be preset to 0
void launch_line_index(const uint8_t* fq, size_t n, const uint64_t* tile_prefix, uint64_t* seq_start,
                       uint64_t* seq_end, uint64_t n_records, int plain_lines, cudaStream_t st);
// rlen[r], rflag[r] (bit0: contains 'N' -> dropped, ped.pl:1436; bit1: byte outside ACGTN);
// len_hist over reads without N; counters[0] += reads without N, counters[1] += kept reads with bad bytes
void launch_read_scan(const uint8_t* fq, const uint64_t* seq_start, const uint64_t* seq_end, uint64_t n_records,
                      uint32_t* rlen, uint8_t* rflag, unsigned long long* len_hist,
                      unsigned long long* counters, cudaStream_t st);
// chunks[r] for r in [0, n_records): number of clip-length pieces read r contributes
//   r >= r_limit -> 0;  flagged -> 0;  whole==1: (len == clip ? 1 : 0);  else len / clip
void launch_chunk_count(const uint32_t* rlen, const uint8_t* rflag, uint64_t n_records, uint64_t r_limit,
                        uint32_t clip, int whole, uint32_t* chunks, cudaStream_t st);
// 2-bit pack every chunk, take min(read, revcomp(read)) (ped.pl:1460-1464 emits both;
// we keep one representative), pal[i] = 1 iff the chunk equals its own reverse complement
void launch_pack(const uint8_t* fq, const uint64_t* seq_start, const uint32_t* chunks, const uint64_t* chunk_base,
                 uint64_t n_records, uint32_t clip, int W, uint64_t* words, uint8_t* pal, cudaStream_t st);

// minmax[0] = min, minmax[1] = max sequence-line length over all records (preset 0xffffffff, 0)
void launch_len_stats(const uint64_t* seq_start, const uint64_t* seq_end, uint64_t n_records, unsigned* minmax,
                      cudaStream_t st);
// fixed-length fast path: every record has L bases; slot = record index; valid[r] = 0 for reads
// with an N (dropped, ped.pl:1436); counters[0] += reads without N, counters[1] += of those
// the reads with a byte outside ACGT
void launch_pack_fixed(const uint8_t* fq, const uint64_t* seq_start, uint64_t n_records, int L, int W,
                       uint64_t* words, uint8_t* pal, uint8_t* valid, unsigned long long* counters,
                       cudaStream_t st);

// ---------------- S1: exact read dedup (dedup.cu) ----------------
// table: capacity slots (power of two) preset to 0xFFFFFFFF; uniq[i] = 1 iff chunk i is
// the representative of its value
// valid (may be null): slots with valid[i] == 0 are skipped
void launch_dedup(const uint64_t* words, int W, uint32_t n, uint32_t* table, uint32_t capacity, const uint8_t* valid,
                  uint8_t* uniq, cudaStream_t st);
// f2: hits[s] += number of the L-base windows of string s (text[start[s] .. +length[s])) that are reads of the
// sample whose unique canonical reads are `words` and whose k_dedup table is `table`
// the same result from one warp per string (rolling windows over the packed string) and the table of
// launch_read_table (its own hash); hits[s] is written, not added to
void launch_read_table(const uint64_t* words, int W, uint32_t n, uint32_t* table, uint32_t capacity, cudaStream_t st);
void launch_probe_strings(const uint8_t* text, const uint64_t* start, const uint32_t* length, uint64_t n_strings, int L,
                          int W, const uint64_t* words, const uint32_t* table, uint32_t capacity, uint32_t* hits,
                          cudaStream_t st);
void launch_probe_windows(const uint8_t* text, const uint64_t* start, const uint32_t* length, uint64_t n_strings, int L,
                          int W, const uint64_t* words, const uint32_t* table, uint32_t capacity, uint32_t* hits,
                          cudaStream_t st);
void launch_compact_reads(const uint64_t* words, const uint8_t* pal, const uint8_t* uniq, const uint64_t* pos,
                          int W, uint32_t n, uint64_t* out_words, uint8_t* out_pal, cudaStream_t st);

// *out += number of non-zero flags
void launch_count_flags(const uint8_t* flags, uint64_t n, unsigned long long* out, cudaStream_t st);

// ---------------- S2: k-mer extraction (kmer.cu) ----------------
// out[r*nwin + i] = canonical k-mer of window i of unique read r; digit_hist[p*256 + d]
// accumulates the 8-bit digit histograms for passes p < n_pass.  When lo < hi only
// k-mers with lo <= canon < hi are wanted: others are written as ~0 (sorts last).
void launch_extract(const uint64_t* words, int W, int L, int k, uint64_t n_reads, uint64_t* out,
                    unsigned long long* digit_hist, int n_pass, cudaStream_t st);
// canonical k-mers of every window of the palindromic reads only (count corrections)
void launch_extract_pal(const uint64_t* words, const uint8_t* pal, int W, int L, int k, uint64_t n_reads,
                        uint64_t* out, unsigned long long* cursor, cudaStream_t st);

// Peer memory of a multi-GPU job: rank r's receive buffer is p[r] (NVLink peer memory mapped
// through CUDA IPC for r != me).
constexpr int PEDK_MAX_PEERS = 8;
struct PeerPtrs {
    uint64_t* p[PEDK_MAX_PEERS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// ---------------- S2-S4 bucketed: bin, bin again, count per bucket (bucket.cu) ----------------
// a, b = bits of the two binning levels (common.cuh kmer_bits_a/_b); wide != 0: the stored
// remainder (2k - a bits of kmer_hash) needs 64-bit words (k > 20), else 32-bit words.
// histA[bin] += canonical k-mers of the reads whose hash falls into level-A bin `bin`
void launch_part_count(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a,
                       unsigned long long* histA, cudaStream_t st);
// every canonical k-mer of the reads -> low 2k-a bits of its hash, appended to its bin's run:
// cur[bin] = next free element of that bin inside the buffer of the rank that owns it
// (dst.p[owner], owner = bin * nranks >> a; peer memory over NVLink for owner != me).
// lim != null (optimistic sizing): lim[bin] = end of the segment this rank may fill; a run that
// does not fit is dropped and *overflow set.  n_pass (a power of two) > 1: only the k-mers of the
// bins with bin % n_pass == pass are kept (a sample whose instances do not fit HBM at once).
void launch_part_scatter(const uint64_t* words, int W, int L, int k, uint64_t n_reads, int a, int nranks, int wide,
                         unsigned long long* cur, const unsigned long long* lim, int* overflow, const PeerPtrs& dst,
                         int n_pass, int pass, cudaStream_t st);
unsigned sub_tile_keys();
// optimistic sizing: cnt[i] = cur[i] - (lim[i] - cap), what really went into segment i
void launch_cursor_counts(const unsigned long long* cur, const unsigned long long* lim, unsigned long long cap, int n,
                          unsigned long long* cnt, cudaStream_t st);
// A level-A bin is held as one or more segments: seg[s] = first element, seg[n_seg + s] = end,
// seg[2 n_seg + s] = bin; tile_prefix[n_seg + 1] = tiles of sub_tile_keys() before each segment.
// -> histB[bin * 256 + sub]
void launch_sub_hist(const void* keys, int wide, const unsigned long long* seg, const unsigned* tile_prefix,
                     int n_seg, unsigned n_tiles, int k, int a, int b, unsigned* histB, cudaStream_t st);
// curB[bin * 256 + sub] = next free element of that bucket in `out`
// limB != null (buckets sized optimistically): limB[bin * 256 + sub] = end of that bucket's room; a run
// that does not fit is dropped and *overflow set
void launch_sub_scatter(const void* in, void* out, int wide, const unsigned long long* seg,
                        const unsigned* tile_prefix, int n_seg, unsigned n_tiles, int k, int a, int b,
                        unsigned long long* curB, const unsigned long long* limB, int* overflow, cudaStream_t st);
// optimistic bucket layout: bin_index[bin] = index of the bin among this rank's bins of the pass (-1: not
// held here); every bucket gets `cap` elements
void launch_bucket_layout(const int* bin_index, unsigned n_buckets, unsigned long long cap, unsigned long long* offB,
                          unsigned long long* limB, cudaStream_t st);
// buckets [offB[i], endB[i]) -> table rows, as launch_count_rows below (exact sizes: endB = offB + 1);
// ticket = 1 device word, err = device flag raised when a bucket cannot be split any further
void launch_bucket_count(const void* keys, int wide, const unsigned long long* offB, const unsigned long long* endB,
                         unsigned n_buckets, int k, int a,
                         int b, int log_slots, int strand, const uint64_t* corr_key, const uint32_t* corr_cnt,
                         uint64_t n_corr, uint64_t* row_key, uint32_t* row_val, uint64_t cap,
                         unsigned long long* counters, unsigned* ticket, int* err, cudaStream_t st);

// ---------------- S3: LSD radix sort (radix.cu) ----------------
struct RadixWorkspace {
    uint64_t* lookback;      // tiles * 256 words
    unsigned int* tile_ctr;  // 1 word
    int* err;                // 1 word
    size_t tiles_cap;
};
size_t radix_tiles(size_t n);
// one onesweep pass on digit `pass` (bits [8*pass, 8*pass+8)); bin_base[256] = exclusive
// scan of that digit's global histogram
void launch_radix_pass(const uint64_t* in, uint64_t* out, const uint32_t* vin, uint32_t* vout, size_t n, int pass,
                       const unsigned long long* digit_hist, RadixWorkspace ws, cudaStream_t st);
// digit histograms of an existing key array (pairs sort; the instance sort gets its
// histograms from the extraction kernel)
void launch_digit_hist(const uint64_t* keys, size_t n, unsigned long long* digit_hist, int n_pass,
                       cudaStream_t st);
void launch_hist_to_base(unsigned long long* digit_hist, int n_pass, cudaStream_t st);

// ---------------- S4: run-length count -> table rows (kmer.cu) ----------------
// keys: sorted canonical instances.  strand == 0: rows (K, run length).  strand != 0: rows
// (K, c | sample<<31) and (revcomp K, c | sample<<31) with c the reference's count; corr =
// sorted (key, windows of palindromic reads) table, may be empty.  Rows are appended in no
// particular order.  counters[0] = rows appended (keeps counting past cap), counters[1] +=
// distinct canonical k-mers; both must be preset by the caller.
void launch_count_rows(const uint64_t* keys, size_t n, int k, int strand, int sample, const uint64_t* corr_key,
                       const uint32_t* corr_cnt, uint64_t n_corr, uint64_t* row_key, uint32_t* row_val, uint64_t cap,
                       unsigned long long* counters, cudaStream_t st);

// ---------------- S5/S6: lbc groups and edges (edges.cu) ----------------
// dst rows = src rows with or_bits set in the value (bit 31 marks the target sample)
// digit_hist != null: also accumulates the 8-bit digit histograms of the keys for n_pass passes
void launch_copy_rows(const uint64_t* src_key, const uint32_t* src_val, uint64_t n, uint64_t* dst_key,
                      uint32_t* dst_val, uint32_t or_bits, unsigned long long* digit_hist, int n_pass,
                      cudaStream_t st);
// first[(sample*64 + tag)] = smallest key of that sample in 3-nt bucket `tag` (preset ~0)
void launch_first_of_bucket(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int k,
                            unsigned long long* first, cudaStream_t st);

struct EdgeRec {
    uint64_t p;      // the (k-1)-mer
    uint32_t c[4];   // control counts by last base (ped.pl:777 columns 4-7)
    uint32_t t[4];   // target counts (columns 8-11)
    uint8_t cont;    // bit x set: base x in the control allele set (>25 %)
    uint8_t alt;     // same for the target
    uint8_t pad[6];
};
// joinKmerSub (ped.pl:703-782) on every (k-1)-mer group of the merged table that is not the
// first group of its bucket in either sample (those rows carry the F7 text quirk and are
// evaluated on the host).  Edges are appended through *n_edges; at most cap are stored.
// pass1_bits: device copy of the 2048-word table build_edge_pass1_table fills (host)
void build_edge_pass1_table(uint32_t* bits);
void launch_edges(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int k,
                  const unsigned long long* first, const uint32_t* pass1_bits, EdgeRec* edges, uint64_t cap,
                  unsigned long long* n_edges, cudaStream_t st);
struct GroupRec {
    uint64_t p;
    uint32_t c[4], t[4];
    uint8_t hasc[4], hast[4];  // row present in the table
};
// look up the groups of n_p given (k-1)-mers (host-side quirk rows)
void launch_lookup_groups(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, const uint64_t* p_list,
                          int n_p, GroupRec* out, cudaStream_t st);
// ---- hash join (edges.cu): rows -> join records binned by kmer_hash(K >> 2) -> per-bucket tables ----
// a' = kmer_bits_a(k-1), b' = kmer_bits_b(k-1) bits for the two binning levels of the (k-1)-mer hash
// hist256[bin] += rows of that level-A' bin; first[sample * 64 + tag] = min(K) (preset ~0)
void launch_rows_count(const uint64_t* row_key, uint64_t n, int sample, int k, unsigned long long* hist256,
                       unsigned long long* first, cudaStream_t st);
// rows -> 8-byte join records appended at cur[bin] inside dst.p[owner of bin]
void launch_rows_scatter(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, int sample, int k, int nranks,
                         unsigned long long* cur, const PeerPtrs& dst, cudaStream_t st);
// bit position of the level-B' digit inside a join record (for launch_sub_hist_bits / _scatter_bits)
int join_record_shift_b(int k, int b);
void launch_sub_hist_bits(const void* keys, int wide, const unsigned long long* seg, const unsigned* tile_prefix,
                          int n_seg, unsigned n_tiles, int shiftB, unsigned maskB, unsigned* histB, cudaStream_t st);
void launch_sub_scatter_bits(const void* in, void* out, int wide, const unsigned long long* seg,
                             const unsigned* tile_prefix, int n_seg, unsigned n_tiles, int shiftB, unsigned maskB,
                             unsigned long long* curB, const unsigned long long* limB, int* overflow,
                             cudaStream_t st);
// buckets of join records (offB[n_buckets + 1]) -> edges; groups that are the first of a 3-nt
// bucket in either sample (first[128]) are left to the host (F7)
void launch_join_bucket(const uint64_t* recs, const unsigned long long* offB, unsigned n_buckets, int k, int b,
                        int log_slots,
                        const unsigned long long* first, const uint32_t* pass1_bits, EdgeRec* edges, uint64_t cap,
                        unsigned long long* n_edges, unsigned* ticket, int* err, cudaStream_t st);
void launch_join_lookup(const uint64_t* recs, const unsigned long long* offB, int k, int b, const uint64_t* p_list,
                        int n_p, GroupRec* out, cudaStream_t st);

// *out += sum over rows of mix64(K * 0x9E3779B97F4A7C15 + count): order-independent table digest
void launch_rows_digest(const uint64_t* row_key, const uint32_t* row_val, uint64_t n, unsigned long long* out,
                        cudaStream_t st);

// ---------------- synthetic reads (synth.cu) ----------------
struct SynthParams;
void launch_synth_fastq(const SynthParams& p, const char* genome_dev, uint64_t first, uint64_t n, char* out_dev,
                        cudaStream_t st);

}  // namespace pedk
