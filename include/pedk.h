/*
 * pedk.h -- C ABI of libpedk.so, the B200-native k-mer hot path of PED
 * (Polymorphic Edge Detection, akiomiyao/ped).
 *
 * The reference has no library API: its boundary is the ped.pl command line
 * and the files its subs hand to each other.  Every entry point
 * below names the ped.pl sub (file:line in the reference tree) whose work it
 * replaces; perl/PedK.xs binds exactly these symbols and INTEGRATION.md shows
 * the edit a ped.pl maintainer makes.
 *
 * Conventions: plain C types only; the caller owns every pointer it passes and
 * every output buffer; strings are copied before return.  Functions return 0 on
 * success or a negative PEDK_E_* code; pedk_last_error() gives the message.  No
 * exceptions and no exit() cross this boundary.  A context drives ONE CUDA
 * device from one host thread at a time; multi-GPU jobs run one context per
 * process/rank and connect them with pedk_comm_* (NCCL).  There is no CPU
 * fallback: without a usable sm_100 device pedk_open fails.
 *
 * 2-bit code A=0 C=1 G=2 T=3 (byte order of LANG=C, ped.pl:91).  A k-mer is a
 * right-aligned 2k-bit integer, first base most significant, so ascending
 * integers are ascending strings.
 */
#ifndef PEDK_H
#define PEDK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEDK_OK 0
#define PEDK_E_INVAL (-1)       /* bad argument */
#define PEDK_E_CUDA (-2)        /* CUDA runtime error (sticky for the context) */
#define PEDK_E_NOMEM (-3)
#define PEDK_E_IO (-4)
#define PEDK_E_ALPHABET (-5)    /* a kept read holds a byte outside ACGTN */
#define PEDK_E_UNSUPPORTED (-6)
#define PEDK_E_NCCL (-7)
#define PEDK_E_INTERNAL (-8)
#define PEDK_E_NODEVICE (-9)    /* no CUDA device / not sm_100 */

typedef struct pedk_ctx pedk_ctx;
typedef struct pedk_sample pedk_sample;
typedef struct pedk_result pedk_result;

/* ---- context ---------------------------------------------------------- */

/* devices/ndev: CUDA ordinals this context may use; ndev must be 1 (one rank
 * per GPU).  Replaces the process pool of ped.pl:244-255, 3504-3543. */
int pedk_open(pedk_ctx **ctx, const int *devices, int ndev);
int pedk_close(pedk_ctx *ctx);
const char *pedk_last_error(pedk_ctx *ctx);
const char *pedk_strerror(int code);
int pedk_abi_version(void);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on, so a
 * caller can bracket calls with its own CUDA events. */
void *pedk_stream(pedk_ctx *ctx);

/* ---- one sample (target or control) ---------------------------------- */

typedef struct pedk_ingest_info {
    uint64_t fastq_bytes;
    uint64_t records;         /* FASTQ records seen */
    uint64_t reads_kept;      /* sequence lines without 'N' (ped.pl:1436) */
    uint64_t strand_reads;    /* lines ped.pl writes to the TAG.tmp files (reads + reverse complements) */
    uint64_t unique_reads;    /* U: lines in sort_uniq/ after `sort | uniq` (ped.pl:1416) */
    uint64_t palindromes;     /* unique reads equal to their own reverse complement */
    int32_t clipping_used;    /* the "clipping : N" line of <s>.log (ped.pl:1495, 1526) */
    int32_t auto_clipped;     /* 1 if mixed lengths triggered ped.pl:1478-1524 */
    int32_t n_groups;         /* read-length groups held (1, or 2 after auto clipping) */
    int32_t pad;
} pedk_ingest_info;

typedef struct pedk_count_info {
    int32_t k;
    int32_t sort_passes;      /* 0 on the bucketed path (the instance stream is not sorted) */
    uint64_t instances;       /* canonical k-mer instances counted on this rank (half of the reference's M) */
    uint64_t canonical_kmers; /* rows of the canonical table */
    uint64_t correction_rows; /* windows of palindromic reads (usually 0) */
    int32_t count_passes;     /* > 1: the instances did not fit HBM at once and were counted in this many
                                 groups of hash bins (the reference spills to disk instead: sort -T, ped.pl:89) */
    int32_t pad;
} pedk_count_info;

int pedk_sample_create(pedk_ctx *ctx, pedk_sample **s);
int pedk_sample_destroy(pedk_sample *s);
/* Optional size hint for the FASTQ text buffer in HBM. */
int pedk_sample_reserve(pedk_sample *s, uint64_t fastq_bytes);
/* Append FASTQ text.  Chunks may split anywhere; the bytes of successive calls
 * are concatenated exactly like `cat file1 file2 |` (ped.pl:1383-1398).
 * _add_fastq: HOST memory (pinned gives full PCIe speed), copied asynchronously.
 * _add_fastq_device: memory already in HBM. */
int pedk_sample_add_fastq(pedk_sample *s, const void *host_bytes, uint64_t n);
int pedk_sample_add_fastq_device(pedk_sample *s, const void *device_bytes, uint64_t n);
/* The text is already in HBM and may be read IN PLACE (no copy): `capacity` bytes from
 * device_bytes must be addressable, at least n rounded up to 16 KiB plus 16 KiB; the bytes
 * beyond n are overwritten with zeros, and the buffer must stay untouched until
 * pedk_sample_ingest / _ingest_reference returns.  Only for a sample that holds no text yet. */
int pedk_sample_borrow_fastq_device(pedk_sample *s, void *device_bytes, uint64_t n, uint64_t capacity);
/* S0+S1 = mkSortUniq/sortUniqSub/complement/sortUniqSort (ped.pl:1365-1530,
 * 3464-3484): 4-line state machine, drop reads with N, clipping (0 = not
 * given -> fixed length, or ped.pl's automatic choice for mixed lengths),
 * strand closure, exact de-duplication.  Frees the FASTQ text. */
int pedk_sample_ingest(pedk_sample *s, int clipping, pedk_ingest_info *info);
/* a10 = mkChrFromFile + mkControlReadChr + mkControlReadSub (ped.pl:1027-1060, 1089-1136):
 * the control of a `ref=` run is the reference itself.  Feed the FASTA text with
 * pedk_sample_add_fastq (any chunking), then: every sequence is cleaned (upper case,
 * non-ACGT -> N, headers and newlines dropped; a sequence whose name repeats replaces the
 * earlier one, like ped.pl's open ">"), tiled into `window`-base windows every `step` bases
 * (0, 0 = ped.pl's 100 and 2), windows holding an N are dropped, the rest go through strand
 * closure and exact de-duplication.  The sample is then in the state pedk_sample_ingest
 * leaves it in (info->records = windows, reads_kept = windows without N). */
int pedk_sample_ingest_reference(pedk_sample *s, int window, int step, pedk_ingest_info *info);
/* The chromosome files mkChrFromFile writes: name is "chr" + the name ped.pl derives from the
 * header line, seq the cleaned sequence (NULL: length only).  *overwritten = 1 for a sequence
 * that a later one of the same name replaced. */
int pedk_sample_num_chromosomes(pedk_sample *s, int *n);
/* host-only helper (no CUDA call): "chr" + the name ped.pl:1045-1050 derives from a header line */
int pedk_chromosome_file_name(const char *header_line, char *out, uint64_t cap);
int pedk_sample_chromosome(pedk_sample *s, int i, char *name, uint64_t name_cap, uint64_t *length,
                           int *overwritten, char *seq, uint64_t seq_cap);
/* f1, mk20mer + mkUniq (ped.pl:1162-1211, 1138-1160) on a sample that went through
 * pedk_sample_ingest_reference: every k-base window without N of every chromosome (a chromosome
 * called NOP is skipped, ped.pl:1217) on both strands; the k-mers that occur on exactly one of those
 * lines stay.  *rows = lines of all ref20_uniq files together.  The reference hard-codes k = 20.
 * The reference must be shorter than 2^31 bases. */
int pedk_sample_ref20_uniq(pedk_sample *ref, int k, uint64_t *rows);
/* `zcat <ref>/ref20_uniq.TAG.gz`: lines `KMER\tCHR\tPOS\tf|r`, ascending.  Caller frees with pedk_free. */
int pedk_sample_ref20_uniq_text(pedk_sample *ref, int tag, char **text, uint64_t *len);
/* f1, mapKmerSub (ped.pl:641-685): snp_text = lines of <tmpdir>/<t>.snp.TAG (the edge list; any tags,
 * in key order).  Every edge whose control allele is one base is looked up as (k-1)-mer + that base
 * in the unique k-mer index; the result is the text of the <tmpdir>/<t>.map.TAG files, in the same
 * order.  Caller frees with pedk_free. */
int pedk_sample_map_kmer_text(pedk_sample *ref, const char *snp_text, uint64_t snp_len, char **text, uint64_t *len);
/* f3, cnvJoin + cnvJoinSub (ped.pl:934-989): the full outer join of the two count tables, for the
 * k-mers of the unique k-mer index of `ref`: lines `chr\tpos\tstrand\ttarget_count\tcontrol_count`
 * (a k-mer one sample does not have counts 0 there), ordered like `sort -k 1 -n -k 2 -n` -- the text
 * of <target>/<target>.<control>.cnv.  Collective on several ranks (the counts found in every rank's key
 * range of the rows are summed; every rank gets the whole text).  Caller frees with pedk_free. */
int pedk_cnv_text(pedk_sample *ref, pedk_sample *control, pedk_sample *target, char **text, uint64_t *len);
/* f2, snpMkT + sortSeq + joinTarget + snpMkC + sortSeq + joinControl + kmerReadCount (ped.pl:2087-2305,
 * 2576-2641, 2388-2500): map_text = the lines of the <tmpdir>/<t>.map.TAG files in tag order
 * (pedk_sample_map_kmer_text).  For every candidate `chr pos ref alt` of the map the L windows of the
 * reference around it and of the reference carrying the alternative allele are looked up among the
 * unique reads of the target (L = its read length) and of the control (its read length); the numbers of
 * windows found (cw cm tw tm) give the genotype M / H / R / N, and the map lines joined with them are
 * the text of <target>/<target>.kmer.snp of a run with a reference (20 columns).  target and control
 * must be ingested with their reads still in HBM; ref went through pedk_sample_ingest_reference (its
 * chromosomes are the `chrNAME` files ped.pl seeks in).  Orders are bytewise (LC_ALL=C).  On several
 * ranks the call is collective: every rank passes the WHOLE map (all ranks' lines) and its own share of
 * the two samples, looks every window up among its reads, and the counters are summed over the ranks;
 * every rank gets the whole text.  Caller frees with pedk_free. */
int pedk_verify_text(pedk_sample *target, pedk_sample *control, pedk_sample *ref, const char *map_text,
                     uint64_t map_len, char **text, uint64_t *len);
/* The two host-only halves of pedk_verify_text (no CUDA call; for tests and for hosts that bring their own
 * look-up): (1) one line per candidate of the map, `chr pos ref alt \t reference string \t mutant string`
 * (ped.pl:2153-2192 for a read length; both strings empty where ped.pl says `next`), the chromosome texts given
 * as n_chr (file name, sequence, length) triples; (2) the <target>.kmer.snp text from the numbers of windows
 * found, hits[2c] = reference string of candidate c, hits[2c+1] = its mutant string, candidates in the order of (1)
 * (kmerReadCount, ped.pl:2388-2500).  Caller frees with pedk_free. */
int pedk_verify_strings_text(const char *map_text, uint64_t map_len, const char *const *chr_names,
                             const char *const *chr_seqs, const uint64_t *chr_lens, int n_chr, int read_length,
                             char **text, uint64_t *len);
int pedk_verify_counts_text(const char *map_text, uint64_t map_len, const uint32_t *target_hits,
                            const uint32_t *control_hits, uint64_t n_strings, int length, int clength, char **text,
                            uint64_t *len);
/* f4, toVcf for method kmer (ped.pl:1532-1574): <target>.kmer.snp -> the text of <target>.kmer.vcf.
 * Host only (no CUDA call).  Caller frees with pedk_free. */
int pedk_vcf_text(const char *kmer_snp, uint64_t snp_len, const char *target_name, char **text, uint64_t *len);
/* S2-S4 = countKmerSub + countKmerMerge (ped.pl:891-932, 862-889): every k-mer
 * of every unique read, brought together (hash bins, buckets) and counted.  4 <= k <= 32 (the reference
 * hard-codes 20, ped.pl:912-913).  Leaves the canonical table in HBM. */
int pedk_sample_count(pedk_sample *s, int k, pedk_count_info *info);
/* The count table as ped.pl's count/<s>.count.*.gz hold it (both strands,
 * ascending; bucket TAG = the rows whose first three bases are TAG).  Call
 * with kmers == NULL to get *n only. */
int pedk_sample_table(pedk_sample *s, uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n);
/* Order-independent digest of the count table on this rank: *rows = its rows (both strands), *digest =
 * the sum over rows of mix64(kmer * 0x9E3779B97F4A7C15 + count) modulo 2^64 (mix64 = the splitmix64
 * finaliser).  Sums over ranks give the digest of the whole table; no sort and no copy is made, so it
 * checks a table of any size against a known answer (bench.py "parity"). */
int pedk_sample_table_digest(pedk_sample *s, uint64_t *rows, uint64_t *digest);
/* Unique reads of length group g as packed canonical representatives
 * (words_per_read u64 each, first base in the top bits of word 0) plus the
 * palindrome flag; the sort_uniq/ lines are these and their reverse
 * complements.  words == NULL: sizes only. */
int pedk_sample_reads(pedk_sample *s, int group, uint64_t *words, uint8_t *pal, uint64_t cap_reads,
                      uint64_t *n_reads, int32_t *read_len, int32_t *words_per_read);
/* Drop device state that later stages do not need (packed reads). */
int pedk_sample_release_reads(pedk_sample *s);

/* ---- target x control ------------------------------------------------- */

typedef struct pedk_edge {
    uint64_t prefix;       /* the (k-1)-mer */
    uint32_t control[4];   /* counts of prefix+A,C,G,T in the control (columns 4-7 of <t>.kmer.snp) */
    uint32_t target[4];    /* same for the target (columns 8-11) */
    uint8_t cont_mask;     /* bit b: base b is a control allele (column 2) */
    uint8_t alt_mask;      /* bit b: base b is a target allele (column 3) */
    uint8_t quirk;         /* 1: row went through ped.pl's first-row text path;
                              its exact bytes are only in pedk_result_text */
    uint8_t pad[5];
} pedk_edge;

/* S5+S6 = lbcKmerSub + joinKmerSub + the no-ref `cat` (ped.pl:833-860, 703-782,
 * 374-378).  Both samples must have been counted with the same k. */
int pedk_join(pedk_ctx *ctx, pedk_sample *control, pedk_sample *target, pedk_result **out);
uint64_t pedk_result_num_edges(const pedk_result *r);
/* The exact bytes of <target>/<target>.kmer.snp (no-ref mode, ped.pl:375). */
int pedk_result_text(const pedk_result *r, const char **text, uint64_t *len);
int pedk_result_edges(const pedk_result *r, pedk_edge *out, uint64_t cap, uint64_t *n);
/* The lbc table of one sample as text (`zcat <s>/lbc/<s>.lbc.TAG.gz`, with the
 * F7 first row), tag in [0,64).  Caller frees with pedk_free. */
int pedk_sample_lbc_text(pedk_sample *s, int tag, char **text, uint64_t *len);
int pedk_sample_count_text(pedk_sample *s, int tag, char **text, uint64_t *len);
int pedk_sample_sort_uniq_text(pedk_sample *s, int tag, char **text, uint64_t *len);
void pedk_free(void *p);
int pedk_result_destroy(pedk_result *r);

/* Whole hot path in one call: FASTQ text of both samples -> edge list.
 * where = 0: buffers are HOST memory; 1: DEVICE memory (copied); 2: DEVICE memory that is
 * addressable up to its size rounded up to 16 KiB plus 16 KiB and may be read in place
 * (pedk_sample_borrow_fastq_device: the padding is zeroed). */
int pedk_kmer_edges(pedk_ctx *ctx, const void *control_fastq, uint64_t control_bytes, const void *target_fastq,
                    uint64_t target_bytes, int where, int k, int clipping, pedk_result **out);

/* The same with the reference itself as the control (`ped.pl target=T,ref=R`: $control = $ref,
 * ped.pl:193): ref_fasta is the text of the reference FASTA, tiled like mkControlRead does
 * (pedk_sample_ingest_reference). */
int pedk_kmer_edges_ref(pedk_ctx *ctx, const void *ref_fasta, uint64_t ref_bytes, const void *target_fastq,
                        uint64_t target_bytes, int where, int k, int clipping, pedk_result **out);

/* ---- file level: what ped.pl calls ------------------- */

/* mkSortUniq (ped.pl:1365-1410): reads <wd>/<subject>/read/ (gz, bz2, xz and
 * *q files, grouped and ordered like ped.pl), writes
 * <wd>/<subject>/sort_uniq/<subject>.sort_uniq.TAG.gz.  *clipping_used is the
 * value of the "clipping :" log line; *auto_clipped is 1 when ped.pl would have
 * assigned it to its global $clipping (ped.pl:1483), which makes it the
 * explicit clipping of every later mkSortUniq call of the same run. */
int pedk_mk_sort_uniq(pedk_ctx *ctx, const char *wd, const char *subject, int clipping, int *clipping_used,
                      int *auto_clipped);
/* mkChrFromFile + mkControlRead (ped.pl:1027-1060, 1062-1136): reads <wd>/<ref>/<file> (FASTA;
 * .gz and .bz2 like ped.pl), writes <wd>/<ref>/chrNAME and
 * <wd>/<ref>/sort_uniq/<ref>.sort_uniq.TAG.gz.  The reference stays resident as a sample, so
 * pedk_count_kmer(wd, ref) continues from HBM. */
int pedk_mk_control_read(pedk_ctx *ctx, const char *wd, const char *ref, const char *file);
/* mk20 (ped.pl:1213-1249): writes <wd>/<ref>/ref20_uniq.TAG.gz.  The reference is taken from this
 * process's pedk_mk_control_read when it is still resident, else parsed again from <wd>/<ref>/<file>
 * (file may be NULL when it is resident).  k = 20 in the reference. */
int pedk_mk_ref20_uniq(pedk_ctx *ctx, const char *wd, const char *ref, const char *file, int k);
/* mapKmer (ped.pl:625-685): reads <tmpdir>/<target>.snp.TAG (pedk_join_kmer with have_ref) and the
 * unique k-mer index (resident, or <wd>/<ref>/ref20_uniq.TAG.gz), writes <tmpdir>/<target>.map.TAG */
int pedk_map_kmer(pedk_ctx *ctx, const char *wd, const char *target, const char *ref, const char *tmpdir);
/* cnvJoin (ped.pl:934-947) after both countKmer calls of `cnv` (ped.pl:362-366): writes
 * <wd>/<target>/<target>.<control>.cnv from the count tables (resident or count files) and the unique
 * k-mer index (resident or ref20_uniq files) */
int pedk_cnv(pedk_ctx *ctx, const char *wd, const char *target, const char *control, const char *ref, int k);
/* snpMkT, sortSeq, joinTarget, snpMkC, sortSeq, joinControl, kmerReadCount (ped.pl:380-386) in one call:
 * reads <tmpdir>/<target>.map.TAG (pedk_map_kmer), the sort_uniq files of target and control (or the
 * reads still resident from pedk_mk_sort_uniq / pedk_mk_control_read) and <wd>/<ref>/chrNAME, writes
 * <wd>/<target>/<target>.kmer.snp.  control may equal ref (ped.pl:193). */
int pedk_verify_kmer(pedk_ctx *ctx, const char *wd, const char *target, const char *control, const char *ref,
                     const char *tmpdir);
/* toVcf (ped.pl:387, 1532-1574): <wd>/<target>/<target>.kmer.snp -> <wd>/<target>/<target>.kmer.vcf */
int pedk_to_vcf(pedk_ctx *ctx, const char *wd, const char *target);
/* countKmer (ped.pl:784-815): writes <wd>/<subject>/count/<subject>.count.TAG.gz */
int pedk_count_kmer(pedk_ctx *ctx, const char *wd, const char *subject, int k);
/* lbcKmer (ped.pl:817-831): writes <wd>/<subject>/lbc/<subject>.lbc.TAG.gz */
int pedk_lbc_kmer(pedk_ctx *ctx, const char *wd, const char *subject, int k);
/* joinKmer (ped.pl:687-701) + ped.pl:374-378: writes <tmpdir>/<target>.snp.TAG when
 * have_ref != 0, else <wd>/<target>/<target>.kmer.snp */
int pedk_join_kmer(pedk_ctx *ctx, const char *wd, const char *target, const char *control, int k,
                   const char *tmpdir, int have_ref);

/* ---- statistics -------------------------------------------------------- */

enum {
    PEDK_ST_INGEST = 0,   /* newline scan, record index, pack */
    PEDK_ST_DEDUP = 1,
    PEDK_ST_EXTRACT = 2,  /* k-mer extraction + hash + level-A binning (one rank; EXCHANGE on several) */
    PEDK_ST_SORT = 3,     /* level-B histogram + binning (PEDK_COUNT_MODE=sort: radix passes over the instances) */
    PEDK_ST_COUNT = 4,    /* per-bucket tables -> strand rows (sort mode: run-length count) */
    PEDK_ST_EXPAND = 5,   /* (table emission only) strand rows copied for the sort */
    PEDK_ST_ROWSORT = 6,  /* join: rows -> join records -> bins -> buckets; tables: radix passes over the rows */
    PEDK_ST_EDGES = 7,    /* join: per-bucket tables + joinKmerSub */
    PEDK_ST_COPY = 8,     /* host<->device copies issued by the library */
    PEDK_ST_EXCHANGE = 9, /* multi-GPU all-to-all */
    PEDK_N_STAGES = 10
};
typedef struct pedk_stats {
    double stage_ms[PEDK_N_STAGES];       /* CUDA-event time on the context stream */
    uint64_t stage_launches[PEDK_N_STAGES];
    uint64_t kernel_launches;             /* all kernels of this library */
    uint64_t sort_pass_launches;          /* k_radix_pass launches on the instance stream */
    uint64_t sort_pass_keys;              /* keys moved by them (sum over launches) */
    double sort_pass_ms;                  /* their summed duration */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t peak_device_bytes;
    /* per kernel class (PEDK_KC_*): launches, summed duration and the ALGORITHMIC bytes
     * (what the kernel must read and write) of those launches */
    double kc_ms[16];
    uint64_t kc_bytes[16];
    uint64_t kc_launches[16];
} pedk_stats;
enum {
    PEDK_KC_RADIX_KEYS = 0,   /* k_radix_pass over 8-byte keys: 16 B/key */
    PEDK_KC_RADIX_PAIRS = 1,  /* k_radix_pass over (key, count) rows: 24 B/row */
    PEDK_KC_PART = 2,         /* k_part_scatter: packed reads in, hashed remainders out */
    PEDK_KC_SUB_HIST = 3,     /* k_sub_hist: one read of the instance stream */
    PEDK_KC_SUB_SCATTER = 4,  /* k_sub_scatter: one read + one write */
    PEDK_KC_BUCKET_COUNT = 5, /* k_bucket_count: one read + 12 B per row */
    PEDK_KC_PACK = 6,         /* k_pack_fixed: FASTQ text in, packed reads out */
    PEDK_KC_PART_COUNT = 7,   /* k_part_count: packed reads in */
    PEDK_KC_JOIN_COUNT = 8,   /* k_rows_count: the keys of the strand rows in (8 B/row) */
    PEDK_KC_JOIN_SCATTER = 9, /* k_rows_scatter: rows in (12 B), join records out (8 B) */
    PEDK_KC_JOIN_SUB = 10,    /* k_sub_hist + k_sub_scatter over the join records: 2 reads + 1 write of 8 B */
    PEDK_KC_JOIN_TABLE = 11,  /* k_join_bucket: join records in (8 B/row) */
    PEDK_N_KC = 12
};
int pedk_get_stats(pedk_ctx *ctx, pedk_stats *out);   /* synchronises the stream */
/* PEDK_GUARD=1 in the environment puts a guard zone behind every device block of the context and checks
 * it when the block is released: the number of blocks a kernel wrote past the end of (0 = clean). */
uint64_t pedk_guard_hits(pedk_ctx *ctx);
int pedk_reset_stats(pedk_ctx *ctx);

/* ---- multi-GPU (one rank per process) ---------------------------------- */

#define PEDK_COMM_ID_BYTES 128
/* rank 0 makes the id, the launcher (torchrun store, MPI, a file) hands it to
 * the other ranks, every rank calls pedk_comm_init. */
int pedk_comm_unique_id(pedk_ctx *ctx, void *id_bytes);
int pedk_comm_init(pedk_ctx *ctx, const void *id_bytes, int rank, int nranks);
/* Owner functions every rank evaluates identically (host copies of what the kernels use):
 * canonical packed reads go to a rank by hash, canonical k-mers by the top bits of
 * pedk_kmer_hash (contiguous ranges of the level-A bins); table rows go to the
 * rank that holds the 3-nt bucket of the k-mer, buckets being split into contiguous ranges
 * balanced on a 64-bin histogram. */
int pedk_owner_of_read(const uint64_t *words, int words_per_read, int nranks);
int pedk_owner_of_kmer(uint64_t canonical_kmer, int k, int nranks);
/* The bijection on 2k-bit integers whose top bits bin a canonical k-mer (level-A bin = the
 * top min(8,k) bits, owner rank = bin * nranks >> bits) and its inverse. */
uint64_t pedk_kmer_hash(uint64_t kmer, int k);
uint64_t pedk_kmer_unhash(uint64_t hash, int k);
int pedk_plan_bucket_owners(const uint64_t *hist64, int nranks, uint8_t *owner64);
/* host-only (no CUDA call): the plan of the k-mer exchange for rank `rank` of `nranks` -- where this
 * rank's run of every level-A bin starts (cur256) and, when sized optimistically, ends (lim256)
 * inside the OWNER's buffer; the segments of this rank's own bins (seg = n_seg starts, n_seg ends,
 * n_seg bins; room for 3 x 256 x nranks values); the elements its receive buffer needs.
 * exact == 0: every (bin, source) segment holds m_max / bins + 1/8 + 4096 elements (counts, what
 * every rank stored, may be NULL: no segments then); exact != 0: sizes from counts.
 * counts[src * 257 + bin]. */
int pedk_plan_level_a(int k, int nranks, int rank, uint64_t m_max, const uint64_t *counts, int exact,
                      uint64_t *cur256, uint64_t *lim256, uint64_t *seg, int *n_seg, uint64_t *buffer_elems,
                      uint64_t *segment_capacity);

/* the same for pass `pass` of n_pass (a power of two): a sample whose instances do not fit HBM at once is
 * counted in n_pass groups of bins, bin % n_pass == pass; the other bins get no room and no segment */
int pedk_plan_level_a_pass(int k, int nranks, int rank, uint64_t m_max, const uint64_t *counts, int exact,
                           int n_pass, int pass, uint64_t *cur256, uint64_t *lim256, uint64_t *seg, int *n_seg,
                           uint64_t *buffer_elems, uint64_t *segment_capacity);

/* ---- synthetic reads (tests and bench.py) ---------------- */

#ifndef PEDK_SYNTH_PARAMS_DEFINED
#define PEDK_SYNTH_PARAMS_DEFINED
typedef struct pedk_synth_params {
    uint64_t seed;
    uint64_t genome_len;
    uint32_t read_len;
    uint32_t err_ppm;   /* substitution errors per million bases */
    uint32_t n_ppm;     /* reads with one N per million reads */
    uint32_t dup_ppm;   /* reads re-using the placement of the previous read, per million */
} pedk_synth_params;
#endif

/* The bytes libpedk_synth.so's host generator writes (include/pedk_synth.h), generated in HBM:
 * genome_dev and out_dev are device pointers. */
int pedk_synth_fastq_device(pedk_ctx *ctx, const pedk_synth_params *p, const char *genome_dev,
                            uint64_t first_read, uint64_t n_reads, char *out_dev);

/* host-only (no CUDA call): pass 1 of joinKmerSub (ped.pl:720-744) as the truth table the edge
 * kernel uses.  A (k-1)-mer group is coded in 16 bits, two per slot (control A,C,G,T = bits
 * 0-7, target = bits 8-15): 0 = count <= 1 or absent, 1 = count 2, 2 = 3..100, 3 = > 100.
 * bits2048[code >> 5] >> (code & 31) & 1 says whether the group passes. */
int pedk_edge_pass1_table(uint32_t *bits2048);

/* ---- raw kernels, exposed for unit tests (device pointers) ------------- */
int pedk_test_sort_u64(pedk_ctx *ctx, uint64_t *keys_dev, uint64_t n, int key_bits);
int pedk_test_sort_pairs(pedk_ctx *ctx, uint64_t *keys_dev, uint32_t *vals_dev, uint64_t n, int key_bits);
int pedk_test_rle(pedk_ctx *ctx, const uint64_t *sorted_dev, uint64_t n, uint64_t *ukey_dev, uint32_t *count_dev,
                  uint64_t *n_runs);

#ifdef __cplusplus
}
#endif
#endif /* PEDK_H */
