// S0: FASTQ text on the device -> packed canonical reads.
//
// Replaces the serial Perl loop of sortUniqSub (ped.pl:1420-1475) and the
// char-by-char `complement` (ped.pl:3464-3484).  The raw file bytes are shipped
// to HBM untouched; record structure is recovered with a newline count + scan,
// then one warp per read packs 32 bases per 64-bit word with ballots.
#include "kernels.h"
#include "util.cuh"

namespace pedk {

namespace {

constexpr int NL_THREADS = 256;
constexpr int NL_BYTES_PER_THREAD = INGEST_TILE_BYTES / NL_THREADS;  // 64
static_assert(NL_BYTES_PER_THREAD == 64, "tile layout");

__device__ __forceinline__ unsigned count_nl16(uint4 v) {
    // __vcmpeq4 gives 0xff per equal byte
    return (__popc(__vcmpeq4(v.x, 0x0A0A0A0Au)) + __popc(__vcmpeq4(v.y, 0x0A0A0A0Au)) +
            __popc(__vcmpeq4(v.z, 0x0A0A0A0Au)) + __popc(__vcmpeq4(v.w, 0x0A0A0A0Au))) >>
           3;
}

__global__ void __launch_bounds__(NL_THREADS) k_newline_count(const uint4* __restrict__ fq, size_t n_tiles,
                                                             uint32_t* __restrict__ tile_counts) {
    __shared__ unsigned warp_sums[NL_THREADS / 32];
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint4* base = fq + tile * (INGEST_TILE_BYTES / 16);
        unsigned c = 0;
#pragma unroll
        for (int j = 0; j < NL_BYTES_PER_THREAD / 16; ++j) c += count_nl16(__ldg(base + j * NL_THREADS + threadIdx.x));
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane_id() == 0) warp_sums[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned s = 0;
#pragma unroll
            for (int w = 0; w < NL_THREADS / 32; ++w) s += warp_sums[w];
            tile_counts[tile] = s;
        }
        __syncthreads();
    }
}

// newline flags of 16 bytes as a 16-bit mask (bit b = byte b)
__device__ __forceinline__ unsigned nl_mask4(uint32_t x) {
    // 0xff per matching byte -> one bit per byte, gathered into bits 24..27 by one multiply
    return ((__vcmpeq4(x, 0x0A0A0A0Au) & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ unsigned nl_mask16(uint4 v) {
    return nl_mask4(v.x) | (nl_mask4(v.y) << 4) | (nl_mask4(v.z) << 8) | (nl_mask4(v.w) << 12);
}

// Line numbers -> byte spans of the sequence lines.  A warp owns 2 KiB of the tile as four
// 512-byte rows (lane l reads 16 bytes of each: coalesced); newline counts are ranked with
// warp shuffles, so a lane knows the line number of each of its newlines and writes the
// span ends straight away.  One read of the text, no per-byte branches.
__global__ void __launch_bounds__(NL_THREADS) k_line_index(const uint8_t* __restrict__ fq, size_t n, size_t n_tiles,
                                                          const uint64_t* __restrict__ tile_prefix,
                                                          uint64_t* __restrict__ seq_start,
                                                          uint64_t* __restrict__ seq_end, uint64_t n_records,
                                                          int plain_lines) {
    constexpr int WARPS = NL_THREADS / 32;
    constexpr int ROWS = INGEST_TILE_BYTES / (WARPS * 512);  // 4
    __shared__ unsigned warp_cnt[WARPS];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t warp_off = tile * INGEST_TILE_BYTES + (size_t)warp * (ROWS * 512);
        unsigned mask[ROWS], ex[ROWS];
        unsigned row_base = 0;  // newlines of this warp before row j
#pragma unroll
        for (int j = 0; j < ROWS; ++j)
            mask[j] = nl_mask16(__ldg(reinterpret_cast<const uint4*>(fq + warp_off + (size_t)j * 512) + lane));
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const unsigned c = __popc(mask[j]);
            unsigned inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= (unsigned)d) inc += o;
            }
            ex[j] = row_base + inc - c;
            row_base += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncthreads();  // warp_cnt of the previous tile has been read
        if (lane == 0) warp_cnt[warp] = row_base;
        __syncthreads();
        uint64_t line0 = tile_prefix[tile];
#pragma unroll
        for (int w = 0; w < WARPS; ++w)
            if (w < (int)warp) line0 += warp_cnt[w];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            uint64_t line = line0 + ex[j];
            unsigned m = mask[j];
            const size_t off = warp_off + (size_t)j * 512 + (size_t)lane * 16;
            while (m) {
                const size_t pos = off + (__ffs(m) - 1);
                m &= m - 1;
                if (pos < n) {
                    if (plain_lines) {
                        if (line < n_records) seq_end[line] = pos;
                        if (line + 1 < n_records) seq_start[line + 1] = pos + 1;
                    } else {
                        const uint64_t rec = line >> 2;
                        const unsigned phase = (unsigned)(line & 3);
                        if (rec < n_records) {
                            if (phase == 0) seq_start[rec] = pos + 1;  // end of the header line
                            if (phase == 1) seq_end[rec] = pos;        // end of the sequence line
                        }
                    }
                }
                ++line;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_read_scan(const uint8_t* __restrict__ fq,
                                                  const uint64_t* __restrict__ seq_start,
                                                  const uint64_t* __restrict__ seq_end, uint64_t n_records,
                                                  uint32_t* __restrict__ rlen, uint8_t* __restrict__ rflag,
                                                  unsigned long long* __restrict__ len_hist,
                                                  unsigned long long* __restrict__ counters) {
    unsigned lane = lane_id();
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    // each warp takes a contiguous span of reads so that equal lengths aggregate
    uint64_t per = (n_records + n_warps - 1) / n_warps;
    uint64_t lo = warp * per, hi = lo + per < n_records ? lo + per : n_records;
    uint32_t run_len = 0;
    unsigned long long run_cnt = 0, kept = 0, bad_reads = 0;
    for (uint64_t r = lo; r < hi; ++r) {
        uint64_t s = seq_start[r], e = seq_end[r];
        uint32_t len = (uint32_t)(e - s);
        bool has_n = false, bad = false;
        for (uint32_t i = lane; i < len; i += 32) {
            uint32_t c = fq[s + i];
            has_n |= (c == 'N');
            bad |= !(is_acgt(c) || c == 'N');
        }
        has_n = __any_sync(0xffffffffu, has_n);
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) {
            rlen[r] = len;
            rflag[r] = (uint8_t)((has_n ? 1 : 0) | (bad ? 2 : 0));
            if (!has_n) {
                ++kept;
                if (bad) ++bad_reads;
                uint32_t bin = len < LEN_HIST_BINS - 1 ? len : LEN_HIST_BINS - 1;
                if (bin == run_len) {
                    ++run_cnt;
                } else {
                    if (run_cnt) atomicAdd(&len_hist[run_len], run_cnt);
                    run_len = bin;
                    run_cnt = 1;
                }
            }
        }
    }
    if (lane == 0) {
        if (run_cnt) atomicAdd(&len_hist[run_len], run_cnt);
        if (kept) atomicAdd(&counters[0], kept);
        if (bad_reads) atomicAdd(&counters[1], bad_reads);
    }
}

__global__ void __launch_bounds__(256) k_chunk_count(const uint32_t* __restrict__ rlen,
                                                    const uint8_t* __restrict__ rflag, uint64_t n_records,
                                                    uint64_t r_limit, uint32_t clip, int whole,
                                                    uint32_t* __restrict__ chunks) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records) return;
    uint32_t c = 0;
    if (r < r_limit && rflag[r] == 0) {
        uint32_t len = rlen[r];
        c = whole ? (len == clip ? 1u : 0u) : len / clip;
    }
    chunks[r] = c;
}

// spread the 32 bits of x to the even bit positions of a 64-bit word
__device__ __forceinline__ uint64_t spread32(uint32_t x) {
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// One warp packs the L bases at src into W 64-bit words (2 bits per base, base 0 in the top bits of word 0),
// builds the reverse complement and keeps the smaller of the two: lane w returns word w of the canonical
// read.  *pal: the read equals its reverse complement.  CHECK: *acgt = every base is one of ACGT (upper case);
// the reads of a sample were filtered before they get here, the windows of the verification were not.
template <bool CHECK>
__device__ __forceinline__ uint64_t warp_pack_canonical(const uint8_t* __restrict__ src, int L, int W, unsigned lane,
                                                        bool* pal, bool* acgt) {
    // lane w ends up holding word w of the forward read
    uint64_t mine = 0;
    unsigned bad = 0;
    for (int w = 0; w < W; ++w) {
        int i = 32 * w + (int)lane;
        uint32_t c = i < L ? src[i] : (uint32_t)'A';
        uint32_t code = i < L ? base_code(c) : 0u;
        if (CHECK) bad |= __ballot_sync(0xffffffffu, !(c == 'A' || c == 'C' || c == 'G' || c == 'T'));
        // ballot bit t <-> base 32w+t; brev puts base 0 at bit 31
        uint32_t m1 = __brev(__ballot_sync(0xffffffffu, code & 2u));
        uint32_t m0 = __brev(__ballot_sync(0xffffffffu, code & 1u));
        uint64_t word = (spread32(m1) << 1) | spread32(m0);
        if ((int)lane == w) mine = word;
    }
    // reverse complement: lane w needs 32 bases starting at L-32w-32
    int sbase = L - 32 * (int)lane - 32;
    int wi = sbase >> 5, o = sbase & 31;
    uint64_t hi = __shfl_sync(0xffffffffu, mine, wi & 31);
    uint64_t lo = __shfl_sync(0xffffffffu, mine, (wi + 1) & 31);
    if (wi < 0 || wi >= W) hi = 0;
    if (wi + 1 < 0 || wi + 1 >= W) lo = 0;
    uint64_t win = o ? ((hi << (2 * o)) | (lo >> (64 - 2 * o))) : hi;
    uint64_t rc = (int)lane < W ? (revcomp_groups64(win) & tail_mask(L, (int)lane)) : 0ull;
    unsigned diff = __ballot_sync(0xffffffffu, mine != rc);
    bool use_rc = false;
    if (diff) {
        int f = __ffs(diff) - 1;
        use_rc = __shfl_sync(0xffffffffu, rc < mine ? 1 : 0, f) != 0;
    }
    *pal = diff == 0;
    if (CHECK) *acgt = bad == 0;
    return use_rc ? rc : mine;
}

__global__ void __launch_bounds__(256) k_pack(const uint8_t* __restrict__ fq, const uint64_t* __restrict__ seq_start,
                                             const uint32_t* __restrict__ chunks,
                                             const uint64_t* __restrict__ chunk_base, uint64_t n_records,
                                             uint32_t clip, int W, uint64_t* __restrict__ words,
                                             uint8_t* __restrict__ pal) {
    unsigned lane = lane_id();
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int L = (int)clip;
    for (uint64_t r = warp; r < n_records; r += n_warps) {
        uint32_t nch = chunks[r];
        if (nch == 0) continue;
        uint64_t s0 = seq_start[r];
        uint64_t slot0 = chunk_base[r];
        for (uint32_t j = 0; j < nch; ++j) {
            bool is_pal, unused;
            uint64_t word = warp_pack_canonical<false>(fq + s0 + (uint64_t)j * clip, L, W, lane, &is_pal, &unused);
            uint64_t slot = slot0 + j;
            if ((int)lane < W) words[slot * (uint64_t)W + lane] = word;
            if (lane == 0) pal[slot] = is_pal ? 1 : 0;
        }
    }
}

// f2 (ped.pl:2189-2196, 2601-2620): is the window of L bases at text[start[s] + i] one of the sample's reads?
// One warp per window; the window is packed and made canonical like a read, hashed like k_dedup hashes it,
// and looked up in the table k_dedup built over the sample's unique canonical reads (open addressing, the
// packed words decide).  hits[s] counts the windows of string s that are reads.
__global__ void __launch_bounds__(256) k_probe_windows(const uint8_t* __restrict__ text,
                                                      const uint64_t* __restrict__ start,
                                                      const uint32_t* __restrict__ length, uint64_t n_strings,
                                                      int L, int W, const uint64_t* __restrict__ words,
                                                      const uint32_t* __restrict__ table, uint32_t mask,
                                                      uint32_t* __restrict__ hits) {
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_windows = n_strings * (uint64_t)L;
    for (uint64_t p = warp; p < n_windows; p += n_warps) {
        const uint64_t s = p / (unsigned)L;
        const uint32_t i = (uint32_t)(p % (unsigned)L);
        if ((uint64_t)i + (unsigned)L > length[s]) continue;  // substr past the end: a shorter line, never a read
        bool is_pal, acgt;
        const uint64_t word = warp_pack_canonical<true>(text + start[s] + i, L, W, lane, &is_pal, &acgt);
        if (!acgt) continue;
        uint64_t h = 0x243F6A8885A308D3ull;  // the hash of k_dedup (dedup.cu)
        for (int w = 0; w < W; ++w) h = mix64(h ^ __shfl_sync(0xffffffffu, word, w));
        uint32_t slot = (uint32_t)(h >> 20) & mask;
        bool hit = false;
        for (;;) {
            const uint32_t cur = table[slot];
            if (cur == 0xFFFFFFFFu) break;
            const uint64_t other = (int)lane < W ? words[(uint64_t)cur * W + lane] : 0ull;
            if (__ballot_sync(0xffffffffu, (int)lane < W && other != word) == 0) {
                hit = true;
                break;
            }
            slot = (slot + 1) & mask;
        }
        if (hit && lane == 0) atomicAdd(&hits[s], 1u);
    }
}

// min / max of the sequence-line lengths of all records (no byte of the text is read)
__global__ void __launch_bounds__(256) k_len_stats(const uint64_t* __restrict__ seq_start,
                                                  const uint64_t* __restrict__ seq_end, uint64_t n_records,
                                                  unsigned* __restrict__ minmax) {
    unsigned lo = 0xffffffffu, hi = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_records;
         r += (uint64_t)gridDim.x * blockDim.x) {
        unsigned len = (unsigned)(seq_end[r] - seq_start[r]);
        lo = len < lo ? len : lo;
        hi = len > hi ? len : hi;
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane_id() == 0) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}

// 0xff in every byte of x that is one of A, C, G, T
__device__ __forceinline__ uint32_t acgt_bytes(uint32_t x) {
    return __vcmpeq4(x, 0x41414141u) | __vcmpeq4(x, 0x43434343u) | __vcmpeq4(x, 0x47474747u) |
           __vcmpeq4(x, 0x54545454u);
}

// Fixed-length fast path (all records carry L bases, no clipping=): the N test, the alphabet
// test, the 2-bit pack and the canonical choice in one pass.  GS lanes share a read, lane g
// builds word g from 32 bytes (nine aligned 32-bit loads + funnel shifts; four bases become
// eight bits with one multiply).  Slot = record index; dropped reads get valid = 0.
template <int GS>
__global__ void __launch_bounds__(256) k_pack_fixed(const uint8_t* __restrict__ fq,
                                                   const uint64_t* __restrict__ seq_start, uint64_t n_records,
                                                   int L, int W, uint64_t* __restrict__ words,
                                                   uint8_t* __restrict__ pal, uint8_t* __restrict__ valid,
                                                   unsigned long long* __restrict__ counters) {
    constexpr int RPW = 32 / GS;  // reads per warp
    const unsigned lane = lane_id();
    const unsigned g = lane % GS, grp = lane / GS, gbase = grp * GS;
    const unsigned gmask = (GS == 32 ? 0xffffffffu : ((1u << GS) - 1u));
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long kept = 0, bad_reads = 0;
    // byte masks of the eight four-byte groups of this lane's word: what lies beyond the read
    // (the newline, the '+' line) is replaced by 'A' (code 0).  Fixed for the whole launch.
    uint32_t bm[8];
    {
        const int nvalid = L - 32 * (int)g;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int nb = nvalid - 4 * i;
            bm[i] = nb >= 4 ? 0xffffffffu : nb <= 0 ? 0u : (0xffffffffu >> (32 - 8 * nb));
        }
    }
    for (uint64_t r0 = warp * RPW; r0 < n_records; r0 += n_warps * RPW) {
        const uint64_t r = r0 + grp;
        const bool live = r < n_records;
        uint64_t mine = 0;
        bool has_n = false, bad = false;
        if (live && (int)g < W) {
            const uint64_t addr = seq_start[r] + 32ull * g;
            const uint32_t* p = reinterpret_cast<const uint32_t*>(fq + (addr & ~3ull));
            const unsigned sh = (unsigned)(addr & 3) * 8;
            uint32_t u[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) u[i] = __ldg(p + i);
            // the alu pipe bounds this kernel: the alphabet test is "does the 2-bit code map back to
            // the byte" (two permutes through the table "ACGT"), settled exactly only for the rare
            // word that fails it
            uint32_t nonacgt = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t x = __funnelshift_r(u[i], u[i + 1], sh);
                x = (x & bm[i]) | (0x41414141u & ~bm[i]);
                const uint32_t c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
                const uint32_t t = c | (c >> 4);
                const uint32_t sel = __byte_perm(t, 0u, 0x4420u);  // the four codes as selector nibbles
                nonacgt |= __byte_perm(0x54474341u, 0u, sel) ^ x;
                mine |= (uint64_t)((c * 0x40100401u) >> 24) << (56 - 8 * i);
            }
            if (nonacgt) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t x = __funnelshift_r(u[i], u[i + 1], sh);
                    x = (x & bm[i]) | (0x41414141u & ~bm[i]);
                    has_n |= __vcmpeq4(x, 0x4E4E4E4Eu) != 0;
                    bad |= (acgt_bytes(x) | __vcmpeq4(x, 0x4E4E4E4Eu)) != 0xffffffffu;
                }
            }
        }
        // flags of the whole read
        const unsigned any_n = (__ballot_sync(0xffffffffu, has_n) >> gbase) & gmask;
        const unsigned any_bad = (__ballot_sync(0xffffffffu, bad) >> gbase) & gmask;
        // reverse complement: lane g needs 32 bases starting at L-32g-32
        const int sbase = L - 32 * (int)g - 32;
        const int wi = sbase >> 5, o = sbase & 31;
        uint64_t hi = __shfl_sync(0xffffffffu, mine, gbase + ((unsigned)wi & (GS - 1)));
        uint64_t lo = __shfl_sync(0xffffffffu, mine, gbase + ((unsigned)(wi + 1) & (GS - 1)));
        if (wi < 0 || wi >= W) hi = 0;
        if (wi + 1 < 0 || wi + 1 >= W) lo = 0;
        const uint64_t win = o ? ((hi << (2 * o)) | (lo >> (64 - 2 * o))) : hi;
        const uint64_t rc = (int)g < W ? (revcomp_groups64(win) & tail_mask(L, (int)g)) : 0ull;
        const unsigned diff = (__ballot_sync(0xffffffffu, mine != rc) >> gbase) & gmask;
        // groups of one warp differ in `diff`: the shuffle must not sit under a divergent branch
        const int f = diff ? __ffs(diff) - 1 : 0;
        const bool use_rc = (__shfl_sync(0xffffffffu, rc < mine ? 1 : 0, gbase + f) != 0) && diff != 0;
        if (live) {
            const bool keep = !any_n && !any_bad;
            if ((int)g < W) words[r * (uint64_t)W + g] = use_rc ? rc : mine;
            if (g == 0) {
                pal[r] = diff == 0 ? 1 : 0;
                valid[r] = keep ? 1 : 0;
                if (!any_n) {
                    ++kept;
                    if (any_bad) ++bad_reads;
                }
            }
        }
    }
    kept = __reduce_add_sync(0xffffffffu, (unsigned)kept);
    bad_reads = __reduce_add_sync(0xffffffffu, (unsigned)bad_reads);
    if (lane == 0) {
        if (kept) atomicAdd(&counters[0], kept);
        if (bad_reads) atomicAdd(&counters[1], bad_reads);
    }
}

}  // namespace

void launch_newline_count(const uint8_t* fq, size_t n, uint32_t* tile_counts, cudaStream_t st) {
    size_t n_tiles = (n + INGEST_TILE_BYTES - 1) / INGEST_TILE_BYTES;
    if (!n_tiles) return;
    unsigned grid = (unsigned)(n_tiles < (size_t)kNumSMs * 8 ? n_tiles : (size_t)kNumSMs * 8);
    k_newline_count<<<grid, NL_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(fq), n_tiles, tile_counts);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_line_index(const uint8_t* fq, size_t n, const uint64_t* tile_prefix, uint64_t* seq_start,
                       uint64_t* seq_end, uint64_t n_records, int plain_lines, cudaStream_t st) {
    size_t n_tiles = (n + INGEST_TILE_BYTES - 1) / INGEST_TILE_BYTES;
    if (!n_tiles) return;
    unsigned grid = (unsigned)(n_tiles < (size_t)kNumSMs * 8 ? n_tiles : (size_t)kNumSMs * 8);
    k_line_index<<<grid, NL_THREADS, 0, st>>>(fq, n, n_tiles, tile_prefix, seq_start, seq_end, n_records, plain_lines);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_read_scan(const uint8_t* fq, const uint64_t* seq_start, const uint64_t* seq_end, uint64_t n_records,
                      uint32_t* rlen, uint8_t* rflag, unsigned long long* len_hist,
                      unsigned long long* counters, cudaStream_t st) {
    if (!n_records) return;
    uint64_t warps_wanted = (n_records + 15) / 16;  // >= 16 reads per warp when there are many
    uint64_t blocks = (warps_wanted + 7) / 8;
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_read_scan<<<grid, 256, 0, st>>>(fq, seq_start, seq_end, n_records, rlen, rflag, len_hist, counters);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_chunk_count(const uint32_t* rlen, const uint8_t* rflag, uint64_t n_records, uint64_t r_limit,
                        uint32_t clip, int whole, uint32_t* chunks, cudaStream_t st) {
    if (!n_records) return;
    unsigned grid = (unsigned)((n_records + 255) / 256);
    k_chunk_count<<<grid, 256, 0, st>>>(rlen, rflag, n_records, r_limit, clip, whole, chunks);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_pack(const uint8_t* fq, const uint64_t* seq_start, const uint32_t* chunks, const uint64_t* chunk_base,
                 uint64_t n_records, uint32_t clip, int W, uint64_t* words, uint8_t* pal, cudaStream_t st) {
    if (!n_records) return;
    uint64_t blocks = (n_records + 7) / 8;
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_pack<<<grid, 256, 0, st>>>(fq, seq_start, chunks, chunk_base, n_records, clip, W, words, pal);
    PEDK_CUDA_TRY(cudaGetLastError());
}

// ---- f2, second form: one warp per STRING, its windows as a rolling view of the packed string ----
//
// The L windows of a string overlap in all but one base, so the string is packed ONCE (lane w holds word w
// of the first 2L-1 bases, and word w of their reverse complement) and window i is two funnel shifts of two
// shuffled words per strand: 64 bits at base offset i of the forward words, and at offset n-L-i of the
// reverse-complement words.  The smaller of the two is the canonical window.  The read table of this form is
// its own (k_read_table): the hash is a sum over the words, so every lane hashes its word and five shuffles
// add them up -- k_dedup's chained hash would put the W words through one lane.
__device__ __forceinline__ uint64_t word_hash(uint64_t word, int w) {
    return mix64(word ^ (0x9E3779B97F4A7C15ull * (uint64_t)(w + 1)));
}

__global__ void __launch_bounds__(256) k_read_table(const uint64_t* __restrict__ words, int W, uint32_t n,
                                                   uint32_t* __restrict__ table, uint32_t mask) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t h = 0;
    for (int w = 0; w < W; ++w) h += word_hash(words[(uint64_t)i * W + w], w);
    uint32_t slot = (uint32_t)(mix64(h) >> 20) & mask;
    // the reads are unique already: the first free slot is this read's
    while (atomicCAS(&table[slot], 0xFFFFFFFFu, i) != 0xFFFFFFFFu) slot = (slot + 1) & mask;
}

// 64 bits of a packed string starting at base `at` (0 <= at): lane w gets the bits for its word w
__device__ __forceinline__ uint64_t rolling_word(uint64_t mine, int at, unsigned lane) {
    const int q = (at >> 5) + (int)lane, o = at & 31;
    uint64_t hi = __shfl_sync(0xffffffffu, mine, q & 31);
    uint64_t lo = __shfl_sync(0xffffffffu, mine, (q + 1) & 31);
    if (q > 31) hi = 0;
    if (q + 1 > 31) lo = 0;
    return o ? ((hi << (2 * o)) | (lo >> (64 - 2 * o))) : hi;
}

__global__ void __launch_bounds__(256) k_probe_strings(const uint8_t* __restrict__ text,
                                                      const uint64_t* __restrict__ start,
                                                      const uint32_t* __restrict__ length, uint64_t n_strings,
                                                      int L, int W, const uint64_t* __restrict__ words,
                                                      const uint32_t* __restrict__ table, uint32_t mask,
                                                      uint32_t* __restrict__ hits) {
    const unsigned lane = lane_id();
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t s = warp; s < n_strings; s += n_warps) {
        const int len = (int)length[s];
        if (len < L) {  // substr past the end: shorter lines, never a read
            if (lane == 0) hits[s] = 0;
            continue;
        }
        const int n = len < 2 * L - 1 ? len : 2 * L - 1;  // the windows 0 .. L-1 only see these bases
        const int last = n - L;                           // last window that is whole
        const uint8_t* src = text + start[s];
        // ---- pack: lane w keeps word w of the string and the not-ACGT flags of its 32 bases ----
        uint64_t fwd = 0;
        uint32_t bad = 0;
        const int Wn = (n + 31) >> 5;
        for (int w = 0; w < Wn; ++w) {
            const int i = 32 * w + (int)lane;
            const uint32_t c = i < n ? src[i] : (uint32_t)'A';
            const uint32_t code = i < n ? base_code(c) : 0u;
            const uint32_t b = __ballot_sync(0xffffffffu, !(c == 'A' || c == 'C' || c == 'G' || c == 'T'));
            const uint32_t m1 = __brev(__ballot_sync(0xffffffffu, code & 2u));
            const uint32_t m0 = __brev(__ballot_sync(0xffffffffu, code & 1u));
            if ((int)lane == w) {
                fwd = (spread32(m1) << 1) | spread32(m0);
                bad = b;  // bit t <-> base 32w + t
            }
        }
        const bool any_bad = __ballot_sync(0xffffffffu, bad != 0) != 0;
        // ---- its reverse complement: word w = revcomp of the 32 bases that end at n - 32w ----
        uint64_t rc;
        {
            const int sbase = n - 32 * (int)lane - 32;
            const int wi = sbase >> 5, o = sbase & 31;
            uint64_t hi = __shfl_sync(0xffffffffu, fwd, wi & 31);
            uint64_t lo = __shfl_sync(0xffffffffu, fwd, (wi + 1) & 31);
            if (wi < 0 || wi > 31) hi = 0;
            if (wi + 1 < 0 || wi + 1 > 31) lo = 0;
            const uint64_t win = o ? ((hi << (2 * o)) | (lo >> (64 - 2 * o))) : hi;
            rc = (int)lane < Wn ? (revcomp_groups64(win) & tail_mask(n, (int)lane)) : 0ull;
        }
        const uint64_t keep = tail_mask(L, (int)lane);  // 0 for lanes >= W
        uint32_t found = 0;
        for (int i = 0; i <= last; ++i) {
            if (any_bad) {
                // bases [i, i+L) against this lane's 32: [32 lane, 32 lane + 32)
                const int a = max(i, 32 * (int)lane), b = min(i + L, 32 * (int)lane + 32);
                uint32_t m = 0;
                if (a < b) m = (b - a == 32 ? 0xffffffffu : ((1u << (b - a)) - 1u)) << (a - 32 * (int)lane);
                if (__ballot_sync(0xffffffffu, (bad & m) != 0)) continue;
            }
            const uint64_t f = rolling_word(fwd, i, lane) & keep;
            const uint64_t r = rolling_word(rc, n - L - i, lane) & keep;
            const unsigned diff = __ballot_sync(0xffffffffu, f != r);
            bool use_rc = false;
            if (diff) use_rc = __shfl_sync(0xffffffffu, r < f ? 1 : 0, __ffs(diff) - 1) != 0;
            const uint64_t word = use_rc ? r : f;
            uint64_t h = (int)lane < W ? word_hash(word, (int)lane) : 0ull;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
            uint32_t slot = (uint32_t)(mix64(h) >> 20) & mask;
            for (;;) {
                const uint32_t cur = table[slot];
                if (cur == 0xFFFFFFFFu) break;
                const uint64_t other = (int)lane < W ? words[(uint64_t)cur * W + lane] : 0ull;
                if (__ballot_sync(0xffffffffu, (int)lane < W && other != word) == 0) {
                    ++found;
                    break;
                }
                slot = (slot + 1) & mask;
            }
        }
        if (lane == 0) hits[s] = found;
    }
}

void launch_read_table(const uint64_t* words, int W, uint32_t n, uint32_t* table, uint32_t capacity, cudaStream_t st) {
    if (!n) return;
    k_read_table<<<(n + 255) / 256, 256, 0, st>>>(words, W, n, table, capacity - 1);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_probe_strings(const uint8_t* text, const uint64_t* start, const uint32_t* length, uint64_t n_strings, int L,
                          int W, const uint64_t* words, const uint32_t* table, uint32_t capacity, uint32_t* hits,
                          cudaStream_t st) {
    if (!n_strings || L <= 0) return;
    const uint64_t blocks = (n_strings + 7) / 8;
    const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_probe_strings<<<grid, 256, 0, st>>>(text, start, length, n_strings, L, W, words, table, capacity - 1, hits);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_probe_windows(const uint8_t* text, const uint64_t* start, const uint32_t* length, uint64_t n_strings, int L,
                          int W, const uint64_t* words, const uint32_t* table, uint32_t capacity, uint32_t* hits,
                          cudaStream_t st) {
    if (!n_strings || L <= 0) return;
    const uint64_t blocks = (n_strings * (uint64_t)L + 7) / 8;
    const unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_probe_windows<<<grid, 256, 0, st>>>(text, start, length, n_strings, L, W, words, table, capacity - 1, hits);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk

namespace pedk {

void launch_len_stats(const uint64_t* seq_start, const uint64_t* seq_end, uint64_t n_records, unsigned* minmax,
                      cudaStream_t st) {
    if (!n_records) return;
    uint64_t blocks = (n_records + 256 * 16 - 1) / (256 * 16);
    unsigned grid = (unsigned)(blocks < (uint64_t)kNumSMs * 8 ? blocks : (uint64_t)kNumSMs * 8);
    k_len_stats<<<grid, 256, 0, st>>>(seq_start, seq_end, n_records, minmax);
    PEDK_CUDA_TRY(cudaGetLastError());
}

void launch_pack_fixed(const uint8_t* fq, const uint64_t* seq_start, uint64_t n_records, int L, int W,
                       uint64_t* words, uint8_t* pal, uint8_t* valid, unsigned long long* counters,
                       cudaStream_t st) {
    if (!n_records) return;
    const int gs = W <= 8 ? 8 : 16;
    uint64_t warps = (n_records + (32 / gs) * 16 - 1) / ((32 / gs) * 16);  // ~16 rounds per warp
    uint64_t blocks = (warps + 7) / 8;
    unsigned grid = (unsigned)(blocks < 1 ? 1 : blocks < (uint64_t)kNumSMs * 16 ? blocks : (uint64_t)kNumSMs * 16);
    if (gs == 8)
        k_pack_fixed<8><<<grid, 256, 0, st>>>(fq, seq_start, n_records, L, W, words, pal, valid, counters);
    else
        k_pack_fixed<16><<<grid, 256, 0, st>>>(fq, seq_start, n_records, L, W, words, pal, valid, counters);
    PEDK_CUDA_TRY(cudaGetLastError());
}

}  // namespace pedk
