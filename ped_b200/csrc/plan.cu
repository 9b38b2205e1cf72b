// Host-only index arithmetic of the level-A exchange (no kernel, no CUDA call): where every
// (source rank, bin) run goes inside the buffer of the rank that owns the bin, and which
// segments the owner hands to the level-B passes.  Kept apart from sample.cu so that the CPU
// tests can replay a whole multi-rank exchange on it (tests/test_host_helpers.py).
#include "plan.h"

#include "../../include/pedk.h"
#include "common.cuh"

namespace pedk {

static int owner_of_bin(int bin, int R, int a) { return (int)(((uint64_t)bin * (uint64_t)R) >> a); }

void plan_level_a_optimistic(int a, int R, int me, uint64_t m_max, LevelAPlan* p, int n_pass, int pass) {
    const int nbA = 1 << a;
    // expected size of a (bin, source) run + 1/8 + 4096, a multiple of 32 elements
    p->C = ((m_max / nbA) + (m_max / nbA) / 8 + 4096 + 31) / 32 * 32;
    p->cur.assign(LEVEL_A_BINS, 0);
    p->lim.assign(LEVEL_A_BINS, 0);
    p->local_index.assign(nbA, -1);
    std::vector<int> owned(R, 0);
    for (int bin = 0; bin < nbA; ++bin) {
        if (bin % n_pass != pass) continue;  // not in this pass: cur == lim == 0, nothing fits, nothing is stored
        const int o = owner_of_bin(bin, R, a);
        p->local_index[bin] = owned[o]++;
        // segment of (bin, src) inside the owner's buffer: ((bin's index among the owner's bins) * R + src) * C
        p->cur[bin] = ((unsigned long long)p->local_index[bin] * R + me) * p->C;
        p->lim[bin] = p->cur[bin] + p->C;
    }
    p->buffer_elems = (unsigned long long)owned[me] * R * p->C;
    p->seg_start.clear();
    p->seg_end.clear();
    p->seg_bin.clear();
}

void level_a_segments_optimistic(int a, int R, int me, const unsigned long long* counts, LevelAPlan* p) {
    const int nbA = 1 << a;
    p->seg_start.clear();
    p->seg_end.clear();
    p->seg_bin.clear();
    for (int bin = 0; bin < nbA; ++bin) {
        if (owner_of_bin(bin, R, a) != me || p->local_index[bin] < 0) continue;
        for (int src = 0; src < R; ++src) {
            const unsigned long long st0 = ((unsigned long long)p->local_index[bin] * R + src) * p->C;
            p->seg_start.push_back(st0);
            p->seg_end.push_back(st0 + counts[(size_t)src * (LEVEL_A_BINS + 1) + bin]);
            p->seg_bin.push_back((unsigned long long)bin);
        }
    }
}

void plan_level_a_exact(int a, int R, int me, const unsigned long long* counts, LevelAPlan* p, int n_pass, int pass) {
    const int nbA = 1 << a;
    p->C = 0;
    p->cur.assign(LEVEL_A_BINS, 0);
    p->lim.clear();
    p->seg_start.clear();
    p->seg_end.clear();
    p->seg_bin.clear();
    std::vector<unsigned long long> owner_fill(R, 0);
    for (int bin = 0; bin < nbA; ++bin) {
        if (bin % n_pass != pass) continue;
        const int owner = owner_of_bin(bin, R, a);
        unsigned long long tot = 0, before_me = 0;
        for (int src = 0; src < R; ++src) {
            const unsigned long long c = counts[(size_t)src * (LEVEL_A_BINS + 1) + bin];
            if (src < me) before_me += c;
            tot += c;
        }
        p->cur[bin] = owner_fill[owner] + before_me;  // sources in rank order inside the bin
        if (owner == me) {
            p->seg_start.push_back(owner_fill[owner]);
            p->seg_end.push_back(owner_fill[owner] + tot);
            p->seg_bin.push_back((unsigned long long)bin);
        }
        owner_fill[owner] += tot;
    }
    p->buffer_elems = owner_fill[me];
}

}  // namespace pedk

// host-only, for tests: the plan of rank `rank` of `nranks`.  exact == 0: optimistic plan for
// m_max instances per rank (counts may be NULL: no segments then); exact != 0: from the counts.
// counts[src * 257 + bin].  seg holds 3 x n_seg values (starts, ends, bins), at most 3 x 256 x nranks.
extern "C" int pedk_plan_level_a(int k, int nranks, int rank, uint64_t m_max, const uint64_t* counts, int exact,
                                 uint64_t* cur256, uint64_t* lim256, uint64_t* seg, int* n_seg,
                                 uint64_t* buffer_elems, uint64_t* segment_capacity) {
    return pedk_plan_level_a_pass(k, nranks, rank, m_max, counts, exact, 1, 0, cur256, lim256, seg, n_seg, buffer_elems,
                                  segment_capacity);
}

extern "C" int pedk_plan_level_a_pass(int k, int nranks, int rank, uint64_t m_max, const uint64_t* counts, int exact,
                                      int n_pass, int pass, uint64_t* cur256, uint64_t* lim256, uint64_t* seg,
                                      int* n_seg, uint64_t* buffer_elems, uint64_t* segment_capacity) {
    if (k < 4 || k > 32 || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks || !cur256 || !n_seg) return PEDK_E_INVAL;
    if (n_pass < 1 || (n_pass & (n_pass - 1)) || pass < 0 || pass >= n_pass) return PEDK_E_INVAL;
    if (exact && !counts) return PEDK_E_INVAL;
    const int a = pedk::kmer_bits_a(k);
    pedk::LevelAPlan p;
    const unsigned long long* c = reinterpret_cast<const unsigned long long*>(counts);
    if (exact) {
        pedk::plan_level_a_exact(a, nranks, rank, c, &p, n_pass, pass);
    } else {
        pedk::plan_level_a_optimistic(a, nranks, rank, m_max, &p, n_pass, pass);
        if (c) pedk::level_a_segments_optimistic(a, nranks, rank, c, &p);
    }
    for (int i = 0; i < pedk::LEVEL_A_BINS; ++i) {
        cur256[i] = p.cur[(size_t)i];
        if (lim256) lim256[i] = exact ? 0 : p.lim[(size_t)i];
    }
    *n_seg = (int)p.seg_start.size();
    if (seg)
        for (int s = 0; s < *n_seg; ++s) {
            seg[s] = p.seg_start[(size_t)s];
            seg[*n_seg + s] = p.seg_end[(size_t)s];
            seg[2 * *n_seg + s] = p.seg_bin[(size_t)s];
        }
    if (buffer_elems) *buffer_elems = p.buffer_elems;
    if (segment_capacity) *segment_capacity = p.C;
    return PEDK_OK;
}
