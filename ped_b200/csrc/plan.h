// Level-A exchange plan (host-only; plan.cu).
#pragma once
#include <stdint.h>

#include <vector>

namespace pedk {

constexpr int LEVEL_A_BINS = 256;  // bins of the first binning level (= NB of bucket.cu)

struct LevelAPlan {
    unsigned long long C = 0;             // optimistic: slots per (bin, source rank) segment; 0: exact plan
    unsigned long long buffer_elems = 0;  // elements this rank's receive buffer needs
    std::vector<unsigned long long> cur;  // [256] where this rank's run of a bin starts in the OWNER's buffer
    std::vector<unsigned long long> lim;  // [256] optimistic: end of that segment
    std::vector<int> local_index;         // bin -> its index among its owner's bins
    // the segments of this rank's own bins, in bin order (then source order)
    std::vector<unsigned long long> seg_start, seg_end, seg_bin;
};

// A sample whose instances do not fit HBM at once is counted in n_pass passes (a power of two):
// pass `pass` takes the level-A bins with bin % n_pass == pass, so every rank owns an equal share
// of every pass.  Bins outside the pass get no room (local_index -1) and no segment.
// every (bin, source) run gets a segment of the expected size + 1/8 + 4096
void plan_level_a_optimistic(int a, int R, int me, uint64_t m_max, LevelAPlan* p, int n_pass = 1, int pass = 0);
// counts[src * 257 + bin] = what rank src stored into bin -> the segments of my bins
void level_a_segments_optimistic(int a, int R, int me, const unsigned long long* counts, LevelAPlan* p);
// exact sizes known up front: bins back to back, inside a bin the sources in rank order
void plan_level_a_exact(int a, int R, int me, const unsigned long long* counts, LevelAPlan* p, int n_pass = 1,
                        int pass = 0);

}  // namespace pedk
