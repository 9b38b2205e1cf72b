// Host build of the shared host/device helpers (ped_b200/csrc/common.cuh) so the
// bit tricks the kernels rely on can be checked on the CPU.
#include "../ped_b200/csrc/common.cuh"
extern "C" {
uint64_t hd_kmer_revcomp(uint64_t k, int n) { return pedk::kmer_revcomp(k, n); }
uint64_t hd_kmer_at(const uint64_t* w, int W, int i, int k) { return pedk::kmer_at(w, W, i, k); }
void hd_read_revcomp(const uint64_t* w, int W, int L, uint64_t* out) { pedk::read_revcomp(w, W, L, out); }
int hd_words_cmp(const uint64_t* a, const uint64_t* b, int W) { return pedk::words_cmp(a, b, W); }
uint32_t hd_base_code(uint32_t c) { return pedk::base_code(c); }
}
