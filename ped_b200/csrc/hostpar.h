// Host-side helpers for the text stages (verify.cu, refindex.cu): a loop and a sort on a few threads.
#pragma once
#include <algorithm>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

namespace pedk {

// fn(lo, hi) over [0, n) on a few host threads (the text work is per line / per candidate); fewer than
// `serial_below` items run on the calling thread
template <class F>
void parallel_for(size_t n, size_t serial_below, F fn) {
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : nt > 16 ? 16 : nt;
    if (n < serial_below) nt = 1;
    if ((size_t)nt > n) nt = n ? (unsigned)n : 1;
    if (nt == 1) {
        fn((size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    std::exception_ptr err;
    std::mutex mu;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            try {
                fn(n * t / nt, n * (t + 1) / nt);
            } catch (...) {
                std::lock_guard<std::mutex> g(mu);
                if (!err) err = std::current_exception();
            }
        });
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

template <class F>
void parallel_for(size_t n, F fn) {
    parallel_for(n, (size_t)2048, fn);
}

// std::sort's result on a few host threads: sorted runs, then pairwise merges (input that is in order
// already -- the usual case for the candidate keys -- costs one pass)
template <class T, class Less>
void parallel_sort(std::vector<T>& v, Less less) {
    if (std::is_sorted(v.begin(), v.end(), less)) return;
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : nt > 16 ? 16 : nt;
    if (v.size() < 8192 || nt == 1) {
        std::sort(v.begin(), v.end(), less);
        return;
    }
    unsigned runs = 1;
    while (runs * 2 <= nt) runs *= 2;
    const size_t n = v.size();
    auto bound = [&](unsigned r) { return n * r / runs; };
    {
        std::vector<std::thread> th;
        for (unsigned r = 0; r < runs; ++r)
            th.emplace_back([&, r]() { std::sort(v.begin() + (long)bound(r), v.begin() + (long)bound(r + 1), less); });
        for (auto& x : th) x.join();
    }
    for (unsigned width = 1; width < runs; width *= 2) {
        std::vector<std::thread> th;
        for (unsigned r = 0; r + width < runs; r += 2 * width)
            th.emplace_back([&, r]() {
                std::inplace_merge(v.begin() + (long)bound(r), v.begin() + (long)bound(r + width),
                                   v.begin() + (long)bound(std::min(runs, r + 2 * width)), less);
            });
        for (auto& x : th) x.join();
    }
}

}  // namespace pedk
