// parallel.h -- the host compiler's and loader's thread fan-out (std::thread; no OpenMP).
//
// The reference parallelises loading with one thread per input file (src/binary_format.cc:38-46)
// and builds its index on one thread per ~1/n of the variables (src/factor_graph.cc:103-128).
// Here every O(edges) phase of compile.cc is split into contiguous ranges; all results are
// independent of the thread count (ranges only write their own slots, reductions are over
// integers or are order-fixed afterwards), so the compiled image is deterministic.
#ifndef DWG_PARALLEL_H_
#define DWG_PARALLEL_H_
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <thread>
#include <vector>

namespace dwg {

// threads used by the host phases: DWG_COMPILE_THREADS, else the hardware's, at most 32
inline unsigned host_threads() {
  if (const char *e = getenv("DWG_COMPILE_THREADS")) return (unsigned)std::max(1, atoi(e));
  return std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
}

// f(begin, end, part) over `parts` contiguous ranges of [0, n); part p covers
// [n * p / parts, n * (p + 1) / parts).  Runs inline when one part suffices or when there are
// fewer than `min_parallel` items (pass 1 when the items are coarse tasks).
template <typename F>
void parallel_ranges(uint64_t n, unsigned parts, F f, uint64_t min_parallel = 4096) {
  if (parts <= 1 || n < min_parallel) {
    if (parts <= 1) { f((uint64_t)0, n, 0u); return; }
    for (unsigned p = 0; p < parts; ++p) f(n * p / parts, n * (p + 1) / parts, p);
    return;
  }
  // an exception in a worker (std::bad_alloc on a graph too large for the host) is carried back to
  // the caller, whose C-ABI entry point turns it into an error code; left in the thread it would
  // be std::terminate
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(parts);
  th.reserve(parts);
  for (unsigned p = 0; p < parts; ++p)
    th.emplace_back([=, &err] {
      try { f(n * p / parts, n * (p + 1) / parts, p); } catch (...) { err[p] = std::current_exception(); }
    });
  for (auto &t : th) t.join();
  for (auto &e : err)
    if (e) std::rethrow_exception(e);
}

template <typename F>
void parallel_for(uint64_t n, F f) {
  parallel_ranges(n, host_threads(), [&](uint64_t a, uint64_t b, unsigned) { f(a, b); });
}

}  // namespace dwg
#endif
