"""host/parallel.h: the thread fan-out of the loader / compiler.  A worker's exception (std::bad_alloc
on a graph too large for the host) must reach the caller -- the C-ABI entry points turn it into an
error code -- instead of ending the process in std::terminate."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_worker_exception_reaches_the_caller(tmp_path):
    src = tmp_path / "t.cc"
    src.write_text(textwrap.dedent('''
        #include <cstdio>
        #include <new>
        #include <stdexcept>
        #include <atomic>
        #include "%s/sampler_b200/csrc/host/parallel.h"
        int main() {
          std::atomic<unsigned long long> sum{0};
          dwg::parallel_ranges(100000, 8, [&](uint64_t a, uint64_t b, unsigned) { for (uint64_t i = a; i < b; ++i) sum += i; });
          if (sum != 100000ull * 99999ull / 2) return 2;
          try {
            dwg::parallel_ranges(100000, 8, [&](uint64_t a, uint64_t, unsigned p) { if (p == 5) throw std::bad_alloc(); (void)a; });
          } catch (const std::bad_alloc &) {
            std::puts("caught");
            return 0;
          }
          return 3;
        }
    ''' % ROOT))
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", str(exe), str(src)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "caught" in r.stdout, (r.returncode, r.stdout, r.stderr)
