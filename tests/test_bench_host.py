"""Host-side logic of bench.py that needs no GPU: the nvidia-smi clock sampler's reduction of its
samples to the line's `clocks` object (a timed region of a few milliseconds usually holds no sample
of a 50 ms sampler: the nearest ones are reported, never a made-up 0 MHz)."""
import importlib.util
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Proc:
    def terminate(self):
        pass


def _sampler(bench, lines):
    c = bench.ClockSampler.__new__(bench.ClockSampler)
    c.proc, c.lines = _Proc(), lines
    return c


def test_clock_sampler_reduction():
    bench = _bench()
    now = time.time()
    idle = "0, 1965, 1965, 300.0, 0x0, Not Active, Not Active, Not Active, Active"
    busy = "0, 1950, 1965, 300.0, 0x0, Not Active, Not Active, Not Active, Not Active"
    lines = [(now - 1.0, idle), (now - 0.5, busy), (now - 0.2, "garbage")]
    # a sample inside the region: that one, and only its throttle reasons
    r = _sampler(bench, lines).stop(now - 0.52, now - 0.49)
    assert r["sm_mhz"] == 1950.0 and r["sm_max_mhz"] == 1965.0 and r["reasons"] == [] and r["samples"] == 1 and "note" not in r
    # none inside: the nearest ones, flagged
    r = _sampler(bench, lines).stop(now + 5.0, now + 5.005)
    assert r["sm_mhz"] == 1957.5 and r["samples"] == 2 and r["reasons"] == ["sw_power_cap"] and "nearest" in r["note"]
    # nothing at all: no clock is invented
    r = _sampler(bench, []).stop(now, now)
    assert r["sm_mhz"] is None and r["samples"] == 0 and r["reasons"]
    # nvidia-smi missing
    c = bench.ClockSampler.__new__(bench.ClockSampler)
    c.proc, c.lines = None, []
    assert c.stop(now, now)["sm_mhz"] is None


def test_bench_defaults_and_workloads():
    bench = _bench()
    assert {"c2", "c3", "c4", "c5"} <= set(bench.WORKLOADS)
    from sampler_b200.replicas import compute_n_epochs
    assert compute_n_epochs(20, 8) == 3 and compute_n_epochs(20, 1) == 20  # src/dimmwitted.cc:280-282
