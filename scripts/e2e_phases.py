"""Phase times of the host-buffer path (dwg_create -> sweeps -> dwg_get_tallies32) on the C2 graph, with
the library's own lap timers (DWG_TIMING=1 on stderr) and three destinations for the per-variable
sample counts: a pageable array, none, a page-locked array (what bench.py's e2e leg passes) -- the
last also without a synchronisation between the sweeps and the getter (the counts then travel while
the sweeps run).
    python scripts/e2e_phases.py [nvar] [sweeps]
"""
import os
import sys
import time

os.environ["DWG_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import sampler_b200 as sb  # noqa: E402
from sampler_b200 import graphgen  # noqa: E402
from sampler_b200.api import _ptr, load  # noqa: E402


def main():
    nvar = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    g = sb.Graph.from_raw(graphgen.c2_boolean(nvar), sb.SCHED_AUTO)
    info = g.info
    t_host = torch.empty(int(info.num_values), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    n_page = np.zeros(int(info.num_variables), np.uint64)
    n_pin = torch.zeros(int(info.num_variables), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    for rep in range(4):
        for mode in ("pageable", "none", "page-locked", "page-locked, sweeps in flight"):
            t0 = time.time()
            s = sb.Sampler(g, sample_evidence=True, seed=rep)
            t1 = time.time()
            s.infer_epochs(sweeps)
            if "in flight" not in mode:
                s.sync()
            t2 = time.time()
            if mode == "none":
                rc = load().dwg_get_tallies32(s._h, _ptr(t_host), None)
                assert rc == 0
            else:
                s.tallies(out=(t_host, n_page if mode == "pageable" else n_pin))
            t3 = time.time()
            print("rep %d nsamples -> %-30s create %.2f ms, sweeps %.2f ms, tallies %.2f ms" % (
                rep, mode, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), file=sys.stderr, flush=True)
            s.close()


if __name__ == "__main__":
    main()
