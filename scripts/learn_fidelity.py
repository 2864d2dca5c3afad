#!/usr/bin/env python
"""Learning fidelity at mid size: learned weights of the parallel schedules vs the UNMODIFIED
reference's (tests/golden/c{2,3,4}_small.npz: mean over 5 reference seeds), for several
--learn_chunks settings.  Needs a GPU.  Prints one line per (graph, chunks)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sampler_b200 as sb  # noqa: E402
from helpers import GOLDEN, parse_dw_args  # noqa: E402
from sampler_b200 import graphgen  # noqa: E402


def run(raw, a, seed, chunks, sched=sb.SCHED_AUTO):
    s = sb.Sampler(sb.Graph.from_raw(raw, sched), sample_evidence=a["sample_evidence"], reg_param=a["reg_param"],
                   seed=seed, learn_chunks=chunks)
    step = a["stepsize"]
    for _ in range(a["n_learn"]):
        s.learn_epoch(step)
        step *= a["decay"]
    return s.weights()


def main():
    names = sys.argv[1:] or ["c2_small", "c3_small", "c4_small"]
    for name in names:
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        raw = getattr(graphgen, str(z["generator"]))(**json.loads(str(z["kwargs"])))
        a = parse_dw_args(str(z["dw_args"]).split())
        wref, wstd = z["weights_mean"], z["weights_std"]
        tol = np.maximum(0.05 * np.abs(wref), 0.03 + 2 * wstd)
        nref = float(z["nseeds"])
        for chunks in (0, 1, 4, 16, 64):
            runs = np.array([run(raw, a, seed, chunks) for seed in range(1, 17)])
            w = runs.mean(0)
            own_std = runs.std(0).mean()
            floor = np.sqrt(2 / np.pi) * np.sqrt(wstd ** 2 / nref + runs.std(0) ** 2 / len(runs)).mean()
            err = np.abs(w - wref)
            big = np.abs(wref) > 0.1
            slope = float(np.dot(w[big], wref[big]) / np.dot(wref[big], wref[big])) if big.any() else float("nan")
            print("%-9s chunks=%-3d bad=%.3f  max_err=%.3f  mean_err=%.4f (noise floor %.4f)  shrink=%.3f  "
                  "seed-std ours %.4f ref %.4f  (|wref| mean %.3f)"
                  % (name, chunks, float((err > tol).mean()), err.max(), err.mean(), floor, slope, own_std,
                     wstd.mean(), np.abs(wref).mean()), flush=True)


if __name__ == "__main__":
    main()
