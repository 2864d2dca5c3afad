#!/usr/bin/env python
"""parallel vs sequential learning on the GPU, same graph / args, many seeds (scratch tool)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sampler_b200 as sb
from helpers import GOLDEN, parse_dw_args
from sampler_b200 import graphgen
from scripts.learn_debug import run  # noqa


def stats(label, runs, base):
    w = runs.mean(0)
    print("%-22s mean|w - base| %.4f  max %.4f  seed-std %.4f" % (label, np.abs(w - base).mean(), np.abs(w - base).max(), runs.std(0).mean()), flush=True)
    return w


if __name__ == "__main__":
    name = "c2_small"
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = json.loads(str(z["kwargs"]))
    raw = getattr(graphgen, str(z["generator"]))(**kw)
    a = parse_dw_args(str(z["dw_args"]).split())
    for reg in (0.01, 0.0):
        print("reg", reg)
        seq = np.array([run(raw, a, s, 0, sb.SCHED_SEQUENTIAL, reg=reg) for s in range(1, 13)])
        base = seq.mean(0)
        if reg == 0.01:
            stats("sequential vs golden", seq, z["weights_mean"])
        for label, sched, chunks in (("chromatic/auto", sb.SCHED_CHROMATIC, 0), ("hogwild/100000", sb.SCHED_HOGWILD, 100000),
                                     ("chromatic/100000", sb.SCHED_CHROMATIC, 100000)):
            runs = np.array([run(raw, a, s, chunks, sched, reg=reg) for s in range(101, 117)])
            stats(label + " vs seq", runs, base)
