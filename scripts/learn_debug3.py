#!/usr/bin/env python
"""parallel vs sequential learning: unary-only graph, and step-size dependence (scratch tool)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sampler_b200 as sb
from sampler_b200 import graphgen


def run(raw, a, seed, chunks, sched):
    s = sb.Sampler(sb.Graph.from_raw(raw, sched), sample_evidence=True, reg_param=a["reg_param"], seed=seed, learn_chunks=chunks)
    step = a["stepsize"]
    for _ in range(a["n_learn"]):
        s.learn_epoch(step)
        step *= a["decay"]
    return s.weights()


def compare(label, raw, a, nseq=12, npar=16):
    seq = np.array([run(raw, a, s, 0, sb.SCHED_SEQUENTIAL) for s in range(1, nseq + 1)])
    base = seq.mean(0)
    for plabel, sched, chunks in (("auto", sb.SCHED_AUTO, 0), ("chunk=slice", sb.SCHED_AUTO, 100000)):
        runs = np.array([run(raw, a, s, chunks, sched) for s in range(101, 101 + npar)])
        w = runs.mean(0)
        floor = np.sqrt(2 / np.pi) * np.sqrt(seq.std(0) ** 2 / nseq + runs.std(0) ** 2 / npar).mean()
        print("%-34s %-12s mean|par - seq| %.4f (noise floor %.4f)  max %.4f  slope %.3f" %
              (label, plabel, np.abs(w - base).mean(), floor, np.abs(w - base).max(), np.dot(w, base) / np.dot(base, base)), flush=True)


if __name__ == "__main__":
    base_args = dict(n_learn=20, stepsize=0.01, decay=0.95, reg_param=0.01)
    compare("unary deg 10, a=0.01", graphgen.unary_boolean(20000, 10, 50), base_args)
    compare("c2 (8 unary + IMPLY), a=0.01", graphgen.c2_boolean(20000, nweight=50), base_args)
    compare("c2, a=0.002 x 100 epochs", graphgen.c2_boolean(20000, nweight=50), dict(base_args, stepsize=0.002, n_learn=100, decay=0.99), nseq=6)
    compare("c2 without evidence noise: 8 unary only", graphgen.unary_boolean(20000, 8, 50), base_args)
