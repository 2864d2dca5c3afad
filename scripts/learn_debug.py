#!/usr/bin/env python
"""Where does the parallel learning differ from the sequential reference?  (GPU scratch tool)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sampler_b200 as sb
from helpers import GOLDEN, parse_dw_args
from sampler_b200 import graphgen


def run(raw, a, seed, chunks, sched, n_learn=None, reg=None):
    s = sb.Sampler(sb.Graph.from_raw(raw, sched), sample_evidence=a["sample_evidence"],
                   reg_param=a["reg_param"] if reg is None else reg, seed=seed, learn_chunks=chunks)
    step = a["stepsize"]
    for _ in range(n_learn or a["n_learn"]):
        s.learn_epoch(step)
        step *= a["decay"]
    return s.weights()


name = sys.argv[1] if len(sys.argv) > 1 else "c2_small"
z = np.load(os.path.join(GOLDEN, name + ".npz"))
raw = getattr(graphgen, str(z["generator"]))(**json.loads(str(z["kwargs"])))
a = parse_dw_args(str(z["dw_args"]).split())
wref = z["weights_mean"]
res = {}
for label, sched, chunks, seeds in (("sequential", sb.SCHED_SEQUENTIAL, 0, 4), ("chromatic/auto", sb.SCHED_CHROMATIC, 0, 16),
                                    ("hogwild/auto", sb.SCHED_HOGWILD, 0, 16), ("hogwild/256", sb.SCHED_HOGWILD, 256, 16)):
    runs = np.array([run(raw, a, seed, chunks, sched) for seed in range(1, seeds + 1)])
    w = runs.mean(0)
    res[label] = w - wref
    print("%-16s mean|err| %.4f  max %.4f  seed-std %.4f  corr(err, wref) %.3f" %
          (label, np.abs(w - wref).mean(), np.abs(w - wref).max(), runs.std(0).mean(), np.corrcoef(w - wref, wref)[0, 1]), flush=True)
ks = list(res)
for i in range(len(ks)):
    for j in range(i + 1, len(ks)):
        print("corr(err %s, err %s) = %.3f" % (ks[i], ks[j], np.corrcoef(res[ks[i]], res[ks[j]])[0, 1]))
# which weights sit on multi-variable factors?
nf = raw.nfac
ar = np.diff(raw.fac_ptr.astype(np.int64))
share = np.bincount(raw.fac_wid[ar > 1].astype(np.int64), minlength=raw.nweight) / np.maximum(1, np.bincount(raw.fac_wid.astype(np.int64), minlength=raw.nweight))
for k in ks:
    print("%-16s corr(|err|, share of multi factors) %.3f   corr(err, share) %.3f" % (k, np.corrcoef(np.abs(res[k]), share)[0, 1], np.corrcoef(res[k], share)[0, 1]))
print("first 10 weights: ref", np.round(wref[:10], 3))
for k in ks:
    print("   err %-14s" % k, np.round(res[k][:10], 3))
