#!/usr/bin/env python
"""Learning fidelity at mid size: learned weights / marginals of the parallel learning sweeps vs the
UNMODIFIED reference's -t 1 results (tests/golden/c{2,3,4}_small.npz), for several settings of the
epoch plan (environment knobs of launch_learn_sell, dwg_api.cu).  Needs a GPU.

  python scripts/learn_fidelity.py [graph ...] [--variants "K=V,K=V;K=V"] [--seeds N] [--top N]
"""
import argparse
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


def run(graph, a, seed, infer):
    s = sb.Sampler(graph, sample_evidence=a["sample_evidence"], reg_param=a["reg_param"], seed=seed)
    step = a["stepsize"]
    for _ in range(a["n_learn"]):
        s.learn_epoch(step)
        step *= a["decay"]
    w = s.weights()
    t = n = None
    if infer:
        s.clear_tallies()
        s.infer_epochs(a["n_infer"])
        t, n = s.tallies()
    s.close()
    return w, t, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("graphs", nargs="*", default=["c2_small", "c4_small"])
    ap.add_argument("--variants", default="")
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--top", type=int, default=0, help="print the N most used weights")
    args = ap.parse_args()
    variants = [dict(kv.split("=") for kv in v.split(",") if kv) for v in args.variants.split(";")] if args.variants else [{}]
    for name in args.graphs:
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        raw = getattr(graphgen, str(z["generator"]))(**json.loads(str(z["kwargs"])))
        a = parse_dw_args(str(z["dw_args"]).split())
        wref, wstd = z["weights_mean"], z["weights_std"]
        mref = z["marginals_mean"].astype(np.float64)
        nref = float(z["nseeds"])
        graph = sb.Graph.from_raw(raw)
        boolean = raw.var_dtype.max() == 0
        use = np.bincount(raw.fac_wid.astype(np.int64), minlength=raw.nweight)
        for env in variants:
            for k in list(os.environ):
                if k.startswith("DWG_LEARN_"):
                    del os.environ[k]
            os.environ.update(env)
            ws, ms = [], []
            for seed in range(1, args.seeds + 1):
                w, t, n = run(graph, a, seed, boolean)
                ws.append(w)
                if boolean:
                    vid = z["marginal_keys"][:, 0]
                    _, _, base, _ = graph.var_meta()
                    ms.append(t[base[vid].astype(np.int64)] / n[vid].astype(np.float64))
            w = np.mean(ws, 0)
            sig = wstd * np.sqrt(1.0 / nref + 1.0 / args.seeds)
            tol = np.maximum(0.05 * np.abs(wref), 0.03) + 3 * sig
            err = np.abs(w - wref)
            big = np.abs(wref) > 0.1
            slope = float(np.dot(w[big], wref[big]) / np.dot(wref[big], wref[big])) if big.any() else float("nan")
            line = "%-9s %-40s bad=%.4f mean|dw|=%.4f (noise %.4f) slope=%.3f" % (
                name, json.dumps(env), float((err > tol).mean()), err.mean(), np.sqrt(2 / np.pi) * sig.mean(), slope)
            if ms:
                m = np.mean(ms, 0)
                per_seed = np.array([x.mean() for x in ms])
                line += " mean|dm|=%.4f shift=%+.4f (per-seed mean marginal std %.4f)" % (
                    np.abs(m - mref).mean(), m.mean() - mref.mean(), per_seed.std())
            print(line, flush=True)
            if args.top:
                for k in np.argsort(-use)[:args.top]:
                    print("    w[%d] uses %d: ours %+.4f ref %+.4f (t8 %+.4f) std %.4f tol %.4f %s" % (
                        k, use[k], w[k], wref[k], z["weights_mean_t8"][k], wstd[k], tol[k], "BAD" if err[k] > tol[k] else ""))


if __name__ == "__main__":
    main()
