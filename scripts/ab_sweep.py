#!/usr/bin/env python
"""A/B the inference-sweep kernel variants on one compiled graph (GPU box only).

  python scripts/ab_sweep.py [--nvar N] [--steps K] VARIANT...

A VARIANT is a comma list of ENV=VALUE pairs read by dwg_create (e.g. DWG_SELL_DIRECT=1,
DWG_SELL_STAGES=3) or the word `default`.  For every variant: same graph, same seed, a few
learning epochs, W warm-up sweeps, K sweeps timed with CUDA events on the sampler's stream.
Prints ms per sweep, variable-samples/s and a checksum of the tallies (variants must agree:
the draws do not depend on the kernel that makes them).
"""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--nvar", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--learn", type=int, default=3)
    ap.add_argument("--no-sample-evidence", action="store_true")
    ap.add_argument("variants", nargs="*", default=["default", "DWG_SELL_DIRECT=1"])
    args = ap.parse_args()
    import torch
    import sampler_b200 as sb
    from sampler_b200 import graphgen

    sb.load()
    gen = {"c2": graphgen.c2_boolean, "c3": graphgen.c3_categorical, "c4": graphgen.c4_logistic,
           "c5": graphgen.c5_large}[args.workload]
    t0 = time.time()
    raw = gen(args.nvar) if args.workload != "c4" else gen(args.nvar, nweight=max(1000, args.nvar // 50))
    t1 = time.time()
    graph = sb.Graph.from_raw(raw, sb.SCHED_AUTO)
    del raw
    info = graph.info
    print(json.dumps({"workload": args.workload, "nvar": args.nvar, "gen_s": t1 - t0, "compile_s": time.time() - t1,
                      "colors": info.num_colors, "schedule": info.schedule, "device_mib": info.device_bytes / 2**20,
                      "bytes_per_var_infer": info.bytes_per_sweep_infer / args.nvar,
                      "bytes_per_var_learn": info.bytes_per_sweep_learn / args.nvar}), flush=True)
    out = []
    for v in args.variants:
        env = {} if v == "default" else dict(kv.split("=", 1) for kv in v.split(","))
        for k in list(os.environ):
            if k.startswith("DWG_SELL_"):
                del os.environ[k]
        os.environ.update(env)
        smp = sb.Sampler(graph, device=0, sample_evidence=not args.no_sample_evidence, reg_param=0.01, seed=1234)
        stream = torch.cuda.ExternalStream(smp.stream(), device=torch.device("cuda", 0))
        step = 0.01
        learn_ms = None
        if args.learn:
            smp.learn_epoch(step)
            smp.sync()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                l0.record(stream)
                for _ in range(args.learn):
                    step *= 0.95
                    smp.learn_epoch(step)
                l1.record(stream)
            smp.sync()
            learn_ms = l0.elapsed_time(l1) / args.learn
        smp.infer_epochs(args.warmup)
        smp.clear_tallies()
        smp.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            smp.infer_epochs(args.steps)
            e1.record(stream)
        smp.sync()
        ms = e0.elapsed_time(e1) / args.steps
        t, n = smp.tallies()
        row = {"variant": v, "ms_per_sweep": ms, "gsamples_per_s": args.nvar / ms * 1e-6,
               "algorithmic_gbs": float(info.bytes_per_sweep_infer) / ms * 1e-6,
               "learn_ms_per_epoch": learn_ms,
               "tally_crc": zlib.crc32(np.ascontiguousarray(t).tobytes()), "tally_sum": int(t.sum())}
        print(json.dumps(row), flush=True)
        out.append(row)
        smp.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
