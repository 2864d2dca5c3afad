#!/usr/bin/env python
"""Wall time of the whole `dw gibbs` process on the C2 workload from binary files (GPU box):
load + compile + 20 learning epochs + 100 inference sweeps + result dump, with the per-phase
times the host prints under DWG_COMPILE_TIMING.  Prints one JSON line."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bundled():
    """configs[0] (C1): whole-process wall time of `dw gibbs` on the reference's bundled graphs with
    their own dw-args -- ours (one GPU, the multi-epoch sequential launches) beside the unmodified
    reference binary oracle/_ref/dw on the host cores (-t 1, the schedule ours reproduces bit for bit,
    and its default thread count).  Median of 5 runs each; prints one JSON line per graph."""
    from sampler_b200 import _build
    gdir = os.path.join(ROOT, "tests", "golden", "graphs")
    ref = os.path.join(ROOT, "oracle", "_ref", "dw")
    for name in sorted(os.listdir(gdir)):
        g = os.path.join(gdir, name)
        dwargs = open(os.path.join(g, "dw-args")).read().split()
        files = ["-m", g + "/graph.meta", "-v", g + "/graph.variables", "-f", g + "/graph.factors", "-w", g + "/graph.weights"]
        if os.path.getsize(g + "/graph.domains"):
            files += ["--domains", g + "/graph.domains"]
        row = {"graph": name, "dw_args": " ".join(dwargs)}
        for label, exe, extra in (("ours_gpu", _build.DW, ["-c", "1"]), ("reference_t1", ref, ["-t", "1", "-c", "1"]), ("reference_default", ref, [])):
            if not os.path.exists(exe):
                row[label + "_s"] = None
                continue
            times = []
            for _ in range(5):
                d = tempfile.mkdtemp(prefix="dwc1_")
                try:
                    t0 = time.time()
                    r = subprocess.run([exe, "gibbs"] + files + ["-o", d, "-q"] + dwargs + extra, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
                    times.append(time.time() - t0)
                    if r.returncode:
                        row[label + "_error"] = r.stderr[-300:]
                finally:
                    shutil.rmtree(d, ignore_errors=True)
            row[label + "_s"] = sorted(times)[len(times) // 2]
        print(json.dumps(row), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nvar", type=int, default=10_000_000)
    ap.add_argument("--bundled", action="store_true", help="C1: the reference's bundled graphs, ours vs oracle/_ref/dw")
    args = ap.parse_args()
    if args.bundled:
        return bundled()
    from sampler_b200 import _build, graphgen, graphio
    raw = graphgen.c2_boolean(args.nvar)
    d = tempfile.mkdtemp(prefix="dwcli_")
    try:
        t0 = time.time()
        paths = graphio.write_binary(raw, d)
        t_write = time.time() - t0
        del raw
        cmd = [_build.DW, "gibbs", "-m", paths["meta"], "-v", paths["variables"], "-f", paths["factors"],
               "-w", paths["weights"], "-o", d, "-l", "20", "-i", "100", "--alpha", "0.01", "--diminish", "0.95",
               "--reg_param", "0.01", "--sample_evidence", "-q"]
        env = dict(os.environ, DWG_COMPILE_TIMING="1", DWG_TIMING="1")
        t0 = time.time()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        wall = time.time() - t0
        phases = {m.group(1) + ": " + m.group(2).strip(): float(m.group(3)) for m in
                  re.finditer(r"\[(load|compile|app|create)\]\s+(.*?)\s+([0-9.]+) ms", r.stderr)}
        out = r.stdout
        learn = re.search(r"TOTAL LEARNING TIME: ([0-9.eE+-]+)", out)
        infer = re.search(r"TOTAL INFERENCE TIME: ([0-9.eE+-]+)", out)
        sizes = {k: os.path.getsize(v) for k, v in paths.items() if os.path.exists(v)}
        res = os.path.join(d, "inference_result.out.text")
        print(json.dumps({
            "workload": "C2, %d variables, -l 20 -i 100 --sample_evidence" % args.nvar, "rc": r.returncode,
            "wall_s": wall, "input_bytes": sum(sizes.values()), "write_inputs_s": t_write,
            "load_ms": sum(v for k, v in phases.items() if k.startswith("load:")),
            "compile_ms": sum(v for k, v in phases.items() if k.startswith("compile:")),
            "app_ms": {k[5:]: v for k, v in phases.items() if k.startswith("app:")},
            "learning_s": float(learn.group(1)) if learn else None,
            "inference_s": float(infer.group(1)) if infer else None,
            "marginal_lines": sum(1 for _ in open(res)) if os.path.exists(res) else 0,
            "phases_ms": phases}))
        if r.returncode:
            sys.stderr.write(r.stderr[-2000:])
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
