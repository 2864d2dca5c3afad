#!/usr/bin/env python
"""bench.py -- Gibbs variable-samples/s of the sweep hot path on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A "step" is one inference sweep (GibbsSamplerThread::sample over every variable,
--sample_evidence) of the synthetic graph BASELINE.json's metric is quoted on that fits one
GPU: configs[1], the Boolean graph of 10 M variables x (8 unary ISTRUE + 1 IMPLY) with 1 k
shared weights, after the configured learning epochs.  One replica per GPU, no data-path
collective in the timed region except the final tally reduction ("scaling": "weak").

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      algorithmic bytes of a sweep / CUDA-event duration vs the
                measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (oracle/_ref/dw, built from the unmodified reference
                sources) on this host's cores, on a bounded sample of the same workload
  e2e           the same metric through the C ABI with HOST buffers: dwg_create (H2D upload of
                the compiled graph) + K sweeps + dwg_get_tallies (D2H), wall-clocked
  learn         learning-epoch throughput (free-chain draw + evid clamp/draw + SGD per variable)

--impl reference times the reference CPU implementation only (rank 0; other ranks exit 0).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_DW = os.path.join(ROOT, "oracle", "_ref", "dw")
WORKLOADS = {
    # name: generator, default size, dw-args of the config (SURVEY.md 8(d) "Concrete synthetic inputs")
    "c2": dict(gen="c2_boolean", nvar=10_000_000,
               desc="configs[1]: Boolean 10M vars x (8 unary ISTRUE + 1 IMPLY_NATURAL), 1k shared weights, 50% evidence, "
                    "--sample_evidence",
               learn=dict(n_epoch=20, alpha=0.01, diminish=0.95, reg_param=0.01), kernel="k_sell_infer_staged<MULTI=1,FV=0>"),
    "c3": dict(gen="c3_categorical", nvar=5_000_000,
               desc="configs[2]: categorical 5M vars, cardinality 10, sparse domains, per variable 10 unary + 2 pairwise "
                    "AND_CATEGORICAL + 1 EQUAL, 1k shared weights, 50% evidence, --sample_evidence",
               learn=dict(n_epoch=10, alpha=0.01, diminish=1.0, reg_param=0.01), kernel="k_cat_infer_staged"),
    # the two graphs too large for any record array: streamed by the library (csrc/host/synth.cc); the
    # reference's CPU run holds 72 B per edge, so it gets the same generator at 1/10 resp. 1/30 of
    # the variables ("run the reference on a /10 (C4) / /30 (C5) graph ... and say so")
    "c4": dict(synth=True, nvar=50_000_000, ref_div=10,
               desc="configs[3]: logistic-style Boolean 50M vars x 20 unary ISTRUE features (f64 values ~N(0,1)), 1M tied "
                    "weights (Zipf 1.1), 90% evidence from a planted logistic model, SGD with decay, --sample_evidence",
               learn=dict(n_epoch=10, alpha=0.001, diminish=0.9, reg_param=0.01), kernel="k_sell_infer_staged<MULTI=0,FV=f64>"),
    "c5": dict(synth=True, nvar=300_000_000, ref_div=30,
               desc="configs[4]: README-scale Boolean 300M vars x 10 unary ISTRUE features per replica, 1k shared weights, "
                    "50% evidence, inference only, --sample_evidence",
               learn=dict(n_epoch=0, alpha=0.01, diminish=0.95, reg_param=0.01), kernel="k_sell_infer_staged<MULTI=0,FV=0>"),
}
LEARN = WORKLOADS["c2"]["learn"]
# the final tally aggregation: reduce onto rank 0 (default; the faster one on 8 B200s) or all-reduce (DWG_BENCH_TALLY_REDUCE=all)
TO_ALL = os.environ.get("DWG_BENCH_TALLY_REDUCE", "root") == "all"


def make_raw(workload, nvar):
    """the workload's records (numpy): file-backed sizes only"""
    from sampler_b200 import graphgen
    wl = WORKLOADS[workload]
    if wl.get("synth"):
        spec = dict(graphgen.SYNTH_SPECS[workload], nvar=nvar)
        return graphgen.synth_unary(seed=1234, **spec)
    return getattr(graphgen, wl["gen"])(nvar)


def common_config(args, world):
    """the part of `config` both arms share word for word (the driver compares them)"""
    wl = WORKLOADS[args.workload]
    nvar = args.nvar or wl["nvar"]
    return {"workload": wl["desc"] if nvar == wl["nvar"] else wl["desc"] + " [nvar overridden: %d]" % nvar,
            "nvar_per_replica": nvar, "replicas": world, "graph_seed": 1234, "sample_evidence": True,
            "l2_policy": "inputs larger than L2 (the compiled graph streamed per sweep exceeds the 126 MB L2); no flush"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        try:
            self.proc = subprocess.Popen([exe, "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        for _ in range(60):  # (nvidia-smi takes a few hundred ms to print its first line)
            if self.lines:
                break
            time.sleep(0.05)
        self.proc.terminate()
        mx, parsed = None, []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            parsed.append((ts, clk, [n for n, val in zip(names, f[5:9]) if val.lower().startswith("active")]))
        if not parsed:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": ["nvidia-smi printed no sample"], "samples": 0}
        used = [p for p in parsed if t0 - 0.05 <= p[0] <= t1 + 0.15]
        out = {}
        if not used:
            # region shorter than the sampling period: the samples nearest to it (the sampler has been running
            # since before the learning epochs and the warm-up sweeps, i.e. under the same kind of load)
            mid = 0.5 * (t0 + t1)
            used = sorted(parsed, key=lambda p: abs(p[0] - mid))[:3]
            out["note"] = "no sample inside the %.1f ms timed region: the %d nearest ones" % (1e3 * (t1 - t0), len(used))
        out.update({"sm_mhz": float(np.median([p[1] for p in used])), "sm_max_mhz": mx,
                    "reasons": sorted({r for p in used for r in p[2]}), "samples": len(used)})
        return out


# --------------------------------------------------------------------------------------
# the reference's CPU path, on a bounded sample of the workload
# --------------------------------------------------------------------------------------
def run_reference_cpu(workload: str, nvar: int, n_infer: int, n_learn: int, tmpdir: str):
    """oracle/_ref/dw gibbs (the unmodified reference, -Ofast, all host threads) on the SAME graph
    the GPU arm sweeps (same generator, same size, written in the reference's binary format).
    Returns per-epoch vars/sec lists (inference, learning) as the reference prints them."""
    from sampler_b200 import graphgen, graphio
    if not os.path.exists(REF_DW):
        return None
    wl = WORKLOADS[workload]
    t0 = time.time()
    raw = make_raw(workload, nvar)
    paths = graphio.write_binary(raw, tmpdir)
    del raw
    t_files = time.time() - t0
    threads = os.cpu_count() or 1
    L = wl["learn"]
    cmd = [REF_DW, "gibbs", "-m", paths["meta"], "-v", paths["variables"], "-f", paths["factors"],
           "-w", paths["weights"], "-o", tmpdir, "-l", str(n_learn), "-i", str(n_infer),
           "--alpha", str(L["alpha"]), "--diminish", str(L["diminish"]),
           "--reg_param", str(L["reg_param"]), "--sample_evidence", "-t", str(threads), "-c", "1"]
    if os.path.exists(paths.get("domains", "")) and os.path.getsize(paths["domains"]):
        cmd += ["--domains", paths["domains"]]
    t0 = time.time()
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    wall = time.time() - t0
    inf = [float(x) for x in re.findall(r"INFERENCE EPOCH.*?sec\.,([0-9.eE+]+) vars/sec", out)]
    lrn = [float(x) for x in re.findall(r"LEARNING EPOCH.*?sec\.,([0-9.eE+]+) vars/sec", out)]
    for k in ("variables", "factors", "weights", "domains"):
        try:
            os.remove(paths[k])
        except (OSError, KeyError):
            pass
    log("reference: files %.1fs, dw gibbs %.1fs (%d learning + %d inference epochs)" % (t_files, wall, len(lrn), len(inf)))
    return dict(infer=inf, learn=lrn, threads=threads, wall=wall, nvar=nvar, tail=out[-400:])


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    nvar_full = args.nvar or wl["nvar"]
    nvar = nvar_full // wl.get("ref_div", 1)
    tmp = tempfile.mkdtemp(prefix="dwref_")
    try:
        r = run_reference_cpu(args.workload, nvar, args.steps + args.warmup, 0, tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if r is None or len(r["infer"]) < args.warmup + 1:
        args.emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/dw missing or produced no epochs: %s"
                              % (r["tail"] if r else "no binary")}))
        return 0
    timed = r["infer"][args.warmup:]
    # whole-run throughput over the timed epochs: total samples / total time
    secs = [nvar / v for v in timed]
    value = nvar * len(timed) / sum(secs)
    sample = "the whole workload: %d variables, %d inference sweeps after %d warm-up sweeps" % (nvar, len(timed), args.warmup)
    if nvar != nvar_full:
        sample = ("the workload's generator at 1/%d of the variables (%d; the reference holds 72 B per edge and cannot load "
                  "the full graph), %d inference sweeps after %d warm-up sweeps" % (wl["ref_div"], nvar, len(timed), args.warmup))
    line = {
        "impl": "reference", "metric": "gibbs_variable_samples_per_sec", "value": value, "unit": "variable-samples/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(args, 1),
        "impl_config": {"binary": "oracle/_ref/dw (unmodified reference, -Ofast)", "threads": r["threads"], "n_datacopy": 1},
        "cpu_baseline": {"value": value, "unit": "variable-samples/s", "cores": r["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "variable-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.emit(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def ours(args):
    import torch
    import sampler_b200 as sb
    from sampler_b200 import graphgen
    from sampler_b200.replicas import ReplicaGroup, compute_n_epochs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus)
            return 2
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    sb.load(build=not os.path.exists(sb.api.lib_path()))
    L = sb.load()
    wl = WORKLOADS[args.workload]
    LEARN = wl["learn"]
    nvar = args.nvar or wl["nvar"]
    t0 = time.time()
    if wl.get("synth"):
        t_gen = 0.0
        graph = sb.Graph.synthetic_unary(seed=1234, **dict(graphgen.SYNTH_SPECS[args.workload], nvar=nvar))
    else:
        raw = make_raw(args.workload, nvar)
        t_gen = time.time() - t0
        t0 = time.time()
        graph = sb.Graph.from_raw(raw, sb.SCHED_AUTO)
        del raw
    t_compile = time.time() - t0
    info = graph.info
    if rank == 0:
        log("graph: %d vars, %d edge records, %d colours, %.1f MiB/replica; gen %.1fs compile %.1fs"
            % (info.num_variables, info.num_index_entries, info.num_colors, info.device_bytes / 2**20, t_gen, t_compile))
    clocks = ClockSampler(local) if rank == 0 else None  # (started here: running well before the timed region)
    smp = sb.Sampler(graph, device=local, sample_evidence=True, reg_param=LEARN["reg_param"], seed=1234, replica=rank)
    stream = torch.cuda.ExternalStream(smp.stream(), device=torch.device("cuda", local))
    # one replica per rank: DimmWitted's learn / inference schedule over torch.distributed (NCCL)
    group = ReplicaGroup(smp, dist, device=local)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- learning phase: ceil(l / N) epochs per replica, weights averaged every epoch ----
    n_learn = compute_n_epochs(LEARN["n_epoch"], world)
    step = LEARN["alpha"]
    learn_ms, learn_launches = None, 0.0
    if n_learn:
        smp.learn_epoch(step)  # warm-up epoch (JIT-free, but first touch of everything)
        smp.sync()
        barrier()
        e0, e1 = ev(), ev()
        l0 = L.dwg_launch_count()
        with torch.cuda.stream(stream):
            e0.record(stream)
            group.learn(LEARN["n_epoch"], step, LEARN["diminish"])  # every epoch: sweep, allreduce of the weights, average
            e1.record(stream)
        smp.sync()
        learn_launches = (L.dwg_launch_count() - l0) / float(n_learn)  # sweep launches + the averaging epilogue
        learn_ms = e0.elapsed_time(e1) / n_learn
    lt = torch.tensor([learn_ms or 0.0], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(lt, op=dist.ReduceOp.MAX)
    learn_ms = float(lt.item()) or None

    # ---- inference: W warm-up sweeps, K timed ----
    smp.clear_tallies()
    smp.infer_epochs(args.warmup)
    if dist is not None:
        group.reduce_tallies_device(TO_ALL, max_count=args.warmup * world)  # warm-up of the collective too (its first call sets up the channels)
    smp.clear_tallies()
    smp.sync()
    torch.cuda.synchronize()
    barrier()
    launches0 = L.dwg_launch_count()
    e0, e1, e2 = ev(), ev(), ev()
    t_wall0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        smp.infer_epochs(args.steps)
        e1.record(stream)
        if dist is not None:
            group.reduce_tallies_device(TO_ALL, max_count=args.steps * world)  # sample tallies reduced at the end of inference (narrow counters)
        e2.record(stream)
    smp.sync()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier()
    launches = L.dwg_launch_count() - launches0
    sweep_ms = e0.elapsed_time(e1) / args.steps
    total_ms = e0.elapsed_time(e2)
    tt = torch.tensor([total_ms, sweep_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, sweep_ms = float(tt[0].item()), float(tt[1].item())
    value = nvar * world * args.steps / (total_ms * 1e-3)
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None

    t, ns = smp.tallies()
    assert int(ns[0]) == args.steps
    # Boolean: share of draws that were 1; categorical: every draw tallies one value (must be exactly 1)
    frac1 = float(t.sum()) / (float(nvar) * args.steps * world) if rank == 0 else 0.0

    # ---- e2e through the C ABI from HOST buffers, whole job: the compiled graph goes host -> device
    # once per node (rank 0: dwg_create from page-locked memory) and reaches the other replicas over
    # NVLink (defer_upload + broadcast of the graph segments, the analogue of the reference's per-NUMA-
    # node copies, src/dimmwitted.cc:113-116); K sweeps per replica; tallies reduced onto rank 0 and
    # read back into a page-locked host array (u32, the device's counter width).  Whole-job value =
    # all replicas' samples / the slowest rank's wall time. ----
    smp.close()
    torch.cuda.synchronize()
    t_host = torch.empty(int(info.num_values), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    # (page-locked as well: the per-variable sample counts then arrive by DMA instead of being written by the host's threads)
    n_host = torch.zeros(int(info.num_variables), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    if dist is not None:  # the first broadcast on a communicator sets up its channels: not part of a sweep
        warm = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        dist.broadcast(warm, src=0)
        torch.cuda.synchronize()
    reps = []
    for rep in range(5):  # five whole repetitions, the median is reported (the first maps the device memory
        barrier()         # of the replica; later ones find it in the library's pool)
        n_host[0] = n_host[-1] = 0  # (the check below must see this repetition's counts)
        t0 = time.time()
        s2 = sb.Sampler(graph, device=local, sample_evidence=True, reg_param=LEARN["reg_param"], seed=99 + rep, replica=rank,
                        defer_upload=(world > 1 and rank != 0))
        if world > 1:
            g2 = ReplicaGroup(s2, dist, device=local)
            g2.broadcast_graph(0)
        t_up = time.time()
        s2.infer_epochs(args.steps)  # asynchronous: the sweeps are queued on the replica's stream
        if world > 1:
            g2.reduce_tallies_device(TO_ALL, max_count=args.steps * world)
        if rank == 0:  # the getter synchronises; the sample counts travel while the sweeps still run
            s2.tallies(out=(t_host, n_host))
        else:
            s2.sync()
        t_sw = t_end = time.time()
        et = torch.tensor([t_end - t0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        reps.append((float(et.item()), t_up - t0, t_sw - t_up, t_end - t_sw))
        if rank == 0:
            assert int(t_host.astype(np.uint64).sum()) > 0 and int(n_host[0]) == args.steps and int(n_host[-1]) == args.steps
        s2.close()
        torch.cuda.synchronize()
    reps.sort()
    e2e_s, p_up, p_sw, p_dn = reps[len(reps) // 2]
    e2e_value = nvar * world * args.steps / e2e_s
    e2e_parts = "median of %d repetitions (%s ms): graph to the device(s) %.1f ms, sweeps%s + tallies and sample counts to the host %.1f ms (rank 0)" % (
        len(reps), ", ".join("%.1f" % (1e3 * r[0]) for r in reps), 1e3 * p_up, " + tally reduction" if world > 1 else "", 1e3 * (p_sw + p_dn))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- the reference's CPU path on this host, bounded sample ----
    cpu = None
    if not args.no_cpu_baseline:
        tmp = tempfile.mkdtemp(prefix="dwref_")
        try:
            r = run_reference_cpu(args.workload, args.ref_nvar or nvar // wl.get("ref_div", 1), 6, 3 if LEARN["n_epoch"] else 0, tmp)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        if r and len(r["infer"]) >= 3:
            cpu = {"value": float(np.median(r["infer"][1:])), "unit": "variable-samples/s", "cores": r["threads"],
                   "kind": "reference",
                   "sample": "oracle/_ref/dw (unmodified reference, -Ofast) -t %d -c 1 on %s (%d variables, the "
                             "files the GPU arm's generator writes): 3 learning + 6 inference epochs, median of inference "
                             "epochs 2..6; learning epochs %s vars/s"
                             % (r["threads"], "the same graph" if r["nvar"] == nvar else "the same generator at 1/%d of the variables" % wl.get("ref_div", 1), r["nvar"], ("%.3g" % np.median(r["learn"][1:])) if len(r["learn"]) > 1 else "n/a")}
        else:
            cpu = {"value": None, "unit": "variable-samples/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": "oracle/_ref/dw unavailable on this box"}

    peak, peak_src = measured_peak()
    bytes_sweep = float(info.bytes_per_sweep_infer)
    achieved = bytes_sweep / (sweep_ms * 1e-3) / 1e9
    # one launch = one colour class: the dominant launch is the largest class
    launch_share = 1.0
    if info.num_colors > 1:
        _, cstart = graph.schedule()
        class_sizes = np.diff(cstart.astype(np.int64))
        launch_share = float(class_sizes.max()) / float(max(1, class_sizes.sum()))
    traffic = frac_dram = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and nvar == wl["nvar"]:
        try:  # DRAM bytes of the sweep kernel's launches from the committed `ncu --set full` capture of this workload
            t = json.load(open(tpath))[args.workload]
            traffic = t["dram_bytes_per_launch"]
            frac_dram = t["dram_bytes_per_sweep"] / (sweep_ms * 1e-3) / 1e9 / peak
            traffic_src = t["source"]
        except Exception:
            traffic = frac_dram = None
    cfg = common_config(args, world)
    line = {
        "metric": "gibbs_variable_samples_per_sec", "value": value, "unit": "variable-samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "impl_config": {"schedule": {2: "chromatic", 1: "hogwild", 3: "sequential"}.get(info.schedule), "colors": info.num_colors,
                        "device_mib_per_replica": info.device_bytes / 2**20, "learn_epochs_before": n_learn + 1, "sampler_seed": 1234},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "frac_dram": frac_dram, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "%s (one launch per colour class; achieved = algorithmic bytes of a whole sweep / CUDA-event "
                               "time of the sweep, i.e. the launch-weighted average; frac_dram = DRAM bytes of a sweep "
                               "(ncu) / the same time / peak)" % wl["kernel"],
                     "algorithmic_bytes_per_sweep": bytes_sweep, "bytes_per_variable_sample": bytes_sweep / nvar,
                     "algorithmic_bytes_per_launch": bytes_sweep * launch_share,
                     "launches_per_sweep": info.num_colors},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "variable-samples/s",
                "h2d_bytes_per_step": int(info.device_bytes / args.steps), "d2h_bytes_per_step": int((4 * info.num_values + 8 * info.num_variables) / args.steps),
                "what": "dwg_create (upload of the compiled graph from page-locked host memory, once per node; the other "
                        "replicas receive it over NVLink) + %d sweeps per replica + tally reduction onto rank 0 + "
                        "dwg_get_tallies32 (u32 tallies and u64 sample counts, D2H into page-locked host arrays), wall clock: %s" % (args.steps, e2e_parts)},
        "learn": None if not learn_ms else
                 {"value": nvar * world / (learn_ms * 1e-3), "unit": "variable-updates/s", "ms_per_epoch": learn_ms,
                  "epochs": n_learn, "hbm_gbs": float(info.bytes_per_sweep_learn) / (learn_ms * 1e-3) / 1e9,
                  "frac": float(info.bytes_per_sweep_learn) / (learn_ms * 1e-3) / 1e9 / peak, "launches_per_epoch": learn_launches,
                  "what": "one epoch = a learning sweep of every variable + the allreduce of the weight vector (%d B) over "
                          "the replicas + the averaging epilogue, all inside the timed region" % (8 * info.num_weights)},
        "gpu_launches": int(launches), "clocks": clk, "frac_true": frac1,
    }
    args.emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 while the bench runs
    (NCCL's version banner, library chatter) goes to stderr.  Returns a writer for the line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        sys.stdout.flush()
        os.write(saved, (line + "\n").encode())
    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--nvar", type=int, default=0, help="override variables per replica (debug)")
    ap.add_argument("--ref-nvar", type=int, default=0, help="cpu_baseline only: variables of the reference's sample "
                    "(default: the workload's own size, i.e. the same files)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.emit = _claim_stdout()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
