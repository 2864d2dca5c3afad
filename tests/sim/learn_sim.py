"""Driver of learn_sim.c (test infrastructure): statistical fidelity of a parallel learning schedule,
simulated on the CPU oracle, against the reference's -t 1 goldens (tests/golden/c*_small.npz and the
bundled graphs' expected.json).

A schedule = (layout, plan):
  layout   how variables are grouped into 32-wide class-pure slices and in which order a learning
           sweep visits the slices ("class": class-major as the inference sweep; "blocked": slices
           formed inside (colour, evidence, id block) groups, visited in ascending id block)
  plan     per epoch a list of launches (fraction range of the visit order, warps in flight,
           per-weight step scale rule)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
import sampler_b200 as sb  # noqa: E402
from helpers import GOLDEN, load_golden_graph, parse_dw_args  # noqa: E402
from sampler_b200 import graphgen  # noqa: E402


class SimCfg(C.Structure):
    _fields_ = [("R", C.c_int), ("WPC", C.c_int), ("merge_per_class", C.c_int), ("batch_cta", C.c_int),
                ("merge_mode", C.c_int), ("SW", C.c_int)]


def lib():
    so = os.path.join(HERE, "liblearnsim.so")
    src = os.path.join(HERE, "learn_sim.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        import subprocess
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src, "-lm"])
    L = C.CDLL(so)
    L.sim_learn_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.sim_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    L.sim_set_step_scale.argtypes = [C.c_void_p]
    return L


def colouring(raw):
    g = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
    perm, cstart = g.schedule()
    color = np.zeros(raw.nvar, np.int64)
    for c in range(len(cstart) - 1):
        color[perm[int(cstart[c]):int(cstart[c + 1])].astype(np.int64)] = c
    return color, perm.astype(np.uint64), cstart.astype(np.uint64)


def make_layout(raw, kind, block=256, sw=32):
    """-> (order, sl_ptr): variables concatenated slice by slice in VISIT order"""
    color, perm, cstart = colouring(raw)
    nvar = raw.nvar
    slices = []
    if kind == "class":      # the inference layout: class-major, as compiled
        for c in range(len(cstart) - 1):
            b, e = int(cstart[c]), int(cstart[c + 1])
            for s0 in range(b, e, sw):
                slices.append(perm[s0:min(e, s0 + sw)])
    elif kind == "id":       # plain id order, 32 at a time (ignores classes: hogwild)
        ids = np.arange(nvar, dtype=np.uint64)
        for s0 in range(0, nvar, sw):
            slices.append(ids[s0:s0 + sw])
    elif kind == "blocked_shape":  # like "blocked", inside an id block sorted by incident-factor count descending
        evid = (raw.var_role > 0).astype(np.int64)
        ids = np.arange(nvar, dtype=np.int64)
        degv = np.bincount(raw.edge_vid.astype(np.int64), minlength=nvar)
        key = np.lexsort((ids, -degv, ids // block, evid, color))
        srt = ids[key]
        grp = color[srt] * 2 + evid[srt]
        bounds = np.flatnonzero(np.diff(grp)) + 1
        tmp = []
        for seg in np.split(srt, bounds):
            for s0 in range(0, len(seg), sw):
                tmp.append(seg[s0:s0 + sw].astype(np.uint64))
        tmp.sort(key=lambda a: (int(a[0]) // block, int(color[int(a[0])])))
        slices = tmp
    elif kind == "blocked":  # class-pure slices inside (colour, evid) in ascending id; visit by first id block
        evid = (raw.var_role > 0).astype(np.int64)
        ids = np.arange(nvar, dtype=np.int64)
        key = np.lexsort((ids, evid, color))          # colour-major, then evid, then id
        srt = ids[key]
        grp = color[srt] * 2 + evid[srt]
        bounds = np.flatnonzero(np.diff(grp)) + 1
        tmp = []
        for seg in np.split(srt, bounds):
            for s0 in range(0, len(seg), sw):
                tmp.append(seg[s0:s0 + sw].astype(np.uint64))
        tmp.sort(key=lambda a: (int(a[0]) // block, int(color[int(a[0])])))
        slices = tmp
    elif kind == "compiled":  # exactly what the product compiles: learn_visit over the padded schedule order
        g = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
        visit = g.image("learn_visit")
        perm_pad = g.image("perm")
        for sl in visit:
            v = perm_pad[int(sl) * 32:int(sl) * 32 + 32]
            slices.append(v[v != 0xffffffff].astype(np.uint64))
    else:
        raise ValueError(kind)
    order = np.concatenate(slices).astype(np.uint64)
    sl_ptr = np.zeros(len(slices) + 1, np.uint64)
    np.cumsum([len(a) for a in slices], out=sl_ptr[1:])
    return order, sl_ptr


def load_case(name):
    p = os.path.join(GOLDEN, name + ".npz")
    if os.path.exists(p):
        z = np.load(p)
        raw = getattr(graphgen, str(z["generator"]))(**json.loads(str(z["kwargs"])))
        a = parse_dw_args(str(z["dw_args"]).split())
        return raw, a, dict(w=z["weights_mean"], wstd=z["weights_std"], keys=z["marginal_keys"],
                            m=z["marginals_mean"].astype(np.float64), n=float(z["nseeds"]),
                            w8=z["weights_mean_t8"], m8=z["marginals_mean_t8"].astype(np.float64))
    raw = load_golden_graph(name)
    gdir = os.path.join(GOLDEN, "graphs", name)
    a = parse_dw_args(open(os.path.join(gdir, "dw-args")).read().split())
    e = json.load(open(os.path.join(gdir, "expected.json")))
    keys = np.array([[int(x) for x in k.split()] for k in e["marginal_keys"]], np.int64).reshape(-1, 2)
    return raw, a, dict(w=np.array(e["weights_mean"]), wstd=np.array(e["weights_std"]), keys=keys,
                        m=np.array(e["marginals_mean"]), n=float(e["nseeds"]), w8=None, m8=None)


def auto_plan(p, step, Nk, nvar):
    """{"c": tail length in relaxation times 1/step (updates of a weight), "cap_tail": bound on
    step * in-flight updates in the tail, "wide": warps of the wide phase, "cap": its cap}"""
    fast = Nk[Nk * step > 1.0]
    if len(fast) == 0:
        return [dict(lo=0.0, hi=1.0, warps=p["wide"], cap=p["cap"])]
    # tail long enough for the median fast weight to make c / step updates
    f = min(1.0, p["c"] / (step * float(np.quantile(fast, p.get("q", 0.5)))))
    warps = max(1, int(p["cap_tail"] / (step * 32.0 * float(Nk.max()) / nvar)))
    return [dict(lo=0.0, hi=1.0 - f, warps=p["wide"], cap=p["cap"]), dict(lo=1.0 - f, hi=1.0, warps=warps, cap=0.0)]


def run(name, layout="blocked", block=256, plan=None, nseeds=8, infer=True, batch_cta=1, sw=32):
    """plan: list of launches per epoch: dicts {lo, hi (fractions of the visit order), warps, cap}
    cap = bound on step * in-flight updates of a weight (the per-weight step is scaled down to meet
    it; 0 = no scaling)."""
    raw, a, gold = load_case(name)
    order, sl_ptr = make_layout(raw, layout, block, sw)
    nsl = len(sl_ptr) - 1
    plan = plan or [dict(lo=0.0, hi=1.0, warps=8, cap=0.0)]
    L = lib()
    ws, ms = [], []
    o0 = ol.Oracle(raw, sample_evidence=a["sample_evidence"], reg_param=a["reg_param"])
    # updates per weight per learning variable visit: N_k over the graph
    evid = raw.var_role > 0
    Nk = np.zeros(o0.nweight)
    fac_arity = np.diff(raw.fac_ptr.astype(np.int64))
    fac_of_edge = np.repeat(np.arange(raw.nfac), fac_arity)
    learn_edge = evid[raw.edge_vid.astype(np.int64)]
    np.add.at(Nk, raw.fac_wid[fac_of_edge[learn_edge]].astype(np.int64), 1.0)
    for seed in range(1, nseeds + 1):
        o = ol.Oracle(raw, sample_evidence=a["sample_evidence"], reg_param=a["reg_param"])
        nw = o.nweight
        W = o.weights.copy()
        cnt = np.zeros(nw)
        w0 = o.weights.copy()
        sd = (C.c_ushort * 3)(seed, 7 * seed + 1, 13 * seed + 5)
        step = a["stepsize"]
        for _ in range(a["n_learn"]):
            for ln in (auto_plan(plan, step, Nk, raw.nvar) if isinstance(plan, dict) else plan):
                s0, s1 = int(round(ln["lo"] * nsl)), int(round(ln["hi"] * nsl))
                if s1 <= s0:
                    continue
                cfg = SimCfg(1, int(ln["warps"]), 0, batch_cta, 0, 32)
                scale = None
                if ln.get("cap", 0) > 0:
                    # in-flight updates of weight k: warps * 32 variables * N_k / nvar
                    m = ln["warps"] * 32.0 * Nk / raw.nvar
                    scale = np.minimum(1.0, ln["cap"] / np.maximum(step * m, 1e-300))
                    L.sim_set_step_scale(scale.ctypes.data)
                else:
                    L.sim_set_step_scale(None)
                L.sim_learn_launch(o.g, C.byref(o.opts), o.s, step, order.ctypes.data, sl_ptr[s0:].ctypes.data, s1 - s0,
                                   C.byref(cfg), W.ctypes.data, cnt.ctypes.data, w0.ctypes.data, sd, 0)
                o.weights[:] = W  # R = 1: the replica copy is the weight vector
            step *= a["decay"]
        ws.append(W.copy())
        if infer and a["n_infer"]:
            o.weights[:] = W
            o.clear_tallies()
            ids = np.arange(o.nvar, dtype=np.uint64)
            L.sim_infer(o.g, C.byref(o.opts), o.s, ids.ctypes.data, o.nvar, a["n_infer"], sd)
            mm = o.marginals()
            ms.append(np.array([mm[(int(v), int(k))] for v, k in gold["keys"]]))
    w = np.mean(ws, 0)
    wref = gold["w"]
    big = np.abs(wref) > 0.1
    slope = float(np.dot(w[big], wref[big]) / np.dot(wref[big], wref[big])) if big.any() else float("nan")
    rel = np.abs(w - wref) / np.maximum(np.abs(wref), 1e-9)
    out = dict(layout=layout, block=block, plan=plan, dw=float(np.abs(w - wref).mean()), dw_signed=float((w - wref).mean()),
               relmax=float(rel[np.abs(wref) > 0.05].max()) if (np.abs(wref) > 0.05).any() else None, slope=slope,
               noise=float(np.sqrt(2 / np.pi) * (gold["wstd"] * np.sqrt(1 / gold["n"] + 1 / nseeds)).mean()))
    if gold["w8"] is not None:
        out["dw_t8"] = float(np.abs(gold["w8"] - wref).mean())
    if ms:
        m = np.mean(ms, 0)
        out.update(dm=float(np.abs(m - gold["m"]).mean()), dmmax=float(np.abs(m - gold["m"]).max()),
                   mshift=float(m.mean() - gold["m"].mean()))
        if gold["m8"] is not None:
            out["mshift_t8"] = float(gold["m8"].mean() - gold["m"].mean())
    return out


if __name__ == "__main__":
    name = sys.argv[1]
    for spec in sys.argv[2:]:
        kw = json.loads(spec)
        print(json.dumps(run(name, **kw)), flush=True)
