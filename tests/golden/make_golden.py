#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and oracle/_ref/, built by
`make -C oracle`); the outputs are committed so that neither the CPU tests nor the
GPU box ever read /root/reference.

  graphs/<name>/graph.{meta,variables,domains,factors,weights}, dw-args
        the reference's seven bundled end-to-end graphs (test/<name>/*.tsv) converted by
        the reference's own `dw text2bin`, multi-file kinds concatenated in `cat graph.x*`
        order exactly as test/run_end_to_end.sh:18-23 feeds them to `dw gibbs`
  graphs/<name>/expected.json
        weights / marginals of the UNMODIFIED reference (`oracle/_ref/dw`, -Ofast, -t 1 -c 1)
        over DW_SEED=1..NSEEDS: per-seed values, mean and std          -> parity check (3)
  graphs/<name>/exact_seed<k>.json
        outputs of the -O2 build (`oracle/_ref/dw_exact`) for DW_SEED=k -> bit-for-bit oracle pin
  factor_kat.json    the 62 EXPECT_NEARs of test/factor_test.cc:11-237, transcribed, and
                     re-verified here against the reference's Factor::potential
  factor_random.npz  random factors (all 10 function ids, arity 1..6) with the reference's
                     potentials                                        -> parity check (1)
  logadd.npz, draw_bool.npz, draw_cat.npz
                     reference logadd / draw_sample answers with their erand48 uniforms
                                                                       -> parity check (2)
  sampler_kat.json   test/sampler_test.cc:54-86 and test/factor_graph_test.cc:54-66
"""
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from sampler_b200 import graphio as gio  # noqa: E402

REF_TEST = "/root/reference/test"
GRAPHS = ["biased_coin", "biased_coin_continuous", "biased_coin_truthiness",
          "biased_coin_with_multinomial", "partial_observation", "sparse_domains",
          "sparse_multinomial2"]
NSEEDS = 200
KIND = {"variable": "variables", "domain": "domains", "factor": "factors", "weight": "weights"}


def text2bin_dir(src, dst):
    """test/run_end_to_end.sh:6-14 with the reference's own text2bin."""
    os.makedirs(dst, exist_ok=True)
    tmp = tempfile.mkdtemp()
    for what, plural in KIND.items():
        parts = []
        for tsv in sorted(f for f in os.listdir(src) if f.startswith(plural) and f.endswith(".tsv")):
            stem = tsv[:-4]
            out = os.path.join(tmp, "graph." + stem)
            args = []
            a = os.path.join(src, stem + ".text2bin-args")
            if os.path.exists(a):
                args = open(a).read().split()
            subprocess.check_call([ol.REF_DW, "text2bin", what, os.path.join(src, tsv), out,
                                   "/dev/null"] + args)
            parts.append(out)
        # `cat graph.<plural>*` order = lexicographic glob order of the OUTPUT names
        parts.sort()
        with open(os.path.join(dst, "graph." + plural), "wb") as f:
            for p in parts:
                f.write(open(p, "rb").read())
    shutil.copy(os.path.join(src, "graph.meta"), os.path.join(dst, "graph.meta"))
    shutil.copy(os.path.join(src, "dw-args"), os.path.join(dst, "dw-args"))
    shutil.rmtree(tmp)


def run_dw(binary, gdir, seed, extra=(), threads=1):
    out = tempfile.mkdtemp()
    args = open(os.path.join(gdir, "dw-args")).read().split()
    cmd = [binary, "gibbs", "-m", os.path.join(gdir, "graph.meta"),
           "-v", os.path.join(gdir, "graph.variables"), "-f", os.path.join(gdir, "graph.factors"),
           "-w", os.path.join(gdir, "graph.weights"), "--domains", os.path.join(gdir, "graph.domains"),
           "-o", out, "--quiet", "-t", str(threads), "-c", "1"] + args + list(extra)
    env = dict(os.environ, DW_SEED=str(seed))
    subprocess.check_call(cmd, env=env, stdout=subprocess.DEVNULL)
    w = open(os.path.join(out, "inference_result.out.weights.text")).read()
    mpath = os.path.join(out, "inference_result.out.text")
    m = open(mpath).read() if os.path.exists(mpath) else ""
    shutil.rmtree(out)
    return w, m


def parse_w(txt):
    return [float(l.split()[1]) for l in txt.splitlines() if l.strip()]


def parse_m(txt):
    return {"%s %s" % tuple(l.split()[:2]): float(l.split()[2]) for l in txt.splitlines() if l.strip()}


def make_graphs():
    for g in GRAPHS:
        dst = os.path.join(HERE, "graphs", g)
        text2bin_dir(os.path.join(REF_TEST, g), dst)
        ws, ms = [], []
        for seed in range(1, NSEEDS + 1):
            w, m = run_dw(ol.REF_DW, dst, seed)
            ws.append(parse_w(w))
            ms.append(parse_m(m))
        ws = np.array(ws)
        keys = list(ms[0].keys())
        mm = np.array([[m[k] for k in keys] for m in ms])
        exp = {
            "nseeds": NSEEDS, "binary": "oracle/_ref/dw (-Ofast) -t 1 -c 1, DW_SEED=1..%d" % NSEEDS,
            "weights_mean": ws.mean(0).tolist(), "weights_std": ws.std(0).tolist(),
            "weights_per_seed": ws.tolist(),
            "marginal_keys": keys, "marginals_mean": mm.mean(0).tolist(),
            "marginals_std": mm.std(0).tolist(),
        }
        json.dump(exp, open(os.path.join(dst, "expected.json"), "w"), indent=1)
        for seed in (1, 2):
            w, m = run_dw(os.path.join(ROOT, "oracle/_ref/dw_exact"), dst, seed)
            json.dump({"seed": seed, "weights_text": w, "marginals_text": m},
                      open(os.path.join(dst, "exact_seed%d.json" % seed), "w"), indent=1)
        print(g, "weights", ws.mean(0)[:4], "+-", ws.std(0)[:4])


def factor_kat():
    """test/factor_test.cc, transcribed: (func, vids, eqs, assignments, vid, proposal, expected)."""
    A, O, E, IN, IM, LI, RA, LO = 2, 1, 3, 0, 13, 7, 8, 9
    L3 = math.log2(3.0)
    rows = []

    def add(vids, eqs, vals, vid, p, table):
        for func, exp in table:
            rows.append(dict(func=func, vids=vids, eqs=eqs, assignments=vals, vid=vid,
                             proposal=p, expected=exp))
    seven = lambda a, o, e, i, l, r, g: [(A, a), (O, o), (E, e), (IN, i), (LI, l), (RA, r), (LO, g)]
    five = lambda i, m, l, r, g: [(IN, i), (IM, m), (LI, l), (RA, r), (LO, g)]
    # ONE_VAR_FACTORS :11-60
    add([0], [1], [0], 0, 1, seven(1, 1, 1, 1, 1, 1, 1))
    add([0], [1], [0], 0, 0, seven(-1, -1, 1, -1, 0, 0, 0))
    # TWO_VAR_FACTORS :62-157
    add([0, 1], [1, 1], [1, 1], 1, 1, seven(1, 1, 1, 1, 1, 1, 1))
    add([0, 1], [1, 1], [1, 1], 1, 0, seven(-1, 1, -1, -1, 0, 0, 0))
    add([0, 1], [1, 1], [0, 0], 1, 1, seven(-1, 1, -1, 0, 1, 1, 1))
    add([0, 1], [1, 1], [0, 0], 1, 0, seven(-1, -1, 1, 0, 1, 1, 1))
    # THREE_VAR_IMPLY :159-237
    add([0, 1, 2], [1, 1, 1], [1, 1, 1], 1, 0, five(0, 1, 2, L3, 1))
    add([0, 1, 2], [1, 1, 1], [1, 1, 1], 1, 1, five(1, 1, 2, L3, 1))
    add([0, 1, 2], [1, 1, 1], [1, 1, 1], 2, 0, five(-1, 0, 0, 0, 0))
    add([0, 1, 2], [1, 1, 1], [1, 1, 1], 2, 1, five(1, 1, 2, L3, 1))
    assert len(rows) == 62
    R = ol.ref()
    for r in rows:  # the transcription must agree with the reference itself
        got = R.ref_factor_potential(r["func"], len(r["vids"]), np.array(r["vids"], np.uint64),
                                     np.array(r["eqs"], np.uint64), 1.0,
                                     np.array(r["assignments"], np.uint64), r["vid"], r["proposal"])
        assert abs(got - r["expected"]) < 1e-5, (r, got)
    json.dump({"source": "test/factor_test.cc:11-237", "tol": 1e-5, "cases": rows},
              open(os.path.join(HERE, "factor_kat.json"), "w"), indent=0)


def factor_random():
    rng = np.random.default_rng(20260517)
    R = ol.ref()
    N, MAXA, NV = 6000, 6, 8
    func = np.zeros(N, np.int32); arity = np.zeros(N, np.int64)
    vids = np.zeros((N, MAXA), np.uint64); eqs = np.zeros((N, MAXA), np.uint64)
    assign = np.zeros((N, NV), np.uint64); fv = np.zeros(N); vid = np.zeros(N, np.uint64)
    prop = np.zeros(N, np.uint64); exp = np.zeros(N); exp_noprop = np.zeros(N)
    for i in range(N):
        func[i] = gio.ALL_FUNCS[i % len(gio.ALL_FUNCS)]
        n = arity[i] = 1 + (i // len(gio.ALL_FUNCS)) % MAXA
        card = 2 if rng.random() < 0.6 else 4
        # repeated variables inside one factor are legal and exercised on purpose
        vids[i, :n] = rng.integers(0, NV, n) if rng.random() < 0.3 else rng.permutation(NV)[:n]
        eqs[i, :n] = rng.integers(0, card, n)
        assign[i] = rng.integers(0, card, NV)
        fv[i] = 1.0 if rng.random() < 0.5 else rng.normal()
        vid[i] = vids[i, rng.integers(0, n)]
        prop[i] = rng.integers(0, card)
        exp[i] = R.ref_factor_potential(int(func[i]), int(n), vids[i, :n].copy(), eqs[i, :n].copy(),
                                        fv[i], assign[i].copy(), int(vid[i]), int(prop[i]))
        exp_noprop[i] = R.ref_factor_potential(int(func[i]), int(n), vids[i, :n].copy(),
                                               eqs[i, :n].copy(), fv[i], assign[i].copy(),
                                               2**64 - 1, 2**64 - 1)
    np.savez_compressed(os.path.join(HERE, "factor_random.npz"), func=func, arity=arity, vids=vids,
                        eqs=eqs, assign=assign, fv=fv, vid=vid, proposal=prop, expected=exp,
                        expected_noproposal=exp_noprop)


def logadd_vectors():
    rng = np.random.default_rng(7)
    R = ol.ref()
    a = np.concatenate([rng.normal(0, 20, 3000), [-100000.0] * 50, rng.normal(0, 1, 500), [0.0, 1.5, -3.0]])
    b = np.concatenate([rng.normal(0, 20, 3000), rng.normal(0, 30, 50), rng.normal(0, 1, 500), [0.0, 1.5, -3.0]])
    # straddle the -18.42 threshold (src/common.h:14,130)
    a = np.concatenate([a, np.zeros(200)])
    b = np.concatenate([b, -18.42 + rng.normal(0, 0.01, 200)])
    out = np.array([R.ref_logadd(float(x), float(y)) for x, y in zip(a, b)])
    np.savez_compressed(os.path.join(HERE, "logadd.npz"), a=a, b=b, expected=out)


def _unary_graph(dtype, card, pots_or_w):
    """one variable whose potential(k) is exactly a given number: K unary factors, own weights."""
    if dtype == 0:  # Boolean: one ISTRUE factor, weight w: pos=+w, neg=-w
        w = pots_or_w
        return gio.RawGraph([0], [0], [0], [2], *gio.empty_domains(), [gio.FUNC_ISTRUE], [0, 1], [0], [1],
                            [0], [1.0], [w], [1]).normalized()
    K = card
    return gio.RawGraph([0], [0], [1], [K], *gio.empty_domains(), [gio.FUNC_AND_CATEGORICAL] * K,
                        np.arange(K + 1), [0] * K, np.arange(K), np.arange(K), [1.0] * K,
                        pots_or_w, [1] * K).normalized()


def draw_vectors():
    """reference draw_sample (src/gibbs_sampler.h:192-254) on explicit potentials."""
    rng = np.random.default_rng(11)
    tmp = tempfile.mkdtemp()
    # Boolean: pos=w, neg=-w
    N = 400
    w = rng.normal(0, 1.5, N); seeds = rng.integers(0, 65536, (N, 3)).astype(np.uint16)
    r = np.zeros(N); out = np.zeros(N, np.uint64)
    g = ol.RefGraph(gio.write_binary(_unary_graph(0, 2, 0.0), tmp))
    for i in range(N):
        g.set_weights(np.array([w[i]]))
        g.L.ref_set_seed(g.h, *[int(x) for x in seeds[i]])
        r[i] = ol.uniforms([int(x) for x in seeds[i]], 1)[0]
        out[i] = g.L.ref_draw_sample(g.h, 0, 1)
    np.savez_compressed(os.path.join(HERE, "draw_bool.npz"), pos=w, neg=-w, r=r, expected=out)
    # categorical
    Ks = [1, 2, 3, 4, 7, 10, 33, 100]
    pots, rs, outs, ks = [], [], [], []
    for K in Ks:
        g = ol.RefGraph(gio.write_binary(_unary_graph(1, K, np.zeros(K)), tmp))
        for i in range(120):
            scale = [0.5, 3.0, 30.0][i % 3]
            p = rng.normal(0, scale, K)
            if i % 10 == 0:
                p[:] = p[0]  # exact ties: logadd's a==b branch
            s = [int(x) for x in rng.integers(0, 65536, 3)]
            g.set_weights(p)
            g.L.ref_set_seed(g.h, *s)
            rs.append(ol.uniforms(s, 1)[0])
            outs.append(int(g.L.ref_draw_sample(g.h, 0, 1)))
            pp = np.zeros(max(Ks)); pp[:K] = p
            pots.append(pp); ks.append(K)
    np.savez_compressed(os.path.join(HERE, "draw_cat.npz"), pots=np.array(pots), card=np.array(ks),
                        r=np.array(rs), expected=np.array(outs, np.uint64))
    shutil.rmtree(tmp)


def sampler_kat():
    u = ol.uniforms([1, 1, 1], 8).tolist()
    json.dump({
        "source": "test/sampler_test.cc:54-86, test/factor_graph_test.cc:54-66 on graphs/biased_coin",
        "erand48_seed_1_1_1": u,
        "sgd_on_variable": {"args": "--alpha 0.1 --reg_param 0", "free0": 0, "var": 0, "stepsize": 0.1,
                            "weight_after": 0.2, "then_var10_twice": 0.2},
        "sample_sgd_single_variable": {"seed": [1, 1, 1], "free0": 1, "var": 0, "stepsize": 0.1,
                                       "weights_after_each_of_3": [0.2, 0.2, 0.2]},
        "sample_single_variable": {"seed": [1, 1, 1], "steps": [
            {"weight": 2, "var": 0, "evid_after": 1}, {"weight": 2, "var": 10, "evid_after": 1},
            {"weight": 20, "var": 11, "evid_after": 1}, {"weight": 20, "var": 12, "evid_after": 1}]},
    }, open(os.path.join(HERE, "sampler_kat.json"), "w"), indent=1)


SCALE = {
    # name: (generator, kwargs, dw args)  -- mid-size graphs of the BASELINE.json shapes
    "c2_small": ("c2_boolean", dict(nvar=20000, nweight=50),
                 "-l 20 -i 200 --alpha 0.01 --diminish 0.95 --reg_param 0.01 --sample_evidence"),
    "c3_small": ("c3_categorical", dict(nvar=5000, card=10, nweight=50),
                 "-l 10 -i 200 --alpha 0.01 --reg_param 0.01 --sample_evidence"),
    "c4_small": ("unary_boolean", dict(nvar=20000, degree=20, nweight=2000, p_evid=0.9, zipf=1.1, real_features=True),
                 "-l 10 -i 50 --alpha 0.001 --diminish 0.9 --reg_param 0.01"),
}


SCALE_SEEDS = 16


def make_scale():
    """reference results on mid-size synthetic graphs (regenerated in the tests from the same
    seeded generator, sampler_b200/graphgen.py), all from the UNMODIFIED reference -t 1 -c 1:
      weights_mean/std, marginals_mean/std   learn + infer with the graph's dw-args, DW_SEED=1..16
      weights_mean_t8, marginals_mean_t8     the same with -t 8 (8 Hogwild threads): how far the reference
                                             moves from its own -t 1 results when run in parallel
      marginals_fixed                        inference only (-l 0) under FIXED weights = weights_mean,
                                             4 seeds x 500 sweeps: the target of the inference sweep
                                             alone, free of learning noise"""
    from sampler_b200 import graphgen
    for name, (gen, kw, dwargs) in SCALE.items():
        raw = getattr(graphgen, gen)(**kw)
        tmp = tempfile.mkdtemp()
        gio.write_binary(raw, tmp)
        open(os.path.join(tmp, "dw-args"), "w").write(dwargs)
        ws, ms, keys = [], [], None
        for seed in range(1, SCALE_SEEDS + 1):
            w, m = run_dw(ol.REF_DW, tmp, seed)
            ws.append(parse_w(w))
            mm = parse_m(m)
            keys = keys or list(mm.keys())
            ms.append([mm[k] for k in keys])
        # the reference's own spread between thread counts: the same runs with -t 8 (Hogwild)
        ws8, ms8 = [], []
        for seed in range(1, SCALE_SEEDS + 1):
            w, m = run_dw(ol.REF_DW, tmp, seed, threads=8)
            ws8.append(parse_w(w))
            mm = parse_m(m)
            ms8.append([mm[k] for k in keys])
        # inference under fixed weights
        wmean = np.mean(ws, 0)
        raw2 = getattr(graphgen, gen)(**kw)
        raw2.w_init = wmean.copy()
        gio.write_binary(raw2, tmp)
        fixed_args = [a for a in dwargs.split()]
        for flag in ("-l", "-i"):
            i = fixed_args.index(flag)
            fixed_args[i + 1] = "0" if flag == "-l" else "500"
        open(os.path.join(tmp, "dw-args"), "w").write(" ".join(fixed_args))
        mf = []
        for seed in range(1, 5):
            _, m = run_dw(ol.REF_DW, tmp, seed)
            mm = parse_m(m)
            mf.append([mm[k] for k in keys])
        shutil.rmtree(tmp)
        kv = np.array([[int(x) for x in k.split()] for k in keys], np.int64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), generator=gen, kwargs=json.dumps(kw),
                            dw_args=dwargs, weights_mean=wmean, weights_std=np.std(ws, 0),
                            marginal_keys=kv, marginals_mean=np.mean(ms, 0).astype(np.float32),
                            marginals_std=np.std(ms, 0).astype(np.float32), nseeds=SCALE_SEEDS,
                            marginals_fixed=np.mean(mf, 0).astype(np.float32), fixed_nsamples=4 * 500,
                            weights_mean_t8=np.mean(ws8, 0), weights_std_t8=np.std(ws8, 0),
                            marginals_mean_t8=np.mean(ms8, 0).astype(np.float32))
        print("   -t 8 vs -t 1: weights mean|diff|", np.abs(np.mean(ws8, 0) - wmean).mean(), "marginals",
              np.abs(np.mean(ms8, 0) - np.mean(ms, 0)).mean())
        print(name, "weights", wmean[:4], "+-", np.std(ws, 0)[:4], "mean marginal", np.mean(ms), np.mean(mf))


def text2bin_fixtures():
    """text2bin/: the reference's converter fixtures (test/text2bin/*.txt inputs with their xxd dumps
    *.bin.txt, test/text2bin.bats:11-27) and, per bundled graph, the TSV inputs of test/<name>/ with
    their text2bin-args, plus the reference's own bin2text rendering of the binary files committed
    under graphs/<name>/ -> converter parity both ways."""
    dst = os.path.join(HERE, "text2bin")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF_TEST, "text2bin"))):
        shutil.copy(os.path.join(REF_TEST, "text2bin", f), os.path.join(dst, f))
    for g in GRAPHS:
        src, out = os.path.join(REF_TEST, g), os.path.join(dst, g)
        os.makedirs(out, exist_ok=True)
        for f in sorted(os.listdir(src)):
            if f.endswith(".tsv") or f.endswith(".text2bin-args"):
                shutil.copy(os.path.join(src, f), os.path.join(out, f))
        # the reference's text2bin output per TSV (count on the last argument's file) and its
        # bin2text rendering of that output
        for what, plural in KIND.items():
            for tsv in sorted(f for f in os.listdir(src) if f.startswith(plural) and f.endswith(".tsv")):
                stem = tsv[:-4]
                args = []
                a = os.path.join(src, stem + ".text2bin-args")
                if os.path.exists(a):
                    args = open(a).read().split()
                binf, cnt = os.path.join(out, stem + ".bin"), os.path.join(out, stem + ".count")
                subprocess.check_call([ol.REF_DW, "text2bin", what, os.path.join(src, tsv), binf, cnt] + args)
        # the reference's bin2text rendering of the whole graph as committed under graphs/<name>/
        gdir, b2t = os.path.join(HERE, "graphs", g), os.path.join(out, "bin2text")
        os.makedirs(b2t, exist_ok=True)
        subprocess.check_call([ol.REF_DW, "bin2text", "-m", os.path.join(gdir, "graph.meta"),
                               "-v", os.path.join(gdir, "graph.variables"), "-f", os.path.join(gdir, "graph.factors"),
                               "-w", os.path.join(gdir, "graph.weights"), "--domains", os.path.join(gdir, "graph.domains"),
                               "-o", b2t], stdout=subprocess.DEVNULL)
        for junk in ("domains.tsv", "graph.meta"):  # dump_domains prints uninitialised memory (src/bin2text.cc:52-69)
            if os.path.exists(os.path.join(b2t, junk)):
                os.remove(os.path.join(b2t, junk))


if __name__ == "__main__":
    assert os.path.isdir(REF_TEST) and ol.have_ref(), "run `make -C oracle` in the build container first"
    what = sys.argv[1:] or ["graphs", "factor", "logadd", "draw", "sampler"]
    if "factor" in what:
        factor_kat(); factor_random()
    if "logadd" in what:
        logadd_vectors()
    if "draw" in what:
        draw_vectors()
    if "sampler" in what:
        sampler_kat()
    if "graphs" in what:
        make_graphs()
    if "scale" in what:
        make_scale()
    if "text2bin" in what:
        text2bin_fixtures()
