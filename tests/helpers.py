"""Shared test helpers: golden fixtures, small synthetic graphs."""
import json
import os

import numpy as np

from sampler_b200 import graphio as gio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPHS = ["biased_coin", "biased_coin_continuous", "biased_coin_truthiness",
          "biased_coin_with_multinomial", "partial_observation", "sparse_domains",
          "sparse_multinomial2"]


def graph_dir(name):
    return os.path.join(GOLDEN, "graphs", name)


def graph_paths(name):
    d = graph_dir(name)
    return {k: os.path.join(d, "graph." + k) for k in
            ("meta", "variables", "domains", "factors", "weights")}


def load_golden_graph(name):
    p = graph_paths(name)
    return gio.read_binary([p["variables"]], [p["domains"]], [p["factors"]], [p["weights"]])


def dw_args(name):
    return open(os.path.join(graph_dir(name), "dw-args")).read().split()


def parse_dw_args(args):
    """the subset of src/cmd_parser.cc:59-171 the bundled graphs use"""
    o = dict(n_learn=0, n_infer=0, stepsize=0.01, decay=0.95, reg_param=0.01,
             sample_evidence=False, learn_non_evidence=False, noise_aware=False, reg_l1=False)
    i = 0
    while i < len(args):
        a = args[i]
        if a in ("-l", "--n_learning_epoch"): o["n_learn"] = int(args[i + 1]); i += 2
        elif a in ("-i", "--n_inference_epoch"): o["n_infer"] = int(args[i + 1]); i += 2
        elif a in ("-a", "--alpha", "-p", "--stepsize"): o["stepsize"] = float(args[i + 1]); i += 2
        elif a in ("-d", "--diminish"): o["decay"] = float(args[i + 1]); i += 2
        elif a in ("-b", "--reg_param"): o["reg_param"] = float(args[i + 1]); i += 2
        elif a in ("-c", "--n_datacopy", "-t", "--n_threads"): i += 2
        elif a == "--sample_evidence": o["sample_evidence"] = True; i += 1
        elif a == "--learn_non_evidence": o["learn_non_evidence"] = True; i += 1
        elif a == "--noise_aware": o["noise_aware"] = True; i += 1
        elif a == "--regularization": o["reg_l1"] = args[i + 1] == "l1"; i += 2
        else: raise ValueError(a)
    return o


def expected(name):
    return json.load(open(os.path.join(graph_dir(name), "expected.json")))


def cxx_g(x):
    """C++ default ostream formatting of a double (precision 6, %g)."""
    return "%g" % x


def random_graph(rng, nvar=40, nfac=120, nweight=12, p_cat=0.4, max_arity=4, p_evid=0.5,
                 with_domains=True, funcs=None, truthiness=False):
    """Small random mixed Boolean/categorical graph exercising every factor function,
    repeated variables inside a factor, sparse domains (file order = dense order), evidence
    value 0 quirk, fixed weights and non-unit feature values."""
    funcs = list(funcs or gio.ALL_FUNCS)
    dtype = (rng.random(nvar) < p_cat).astype(np.uint16)
    card = np.where(dtype == 1, rng.integers(1, 6, nvar), 2).astype(np.uint64)
    role = (rng.random(nvar) < p_evid).astype(np.uint8) * rng.integers(1, 3, nvar).astype(np.uint8)
    dom_vid, dom_ptr, dom_val, dom_truth = [], [0], [], []
    domain = {}
    for v in range(nvar):
        if dtype[v] == 1 and with_domains and rng.random() < 0.7:
            vals = rng.permutation(20)[: int(card[v])]  # NOT sorted: file order is dense order
            domain[v] = vals
            dom_vid.append(v)
            dom_val.extend(vals.tolist())
            if truthiness and rng.random() < 0.6:
                t = rng.random(int(card[v])); t = t / t.sum() * rng.random()
            else:
                t = np.zeros(int(card[v]))
            dom_truth.extend(t.tolist())
            dom_ptr.append(len(dom_val))
        else:
            domain[v] = np.arange(int(card[v]))
    init = np.array([domain[v][rng.integers(0, len(domain[v]))] for v in range(nvar)], dtype=np.uint64)
    func, ptr, evid, eeq, wid, fv = [], [0], [], [], [], []
    for _ in range(nfac):
        f = funcs[rng.integers(0, len(funcs))]
        n = int(rng.integers(1, max_arity + 1))
        vs = rng.integers(0, nvar, n) if rng.random() < 0.2 else rng.permutation(nvar)[:n]
        for v in vs:
            evid.append(int(v))
            eeq.append(int(domain[int(v)][rng.integers(0, len(domain[int(v)]))]))
        func.append(f); ptr.append(len(evid)); wid.append(int(rng.integers(0, nweight)))
        fv.append(1.0 if rng.random() < 0.5 else float(rng.normal()))
    w_init = rng.normal(0, 1, nweight)
    w_fixed = (rng.random(nweight) < 0.3).astype(np.uint8)
    return gio.RawGraph(role, init, dtype, card, dom_vid, dom_ptr, dom_val, dom_truth, func, ptr,
                        evid, eeq, wid, fv, w_init, w_fixed).normalized()
