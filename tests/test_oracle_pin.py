"""Pin the CPU oracle (oracle/dw_oracle.c) to the reference: the reference's own
known-answer tests (committed under tests/golden/ by make_golden.py), reference-generated
vectors, bit-for-bit whole runs, and -- where oracle/_ref/libdwref.so exists -- the
reference objects themselves."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from helpers import (GOLDEN, GRAPHS, cxx_g, dw_args, graph_dir, graph_paths, load_golden_graph,
                     parse_dw_args, random_graph)
from sampler_b200 import graphio as gio

INVALID = 2**64 - 1


def test_factor_known_answers():
    """the 62 EXPECT_NEARs of test/factor_test.cc"""
    L = ol.oracle()
    kat = json.load(open(os.path.join(GOLDEN, "factor_kat.json")))
    assert len(kat["cases"]) == 62
    for c in kat["cases"]:
        got = L.orc_factor_potential(c["func"], len(c["vids"]), np.array(c["vids"], np.uint64),
                                     np.array(c["eqs"], np.uint64), 1.0,
                                     np.array(c["assignments"], np.uint64), c["vid"], c["proposal"])
        assert abs(got - c["expected"]) <= kat["tol"], c


def test_factor_random_vs_reference_vectors():
    L = ol.oracle()
    z = np.load(os.path.join(GOLDEN, "factor_random.npz"))
    for i in range(len(z["func"])):
        n = int(z["arity"][i])
        a = (int(z["func"][i]), n, z["vids"][i, :n].copy(), z["eqs"][i, :n].copy(), float(z["fv"][i]),
             z["assign"][i].copy())
        assert L.orc_factor_potential(*a, int(z["vid"][i]), int(z["proposal"][i])) == z["expected"][i]
        assert L.orc_factor_potential(*a, INVALID, INVALID) == z["expected_noproposal"][i]


def test_logadd_bit_exact():
    L = ol.oracle()
    z = np.load(os.path.join(GOLDEN, "logadd.npz"))
    got = np.array([L.orc_logadd(float(a), float(b)) for a, b in zip(z["a"], z["b"])])
    assert np.array_equal(got, z["expected"])


def test_draws_bit_exact():
    L = ol.oracle()
    z = np.load(os.path.join(GOLDEN, "draw_bool.npz"))
    got = [L.orc_draw_boolean(float(p), float(n), float(r)) for p, n, r in zip(z["pos"], z["neg"], z["r"])]
    assert np.array_equal(np.array(got, np.uint64), z["expected"])
    z = np.load(os.path.join(GOLDEN, "draw_cat.npz"))
    for pots, k, r, e in zip(z["pots"], z["card"], z["r"], z["expected"]):
        assert L.orc_draw_categorical(np.ascontiguousarray(pots[:k]), int(k), float(r)) == e


def _coin(**kw):
    return ol.Oracle(load_golden_graph("biased_coin"), reg_param=0.0, **kw)


def test_sgd_known_answer():
    """test/factor_graph_test.cc:54-66"""
    o = _coin()
    o.assign_free[0] = 0
    o.sgd_on_variable(0, 0.1)
    assert o.weights[0] == 0.2
    o.sgd_on_variable(10, 0.1)
    o.sgd_on_variable(10, 0.1)
    assert o.weights[0] == 0.2


def test_sampler_known_answers():
    """test/sampler_test.cc:54-86"""
    kat = json.load(open(os.path.join(GOLDEN, "sampler_kat.json")))
    u = ol.uniforms([1, 1, 1], 4)
    assert np.allclose(u, kat["erand48_seed_1_1_1"][:4], rtol=0, atol=0)
    assert abs(u[0] - 0.77099259800749564) < 1e-15
    o = _coin()
    o.assign_free[0] = 1
    seed = ol.Seed3(1, 1, 1)
    for _ in range(3):
        o.sample_sgd_single_variable(0, 0.1, seed)
        assert o.weights[0] == 0.2
    o = _coin()
    o.assign_free[0] = 1
    seed = ol.Seed3(1, 1, 1)
    o.weights[0] = 2
    o.sample_single_variable(0, seed); assert o.assign_evid[0] == 1
    o.sample_single_variable(10, seed); assert o.assign_evid[10] == 1
    o.weights[0] = 20
    o.sample_single_variable(11, seed); assert o.assign_evid[11] == 1
    o.sample_single_variable(12, seed); assert o.assign_evid[12] == 1


@pytest.mark.parametrize("name", GRAPHS)
def test_whole_run_bit_for_bit(name):
    """oracle `orc_run` == reference `dw gibbs -t 1 -c 1` (O2 build) output files, byte for byte"""
    a = parse_dw_args(dw_args(name))
    for seed in (1, 2):
        gold = json.load(open(os.path.join(graph_dir(name), "exact_seed%d.json" % seed)))
        o = ol.Oracle(load_golden_graph(name), sample_evidence=a["sample_evidence"],
                      learn_non_evidence=a["learn_non_evidence"], noise_aware=a["noise_aware"],
                      reg_l1=a["reg_l1"], reg_param=a["reg_param"])
        o.run(a["n_learn"], a["n_infer"], a["stepsize"], a["decay"], srand_seed=seed)
        wtxt = "".join("%d %s\n" % (i, cxx_g(w)) for i, w in enumerate(o.weights))
        assert wtxt == gold["weights_text"]
        mtxt = "".join("%d %d %s\n" % (v, val, cxx_g(p)) for (v, val), p in o.marginals().items())
        assert mtxt == gold["marginals_text"]


# ---- against the reference objects themselves (build container / prebuilt _ref) ----
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libdwref.so not built")


def _pair(raw, tmp_path, extra=(), **okw):
    paths = gio.write_binary(raw, str(tmp_path))
    return ol.Oracle(raw, **okw), ol.RefGraph(paths, extra)


def _check_structure(o, r):
    rv = r.variables()
    assert np.array_equal(rv[:, 0], o.dtype) and np.array_equal(rv[:, 1], o.is_evid)
    assert np.array_equal(rv[:, 3], o.assignment_dense) and np.array_equal(rv[:, 4], o.var_val_base)
    val, tr, base, ln = r.values()
    assert np.array_equal(val, o.value) and np.array_equal(tr, o.truthiness)
    assert np.array_equal(ln, o.fi_len)
    nz = ln > 0
    assert np.array_equal(base[nz], o.fi_base[nz])
    assert np.array_equal(r.factor_index(o.nindex), o.factor_index)
    a, b = r.vifs()
    assert np.array_equal(a, o.vif_vid) and np.array_equal(b, o.vif_eq)


@needs_ref
@pytest.mark.parametrize("name", GRAPHS)
def test_index_matches_reference_bundled(name, tmp_path):
    o = ol.Oracle(load_golden_graph(name))
    r = ol.RefGraph(graph_paths(name))
    _check_structure(o, r)


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_random_graph_vs_reference(seed, tmp_path):
    """construct_index, potential(), sweeps: bit-exact on random mixed graphs"""
    rng = np.random.default_rng(100 + seed)
    noise = seed % 2 == 1
    raw = random_graph(rng, truthiness=noise)
    extra = ["--reg_param", "0.05", "--sample_evidence"] + (["--noise_aware"] if noise else []) + \
            (["--learn_non_evidence"] if seed % 3 == 0 else []) + \
            (["--regularization", "l1"] if seed == 4 else [])
    o, r = _pair(raw, tmp_path, extra, sample_evidence=True, noise_aware=noise,
                 learn_non_evidence=seed % 3 == 0, reg_l1=seed == 4, reg_param=0.05)
    _check_structure(o, r)
    # potentials on random assignments and weights
    for _ in range(5):
        a = np.array([rng.integers(0, max(1, int(c))) for c in np.where(o.dtype == 0, 2, o.card)], np.uint64)
        w = rng.normal(0, 1, o.nweight)
        o.assign_evid[:] = a; o.weights[:] = w
        r.set_assignments(1, a); r.set_weights(w)
        for v in range(o.nvar):
            for p in range(int(2 if o.dtype[v] == 0 else o.card[v])):
                assert o.potential(v, p) == r.L.ref_potential(r.h, v, p, 1)
    # sweeps with the same erand48 seed
    s = [3, 1000 + seed, 77]
    oseed = ol.Seed3(*s)
    r.L.ref_set_seed(r.h, *s)
    for ep in range(4):
        o.sweep_learn(0.05, oseed)
        r.L.ref_sweep_learn(r.h, 0.05)
        assert np.array_equal(o.weights, r.get_weights())
        assert np.array_equal(o.assign_free, r.get_assignments(0))
        assert np.array_equal(o.assign_evid, r.get_assignments(1))
    o.clear_tallies(); r.L.ref_clear_tallies(r.h)
    for ep in range(4):
        o.sweep_infer(oseed)
        r.L.ref_sweep_infer(r.h)
    t, n = r.get_tallies()
    assert np.array_equal(o.tallies, t) and np.array_equal(o.nsamples, n)
    assert np.array_equal(o.assign_evid, r.get_assignments(1))
