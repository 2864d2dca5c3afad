"""Synthetic factor graphs of the shapes BASELINE.json names (SURVEY.md 8(d) "Concrete
synthetic inputs"), as decoded records (RawGraph).  numpy PCG64, graph seed 1234 by default.
Sizes are parameters so that the same generator feeds the full-size bench, the scaled-down
CPU baseline and the parity tests.  Input tooling, not part of the sweep.
"""
from __future__ import annotations

import numpy as np

from .graphio import (FUNC_AND_CATEGORICAL, FUNC_EQUAL, FUNC_IMPLY_NATURAL, FUNC_ISTRUE, RawGraph,
                      empty_domains)

DEFAULT_SEED = 1234


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def c2_boolean(nvar: int = 10_000_000, n_unary: int = 8, nweight: int = 1000, p_evid: float = 0.5,
               seed: int = DEFAULT_SEED, planted_weights: bool = False) -> RawGraph:
    """configs[1]: Boolean graph, per variable `n_unary` unary ISTRUE (equal_to 1) + one
    IMPLY_NATURAL(u => v) with a uniform random body u != v; mean incident degree n_unary + 2;
    `nweight` shared weights, weight ids uniform; 50 % evidence labelled by the planted unary
    model w* ~ N(0, 0.5^2); feature value 1.0.  Factor order: variable by variable."""
    rng = np.random.default_rng(seed)
    per = n_unary + 1
    nfac = nvar * per
    wid = rng.integers(0, nweight, (nvar, per), dtype=np.uint64)
    wstar = rng.normal(0.0, 0.5, nweight)
    func = np.empty((nvar, per), np.int32)
    func[:, :n_unary] = FUNC_ISTRUE
    func[:, n_unary] = FUNC_IMPLY_NATURAL
    v = np.arange(nvar, dtype=np.uint64)
    u = (v + rng.integers(1, max(2, nvar), nvar, dtype=np.uint64)) % np.uint64(nvar)
    if nvar == 1:
        u[:] = 0
    edge_vid = np.empty((nvar, n_unary + 2), np.uint64)
    edge_vid[:, :n_unary] = v[:, None]
    edge_vid[:, n_unary] = u
    edge_vid[:, n_unary + 1] = v
    arity = np.ones((nvar, per), np.uint64)
    arity[:, n_unary] = 2
    fac_ptr = np.zeros(nfac + 1, np.uint64)
    np.cumsum(arity.ravel(), out=fac_ptr[1:])
    logit = 2.0 * wstar[wid[:, :n_unary].astype(np.int64)].sum(axis=1)
    label = (rng.random(nvar) < _sigmoid(logit)).astype(np.uint64)
    role = (rng.random(nvar) < p_evid).astype(np.uint8)
    return RawGraph(role, np.where(role > 0, label, 0).astype(np.uint64), np.zeros(nvar, np.uint16),
                    np.full(nvar, 2, np.uint64), *empty_domains(), func.ravel(), fac_ptr,
                    edge_vid.ravel(), np.ones(edge_vid.size, np.uint64), wid.ravel(),
                    np.ones(nfac), wstar if planted_weights else np.zeros(nweight),
                    np.zeros(nweight, np.uint8)).normalized()


def c3_categorical(nvar: int = 5_000_000, card: int = 10, nweight: int = 1000, p_evid: float = 0.5,
                   seed: int = DEFAULT_SEED) -> RawGraph:
    """configs[2]: categorical variables of cardinality `card` with sparse domains (a sorted random
    subset of [0, 1000), truthiness 0); per variable one unary AND_CATEGORICAL per value, two
    pairwise AND_CATEGORICAL on (v, v+1) with random value pairs and one EQUAL(v, v+1) on a random
    value pair; 50 % evidence."""
    rng = np.random.default_rng(seed)
    K = card
    # sorted random K-subset of [0,1000) per variable: K sorted distinct offsets via cumulative gaps
    gaps = rng.integers(1, max(2, 1000 // K), (nvar, K), dtype=np.uint64)
    dom = np.cumsum(gaps, axis=1, dtype=np.uint64) - np.uint64(1)
    dom_vid = np.arange(nvar, dtype=np.uint64)
    dom_ptr = np.arange(nvar + 1, dtype=np.uint64) * np.uint64(K)
    nxt = (np.arange(nvar, dtype=np.int64) + 1) % nvar
    per = K + 3
    func = np.empty((nvar, per), np.int32)
    func[:, :K + 2] = FUNC_AND_CATEGORICAL
    func[:, K + 2] = FUNC_EQUAL
    arity = np.ones((nvar, per), np.uint64)
    arity[:, K:] = 2
    fac_ptr = np.zeros(nvar * per + 1, np.uint64)
    np.cumsum(arity.ravel(), out=fac_ptr[1:])
    ne = K + 6
    edge_vid = np.empty((nvar, ne), np.uint64)
    edge_eq = np.empty((nvar, ne), np.uint64)
    v = np.arange(nvar, dtype=np.uint64)
    edge_vid[:, :K] = v[:, None]
    edge_eq[:, :K] = dom
    rows = np.arange(nvar)
    for j in range(3):
        a = rng.integers(0, K, nvar)
        b = rng.integers(0, K, nvar)
        edge_vid[:, K + 2 * j] = v
        edge_vid[:, K + 2 * j + 1] = nxt.astype(np.uint64)
        edge_eq[:, K + 2 * j] = dom[rows, a]
        edge_eq[:, K + 2 * j + 1] = dom[nxt, b]
    wid = rng.integers(0, nweight, (nvar, per), dtype=np.uint64)
    role = (rng.random(nvar) < p_evid).astype(np.uint8)
    init = dom[rows, rng.integers(0, K, nvar)]
    return RawGraph(role, np.where(role > 0, init, 0).astype(np.uint64), np.ones(nvar, np.uint16),
                    np.full(nvar, K, np.uint64), dom_vid, dom_ptr, dom.ravel(), np.zeros(nvar * K),
                    func.ravel(), fac_ptr, edge_vid.ravel(), edge_eq.ravel(), wid.ravel(),
                    np.ones(nvar * per), np.zeros(nweight), np.zeros(nweight, np.uint8)).normalized()


def unary_boolean(nvar: int, degree: int, nweight: int, p_evid: float = 0.5, zipf: float = 0.0,
                  real_features: bool = False, seed: int = DEFAULT_SEED) -> RawGraph:
    """configs[3]/[4] shape: Boolean variables x `degree` unary ISTRUE features.
    C4 = (5e7, 20, 1e6, zipf 1.1, real-valued f64 features, ~all evidence);
    C5 = (3e8, 10, 1000) per replica."""
    rng = np.random.default_rng(seed)
    nfac = nvar * degree
    if zipf > 0:
        # Zipf(s) over nweight ids by inverse-CDF on the truncated distribution
        ranks = np.arange(1, nweight + 1, dtype=np.float64)
        cdf = np.cumsum(ranks ** -zipf)
        cdf /= cdf[-1]
        wid = np.searchsorted(cdf, rng.random(nfac)).astype(np.uint64)
    else:
        wid = rng.integers(0, nweight, nfac, dtype=np.uint64)
    fv = rng.normal(0.0, 1.0, nfac) if real_features else np.ones(nfac)
    wstar = rng.normal(0.0, 0.5, nweight)
    logit = 2.0 * (wstar[wid.astype(np.int64)] * fv).reshape(nvar, degree).sum(axis=1)
    label = (rng.random(nvar) < _sigmoid(logit)).astype(np.uint64)
    role = (rng.random(nvar) < p_evid).astype(np.uint8)
    return RawGraph(role, np.where(role > 0, label, 0).astype(np.uint64), np.zeros(nvar, np.uint16),
                    np.full(nvar, 2, np.uint64), *empty_domains(), np.full(nfac, FUNC_ISTRUE, np.int32),
                    np.arange(nfac + 1, dtype=np.uint64), np.repeat(np.arange(nvar, dtype=np.uint64), degree),
                    np.ones(nfac, np.uint64), wid, fv, np.zeros(nweight),
                    np.zeros(nweight, np.uint8)).normalized()


def c4_logistic(nvar: int = 50_000_000, degree: int = 20, nweight: int = 1_000_000,
                seed: int = DEFAULT_SEED) -> RawGraph:
    return unary_boolean(nvar, degree, nweight, p_evid=0.9, zipf=1.1, real_features=True, seed=seed)


def c5_large(nvar: int = 300_000_000, degree: int = 10, nweight: int = 1000,
             seed: int = DEFAULT_SEED) -> RawGraph:
    return unary_boolean(nvar, degree, nweight, p_evid=0.5, seed=seed)


# --------------------------------------------------------------------------------------
# Counter-based generator (configs[3] / configs[4] at full size).  Every quantity is a pure
# function of (seed, stream, index) through the splitmix64 finaliser, so the C++ builder
# (csrc/host/synth.cc, dwg_graph_synthetic_unary) produces the SAME graph record by record
# without ever holding the records, and this numpy mirror feeds the parity tests and the
# reference's scaled-down CPU run.  Only +, * and comparisons on doubles (no libm calls) except
# the logistic label, so both sides agree bit for bit.
# --------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SYNTH_SQRT3 = 1.7320508075688772


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)"""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _hash(seed, stream, idx):
    with np.errstate(over="ignore"):
        base = _mix64(np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        return _mix64(base + idx.astype(np.uint64))


def _u01(h):
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def zipf_cdf(nweight, s):
    ranks = np.arange(1, nweight + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** -s)
    return cdf / cdf[-1]


def synth_unary(nvar: int, degree: int, nweight: int, zipf: float = 0.0, real_features: bool = False,
                p_evid: float = 0.5, seed: int = DEFAULT_SEED, cdf=None) -> RawGraph:
    """Boolean variables x `degree` unary ISTRUE features, counter-based (see above).
    C4 = (5e7, 20, 1e6, zipf 1.1, real-valued features, p_evid 0.9); C5 = (3e8, 10, 1000, p_evid 0.5).
    `cdf`: the Zipf table to use (tests pass the one the C++ builder computed, so that a last-bit
    difference between two pow() implementations cannot move a weight id)."""
    nfac = nvar * degree
    f = np.arange(nfac, dtype=np.uint64)
    v = np.arange(nvar, dtype=np.uint64)
    if zipf > 0:
        cdf = zipf_cdf(nweight, zipf) if cdf is None else cdf
        wid = np.minimum(np.searchsorted(cdf, _u01(_hash(seed, 3, f)), side="left"), nweight - 1).astype(np.uint64)
    else:
        wid = _hash(seed, 3, f) % np.uint64(nweight)
    if real_features:
        fv = (((_u01(_hash(seed, 4, f)) + _u01(_hash(seed, 5, f))) + _u01(_hash(seed, 6, f))) + _u01(_hash(seed, 7, f)) - 2.0) * SYNTH_SQRT3
    else:
        fv = np.ones(nfac)
    k = np.arange(nweight, dtype=np.uint64)
    wstar = 0.5 * ((((_u01(_hash(seed, 8, k)) + _u01(_hash(seed, 9, k))) + _u01(_hash(seed, 10, k))) + _u01(_hash(seed, 11, k)) - 2.0) * SYNTH_SQRT3)
    term = (wstar[wid.astype(np.int64)] * fv).reshape(nvar, degree)
    acc = np.zeros(nvar)
    for j in range(degree):  # sequential sum, like the builder
        acc = acc + term[:, j]
    logit = 2.0 * acc
    label = (_u01(_hash(seed, 2, v)) < 1.0 / (1.0 + np.exp(-logit))).astype(np.uint64)
    role = (_u01(_hash(seed, 1, v)) < p_evid).astype(np.uint8)
    return RawGraph(role, np.where(role > 0, label, 0).astype(np.uint64), np.zeros(nvar, np.uint16),
                    np.full(nvar, 2, np.uint64), *empty_domains(), np.full(nfac, FUNC_ISTRUE, np.int32),
                    np.arange(nfac + 1, dtype=np.uint64), np.repeat(v, degree), np.ones(nfac, np.uint64), wid, fv,
                    np.zeros(nweight), np.zeros(nweight, np.uint8)).normalized()


SYNTH_SPECS = {
    "c4": dict(nvar=50_000_000, degree=20, nweight=1_000_000, zipf=1.1, real_features=True, p_evid=0.9),
    "c5": dict(nvar=300_000_000, degree=10, nweight=1000, zipf=0.0, real_features=False, p_evid=0.5),
}
