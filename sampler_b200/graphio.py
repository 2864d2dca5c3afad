"""Factor-graph records and the DimmWitted on-disk format, in numpy.

Input-side tooling only (synthetic graph generators, fixture writers, result-file
parsers): the product's loader is C++ (sampler_b200/csrc/host/loader.cc).  The
format is the reference's doc/binary_format.md (big-endian records; reader
src/binary_format.cc:48-226, writer src/text2bin.cc:19-236).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional

import numpy as np

# src/common.h:33-46
FUNC_IMPLY_NATURAL = 0
FUNC_OR = 1
FUNC_AND = 2
FUNC_EQUAL = 3
FUNC_ISTRUE = 4
FUNC_LINEAR = 7
FUNC_RATIO = 8
FUNC_LOGICAL = 9
FUNC_AND_CATEGORICAL = 12
FUNC_IMPLY_MLN = 13
ALL_FUNCS = (0, 1, 2, 3, 4, 7, 8, 9, 12, 13)


@dataclasses.dataclass
class RawGraph:
    """Decoded records, ids dense and in id order (src/factor_graph.cc:210-218)."""

    var_role: np.ndarray   # u8  [nvar]   0 query, >=1 evidence (binary_format.cc:107)
    var_init: np.ndarray   # u64 [nvar]   initial (sparse) value
    var_dtype: np.ndarray  # u16 [nvar]   0 Boolean, 1 categorical
    var_card: np.ndarray   # u64 [nvar]
    dom_vid: np.ndarray    # u64 [ndom]
    dom_ptr: np.ndarray    # u64 [ndom+1]
    dom_val: np.ndarray    # u64 [sum card]
    dom_truth: np.ndarray  # f64 [sum card]
    fac_func: np.ndarray   # i32 [nfac]
    fac_ptr: np.ndarray    # u64 [nfac+1]  CSR over edges
    edge_vid: np.ndarray   # u64 [nedge]
    edge_eq: np.ndarray    # u64 [nedge]  ORIGINAL (sparse) equal_to
    fac_wid: np.ndarray    # u64 [nfac]
    fac_fv: np.ndarray     # f64 [nfac]
    w_init: np.ndarray     # f64 [nweight]
    w_fixed: np.ndarray    # u8  [nweight]

    @property
    def nvar(self) -> int:
        return int(self.var_role.shape[0])

    @property
    def nfac(self) -> int:
        return int(self.fac_func.shape[0])

    @property
    def nedge(self) -> int:
        return int(self.edge_vid.shape[0])

    @property
    def nweight(self) -> int:
        return int(self.w_init.shape[0])

    def var_rec(self) -> np.ndarray:
        """[nvar,4] u64 = role, init, dtype, card (the oracle's input layout)."""
        r = np.empty((self.nvar, 4), dtype=np.uint64)
        r[:, 0] = self.var_role
        r[:, 1] = self.var_init
        r[:, 2] = self.var_dtype
        r[:, 3] = self.var_card
        return r

    def normalized(self) -> "RawGraph":
        c = lambda a, t: np.ascontiguousarray(a, dtype=t)
        return RawGraph(
            c(self.var_role, np.uint8), c(self.var_init, np.uint64),
            c(self.var_dtype, np.uint16), c(self.var_card, np.uint64),
            c(self.dom_vid, np.uint64), c(self.dom_ptr, np.uint64),
            c(self.dom_val, np.uint64), c(self.dom_truth, np.float64),
            c(self.fac_func, np.int32), c(self.fac_ptr, np.uint64),
            c(self.edge_vid, np.uint64), c(self.edge_eq, np.uint64),
            c(self.fac_wid, np.uint64), c(self.fac_fv, np.float64),
            c(self.w_init, np.float64), c(self.w_fixed, np.uint8))


def empty_domains():
    z = np.zeros(0, dtype=np.uint64)
    return z, np.zeros(1, dtype=np.uint64), z.copy(), np.zeros(0, dtype=np.float64)


# --------------------------------------------------------------------------
# binary writers (vectorised; these files feed the reference `dw` and ours)
# --------------------------------------------------------------------------

def _be_bytes(a: np.ndarray, dt: str) -> np.ndarray:
    b = np.ascontiguousarray(a).astype(dt)
    return b.view(np.uint8).reshape(b.shape[0], -1)


def encode_weights(w_init, w_fixed, ids=None) -> bytes:
    n = len(w_init)
    rec = np.zeros(n, dtype=[("id", ">u8"), ("fixed", "u1"), ("init", ">f8")])
    rec["id"] = np.arange(n, dtype=np.uint64) if ids is None else ids
    rec["fixed"] = w_fixed
    rec["init"] = w_init
    assert rec.dtype.itemsize == 17
    return rec.tobytes()


def encode_variables(role, init, dtype, card, ids=None) -> bytes:
    n = len(role)
    rec = np.zeros(n, dtype=[("id", ">u8"), ("role", "u1"), ("init", ">u8"),
                             ("dtype", ">u2"), ("card", ">u8")])
    rec["id"] = np.arange(n, dtype=np.uint64) if ids is None else ids
    rec["role"] = role
    rec["init"] = init
    rec["dtype"] = dtype
    rec["card"] = card
    assert rec.dtype.itemsize == 27
    return rec.tobytes()


def encode_domains(dom_vid, dom_ptr, dom_val, dom_truth) -> bytes:
    out = bytearray()
    for i in range(len(dom_vid)):
        a, b = int(dom_ptr[i]), int(dom_ptr[i + 1])
        out += np.array([dom_vid[i], b - a], dtype=">u8").tobytes()
        rec = np.zeros(b - a, dtype=[("v", ">u8"), ("t", ">f8")])
        rec["v"] = dom_val[a:b]
        rec["t"] = dom_truth[a:b]
        out += rec.tobytes()
    return bytes(out)


def encode_factors(func, fac_ptr, edge_vid, edge_eq, wid, fv) -> bytes:
    """{u16 func, u64 arity, arity x {u64 vid, u64 eq}, u64 wid, f64 fv} per factor."""
    nfac = len(func)
    fac_ptr = np.asarray(fac_ptr, dtype=np.int64)
    arity = np.diff(fac_ptr)
    size = 2 + 8 + 16 * arity + 8 + 8
    off = np.zeros(nfac + 1, dtype=np.int64)
    np.cumsum(size, out=off[1:])
    buf = np.zeros(int(off[-1]), dtype=np.uint8)

    def put(pos, mat):
        idx = pos[:, None] + np.arange(mat.shape[1], dtype=np.int64)[None, :]
        buf[idx.ravel()] = mat.ravel()

    base = off[:-1]
    put(base, _be_bytes(np.asarray(func).astype(np.uint16), ">u2"))
    put(base + 2, _be_bytes(arity.astype(np.uint64), ">u8"))
    if len(edge_vid):
        fac_of_edge = np.repeat(np.arange(nfac, dtype=np.int64), arity)
        pos_in_fac = np.arange(len(edge_vid), dtype=np.int64) - fac_ptr[fac_of_edge]
        epos = base[fac_of_edge] + 10 + 16 * pos_in_fac
        put(epos, _be_bytes(np.asarray(edge_vid, dtype=np.uint64), ">u8"))
        put(epos + 8, _be_bytes(np.asarray(edge_eq, dtype=np.uint64), ">u8"))
    tail = base + 10 + 16 * arity
    put(tail, _be_bytes(np.asarray(wid, dtype=np.uint64), ">u8"))
    put(tail + 8, _be_bytes(np.asarray(fv, dtype=np.float64), ">f8"))
    return buf.tobytes()


def write_binary(g: RawGraph, outdir: str, prefix: str = "graph") -> Dict[str, str]:
    """Write meta/variables/domains/factors/weights; returns the paths."""
    os.makedirs(outdir, exist_ok=True)
    g = g.normalized()
    p = {k: os.path.join(outdir, f"{prefix}.{k}") for k in
         ("meta", "variables", "domains", "factors", "weights")}
    with open(p["meta"], "w") as f:  # src/binary_format.cc:23-36
        f.write(f"{g.nweight},{g.nvar},{g.nfac},{g.nedge}\n")
    with open(p["variables"], "wb") as f:
        f.write(encode_variables(g.var_role, g.var_init, g.var_dtype, g.var_card))
    with open(p["domains"], "wb") as f:
        f.write(encode_domains(g.dom_vid, g.dom_ptr, g.dom_val, g.dom_truth))
    with open(p["factors"], "wb") as f:
        f.write(encode_factors(g.fac_func, g.fac_ptr, g.edge_vid, g.edge_eq,
                               g.fac_wid, g.fac_fv))
    with open(p["weights"], "wb") as f:
        f.write(encode_weights(g.w_init, g.w_fixed))
    return p


# --------------------------------------------------------------------------
# binary readers (for fixtures; the product reads files in C++)
# --------------------------------------------------------------------------

def decode_weights(b: bytes):
    rec = np.frombuffer(b, dtype=[("id", ">u8"), ("fixed", "u1"), ("init", ">f8")])
    return rec["id"].astype(np.uint64), rec["fixed"].astype(np.uint8), rec["init"].astype(np.float64)


def decode_variables(b: bytes):
    rec = np.frombuffer(b, dtype=[("id", ">u8"), ("role", "u1"), ("init", ">u8"),
                                  ("dtype", ">u2"), ("card", ">u8")])
    return (rec["id"].astype(np.uint64), rec["role"].astype(np.uint8),
            rec["init"].astype(np.uint64), rec["dtype"].astype(np.uint16),
            rec["card"].astype(np.uint64))


def decode_domains(b: bytes):
    vids, ptr, vals, truth = [], [0], [], []
    pos = 0
    while pos < len(b):
        vid, card = np.frombuffer(b, dtype=">u8", count=2, offset=pos)
        pos += 16
        rec = np.frombuffer(b, dtype=[("v", ">u8"), ("t", ">f8")], count=int(card), offset=pos)
        pos += 16 * int(card)
        vids.append(int(vid))
        vals.extend(rec["v"].tolist())
        truth.extend(rec["t"].tolist())
        ptr.append(len(vals))
    return (np.array(vids, dtype=np.uint64), np.array(ptr, dtype=np.uint64),
            np.array(vals, dtype=np.uint64), np.array(truth, dtype=np.float64))


def decode_factors(b: bytes):
    func, ptr, evid, eeq, wid, fv = [], [0], [], [], [], []
    pos = 0
    while pos < len(b):
        f = int(np.frombuffer(b, dtype=">u2", count=1, offset=pos)[0])
        n = int(np.frombuffer(b, dtype=">u8", count=1, offset=pos + 2)[0])
        pos += 10
        e = np.frombuffer(b, dtype=">u8", count=2 * n, offset=pos)
        pos += 16 * n
        evid.extend(e[0::2].tolist())
        eeq.extend(e[1::2].tolist())
        w = int(np.frombuffer(b, dtype=">u8", count=1, offset=pos)[0])
        v = float(np.frombuffer(b, dtype=">f8", count=1, offset=pos + 8)[0])
        pos += 16
        func.append(f)
        ptr.append(len(evid))
        wid.append(w)
        fv.append(v)
    return (np.array(func, dtype=np.int32), np.array(ptr, dtype=np.uint64),
            np.array(evid, dtype=np.uint64), np.array(eeq, dtype=np.uint64),
            np.array(wid, dtype=np.uint64), np.array(fv, dtype=np.float64))


def read_meta(path: str):
    """numWeights,numVariables,numFactors,numEdges[,ignored...] (binary_format.cc:23-36)."""
    with open(path) as f:
        parts = f.read().split(",")
    return tuple(int(x.strip() or 0) for x in parts[:4])


def read_binary(variables: List[str], domains: List[str], factors: List[str],
                weights: List[str]) -> RawGraph:
    """Concatenate the given files per kind (the e2e harness does `cat graph.x*`,
    test/run_end_to_end.sh:18-23) and decode; records are placed by id."""
    cat = lambda ps: b"".join(open(p, "rb").read() for p in ps)
    vid, role, init, dt, card = decode_variables(cat(variables))
    order = np.argsort(vid, kind="stable")
    assert np.array_equal(vid[order], np.arange(len(vid), dtype=np.uint64)), "variable ids not dense"
    wid_, wfix, winit = decode_weights(cat(weights))
    worder = np.argsort(wid_, kind="stable")
    assert np.array_equal(wid_[worder], np.arange(len(wid_), dtype=np.uint64)), "weight ids not dense"
    dv, dp, dval, dtr = decode_domains(cat(domains)) if domains else empty_domains()
    ff, fp, ev, ee, fw, fv = decode_factors(cat(factors))
    return RawGraph(role[order], init[order], dt[order], card[order], dv, dp, dval, dtr,
                    ff, fp, ev, ee, fw, fv, winit[worder], wfix[worder]).normalized()


# --------------------------------------------------------------------------
# result files (src/inference_result.cc:101-105, 211-243)
# --------------------------------------------------------------------------

def read_weights_text(path: str) -> np.ndarray:
    out = {}
    with open(path) as f:
        for line in f:
            if line.strip():
                a, b = line.split()
                out[int(a)] = float(b)
    return np.array([out[i] for i in range(len(out))], dtype=np.float64)


def read_marginals_text(path: str) -> Dict[tuple, float]:
    out = {}
    with open(path) as f:
        for line in f:
            if line.strip():
                a, b, c = line.split()
                out[(int(a), int(b))] = float(c)
    return out
