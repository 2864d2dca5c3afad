"""ctypes binding of the C ABI in include/dwg.h -- the stub a maintainer of the reference
would add on the reference side.  Thin by design: tests and bench.py
call the product exactly through the entry points a C/C++ host would.

There is no CPU fallback: if libdwg.so is missing or no CUDA device is present the calls
raise DwgError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _build
from .graphio import RawGraph

SCHED_AUTO, SCHED_HOGWILD, SCHED_CHROMATIC, SCHED_SEQUENTIAL = 0, 1, 2, 3


class DwgError(RuntimeError):
    pass


class Opts(C.Structure):
    _fields_ = [("sample_evidence", C.c_int), ("learn_non_evidence", C.c_int),
                ("noise_aware", C.c_int), ("regularization", C.c_int),
                ("reg_param", C.c_double), ("schedule", C.c_int), ("seed", C.c_uint64),
                ("replica", C.c_uint32), ("learn_chunks", C.c_int), ("defer_upload", C.c_int)]


class Raw(C.Structure):
    _fields_ = [("nvar", C.c_uint64), ("ndom", C.c_uint64), ("nfac", C.c_uint64), ("nweight", C.c_uint64),
                ("var_role", C.c_void_p), ("var_init", C.c_void_p), ("var_dtype", C.c_void_p),
                ("var_card", C.c_void_p), ("dom_vid", C.c_void_p), ("dom_ptr", C.c_void_p),
                ("dom_val", C.c_void_p), ("dom_truth", C.c_void_p), ("fac_func", C.c_void_p),
                ("fac_ptr", C.c_void_p), ("edge_vid", C.c_void_p), ("edge_eq", C.c_void_p),
                ("fac_wid", C.c_void_p), ("fac_fv", C.c_void_p), ("w_init", C.c_void_p),
                ("w_fixed", C.c_void_p)]


class GraphInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("num_variables", "num_factors", "num_weights", "num_edges", "num_values",
                 "num_variables_evidence", "num_variables_query", "num_index_entries",
                 "num_other_refs")] + [
        ("num_colors", C.c_uint32), ("schedule", C.c_int), ("has_categorical", C.c_int),
        ("has_multi_arity", C.c_int), ("fv_mode", C.c_int), ("assign_bytes", C.c_int),
        ("device_bytes", C.c_uint64), ("bytes_per_sweep_infer", C.c_double),
        ("bytes_per_sweep_infer_query", C.c_double), ("bytes_per_sweep_learn", C.c_double),
        ("num_device_rows", C.c_uint64)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build: bool = False):
    """dlopen libdwg.so (optionally building it first).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        _build.build()
    if not os.path.exists(_build.LIB):
        raise DwgError(f"{_build.LIB} is missing: run `python -m sampler_b200._build` "
                       "(there is no CPU fallback for the Gibbs sweep)")
    L = C.CDLL(_build.LIB)
    vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_double
    L.dwg_last_error.restype = C.c_char_p
    L.dwg_version.restype = C.c_char_p
    L.dwg_launch_count.restype = u64
    L.dwg_opts_default.argtypes = [C.POINTER(Opts)]
    L.dwg_graph_from_raw.argtypes = [C.POINTER(Raw), i32, C.POINTER(vp)]
    L.dwg_graph_from_files.argtypes = [C.c_char_p] + [i32, C.POINTER(C.c_char_p)] * 4 + [i32, C.POINTER(vp)]
    L.dwg_graph_free.argtypes = [vp]
    L.dwg_graph_set_page_locking.argtypes = [vp, i32]
    L.dwg_graph_info.argtypes = [vp, C.POINTER(GraphInfo)]
    L.dwg_graph_var_meta.argtypes = [vp, vp, vp, vp, vp]
    L.dwg_debug_schedule.argtypes = [vp, vp, vp]
    L.dwg_graph_synthetic_unary.argtypes = [u64, C.c_uint32, u64, dbl, i32, dbl, u64, vp, C.POINTER(vp)]
    L.dwg_debug_learn_order.argtypes = [vp, vp, vp]
    L.dwg_debug_image.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(u64), C.POINTER(C.c_uint32)]
    L.dwg_create.argtypes = [vp, C.POINTER(Opts), i32, C.POINTER(vp)]
    L.dwg_destroy.argtypes = [vp]
    L.dwg_release_cached_memory.argtypes = [i32]
    L.dwg_graph_segments.argtypes = [vp, i32, vp, vp, C.POINTER(i32)]
    L.dwg_reset_state.argtypes = [vp]
    L.dwg_learn_epoch.argtypes = [vp, dbl]
    L.dwg_infer_epoch.argtypes = [vp]
    L.dwg_infer_epochs.argtypes = [vp, u64]
    L.dwg_learn_epochs.argtypes = [vp, u64, dbl, dbl]
    L.dwg_sync.argtypes = [vp]
    L.dwg_stream.restype = vp
    L.dwg_stream.argtypes = [vp]
    L.dwg_clear_tallies.argtypes = [vp]
    L.dwg_get_tallies.argtypes = [vp, vp, vp]
    L.dwg_get_tallies32.argtypes = [vp, vp, vp]
    L.dwg_add_tallies.argtypes = [vp, vp, vp]
    L.dwg_get_weights.argtypes = [vp, vp]
    L.dwg_set_weights.argtypes = [vp, vp]
    L.dwg_set_weight_sum.argtypes = [vp, vp]
    L.dwg_get_assignments.argtypes = [vp, i32, vp]
    L.dwg_set_assignments.argtypes = [vp, i32, vp]
    L.dwg_weights_device_ptr.restype = vp
    L.dwg_weights_device_ptr.argtypes = [vp]
    L.dwg_tallies_device_ptr.restype = vp
    L.dwg_tallies_device_ptr.argtypes = [vp]
    L.dwg_tallies_pack.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u64)]
    L.dwg_tallies_unpack.argtypes = [vp, i32]
    L.dwg_finish_weight_average.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.dwg_average_weights_nccl.argtypes = [vp, vp, i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.dwg_eval_potentials.argtypes = [vp, u64, vp, vp, i32, vp]
    L.dwg_draw.argtypes = [i32, u64, vp, u64, vp, vp, vp]
    L.dwg_logadd.argtypes = [i32, u64, vp, vp, vp]
    L.dwg_debug_set_uniforms.argtypes = [vp, vp, vp]
    L.dwg_debug_set_sequential.argtypes = [vp, i32]
    L.dwg_dw_main.argtypes = [i32, C.POINTER(C.c_char_p)]
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise DwgError("dwg error %d: %s" % (rc, load().dwg_last_error().decode()))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Graph:
    """dwg_graph: a factor graph compiled into the device layout
    (== FactorGraph after load + construct_index, src/dimmwitted.cc:63-72)."""

    def __init__(self, handle):
        self._h = handle
        self.info = GraphInfo()
        _check(load().dwg_graph_info(self._h, C.byref(self.info)))

    @classmethod
    def from_raw(cls, raw: RawGraph, schedule: int = SCHED_AUTO) -> "Graph":
        L = load()
        r = raw.normalized()
        keep = [r.var_role, r.var_init, r.var_dtype, r.var_card, r.dom_vid, r.dom_ptr, r.dom_val,
                r.dom_truth, r.fac_func, r.fac_ptr, r.edge_vid, r.edge_eq, r.fac_wid, r.fac_fv,
                r.w_init, r.w_fixed]
        c = Raw(r.nvar, len(r.dom_vid), r.nfac, r.nweight, *[_ptr(a) for a in keep])
        out = C.c_void_p()
        _check(L.dwg_graph_from_raw(C.byref(c), schedule, C.byref(out)))
        return cls(out)

    @classmethod
    def synthetic_unary(cls, nvar: int, degree: int, nweight: int, zipf: float = 0.0, real_features: bool = False,
                        p_evid: float = 0.5, seed: int = 1234, want_cdf: bool = False):
        """the compiled image of graphgen.synth_unary(...) streamed by the library (no records held):
        BASELINE.json configs[3] / configs[4] at full size"""
        out = C.c_void_p()
        cdf = np.zeros(nweight if (want_cdf and zipf > 0) else 0, np.float64)
        _check(load().dwg_graph_synthetic_unary(nvar, degree, nweight, float(zipf), int(real_features), float(p_evid),
                                                seed, _ptr(cdf) if len(cdf) else None, C.byref(out)))
        g = cls(out)
        return (g, cdf) if want_cdf else g

    @classmethod
    def from_files(cls, meta: str, variables: Sequence[str], domains: Sequence[str],
                   factors: Sequence[str], weights: Sequence[str], schedule: int = SCHED_AUTO) -> "Graph":
        L = load()

        def arr(xs):
            xs = [x.encode() for x in xs]
            return len(xs), (C.c_char_p * max(1, len(xs)))(*xs)
        out = C.c_void_p()
        _check(L.dwg_graph_from_files(meta.encode(), *arr(variables), *arr(domains), *arr(factors),
                                      *arr(weights), schedule, C.byref(out)))
        return cls(out)

    def __del__(self):
        try:
            if self._h:
                load().dwg_graph_free(self._h)
                self._h = None
        except Exception:
            pass

    def var_meta(self):
        i = self.info
        dtype = np.zeros(i.num_variables, np.uint8)
        is_evid = np.zeros(i.num_variables, np.uint8)
        base = np.zeros(i.num_variables + 1, np.uint64)
        values = np.zeros(i.num_values, np.uint64)
        _check(load().dwg_graph_var_meta(self._h, _ptr(dtype), _ptr(is_evid), _ptr(base), _ptr(values)))
        return dtype, is_evid, base, values

    _IMG_DTYPES = {"e_fv32": np.float32, "e_fv64": np.float64, "s_fv32": np.float32, "s_fv64": np.float64,
                   "truthiness": np.float64, "w_init": np.float64, "cls_N_evid": np.float32,
                   "cls_Q_evid": np.float32, "cls_N_all": np.float32, "cls_Q_all": np.float32,
                   "cls_H_evid": np.float32, "cls_H_all": np.float32}

    def image(self, name: str) -> np.ndarray:
        """copy of one array of the compiled device image (tests)"""
        data, count, es = C.c_void_p(), C.c_uint64(), C.c_uint32()
        _check(load().dwg_debug_image(self._h, name.encode(), C.byref(data), C.byref(count), C.byref(es)))
        dt = self._IMG_DTYPES.get(name) or {1: np.uint8, 4: np.uint32, 8: np.uint64}[es.value]
        if count.value == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (count.value * es.value)).from_address(data.value)
        return np.frombuffer(buf, dtype=dt).copy()

    def schedule(self):
        perm = np.zeros(self.info.num_variables, np.uint64)
        cs = np.zeros(self.info.num_colors + 1, np.uint64)
        _check(load().dwg_debug_schedule(self._h, _ptr(perm), _ptr(cs)))
        return perm, cs

    def learn_order(self) -> np.ndarray:
        """original ids in the order a learning sweep of the parallel schedules visits them"""
        perm = np.zeros(self.info.num_variables, np.uint64)
        n = C.c_uint64()
        _check(load().dwg_debug_learn_order(self._h, _ptr(perm), C.byref(n)))
        assert n.value == len(perm)
        return perm


class Sampler:
    """dwg_handle: one replica on one GPU (== GibbsSampler + InferenceResult,
    src/gibbs_sampler.h:18-57, src/inference_result.h:27-43)."""

    def __init__(self, graph: Graph, device: int = 0, sample_evidence=False, learn_non_evidence=False,
                 noise_aware=False, regularization="l2", reg_param=0.01, seed=0x5eed5eed, replica=0, learn_chunks=0,
                 defer_upload=False):
        L = load()
        self.graph = graph
        self.opts = Opts(int(sample_evidence), int(learn_non_evidence), int(noise_aware),
                         1 if regularization == "l1" else 0, float(reg_param), 0, int(seed), int(replica), int(learn_chunks),
                         int(defer_upload))
        self._h = C.c_void_p()
        _check(L.dwg_create(graph._h, C.byref(self.opts), device, C.byref(self._h)))
        self.device = device
        self.nvar = graph.info.num_variables
        self.nval = graph.info.num_values
        self.nweight = graph.info.num_weights
        self.ndrow = graph.info.num_device_rows

    def close(self):
        if self._h:
            load().dwg_destroy(self._h)
            self._h = None

    def graph_segments(self):
        """[(device pointer, bytes)] of the replica's copy of the graph, in a graph-determined order"""
        n = C.c_int()
        _check(load().dwg_graph_segments(self._h, 0, None, None, C.byref(n)))
        ptrs = (C.c_void_p * max(1, n.value))()
        sizes = (C.c_uint64 * max(1, n.value))()
        _check(load().dwg_graph_segments(self._h, n.value, ptrs, sizes, C.byref(n)))
        return [(int(ptrs[i] or 0), int(sizes[i])) for i in range(n.value)]

    def reset_state(self):
        _check(load().dwg_reset_state(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def learn_epoch(self, stepsize: float):
        _check(load().dwg_learn_epoch(self._h, stepsize))

    def learn_epochs(self, n: int, stepsize: float, decay: float):
        """n learning sweeps, stepsize *= decay after each (one launch under the sequential schedule)"""
        _check(load().dwg_learn_epochs(self._h, n, stepsize, decay))

    def infer_epoch(self):
        _check(load().dwg_infer_epoch(self._h))

    def infer_epochs(self, n: int):
        _check(load().dwg_infer_epochs(self._h, n))

    def sync(self):
        _check(load().dwg_sync(self._h))

    def stream(self) -> int:
        return load().dwg_stream(self._h) or 0

    def clear_tallies(self):
        _check(load().dwg_clear_tallies(self._h))

    def tallies(self, wide: bool = True, out=None):
        """(tallies[num_values], nsamples[num_variables]); wide=False: u32 tallies (the device's
        counter width, no widening pass on the host).  out = (t, n): arrays to fill (a page-locked u32
        t takes the DMA directly)"""
        if out is not None:
            t, n = out
            wide = t.dtype == np.uint64
            assert t.shape[0] == self.nval and n.shape[0] == self.nvar and n.dtype == np.uint64
        else:
            t = np.empty(self.nval, np.uint64 if wide else np.uint32)
            n = np.empty(self.nvar, np.uint64)
        _check((load().dwg_get_tallies if wide else load().dwg_get_tallies32)(self._h, _ptr(t), _ptr(n)))
        return t, n

    def add_tallies(self, t, n):
        t = np.ascontiguousarray(t, np.uint64); n = np.ascontiguousarray(n, np.uint64)
        _check(load().dwg_add_tallies(self._h, _ptr(t), _ptr(n)))

    def weights(self):
        w = np.zeros(self.nweight, np.float64)
        _check(load().dwg_get_weights(self._h, _ptr(w)))
        return w

    def set_weights(self, w):
        w = np.ascontiguousarray(w, np.float64)
        assert w.shape[0] == self.nweight
        _check(load().dwg_set_weights(self._h, _ptr(w)))

    def set_weights_sum(self, w_sum):
        """store the cross-replica SUM (host exchange); follow with finish_weight_average(n)"""
        w = np.ascontiguousarray(w_sum, np.float64)
        assert w.shape[0] == self.nweight
        _check(load().dwg_set_weight_sum(self._h, _ptr(w)))

    def assignments(self, evid_chain: int):
        a = np.zeros(self.nvar, np.uint64)
        _check(load().dwg_get_assignments(self._h, evid_chain, _ptr(a)))
        return a

    def set_assignments(self, evid_chain: int, a):
        a = np.ascontiguousarray(a, np.uint64)
        assert a.shape[0] == self.nvar
        _check(load().dwg_set_assignments(self._h, evid_chain, _ptr(a)))

    def weights_device_ptr(self) -> int:
        return load().dwg_weights_device_ptr(self._h) or 0

    def tallies_device_ptr(self) -> int:
        return load().dwg_tallies_device_ptr(self._h) or 0

    def tallies_pack(self, fields: int):
        """(device pointer, words): the tallies as `fields` narrow counters per u32 word (dwg.h)"""
        p, n = C.c_void_p(), C.c_uint64()
        _check(load().dwg_tallies_pack(self._h, int(fields), C.byref(p), C.byref(n)))
        return p.value or 0, int(n.value)

    def tallies_unpack(self, fields: int):
        _check(load().dwg_tallies_unpack(self._h, int(fields)))

    def finish_weight_average(self, n_replicas: int, want_norms=True):
        a, b = C.c_double(), C.c_double()
        _check(load().dwg_finish_weight_average(self._h, n_replicas, C.byref(a) if want_norms else None,
                                                C.byref(b) if want_norms else None))
        return a.value, b.value

    def eval_potentials(self, vids, proposals, evid_chain=1):
        vids = np.ascontiguousarray(vids, np.uint64); proposals = np.ascontiguousarray(proposals, np.uint64)
        out = np.zeros(len(vids), np.float64)
        _check(load().dwg_eval_potentials(self._h, len(vids), _ptr(vids), _ptr(proposals), evid_chain, _ptr(out)))
        return out

    def set_uniforms(self, r_free, r_evid=None):
        if r_free is None:
            _check(load().dwg_debug_set_uniforms(self._h, None, None))
            return
        a = np.ascontiguousarray(r_free, np.float64)
        b = None if r_evid is None else np.ascontiguousarray(r_evid, np.float64)
        _check(load().dwg_debug_set_uniforms(self._h, _ptr(a), _ptr(b)))

    def set_sequential(self, on: bool):
        _check(load().dwg_debug_set_sequential(self._h, int(on)))

    def marginals(self):
        """{(vid, sparse value): prob} as dump_marginals_in_text prints them
        (src/inference_result.cc:211-243)."""
        dtype, is_evid, base, values = self.graph.var_meta()
        t, n = self.tallies()
        out = {}
        for v in range(self.nvar):
            if is_evid[v] and not self.opts.sample_evidence:
                continue
            b = int(base[v])
            if dtype[v] == 0:
                out[(v, 1)] = float(t[b]) / float(n[v])
            else:
                for j in range(int(base[v + 1]) - b):
                    out[(v, int(values[b + j]))] = float(t[b + j]) / float(n[v])
        return out


def release_cached_memory(device: int = 0) -> None:
    """return the device memory destroyed samplers left in the library's pool to the driver"""
    _check(load().dwg_release_cached_memory(int(device)))


def draw(pots: np.ndarray, card: np.ndarray, r: np.ndarray, device: int = 0) -> np.ndarray:
    pots = np.ascontiguousarray(pots, np.float64)
    card = np.ascontiguousarray(card, np.uint32)
    r = np.ascontiguousarray(r, np.float64)
    out = np.zeros(len(r), np.uint32)
    _check(load().dwg_draw(device, len(r), _ptr(pots), pots.shape[1], _ptr(card), _ptr(r), _ptr(out)))
    return out


def logadd(a: np.ndarray, b: np.ndarray, device: int = 0) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    out = np.zeros(len(a), np.float64)
    _check(load().dwg_logadd(device, len(a), _ptr(a), _ptr(b), _ptr(out)))
    return out


def last_error() -> str:
    return load().dwg_last_error().decode()


def dw_main(argv: Sequence[str]) -> int:
    """the whole `dw` CLI in-process (== dd::dw, src/dimmwitted.cc:20-35)"""
    args = [a.encode() for a in argv]
    arr = (C.c_char_p * len(args))(*args)
    return load().dwg_dw_main(len(args), arr)
