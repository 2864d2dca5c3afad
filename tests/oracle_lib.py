"""ctypes bindings to the CPU oracle (oracle/libdworacle.so) and, when present, to
the reference-built checker (oracle/_ref/libdwref.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libdworacle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libdwref.so")
REF_DW = os.path.join(ORACLE_DIR, "_ref", "dw")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class OrcGraph(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("nvar", "nfac", "nweight", "nedge", "nval", "nindex")] + [
        ("dtype", C.POINTER(C.c_uint8)), ("is_evid", C.POINTER(C.c_uint8)),
        ("card", C.POINTER(C.c_uint64)), ("assignment_dense", C.POINTER(C.c_uint64)),
        ("var_val_base", C.POINTER(C.c_uint64)), ("total_truthiness", C.POINTER(C.c_double)),
        ("value", C.POINTER(C.c_uint64)), ("truthiness", C.POINTER(C.c_double)),
        ("fi_base", C.POINTER(C.c_uint64)), ("fi_len", C.POINTER(C.c_uint64)),
        ("factor_index", C.POINTER(C.c_uint64)),
        ("func", C.POINTER(C.c_int32)), ("num_vars", C.POINTER(C.c_uint64)),
        ("vif_base", C.POINTER(C.c_uint64)), ("wid", C.POINTER(C.c_uint64)),
        ("fv", C.POINTER(C.c_double)),
        ("vif_vid", C.POINTER(C.c_uint64)), ("vif_eq", C.POINTER(C.c_uint64)),
        ("w_init", C.POINTER(C.c_double)), ("w_fixed", C.POINTER(C.c_uint8)),
    ]


class OrcState(C.Structure):
    _fields_ = [("assign_free", C.POINTER(C.c_uint64)), ("assign_evid", C.POINTER(C.c_uint64)),
                ("weights", C.POINTER(C.c_double)), ("tallies", C.POINTER(C.c_uint64)),
                ("nsamples", C.POINTER(C.c_uint64))]


class OrcOpts(C.Structure):
    _fields_ = [("sample_evidence", C.c_int), ("learn_non_evidence", C.c_int),
                ("noise_aware", C.c_int), ("reg_l1", C.c_int), ("reg_param", C.c_double)]


Seed3 = C.c_ushort * 3
_oracle = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libdworacle.so"])


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    L = C.CDLL(ORACLE_SO)
    GP, SP, OP = C.POINTER(OrcGraph), C.POINTER(OrcState), C.POINTER(OrcOpts)
    L.orc_graph_build.restype = GP
    L.orc_graph_build.argtypes = [C.c_uint64, u64p, C.c_uint64, u64p, u64p, u64p, f64p,
                                  C.c_uint64, i32p, u64p, u64p, u64p, u64p, f64p,
                                  C.c_uint64, f64p, u8p]
    L.orc_graph_free.argtypes = [GP]
    L.orc_state_new.restype = SP
    L.orc_state_new.argtypes = [GP]
    L.orc_state_free.argtypes = [SP]
    L.orc_clear_tallies.argtypes = [GP, SP]
    L.orc_logadd.restype = C.c_double
    L.orc_logadd.argtypes = [C.c_double, C.c_double]
    L.orc_factor_potential.restype = C.c_double
    L.orc_factor_potential.argtypes = [C.c_int, C.c_uint64, u64p, u64p, C.c_double, u64p,
                                       C.c_uint64, C.c_uint64]
    L.orc_potential.restype = C.c_double
    L.orc_potential.argtypes = [GP, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64),
                                C.POINTER(C.c_double)]
    L.orc_draw_boolean.restype = C.c_uint64
    L.orc_draw_boolean.argtypes = [C.c_double, C.c_double, C.c_double]
    L.orc_draw_categorical.restype = C.c_uint64
    L.orc_draw_categorical.argtypes = [f64p, C.c_uint64, C.c_double]
    L.orc_draw_sample_r.restype = C.c_uint64
    L.orc_draw_sample_r.argtypes = [GP, C.c_uint64, C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_double), C.c_double]
    L.orc_update_weight.argtypes = [OP, C.POINTER(C.c_double), C.c_uint64, C.c_double, C.c_double]
    L.orc_sgd_on_variable.argtypes = [GP, OP, SP, C.c_uint64, C.c_double]
    L.orc_sample_single_variable.argtypes = [GP, OP, SP, C.c_uint64, Seed3]
    L.orc_sample_sgd_single_variable.argtypes = [GP, OP, SP, C.c_uint64, C.c_double, Seed3]
    L.orc_sweep_infer.argtypes = [GP, OP, SP, C.c_uint64, C.c_uint64, Seed3]
    L.orc_sweep_learn.argtypes = [GP, OP, SP, C.c_double, C.c_uint64, C.c_uint64, Seed3]
    L.orc_sweep_infer_seq.argtypes = [GP, OP, SP, u64p, C.c_uint64, f64p]
    L.orc_sweep_learn_seq.argtypes = [GP, OP, SP, C.c_double, u64p, C.c_uint64, f64p, f64p]
    L.orc_run.argtypes = [GP, OP, SP, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_long]
    _oracle = L
    return L


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),))


class Oracle:
    """One indexed graph + one InferenceResult-like state, on the CPU oracle."""

    def __init__(self, raw, sample_evidence=False, learn_non_evidence=False,
                 noise_aware=False, reg_l1=False, reg_param=0.01):
        self.L = oracle()
        raw = raw.normalized()
        self.raw = raw
        self.g = self.L.orc_graph_build(
            raw.nvar, raw.var_rec().ravel(), len(raw.dom_vid), raw.dom_vid, raw.dom_ptr,
            raw.dom_val, raw.dom_truth, raw.nfac, raw.fac_func, raw.fac_ptr, raw.edge_vid,
            raw.edge_eq, raw.fac_wid, raw.fac_fv, raw.nweight, raw.w_init, raw.w_fixed)
        if not self.g:
            raise ValueError("oracle: malformed graph (the reference would assert)")
        self.s = self.L.orc_state_new(self.g)
        self.opts = OrcOpts(int(sample_evidence), int(learn_non_evidence), int(noise_aware),
                            int(reg_l1), float(reg_param))
        G, S = self.g.contents, self.s.contents
        self.nvar, self.nval, self.nweight = int(G.nvar), int(G.nval), int(G.nweight)
        self.nfac, self.nindex = int(G.nfac), int(G.nindex)
        self.assign_free = _view(S.assign_free, self.nvar, np.uint64)
        self.assign_evid = _view(S.assign_evid, self.nvar, np.uint64)
        self.weights = _view(S.weights, self.nweight, np.float64)
        self.tallies = _view(S.tallies, self.nval, np.uint64)
        self.nsamples = _view(S.nsamples, self.nvar, np.uint64)
        self.var_val_base = _view(G.var_val_base, self.nvar, np.uint64)
        self.dtype = _view(G.dtype, self.nvar, np.uint8)
        self.is_evid = _view(G.is_evid, self.nvar, np.uint8)
        self.card = _view(G.card, self.nvar, np.uint64)
        self.assignment_dense = _view(G.assignment_dense, self.nvar, np.uint64)
        self.value = _view(G.value, self.nval, np.uint64)
        self.truthiness = _view(G.truthiness, self.nval, np.float64)
        self.fi_base = _view(G.fi_base, self.nval, np.uint64)
        self.fi_len = _view(G.fi_len, self.nval, np.uint64)
        self.factor_index = _view(G.factor_index, self.nindex, np.uint64)
        self.vif_vid = _view(G.vif_vid, int(G.nedge), np.uint64)
        self.vif_eq = _view(G.vif_eq, int(G.nedge), np.uint64)

    def __del__(self):
        try:
            self.L.orc_state_free(self.s)
            self.L.orc_graph_free(self.g)
        except Exception:
            pass

    def chain(self, which):
        S = self.s.contents
        return S.assign_evid if which else S.assign_free

    def potential(self, vid, proposal, which=1):
        return self.L.orc_potential(self.g, vid, proposal, self.chain(which), self.s.contents.weights)

    def draw_r(self, vid, r, which=1):
        return self.L.orc_draw_sample_r(self.g, vid, self.chain(which), self.s.contents.weights, r)

    def sgd_on_variable(self, vid, stepsize):
        self.L.orc_sgd_on_variable(self.g, C.byref(self.opts), self.s, vid, stepsize)

    def sample_single_variable(self, vid, seed):
        self.L.orc_sample_single_variable(self.g, C.byref(self.opts), self.s, vid, seed)

    def sample_sgd_single_variable(self, vid, stepsize, seed):
        self.L.orc_sample_sgd_single_variable(self.g, C.byref(self.opts), self.s, vid, stepsize, seed)

    def sweep_infer(self, seed, ith=0, n=1):
        self.L.orc_sweep_infer(self.g, C.byref(self.opts), self.s, ith, n, seed)

    def sweep_learn(self, stepsize, seed, ith=0, n=1):
        self.L.orc_sweep_learn(self.g, C.byref(self.opts), self.s, stepsize, ith, n, seed)

    def sweep_infer_seq(self, order, r0):
        order = np.ascontiguousarray(order, np.uint64)
        self.L.orc_sweep_infer_seq(self.g, C.byref(self.opts), self.s, order, len(order),
                                   np.ascontiguousarray(r0, np.float64))

    def sweep_learn_seq(self, stepsize, order, r0, r1):
        order = np.ascontiguousarray(order, np.uint64)
        self.L.orc_sweep_learn_seq(self.g, C.byref(self.opts), self.s, stepsize, order, len(order),
                                   np.ascontiguousarray(r0, np.float64), np.ascontiguousarray(r1, np.float64))

    def clear_tallies(self):
        self.L.orc_clear_tallies(self.g, self.s)

    def run(self, n_learn, n_infer, stepsize, decay, srand_seed=-1):
        self.L.orc_run(self.g, C.byref(self.opts), self.s, n_learn, n_infer, stepsize, decay, srand_seed)

    def marginals(self):
        """{(vid, sparse value): prob} exactly as dump_marginals_in_text would print
        (src/inference_result.cc:211-243)."""
        out = {}
        for v in range(self.nvar):
            if self.is_evid[v] and not self.opts.sample_evidence:
                continue
            b = int(self.var_val_base[v])
            n = float(self.nsamples[v])
            if self.dtype[v] == 0:
                out[(v, 1)] = float(self.tallies[b]) / n
            else:
                for j in range(int(self.card[v])):
                    out[(v, int(self.value[b + j]))] = float(self.tallies[b + j]) / n
        return out


# --------------------------------------------------------------------------
# reference-built checker (only where oracle/_ref/ was built)
# --------------------------------------------------------------------------
_ref = None


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is not None:
        return _ref
    L = C.CDLL(REF_SO)
    L.ref_logadd.restype = C.c_double
    L.ref_logadd.argtypes = [C.c_double, C.c_double]
    L.ref_factor_potential.restype = C.c_double
    L.ref_factor_potential.argtypes = [C.c_int, C.c_uint64, u64p, u64p, C.c_double, u64p,
                                       C.c_uint64, C.c_uint64]
    L.ref_uniforms.argtypes = [C.c_ushort, C.c_ushort, C.c_ushort, C.c_int, f64p]
    L.ref_open.restype = C.c_void_p
    L.ref_open.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.ref_close.argtypes = [C.c_void_p]
    for n in ("variables", "factors", "weights", "edges", "values"):
        f = getattr(L, "ref_num_" + n)
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p]
    L.ref_get_variables.argtypes = [C.c_void_p, u64p]
    L.ref_get_values.argtypes = [C.c_void_p, u64p, f64p, u64p, u64p]
    L.ref_get_factor_index.argtypes = [C.c_void_p, u64p, C.c_uint64]
    L.ref_get_factors.argtypes = [C.c_void_p, i64p, f64p]
    L.ref_get_vifs.argtypes = [C.c_void_p, u64p, u64p]
    L.ref_get_assignments.argtypes = [C.c_void_p, C.c_int, u64p]
    L.ref_set_assignments.argtypes = [C.c_void_p, C.c_int, u64p]
    L.ref_get_weights.argtypes = [C.c_void_p, f64p]
    L.ref_set_weights.argtypes = [C.c_void_p, f64p]
    L.ref_get_tallies.argtypes = [C.c_void_p, u64p, u64p]
    L.ref_clear_tallies.argtypes = [C.c_void_p]
    L.ref_set_seed.argtypes = [C.c_void_p, C.c_ushort, C.c_ushort, C.c_ushort]
    L.ref_potential.restype = C.c_double
    L.ref_potential.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int]
    L.ref_factor_potential_by_id.restype = C.c_double
    L.ref_factor_potential_by_id.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64]
    L.ref_draw_sample.restype = C.c_uint64
    L.ref_draw_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    L.ref_sample_evid.restype = C.c_uint64
    L.ref_sample_evid.argtypes = [C.c_void_p, C.c_uint64]
    L.ref_sample_single_variable.argtypes = [C.c_void_p, C.c_uint64]
    L.ref_sample_sgd_single_variable.argtypes = [C.c_void_p, C.c_uint64, C.c_double]
    L.ref_sgd_on_variable.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_int]
    L.ref_update_weight.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double]
    L.ref_sweep_infer.argtypes = [C.c_void_p]
    L.ref_sweep_learn.argtypes = [C.c_void_p, C.c_double]
    _ref = L
    return L


class RefGraph:
    """The reference's own FactorGraph/InferenceResult/GibbsSamplerThread, loaded from files."""

    def __init__(self, paths, extra_args=()):
        self.L = ref()
        argv = ["dw", "gibbs", "-m", paths["meta"], "-v", paths["variables"], "-f", paths["factors"],
                "-w", paths["weights"], "-o", "."]
        if os.path.exists(paths.get("domains", "")) and os.path.getsize(paths["domains"]) > 0:
            argv += ["--domains", paths["domains"]]
        argv += list(extra_args)
        if "-l" not in argv:
            argv += ["-l", "1"]
        if "-i" not in argv:
            argv += ["-i", "1"]
        arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
        self.h = self.L.ref_open(len(argv), arr)
        self.nvar = int(self.L.ref_num_variables(self.h))
        self.nval = int(self.L.ref_num_values(self.h))
        self.nweight = int(self.L.ref_num_weights(self.h))
        self.nfac = int(self.L.ref_num_factors(self.h))
        self.nedge = int(self.L.ref_num_edges(self.h))

    def __del__(self):
        try:
            self.L.ref_close(self.h)
        except Exception:
            pass

    def variables(self):
        out = np.zeros((self.nvar, 6), dtype=np.uint64)
        self.L.ref_get_variables(self.h, out.ravel())
        return out

    def values(self):
        v = np.zeros(self.nval, np.uint64); t = np.zeros(self.nval, np.float64)
        b = np.zeros(self.nval, np.uint64); l = np.zeros(self.nval, np.uint64)
        self.L.ref_get_values(self.h, v, t, b, l)
        return v, t, b, l

    def factor_index(self, n):
        out = np.zeros(int(n), np.uint64)
        self.L.ref_get_factor_index(self.h, out, int(n))
        return out

    def vifs(self):
        a = np.zeros(self.nedge, np.uint64); b = np.zeros(self.nedge, np.uint64)
        self.L.ref_get_vifs(self.h, a, b)
        return a, b

    def get_assignments(self, which):
        out = np.zeros(self.nvar, np.uint64)
        self.L.ref_get_assignments(self.h, which, out)
        return out

    def set_assignments(self, which, a):
        self.L.ref_set_assignments(self.h, which, np.ascontiguousarray(a, dtype=np.uint64))

    def get_weights(self):
        out = np.zeros(self.nweight, np.float64)
        self.L.ref_get_weights(self.h, out)
        return out

    def set_weights(self, w):
        self.L.ref_set_weights(self.h, np.ascontiguousarray(w, dtype=np.float64))

    def get_tallies(self):
        t = np.zeros(self.nval, np.uint64); n = np.zeros(self.nvar, np.uint64)
        self.L.ref_get_tallies(self.h, t, n)
        return t, n


def uniforms(seed, n):
    """The glibc erand48 stream from a 3x16-bit seed (what the reference consumes)."""
    if have_ref():
        out = np.zeros(n, np.float64)
        ref().ref_uniforms(seed[0], seed[1], seed[2], n, out)
        return out
    libc = C.CDLL(None)
    libc.erand48.restype = C.c_double
    s = Seed3(*seed)
    return np.array([libc.erand48(s) for _ in range(n)], dtype=np.float64)
