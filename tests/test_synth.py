"""The streamed synthetic builder (csrc/host/synth.cc, dwg_graph_synthetic_unary: BASELINE.json
configs[3] / configs[4] at sizes no record array can hold) against the record path: the numpy
mirror of the generator (graphgen.synth_unary) compiled by compile_graph must give the SAME device
image, array by array.  Host-only."""
import numpy as np
import pytest

import sampler_b200 as sb
from sampler_b200 import graphgen

ARRAYS = ["perm", "inv", "color_start", "var_flags", "sl_off", "blob", "learn_visit", "w_perm", "w_inv", "row_of_value",
          "cls_N_evid", "cls_Q_evid", "cls_H_evid", "cls_N_all", "cls_Q_all", "cls_H_all", "w_init", "w_fixed"]
CASES = {
    "c5_shape": dict(nvar=20_011, degree=10, nweight=1000, zipf=0.0, real_features=False, p_evid=0.5),
    "c4_shape": dict(nvar=9_973, degree=20, nweight=6000, zipf=1.1, real_features=True, p_evid=0.9),
    "tiny": dict(nvar=37, degree=3, nweight=5, zipf=0.0, real_features=True, p_evid=0.3),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads", ["1", "7"])
def test_streamed_image_equals_compiled_records(name, threads, monkeypatch):
    monkeypatch.setenv("DWG_COMPILE_THREADS", threads)
    kw = CASES[name]
    g, cdf = sb.Graph.synthetic_unary(seed=99, want_cdf=True, **kw)
    raw = graphgen.synth_unary(seed=99, cdf=cdf if len(cdf) else None, **kw)
    ref = sb.Graph.from_raw(raw, sb.SCHED_HOGWILD if name == "tiny" else sb.SCHED_AUTO)
    a, b = g.info, ref.info
    for f in ("num_variables", "num_factors", "num_weights", "num_edges", "num_values", "num_variables_evidence",
              "num_variables_query", "num_index_entries", "num_colors", "schedule", "has_categorical", "has_multi_arity",
              "fv_mode", "assign_bytes", "bytes_per_sweep_infer", "bytes_per_sweep_infer_query", "bytes_per_sweep_learn",
              "num_device_rows"):
        assert getattr(a, f) == getattr(b, f), f
    for arr in ARRAYS:
        x, y = g.image(arr), ref.image(arr)
        assert x.dtype == y.dtype and np.array_equal(x, y), arr
    assert np.array_equal(g.learn_order(), ref.learn_order())
    d1, e1, b1, v1 = g.var_meta()
    d2, e2, b2, v2 = ref.var_meta()
    assert np.array_equal(d1, d2) and np.array_equal(e1, e2) and np.array_equal(b1, b2) and np.array_equal(v1, v2)


def test_streamed_image_has_no_csr_stream():
    g = sb.Graph.synthetic_unary(1000, 4, 10)
    assert len(g.image("row_ptr")) == 0 and len(g.image("e_wid")) == 0 and len(g.image("blob")) > 0
