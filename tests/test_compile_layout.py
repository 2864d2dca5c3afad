"""CPU checks of the host side: loader + compile.cc (the device layout) against the oracle,
schedule validity, C-ABI error behaviour.  No GPU needed (nothing here launches a kernel)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
import sampler_b200 as sb
from helpers import GRAPHS, graph_paths, load_golden_graph, random_graph
from image_model import ImageModel, VF_CAT, VF_EVID, VF_TRUTH
from sampler_b200 import graphio as gio


def _check_against_oracle(raw, schedule, rng, nassign=3):
    g = sb.Graph.from_raw(raw, schedule)
    m = ImageModel(g)
    o = ol.Oracle(raw)
    assert g.info.num_values == o.nval and g.info.num_index_entries == o.nindex
    perm = m.perm[m.valid].astype(np.int64)
    inv = m.inv.astype(np.int64)
    assert sorted(perm.tolist()) == list(range(o.nvar))
    assert np.array_equal(m.perm[inv], np.arange(o.nvar))
    assert ((m.var_flags[~m.valid] & 0x80) != 0).all() and ((m.var_flags[m.valid] & 0x80) == 0).all()
    # per-variable metadata
    assert np.array_equal((m.var_flags[inv] & VF_EVID) != 0, o.is_evid != 0)
    assert np.array_equal((m.var_flags[inv] & VF_CAT) != 0, o.dtype != 0)
    assert np.array_equal(m.evid_val[inv], o.assignment_dense)
    # internal weight ids are a permutation, hottest first
    assert sorted(m.w_perm.tolist()) == list(range(o.nweight))
    assert np.array_equal(m.w_perm[m.w_inv.astype(np.int64)], np.arange(o.nweight))
    dtype, is_evid, base, values = g.var_meta()
    assert np.array_equal(base[:-1], o.var_val_base) and np.array_equal(values, o.value)
    if len(m.truthiness):
        assert np.array_equal(m.truthiness[m.row_of_value.astype(np.int64)], o.truthiness)
    else:
        assert not o.truthiness.any()
    # potentials of every (variable, value) on random chains and weights: bit-exact
    card = np.where(o.dtype == 0, 2, o.card).astype(np.int64)
    for _ in range(nassign):
        a = np.array([rng.integers(0, max(1, c)) for c in card], np.uint64)
        w = rng.normal(0, 1, o.nweight)
        o.assign_evid[:] = a
        o.weights[:] = w
        for v in range(o.nvar):
            for p in range(int(card[v])):
                assert m.potential(v, p, a, w) == o.potential(v, p), (v, p)
            if len(m.sl_off):  # sliced-ELL blob of Boolean graphs: unary records first, so the
                # same terms in another order -- equal to fp64 reassociation
                assert np.allclose(m.sell_potentials(v, a, w), (o.potential(v, 1), o.potential(v, 0)),
                                   rtol=1e-12, atol=1e-12), v
    if len(m.sl_off):
        assert m.sell_fixed_bits_ok(raw.normalized().w_fixed)
    assert (len(m.sl_off) > 0) == (not o.dtype.any() and g.info.schedule != sb.SCHED_SEQUENTIAL)
    return g, m, o


@pytest.mark.parametrize("name", GRAPHS)
def test_layout_bundled_graphs(name):
    rng = np.random.default_rng(1)
    _check_against_oracle(load_golden_graph(name), sb.SCHED_AUTO, rng)


@pytest.mark.parametrize("seed", range(8))
def test_layout_random_graphs(seed):
    rng = np.random.default_rng(500 + seed)
    raw = random_graph(rng, truthiness=seed % 2 == 0, p_cat=[0.0, 0.4, 1.0][seed % 3])
    for sched in (sb.SCHED_HOGWILD, sb.SCHED_CHROMATIC):
        _check_against_oracle(raw, sched, rng, nassign=2)


def test_layout_large_cardinality_uses_wide_chains():
    rng = np.random.default_rng(3)
    raw = random_graph(rng, nvar=10, nfac=30, p_cat=1.0, with_domains=False)
    raw.var_card[:] = 300
    g, m, o = _check_against_oracle(raw, sb.SCHED_AUTO, rng, nassign=1)
    assert g.info.assign_bytes == 4


@pytest.mark.parametrize("seed", range(4))
def test_chromatic_classes_are_independent_sets(seed):
    rng = np.random.default_rng(40 + seed)
    raw = random_graph(rng, nvar=60, nfac=150)
    g = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
    perm, cs = g.schedule()
    color = np.zeros(raw.nvar, np.int64)
    for c in range(len(cs) - 1):
        color[perm[int(cs[c]):int(cs[c + 1])].astype(np.int64)] = c
    assert g.info.schedule == sb.SCHED_CHROMATIC and g.info.num_colors == len(cs) - 1
    for f in range(raw.nfac):
        vs = set(int(v) for v in raw.edge_vid[int(raw.fac_ptr[f]):int(raw.fac_ptr[f + 1])])
        cols = [color[v] for v in vs]
        assert len(set(cols)) == len(cols), "two variables of one factor share a colour"


def test_auto_schedule():
    # tiny graphs: one sequential chain in id order (the reference's -t 1 semantics)
    for name in GRAPHS:
        g = sb.Graph.from_raw(load_golden_graph(name))
        assert g.info.schedule == sb.SCHED_SEQUENTIAL and g.info.num_colors == 1
        assert np.array_equal(g.schedule()[0], np.arange(g.info.num_variables))
    assert sb.Graph.from_raw(load_golden_graph("biased_coin"), sb.SCHED_HOGWILD).info.schedule == sb.SCHED_HOGWILD
    g = sb.Graph.from_raw(load_golden_graph("partial_observation"), sb.SCHED_CHROMATIC)
    assert g.info.schedule == sb.SCHED_CHROMATIC and g.info.num_colors == 2
    rng = np.random.default_rng(0)
    big = random_graph(rng, nvar=200, nfac=600)
    assert sb.Graph.from_raw(big).info.schedule == sb.SCHED_CHROMATIC
    unary = random_graph(rng, nvar=200, nfac=600, max_arity=1)
    assert sb.Graph.from_raw(unary).info.schedule == sb.SCHED_HOGWILD


@pytest.mark.parametrize("name", GRAPHS)
def test_loader_matches_raw_records(name):
    p = graph_paths(name)
    a = sb.Graph.from_files(p["meta"], [p["variables"]], [p["domains"]], [p["factors"]], [p["weights"]])
    b = sb.Graph.from_raw(load_golden_graph(name))
    for arr in ("perm", "row_ptr", "row_oth", "e_wid", "e_tag", "e_fv32", "e_fv64", "oth", "evid_val",
                "var_flags", "truthiness", "w_init", "w_fixed", "values", "blob", "sl_off"):
        assert np.array_equal(a.image(arr), b.image(arr)), arr


def test_loader_multiple_files_and_pipes(tmp_path):
    """each kind may come in several files (src/cmd_parser.cc:63-70) and from a FIFO
    (test/run_end_to_end.sh:18-23 feeds `dw` process substitutions)"""
    raw = load_golden_graph("sparse_domains")
    d = str(tmp_path)
    half = raw.nfac // 2
    ep = int(raw.fac_ptr[half])
    f1 = gio.encode_factors(raw.fac_func[:half], raw.fac_ptr[:half + 1], raw.edge_vid[:ep], raw.edge_eq[:ep],
                            raw.fac_wid[:half], raw.fac_fv[:half])
    f2 = gio.encode_factors(raw.fac_func[half:], raw.fac_ptr[half:] - raw.fac_ptr[half], raw.edge_vid[ep:],
                            raw.edge_eq[ep:], raw.fac_wid[half:], raw.fac_fv[half:])
    paths = gio.write_binary(raw, d)
    open(os.path.join(d, "f1"), "wb").write(f1)
    fifo = os.path.join(d, "f2.fifo")
    os.mkfifo(fifo)
    import threading
    t = threading.Thread(target=lambda: open(fifo, "wb").write(f2))
    t.start()
    v = open(paths["variables"], "rb").read()
    open(os.path.join(d, "v1"), "wb").write(v[27 * 7:])  # out of order on purpose: placed by id
    open(os.path.join(d, "v2"), "wb").write(v[:27 * 7])
    g = sb.Graph.from_files(paths["meta"], [os.path.join(d, "v1"), os.path.join(d, "v2")], [paths["domains"]],
                            [os.path.join(d, "f1"), fifo], [paths["weights"]])
    t.join()
    ref = sb.Graph.from_raw(raw)
    for arr in ("row_ptr", "e_wid", "e_tag", "oth", "evid_val"):
        assert np.array_equal(g.image(arr), ref.image(arr)), arr


def test_errors_are_codes_not_aborts(tmp_path):
    raw = load_golden_graph("biased_coin")
    bad = gio.RawGraph(**{**raw.__dict__})
    bad.fac_func = raw.fac_func.copy(); bad.fac_func[3] = 5  # src/factor.h:81-84 would abort
    with pytest.raises(sb.DwgError, match="unsupported FACTOR_FUNCTION_TYPE"):
        sb.Graph.from_raw(bad)
    bad = gio.RawGraph(**{**raw.__dict__})
    bad.edge_vid = raw.edge_vid.copy(); bad.edge_vid[0] = 99
    with pytest.raises(sb.DwgError, match="unknown variable"):
        sb.Graph.from_raw(bad)
    # counts must match the meta file (safety_check, src/factor_graph.cc:203-206)
    paths = gio.write_binary(raw, str(tmp_path))
    open(paths["meta"], "w").write("1,18,17,18")
    with pytest.raises(sb.DwgError, match="number of factors"):
        sb.Graph.from_files(paths["meta"], [paths["variables"]], [], [paths["factors"]], [paths["weights"]])
    open(paths["meta"], "w").write("1,18,19,19")  # announces more than the file holds
    with pytest.raises(sb.DwgError, match="number of factors"):
        sb.Graph.from_files(paths["meta"], [paths["variables"]], [], [paths["factors"]], [paths["weights"]])
    open(paths["meta"], "w").write("1,18,18,18")
    whole = open(paths["factors"], "rb").read()
    open(paths["factors"] + ".cut", "wb").write(whole[:-5])
    with pytest.raises(sb.DwgError, match="truncated record"):
        sb.Graph.from_files(paths["meta"], [paths["variables"]], [], [paths["factors"] + ".cut"], [paths["weights"]])
    with pytest.raises(sb.DwgError, match="cannot open"):
        sb.Graph.from_files(paths["meta"], ["/nonexistent"], [], [paths["factors"]], [paths["weights"]])
    # meta with trailing junk / no newline parses like atoll on 4 comma fields
    open(paths["meta"], "w").write("1,18,18,18,out/x,,")
    sb.Graph.from_files(paths["meta"], [paths["variables"]], [], [paths["factors"]], [paths["weights"]])


def test_algorithmic_bytes_formula():
    """13 B/var + 9 B/unary factor (+5 per other variable); learning +2 B/var"""
    g = sb.Graph.from_raw(load_golden_graph("biased_coin"))
    assert g.info.bytes_per_sweep_infer == 18 * (13 + 9)
    assert g.info.bytes_per_sweep_infer_query == 9 * (13 + 9)
    assert g.info.bytes_per_sweep_learn == 18 * (13 + 9 + 2)
    g = sb.Graph.from_raw(load_golden_graph("partial_observation"))  # 12 vars, 8 EQUAL(2): 16 incidences
    assert g.info.bytes_per_sweep_infer == 12 * 13 + 16 * 14
    assert g.info.bytes_per_sweep_learn == 12 * 15 + 16 * 15


def test_no_device_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    g = sb.Graph.from_raw(load_golden_graph("biased_coin"))
    with pytest.raises(sb.DwgError, match="no CUDA device"):
        sb.Sampler(g)


IMAGE_ARRAYS = ("perm", "inv", "color_start", "var_flags", "evid_val", "var_row", "row_ptr", "row_oth",
                "truthiness", "e_wid", "e_fv32", "e_fv64", "e_tag", "oth", "row_of_value", "values",
                "value_base", "w_perm", "w_inv", "sl_off", "blob", "cls_N_evid", "cls_Q_evid", "cls_H_evid",
                "cls_N_all", "cls_Q_all", "cls_H_all")


@pytest.mark.parametrize("shape", ["c2", "c3", "mixed", "zipf"])
def test_compiled_image_does_not_depend_on_the_thread_count(shape, monkeypatch):
    """compile.cc fans every O(edges) phase out over host threads (host/parallel.h): the image --
    schedule, CSR stream, sliced-ELL blob, learning statistics -- is the same byte for byte with
    1, 3 and 8 threads (graphs large enough that the parallel paths really split)"""
    from sampler_b200 import graphgen
    rng = np.random.default_rng(99)
    if shape == "c2":
        raw = graphgen.c2_boolean(150_000, nweight=300)
    elif shape == "c3":
        raw = graphgen.c3_categorical(40_000, nweight=200)
    elif shape == "zipf":
        raw = graphgen.unary_boolean(100_000, 12, 20_000, p_evid=0.9, zipf=1.1, real_features=True)
    else:
        raw = random_graph(rng, nvar=6_000, nfac=16_000, nweight=300, max_arity=4)
    ref = None
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("DWG_COMPILE_THREADS", threads)
        g = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
        img = {name: g.image(name).copy() for name in IMAGE_ARRAYS}
        info = g.info
        meta = (info.num_colors, info.schedule, info.num_index_entries, info.num_other_refs, info.bytes_per_sweep_infer,
                info.bytes_per_sweep_learn, info.bytes_per_sweep_infer_query)
        if ref is None:
            ref = (img, meta)
            continue
        assert meta == ref[1]
        for name in IMAGE_ARRAYS:
            assert img[name].dtype == ref[0][name].dtype and np.array_equal(img[name].view(np.uint8), ref[0][name].view(np.uint8)), (name, threads)


def test_loader_blocks_match_in_memory_graph(tmp_path, monkeypatch):
    """the file path (mmap, one thread walking the factor arities while the others decode blocks of
    2^16 records, two factor files) builds the image the in-memory records build, with 1 and with 8
    host threads"""
    from sampler_b200 import graphgen
    raw = graphgen.c2_boolean(120_000, nweight=64)
    ref = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
    paths = gio.write_binary(raw, str(tmp_path))
    whole = open(paths["factors"], "rb").read()
    cut = 0  # split the factor file at a record boundary, after 500 000 records
    for _ in range(500_000):
        cut += 26 + 16 * int.from_bytes(whole[cut + 2:cut + 10], "big")
    open(paths["factors"] + ".a", "wb").write(whole[:cut])
    open(paths["factors"] + ".b", "wb").write(whole[cut:])
    for threads in ("1", "8"):
        monkeypatch.setenv("DWG_COMPILE_THREADS", threads)
        for files in ([paths["factors"]], [paths["factors"] + ".a", paths["factors"] + ".b"]):
            g = sb.Graph.from_files(paths["meta"], [paths["variables"]], [], files, [paths["weights"]], sb.SCHED_CHROMATIC)
            for name in IMAGE_ARRAYS:
                assert np.array_equal(g.image(name), ref.image(name)), (name, threads, len(files))


@pytest.mark.parametrize("shape", ["c2", "c3"])
def test_large_chromatic_schedule_is_a_proper_colouring(shape, monkeypatch):
    """the parallel wavefront first-fit pass + sequential balancing pass on graphs large enough to
    split across threads (random couplings: c2; id-ordered chains, the stalled path: c3): no factor
    joins two variables of one colour class, the classes are balanced, every variable is scheduled"""
    from sampler_b200 import graphgen
    monkeypatch.setenv("DWG_COMPILE_THREADS", "8")
    raw = graphgen.c2_boolean(200_000, nweight=64) if shape == "c2" else graphgen.c3_categorical(60_000, nweight=64)
    g = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
    perm, cs = g.schedule()
    nvar = len(raw.var_role)
    color = np.full(nvar, -1, np.int64)
    for c in range(len(cs) - 1):
        members = perm[int(cs[c]):int(cs[c + 1])]
        members = members[members != 0xFFFFFFFF].astype(np.int64)
        color[members] = c
    assert (color >= 0).all()
    ptr = raw.fac_ptr.astype(np.int64)
    arity = np.diff(ptr)
    assert arity.max() <= 2
    two = np.nonzero(arity == 2)[0]
    a, b = raw.edge_vid[ptr[two]].astype(np.int64), raw.edge_vid[ptr[two] + 1].astype(np.int64)
    distinct = a != b
    assert (color[a[distinct]] != color[b[distinct]]).all(), "two variables of one factor share a colour"
    sizes = np.bincount(color)
    big = sizes[sizes > 0.01 * nvar]
    assert big.max() <= 1.05 * big.min() + 64, sizes
    # the variables the balancing pass strands (all planned colours taken around them) are brought
    # back by one-step Kempe moves: no overflow class of a few variables is left
    assert len(big) == len(sizes), sizes


@pytest.mark.parametrize("seed", range(3))
def test_wide_slices_hold_the_same_potentials(seed, monkeypatch):
    """warp-per-variable slices (SF_WIDE: a high-degree variable's records dealt round-robin to the
    32 lanes): the lanes' shares must add up to the oracle's potentials, for unary, two-variable
    and general records alike; ordinary variables are untouched"""
    monkeypatch.setenv("DWG_WIDE_MIN", "12")
    monkeypatch.setenv("DWG_PART_RECORDS", "40")  # the hubs are cut into parts (SF_PART), the lesser wide variables are not
    rng = np.random.default_rng(6000 + seed)
    raw = random_graph(rng, nvar=120, nfac=700, nweight=15, p_cat=0.0, max_arity=[2, 3, 4][seed])
    # two hubs: variables 0 and 1 appear in many more factors
    ev = raw.edge_vid.copy()
    hub = rng.random(len(ev)) < 0.15
    ev[hub] = rng.integers(0, 2, hub.sum())
    raw.edge_vid = ev
    raw = raw.normalized()
    g = sb.Graph.from_raw(raw, sb.SCHED_CHROMATIC)
    m = ImageModel(g)
    o = ol.Oracle(raw)
    flags = [m.slice_header(s)["flags"] for s in range(len(m.sl_off) - 1)]
    assert any(f & 4 and not f & 8 for f in flags) and any(f & 8 for f in flags)
    for trial in range(2):
        a = rng.integers(0, 2, raw.nvar).astype(np.uint64)
        w = rng.normal(0, 1, raw.nweight)
        o.assign_evid[:] = a
        o.weights[:] = w
        for v in range(raw.nvar):
            assert np.allclose(m.sell_potentials(v, a, w), (o.potential(v, 1), o.potential(v, 0)), rtol=1e-12, atol=1e-12), v
