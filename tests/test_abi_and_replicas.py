"""CPU checks of the boundary: libdwg.so exports every symbol include/dwg.h declares (no
compute call is made), and the one-replica-per-rank schedule (sampler_b200/replicas.py, the
N > 1 path of bench.py) follows DimmWitted::learn / inference / aggregate on two gloo ranks.

The replicas of the gloo test are driven by the CPU oracle behind the Sampler interface -- the
oracle is the checker here, standing in for a GPU replica that this box does not have."""
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sampler_b200 as sb

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, "include", "dwg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(dwg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35, names
    lib = sb.load()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.dwg_version().decode().startswith("dwg")


def test_opts_and_info_struct_layout_matches_header():
    """the ctypes mirrors (the binding INTEGRATION.md shows) must track the header field by field"""
    hdr = open(os.path.join(REPO, "include", "dwg.h")).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            first, *rest = decl.split(",")
            out.append(first.split()[-1].lstrip("*"))
            out += [r.strip().lstrip("*") for r in rest]
        return out
    assert fields("dwg_opts") == [f[0] for f in sb.api.Opts._fields_]
    assert fields("dwg_graph_info") == [f[0] for f in sb.api.GraphInfo._fields_]
    assert fields("dwg_raw") == [f[0] for f in sb.api.Raw._fields_]


# ------------------------------------------------------------------ two gloo ranks
class OracleReplica:
    """the Sampler interface (api.py) over the CPU oracle: one sequential chain per rank, seeded
    by rank like GibbsSamplerThread's per-replica seeds (src/gibbs_sampler.cc:49)"""

    def __init__(self, raw, rank, **opts):
        import oracle_lib as ol
        self.o = ol.Oracle(raw, **opts)
        self.rank = rank
        self.nvar, self.nval, self.nweight = self.o.nvar, self.o.nval, self.o.nweight
        self.device = None
        self.calls = []
        self.prev = self.o.weights.copy()
        import ctypes as C
        self.seed = (C.c_ushort * 3)(rank + 1, 1, 1)  # erand48 state, advanced in place

    def learn_epoch(self, stepsize):
        self.calls.append(("learn", stepsize))
        self.o.sweep_learn(stepsize, self.seed)

    def infer_epochs(self, n):
        for _ in range(n):
            self.calls.append(("infer",))
            self.o.sweep_infer(self.seed)

    def clear_tallies(self):
        self.calls.append(("clear",))
        self.o.clear_tallies()

    def weights(self):
        return self.o.weights.copy()

    def set_weights_sum(self, w):
        self.o.weights[:] = w

    def finish_weight_average(self, n, want_norms=True):
        fixed = self.o.raw.w_fixed.astype(bool)
        w = self.o.weights / n
        w[fixed] = self.prev[fixed]  # fixed weights are identical everywhere: averaging is a no-op
        self.o.weights[:] = w
        d = np.abs(w - self.prev)
        self.prev = w.copy()
        return float(d.max()), float((d * d).sum())

    def tallies(self):
        return self.o.tallies.copy(), self.o.nsamples.copy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    for p in (REPO, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from helpers import load_golden_graph
    from sampler_b200.replicas import ReplicaGroup
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raw = load_golden_graph("biased_coin")
        rep = OracleReplica(raw, rank, sample_evidence=True, reg_param=0.0)
        grp = ReplicaGroup(rep, dist)
        assert grp.world == world and not grp.on_device
        n_epoch, step, norms = grp.learn(301, 0.1, 0.99, want_norms=True)
        w = rep.weights()
        n_inf = grp.inference(401)
        t, n = grp.aggregate()
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), w=w, t=t, n=n, n_epoch=n_epoch, n_inf=n_inf, step=step,
                 norms=np.array(norms), n_learn_calls=sum(1 for c in rep.calls if c[0] == "learn"),
                 n_infer_calls=sum(1 for c in rep.calls if c[0] == "infer"),
                 last_step=[c[1] for c in rep.calls if c[0] == "learn"][-1],
                 own_t=rep.tallies()[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_schedule_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(tmp_path, "r%d.npz" % i)) for i in range(world)]
    # compute_n_epochs (src/dimmwitted.cc:280-282): ceil(301 / 2), ceil(401 / 2) per replica
    for x in r:
        assert int(x["n_epoch"]) == 151 and int(x["n_learn_calls"]) == 151
        assert int(x["n_inf"]) == 201 and int(x["n_infer_calls"]) == 201
        assert np.isclose(float(x["last_step"]), 0.1 * 0.99 ** 150)  # stepsize *= decay after every epoch
        assert np.isclose(float(x["step"]), 0.1 * 0.99 ** 151)
    # after every epoch all replicas hold the same averaged weights (copy_weights_to)
    assert np.array_equal(r[0]["w"], r[1]["w"])
    assert np.array_equal(r[0]["norms"], r[1]["norms"]) and r[0]["norms"].shape == (151, 2)
    # biased_coin: 8 of 9 evidence coins are heads -> weight ~ 1.04 (test/biased_coin/check_result)
    assert abs(float(r[0]["w"][0]) - 1.04) < 0.15
    # aggregate_marginals_from: rank 0 holds the sum of both replicas' tallies and sample counts
    assert np.array_equal(r[0]["t"], r[0]["own_t"] + r[1]["own_t"])
    assert (r[0]["n"] == 402).all()
    p = r[0]["t"] / r[0]["n"]
    assert np.abs(p - 0.89).max() < 0.06


def test_single_process_group_is_a_plain_schedule():
    from helpers import load_golden_graph
    from sampler_b200.replicas import ReplicaGroup, compute_n_epochs
    assert [compute_n_epochs(n, c) for n, c in ((10, 1), (10, 3), (0, 4), (1, 8))] == [10, 4, 0, 1]
    rep = OracleReplica(load_golden_graph("biased_coin"), 0, sample_evidence=True, reg_param=0.0)
    grp = ReplicaGroup(rep, None)
    assert grp.world == 1
    n_epoch, step, _ = grp.learn(5, 0.1, 0.5)
    assert n_epoch == 5 and np.isclose(step, 0.1 * 0.5 ** 5)
    assert grp.inference(7, burn_in=2) == 7
    t, n = grp.aggregate()
    assert (n == 7).all()  # burn-in sweeps are discarded by the clear that follows them
