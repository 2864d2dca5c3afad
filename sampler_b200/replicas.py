"""One factor-graph replica per rank: the learn-then-infer schedule of DimmWitted across
`torch.distributed` ranks (one process per GPU).

Mirrors, rank for replica:
  DimmWitted::compute_n_epochs     src/dimmwitted.cc:280-282   n = ceil(requested / n_samplers)
  DimmWitted::learn                src/dimmwitted.cc:162-207   sweep -> average -> stepsize *= decay
  DimmWitted::update_weights       src/dimmwitted.cc:209-243   sum over replicas, / n, lmax / l2
  InferenceResult::copy_weights_to src/inference_result.cc:81-86 (an allreduce leaves every
                                   replica with the average, so there is nothing to copy back)
  DimmWitted::inference            src/dimmwitted.cc:121-160   clear tallies, sweeps
  aggregate_marginals_from         src/inference_result.cc:116-127 tallies / nsamples summed on rank 0

The replica object is anything with the `Sampler` interface of api.py.  With the NCCL
backend the collectives run on the replica's own device arrays and stream (no host copy);
with any other backend (gloo in the CPU tests) they go through the host accessors
dwg_get_weights / dwg_set_weights / dwg_get_tallies of the C ABI.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


def compute_n_epochs(requested: int, n_samplers: int) -> int:
    """src/dimmwitted.cc:280-282"""
    return int(math.ceil(requested / float(n_samplers)))


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


class ReplicaGroup:
    def __init__(self, sampler, dist=None, device: Optional[int] = None):
        self.s = sampler
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.on_device = bool(self.dist) and self.dist.get_backend() == "nccl"
        self._w_t = self._t_t = self._stream = None
        if self.on_device:
            import torch
            dev = torch.device("cuda", sampler.device if device is None else device)
            self._stream = torch.cuda.ExternalStream(sampler.stream(), device=dev)
            self._w_t = torch.as_tensor(_DevArray(sampler.weights_device_ptr(), sampler.nweight, "<f8"), device=dev)
            ndrow = getattr(sampler, "ndrow", None) or sampler.nval
            self._t_t = torch.as_tensor(_DevArray(sampler.tallies_device_ptr(), ndrow, "<i4"), device=dev)

    # -- the NUMA copies of the graph (src/dimmwitted.cc:113-116) ---------------------------
    def broadcast_graph(self, src: int = 0):
        """NCCL only: replicas created with defer_upload receive replica `src`'s device copy of the
        graph over NVLink (one host->device upload per node instead of one per GPU), then start
        from the initial state."""
        assert self.on_device
        import torch
        dev = self._w_t.device
        with torch.cuda.stream(self._stream):
            for ptr, nbytes in self.s.graph_segments():
                t = torch.as_tensor(_DevArray(ptr, nbytes, "|u1"), device=dev)
                self.dist.broadcast(t, src=src)
        if self.rank != src:
            self.s.reset_state()

    # -- DimmWitted::update_weights + copy_weights_to ------------------------------------
    def average_weights(self, want_norms: bool = False):
        """Sum the replicas' weights, divide by the replica count; returns (lmax, l2sq) of the
        change since the previous call when asked (rank-local copy of the same numbers)."""
        if self.dist is None:
            return self.s.finish_weight_average(1, want_norms) if want_norms else (0.0, 0.0)
        if self.on_device:
            import torch
            with torch.cuda.stream(self._stream):
                self.dist.all_reduce(self._w_t)
            return self.s.finish_weight_average(self.world, want_norms)
        import torch
        w = torch.from_numpy(self.s.weights())
        self.dist.all_reduce(w)
        self.s.set_weights_sum(w.numpy())  # un-averaged sum; the epilogue kernel scales it
        return self.s.finish_weight_average(self.world, want_norms)

    # -- DimmWitted::learn -----------------------------------------------------------------
    def learn(self, n_requested: int, stepsize: float, decay: float, want_norms: bool = False):
        n_epoch = compute_n_epochs(n_requested, self.world)
        norms = []
        for _ in range(n_epoch):
            self.s.learn_epoch(stepsize)
            r = self.average_weights(want_norms)
            if want_norms:
                norms.append(r)
            stepsize *= decay
        return n_epoch, stepsize, norms

    # -- DimmWitted::inference -------------------------------------------------------------
    def inference(self, n_requested: int, burn_in: int = 0):
        n_epoch = compute_n_epochs(n_requested, self.world)
        if n_requested > 0 and burn_in:
            self.s.infer_epochs(burn_in)
        self.s.clear_tallies()
        self.s.infer_epochs(n_epoch)
        return n_epoch

    # -- aggregate_marginals_from ------------------------------------------------------------
    def reduce_tallies_device(self, to_all: bool = False, max_count: Optional[int] = None):
        """NCCL only: sum the u32 tally arrays onto rank 0 on the replica's stream (the timed
        exchange of bench.py).  The per-variable sample counts are implicit (sweeps run).
        to_all: an all-reduce instead (every rank ends with the aggregate) -- measured on 8 B200s
        with 40 MB of tallies it is the slower of the two (0.81 ms against 0.27 ms for the reduce,
        profiles/r2b_bench_c2_8gpu*.json), so the reduce onto the root stays the default."""
        assert self.on_device
        import torch
        # max_count: an upper bound of every AGGREGATED tally (sweeps summed over the replicas).  Counters
        # that fit 8 / 16 bits travel four / two to a word (dwg_tallies_pack): a quarter / half of the bytes
        fields = 1 if max_count is None else 4 if max_count < (1 << 8) else 2 if max_count < (1 << 16) else 1
        with torch.cuda.stream(self._stream):
            t = self._t_t
            if fields > 1:
                ptr, nwords = self.s.tallies_pack(fields)
                t = torch.as_tensor(_DevArray(ptr, nwords, "<i4"), device=self._t_t.device)
            if to_all:
                self.dist.all_reduce(t)
            else:
                self.dist.reduce(t, dst=0)
            if fields > 1 and (to_all or self.rank == 0):
                self.s.tallies_unpack(fields)

    def aggregate(self):
        """(tallies[num_values], nsamples[num_variables]) summed over replicas; valid on rank 0."""
        t, n = self.s.tallies()
        if self.dist is None:
            return t, n
        import torch
        tt = torch.from_numpy(t.astype(np.int64))
        nn = torch.from_numpy(n.astype(np.int64))
        if self.on_device:
            # the host copies ride a device tensor: NCCL cannot move host memory
            dev = self._w_t.device
            tt, nn = tt.to(dev), nn.to(dev)
        self.dist.reduce(tt, dst=0)
        self.dist.reduce(nn, dst=0)
        return tt.cpu().numpy().astype(np.uint64), nn.cpu().numpy().astype(np.uint64)
