"""Multi-GPU sharding of the microstate path: one process per GPU (``torch.distributed``), no data-path collective
except ONE all-reduce(sum) per call.

* Monte Carlo: the chain ensemble is split over ranks (every rank holds the whole model and contributes to every pH
  value); chain c of the global ensemble draws the same Philox stream on whichever rank runs it, so the summed
  counts do not depend on the number of ranks.  Exchange: int64 counts[npH, ninst] (+ counters, energy sums).
* Analytic: the chunk range of the enumeration is split over ranks; every rank produces proton-count bins relative
  to the same reference energy, so bins simply add.  Exchange: float64 bins[(1+ninst) * nbins].

``torch`` is plumbing here (process group, device tensors, NCCL over NVLink); the compute is the C-ABI library.
"""

import ctypes as C

import numpy as np

from . import _native


def shard_range(total, rank, world):
    """Contiguous, balanced split of ``range(total)`` -> (offset, count) of ``rank``."""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def _all_reduce_sum(array, group=None, device=None):
    """Sum a numpy array over the process group (NCCL on ``device`` when given, else the group's CPU backend)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return array
    tensor = torch.from_numpy(np.ascontiguousarray(array))
    if device is not None:
        tensor = tensor.to(device)
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor.cpu().numpy()


def mc_probabilities(sampler, pHs, total_chains, rank=0, world=1, group=None, device=None, run_counts=None):
    """Titration-curve probabilities from ``total_chains`` chains per pH value spread over ``world`` ranks.

    ``run_counts(pHs, chain_offset, nchains) -> (counts, counters, esum)`` defaults to the sampler's GPU launch;
    tests substitute the CPU chain specification to exercise the sharding without a GPU."""
    offset, count = shard_range(total_chains, rank, world)
    ph = np.ascontiguousarray(np.atleast_1d(pHs), dtype=np.float64)
    if run_counts is None:
        def run_counts(values, chain_offset, nchains):
            sampler.nchains, sampler.chainOffset = nchains, chain_offset
            return sampler.RunCounts(values)
    if count > 0:
        counts, counters, esum = run_counts(ph, offset, count)
    else:
        ninst = sampler.owner.ninstances
        counts, counters, esum = np.zeros((len(ph), ninst), dtype=np.int64), np.zeros((len(ph), 4), dtype=np.int64), np.zeros(len(ph))
    counts = _all_reduce_sum(counts, group, device)
    counters = _all_reduce_sum(counters, group, device)
    esum = _all_reduce_sum(esum, group, device)
    probabilities = counts / float(total_chains * sampler.nprod)
    return probabilities, counts, counters, esum


def analytic_probabilities(model, pHs, unfolded=False, rank=0, world=1, group=None, device=None):
    """Exhaustive-enumeration probabilities with the chunk range sharded over ranks (GPU only)."""
    import torch

    em = model.energyModel
    lib, handle = _native.library(), em._handle
    ph = np.ascontiguousarray(np.atleast_1d(pHs), dtype=np.float64)
    ninst = model.ninstances
    nbins = int(lib.pce_analytic_bins_size(handle))
    max_protons = nbins // (1 + ninst) - 1
    out = np.zeros((len(ph), ninst), dtype=np.float64)
    for group_indices in em._phGroups(ph, max_protons):
        sub = np.ascontiguousarray(ph[group_indices])
        bins = torch.zeros(nbins, dtype=torch.float64, device=device if device is not None else "cuda")
        _native.check(lib.pce_analytic_bins(handle, sub, len(sub), int(bool(unfolded)), int(rank), int(world),
                                            float(em.analyticStatesLimit), C.c_void_p(bins.data_ptr())))
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=group)
        host = bins.cpu().numpy()
        p = np.zeros((len(sub), ninst), dtype=np.float64)
        _native.check(lib.pce_analytic_finish(handle, host, sub, len(sub), int(bool(unfolded)), 0.0, _native.ptr(p), None))
        out[group_indices] = p
    return out
