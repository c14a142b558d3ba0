"""Multi-GPU sharding of the microstate path: one process per GPU (``torch.distributed``), no data-path collective
except ONE all-reduce(sum) per call.

* Monte Carlo: the chain ensemble is split over ranks (every rank holds the whole model and contributes to every pH
  value); chain c of the global ensemble draws the same Philox stream on whichever rank runs it, so the summed
  counts do not depend on the number of ranks.  Exchange: ONE float64 buffer holding counts[npH, ninst], the four move
  counters per pH value and the energy sums (the integers stay below 2^53, so their float64 sums are exact).  On the GPU
  the kernel writes device buffers, the packing and the all-reduce run on the device, and one copy brings the result back.
* Analytic: the chunk range of the enumeration is split over ranks; every rank produces proton-count bins relative
  to the same reference energy and the same bin scales, so bins simply add.  Exchange: float64 bins[(1+ninst) * nbins].

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


def _world(group=None):
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def mc_probabilities(sampler, pHs, total_chains, rank=0, world=1, group=None, device=None, run_counts=None, phIndexOffset=None, energy=False):
    """Titration-curve probabilities from ``total_chains`` chains per pH value spread over ``world`` ranks.

    -> (probabilities[npH, ninst], counts int64, counters int64 [npH, 4], esum[npH]), identical on every rank.

    ``device``: a torch CUDA device -> the sampler's kernel writes device buffers and the exchange is one NCCL all-reduce;
    ``run_counts(pHs, chain_offset, nchains) -> (counts, counters, esum)`` (host arrays) replaces the launch: the CPU tests
    substitute the chain specification to exercise the sharding and the collective over gloo without a GPU.
    ``phIndexOffset``: explicit random-stream address (replayable); default: the sampler's next unused streams -- every rank
    must have the same call history, as it has when all ranks run the same program.
    ``energy``: also track the chain energies (``esum``; zeros otherwise, like the reference's quiet path)."""
    import torch
    import torch.distributed as dist

    offset, count = shard_range(total_chains, rank, world)
    ph = np.ascontiguousarray(np.atleast_1d(pHs), dtype=np.float64)
    nph, ninst = len(ph), sampler.owner.ninstances
    ncounts, ncounters = nph * ninst, nph * 4
    stream_offset = sampler._takeStreams(nph, phIndexOffset)
    saved = (sampler.nchains, sampler.chainOffset)
    try:
        if run_counts is None and device is not None:
            # device-resident: kernel -> int64 / float64 device buffers -> one packed float64 buffer
            sampler._requireOwner()
            d_counts = torch.zeros(ncounts + ncounters, dtype=torch.int64, device=device)
            d_esum = torch.zeros(nph, dtype=torch.float64, device=device)
            if count > 0:
                sampler.nchains, sampler.chainOffset = count, offset
                sampler._configure()
                _native.check(sampler._lib.pce_mc_run_device(sampler._handle, ph, nph, stream_offset, int(offset), C.c_void_p(d_counts.data_ptr()),
                                                             C.c_void_p(d_counts.data_ptr() + 8 * ncounts), C.c_void_p(d_esum.data_ptr()) if energy else None))
                sampler.owner.energyModel.Synchronize()            # the launch stream may not be torch's current stream
            packed = torch.cat([d_counts.to(torch.float64), d_esum])
        else:
            if run_counts is None:
                def run_counts(values, chain_offset, nchains):
                    sampler.nchains, sampler.chainOffset = nchains, chain_offset
                    return sampler.RunCounts(values, stream_offset, energy=energy)
            if count > 0:
                counts, counters, esum = run_counts(ph, offset, count)
            else:
                counts, counters, esum = np.zeros((nph, ninst), dtype=np.int64), np.zeros((nph, 4), dtype=np.int64), np.zeros(nph)
            packed = torch.from_numpy(np.concatenate([np.asarray(counts, dtype=np.float64).ravel(), np.asarray(counters, dtype=np.float64).ravel(),
                                                      np.asarray(esum, dtype=np.float64).ravel()]))
            if device is not None:
                packed = packed.to(device)
        if _world(group) > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)       # the path's only exchange
        host = packed.cpu().numpy()
    finally:
        sampler.nchains, sampler.chainOffset = saved
    counts = np.rint(host[:ncounts]).astype(np.int64).reshape(nph, ninst)
    counters = np.rint(host[ncounts:ncounts + ncounters]).astype(np.int64).reshape(nph, 4)
    esum = host[ncounts + ncounters:].copy()
    probabilities = counts / float(total_chains * sampler.nprod)
    sampler.lastCounters, sampler.lastEnergySums = counters, esum
    sampler.owner.energyModel.SetProbabilities(probabilities[-1])        # every rank holds the reduced result of the last pH value
    return probabilities, counts, counters, esum


def analytic_probabilities(model, pHs, unfolded=False, rank=0, world=1, group=None, device=None):
    """Exhaustive-enumeration probabilities with the chunk range sharded over ranks (GPU only): one all-reduce of the bins."""
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
