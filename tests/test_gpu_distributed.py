"""The multi-GPU product API (pcetk_b200.distributed) on real devices: two ranks, device-resident counts, ONE packed
all-reduce.  With two or more GPUs visible the ranks take one device each over NCCL; on a one-GPU box both ranks share
cuda:0 and the collective runs over gloo on CUDA tensors -- the device code path (kernel -> device buffers -> packed float64
tensor -> all_reduce -> one copy back) is the same.  Asserted: the summed counts do not depend on the number of ranks, the
sharded enumeration equals the single-device one, the caller's sampler is left as it was."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GRID = [3.0, 7.0, 11.0]
TOTAL_CHAINS = 9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_number, path, backend):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    import pcetk_b200 as P
    from pcetk_b200.distributed import analytic_probabilities, mc_probabilities
    from conftest import load_fixture

    index = rank if backend == "nccl" else 0
    torch.cuda.set_device(index)
    device = torch.device("cuda", index)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_number)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    rec, _ = load_fixture("defensin")
    model = P.MEADModel(rec, log=None, device=index)
    sampler = P.MCModelDefault(nequil=50, nprod=400, randomSeed=4711, nchains=5)
    model.DefineMCModel(sampler, log=None)
    p, counts, counters, esum = mc_probabilities(sampler, GRID, total_chains=TOTAL_CHAINS, rank=rank, world=world, device=device, phIndexOffset=0, energy=True)
    assert sampler.nchains == 5 and sampler.chainOffset == 0                       # the caller's sampler is restored
    assert np.allclose(model.energyModel.GetProbabilities(), p[-1])               # every rank holds the reduced result
    model.DefineMCModel(None, log=None)
    pa = analytic_probabilities(model, GRID, rank=rank, world=world, device=device)
    if rank == 0:
        np.savez(path, p=p, counts=counts, counters=counters, esum=esum, pa=pa)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_on_devices_equal_one_rank(tmp_path):
    import torch
    import torch.multiprocessing as mp

    import pcetk_b200 as P
    from conftest import load_fixture

    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    path = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), path, backend), nprocs=2, join=True)
    got = np.load(path)
    rec, _ = load_fixture("defensin")
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=50, nprod=400, randomSeed=4711, nchains=TOTAL_CHAINS)
    model.DefineMCModel(sampler, log=None)
    counts, counters, esum = sampler.RunCounts(GRID, phIndexOffset=0)
    assert np.array_equal(got["counts"], counts)                                  # independent of the number of ranks
    assert np.array_equal(got["counters"], counters)
    assert np.abs(got["esum"] / esum - 1.0).max() < 1e-12
    assert np.allclose(got["p"], counts / float(TOTAL_CHAINS * 400))
    model.DefineMCModel(None, log=None)
    exact = model.CalculateProbabilitiesBatch(GRID)
    mask = exact > 1e-250
    assert np.abs(got["pa"][mask] / exact[mask] - 1.0).max() <= 1e-13
