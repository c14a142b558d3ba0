"""The N>1 path on CPU: world_size-2 gloo process group, chains sharded over ranks, one all-reduce of the counts.
The per-rank compute is replaced by the CPU chain specification (oracle) so that no GPU is needed; what is under
test is the sharding arithmetic, the stream addressing by GLOBAL chain index, and the collective."""

import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from pcetk_b200.distributed import shard_range


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert sum(c for _o, c in spans) == total
            expect = 0
            for offset, count in spans:
                assert offset == expect
                expect += count
            assert max(c for _o, c in spans) - min(c for _o, c in spans) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_number, path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle
    import pcetk_b200 as P
    from pcetk_b200.distributed import mc_probabilities
    from conftest import load_fixture

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_number)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=20, nprod=200, randomSeed=4711)
    model.DefineMCModel(sampler, log=None)
    port = oracle.Port()
    wsym = port.symmetrize(rec.interactions)
    pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])

    def run_counts(ph, chain_offset, nchains):
        counts = np.zeros((len(ph), rec.ninstances), dtype=np.int64)
        counters = np.zeros((len(ph), 4), dtype=np.int64)
        esum = np.zeros(len(ph))
        for q, value in enumerate(ph):
            for chain in range(chain_offset, chain_offset + nchains):
                out = port.mc_philox_chain(rec, wsym, pairs, sampler.nequil, sampler.nprod, sampler.seed, chain, q, float(value))
                counts[q] += out["counts"]; counters[q] += out["counters"]; esum[q] += out["esum"]
        return counts, counters, esum

    p, counts, counters, esum = mc_probabilities(sampler, [4.0, 7.0], total_chains=5, rank=rank, world=world, run_counts=run_counts)
    if rank == 0:
        np.savez(path, p=p, counts=counts, counters=counters, esum=esum)
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_rank(tmp_path, port):
    from conftest import load_fixture
    import pcetk_b200 as P

    path = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), path), nprocs=2, join=True)
    got = np.load(path)
    rec, _ = load_fixture("twosites")
    model = P.MEADModel(rec, log=None)
    sampler = P.MCModelDefault(nequil=20, nprod=200, randomSeed=4711)
    model.DefineMCModel(sampler, log=None)
    wsym = port.symmetrize(rec.interactions)
    pairs = ([a for a, _b, _w in sampler.GetPairs()], [b for _a, b, _w in sampler.GetPairs()])
    counts = np.zeros((2, rec.ninstances), dtype=np.int64)
    for q, ph in enumerate([4.0, 7.0]):
        for chain in range(5):
            counts[q] += port.mc_philox_chain(rec, wsym, pairs, 20, 200, 4711, chain, q, ph)["counts"]
    assert np.array_equal(got["counts"], counts)            # independent of the number of ranks
    assert np.allclose(got["p"], counts / (5 * 200.0))
    assert got["counts"].sum() == 2 * 5 * 200 * rec.nsites
