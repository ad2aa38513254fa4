"""world_size-2 gloo test of the multi-GPU host logic (sharding, gather, histogram all-reduce) on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc4quantum_b200.ensemble import shard_bounds, gather_results, allreduce_histogram


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_bounds(n_total, rank, world)
    fid = torch.arange(lo, hi, dtype=torch.float64) / n_total          # stand-in for this rank's fidelities
    us = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1, 1).repeat(1, 2, 3)
    all_fid = gather_results(fid, n_total)
    all_us = gather_results(us, n_total)
    hist = torch.from_numpy(np.histogram(fid.numpy(), bins=16, range=(0.0, 1.0))[0].astype(np.int64))
    allreduce_histogram(hist)
    if rank == 0:
        torch.save({'fid': all_fid, 'us': all_us, 'hist': hist}, out)
    dist.destroy_process_group()


def test_shard_gather_histogram_world2(tmp_path):
    n_total = 1001          # ragged: ranks own 501 and 500 members
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), n_total, out), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.equal(got['fid'], torch.arange(n_total, dtype=torch.float64) / n_total)
    assert got['us'].shape == (n_total, 2, 3) and torch.equal(got['us'][:, 0, 0], torch.arange(n_total, dtype=torch.float64))
    ref = np.histogram(np.arange(n_total) / n_total, bins=16, range=(0.0, 1.0))[0]
    assert np.array_equal(got['hist'].numpy(), ref)
