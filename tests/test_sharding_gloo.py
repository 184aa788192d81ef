"""N > 1 path on CPU: two gloo processes shard a batch by rank (fnft_b200.shard) and
gather the results.  The per-signal transform is stood in for by the numpy oracle (the
product has no CPU path); what is tested is the host-side plumbing."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_batch_exactly():
    from fnft_b200.shard import shard_range
    for B in (1, 2, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s, e = shard_range(B, r, world)
                assert 0 <= s <= e <= B
                seen += list(range(s, e))
            assert seen == list(range(B))
            sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, q_all, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from fnft_b200.shard import gather_rows, shard_range
    from oracle import fnft_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard_range(B, rank, world)
    rows = np.stack([O.nsev_contspec(q_all[b], (-4.0, 4.0), 16, (-2.0, 2.0), 1) for b in range(s, e)])
    full = gather_rows(torch.from_numpy(rows), B)
    dist.barrier()
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp
    from oracle import fnft_oracle as O
    B, D = 5, 64  # odd batch: ranks get 3 and 2 signals
    rng = np.random.default_rng(11)
    t = np.linspace(-4, 4, D)
    q_all = rng.uniform(0.5, 2, (B, 1)) / np.cosh(t)[None] * np.exp(1j * rng.uniform(-1, 1, (B, 1)) * t[None])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, B, q_all, out_path), nprocs=2, join=True)
    got = np.load(out_path)
    want = np.stack([O.nsev_contspec(q_all[b], (-4.0, 4.0), 16, (-2.0, 2.0), 1) for b in range(B)])
    assert got.shape == want.shape
    assert np.array_equal(got, want)
