"""world_size-2 gloo test of the multi-GPU host logic (no GPU): each rank holds the Sim(3)
rows of its shard of the pair list; after gather_rows every rank has the identical full
table and the chain accumulation on it is identical to the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from da3slam_b200.pipeline import accumulate_sim3, rows_to_sim3
    from da3slam_b200.sharding import gather_rows, shard_range
    rng = np.random.default_rng(5)                       # same table on every rank; each contributes its shard
    full = rng.normal(size=(n_pairs, 16))
    a, b = shard_range(n_pairs, rank, world)
    table = gather_rows(torch.from_numpy(full[a:b].copy()), n_pairs)
    assert table.shape == (n_pairs, 16)
    acc = accumulate_sim3(rows_to_sim3(table.numpy()))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), table.numpy())
    np.save(os.path.join(out_dir, f"acc{rank}.npy"), np.stack([np.concatenate([[s], R.ravel(), t]) for s, R, t in acc]))
    dist.destroy_process_group()


def test_gather_rows_world2(tmp_path):
    for n_pairs in (7, 18):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_pairs, str(tmp_path)), nprocs=2, join=True)
        t0, t1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
        full = np.random.default_rng(5).normal(size=(n_pairs, 16))
        assert np.array_equal(t0, full) and np.array_equal(t1, full)          # bit-identical on every rank
        assert np.array_equal(np.load(tmp_path / "acc0.npy"), np.load(tmp_path / "acc1.npy"))
