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


def _worker_exchange(rank, world, port, n_submaps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from da3slam_b200.sharding import RowExchange, shard_range, shard_sequence
    # (1) strong-sharded loop candidates: block partition of the pair list, as bench.py's loop512 section uses it
    n_pairs = 11
    full = np.random.default_rng(9).normal(size=(n_pairs, 16))
    a, b = shard_range(n_pairs, rank, world)
    ex = RowExchange(n_pairs, "cpu")
    for _ in range(3):                                   # re-used step after step: no allocation, same answer
        table = ex(torch.from_numpy(full[a:b].copy()))
    np.save(os.path.join(out_dir, f"x{rank}.npy"), table.numpy())
    # (2) one sequence spread over the ranks (global map): rank r owns submaps [a, b) and the pairs starting in them
    sh = shard_sequence(n_submaps, rank, world)
    rows = np.random.default_rng(10).normal(size=(n_submaps - 1, 16))
    mine = [k for k in range(sh["a"], sh["b"]) if k + 1 < n_submaps]
    assert len(mine) == sh["pair_sizes"][rank] and sum(sh["pair_sizes"]) == n_submaps - 1
    assert sh["halo"] == (sh["b"] < n_submaps and sh["b"] > sh["a"])
    ex2 = RowExchange(n_submaps - 1, "cpu", sizes=sh["pair_sizes"])
    got = ex2(torch.from_numpy(rows[mine].copy().reshape(len(mine), 16)))
    np.save(os.path.join(out_dir, f"s{rank}.npy"), got.numpy())
    dist.destroy_process_group()


def test_row_exchange_and_sequence_shards_world2(tmp_path):
    for n_submaps in (8, 3, 2):
        port = _free_port()
        mp.spawn(_worker_exchange, args=(2, port, n_submaps, str(tmp_path)), nprocs=2, join=True)
        full = np.random.default_rng(9).normal(size=(11, 16))
        rows = np.random.default_rng(10).normal(size=(n_submaps - 1, 16))
        for r in (0, 1):
            assert np.array_equal(np.load(tmp_path / f"x{r}.npy"), full)
            assert np.array_equal(np.load(tmp_path / f"s{r}.npy"), rows)


def test_row_exchange_single_process_is_identity():
    from da3slam_b200.sharding import RowExchange, shard_sequence
    t = torch.arange(48, dtype=torch.float64).view(3, 16)
    assert RowExchange(3, "cpu")(t) is t
    sh = shard_sequence(8, 0, 1)
    assert (sh["a"], sh["b"], sh["halo"], sh["pair_sizes"]) == (0, 8, False, [7])
    sizes = [shard_sequence(8, r, 8)["pair_sizes"] for r in range(8)]
    assert all(s_ == [1, 1, 1, 1, 1, 1, 1, 0] for s_ in sizes)
    assert [shard_sequence(8, r, 8)["halo"] for r in range(8)] == [True] * 7 + [False]
