"""Multi-GPU check of the voxel-map merge, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_voxel_check.py

Every rank voxel-hashes its own synthetic cloud, VoxelExchange.merge routes the records to their owners
through CUDA-IPC peer memory (da3s_voxel_send), rank 0 compares the union of all shares with the grid it
builds alone from ALL points: keys, counts, positions and colours must be bit-identical.  Prints one JSON line.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from da3slam_b200 import ops                                   # noqa: E402
from da3slam_b200.sharding import VoxelExchange               # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    n, voxel, slots = int(os.environ.get("POINTS", 4_000_000)), 0.02, 1 << 23
    clouds = []
    for r in range(world):                                     # every rank can rebuild every cloud (rank 0 needs them all)
        g = torch.Generator(device=dev); g.manual_seed(100 + r)
        xyz = torch.randn((n, 3), device=dev, generator=g) * 0.8
        rgb = torch.randint(0, 256, (n, 3), device=dev, generator=g, dtype=torch.uint8)
        clouds.append((xyz, rgb))
    grid = ops.VoxelGrid(dev, slots, slots, True)
    cap = slots // max(1, world) + (1 << 16)
    ex = VoxelExchange(dev, world, rank, cap)
    times = []
    for it in range(3):
        grid.begin()
        grid.insert(clouds[rank][0], clouds[rank][1], None, voxel)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        ex.merge(grid, voxel)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        xyz, rgb, cnt, key = grid.read(sort=True)
    share = [t.cpu().numpy() for t in (xyz, rgb, cnt, key)]
    sent = ex.counts.cpu().numpy().tolist()
    gathered = [None] * world
    dist.all_gather_object(gathered, share)
    ok, n_vox = True, 0
    if rank == 0:
        grid.begin()
        for xyz_r, rgb_r in clouds:
            grid.insert(xyz_r, rgb_r, None, voxel)
        grid.finish(voxel)
        e_xyz, e_rgb, e_cnt, e_key = [t.cpu().numpy() for t in grid.read(sort=True)]
        key = np.concatenate([s[3] for s in gathered]); order = np.argsort(key)
        ok = (np.array_equal(key[order], e_key) and np.array_equal(np.concatenate([s[2] for s in gathered])[order], e_cnt)
              and np.array_equal(np.concatenate([s[0] for s in gathered])[order], e_xyz)
              and np.array_equal(np.concatenate([s[1] for s in gathered])[order], e_rgb))
        n_vox = int(len(e_key))
        print(json.dumps({"check": "multi_gpu_voxel_merge", "world": world, "points_per_rank": n, "voxels": n_vox,
                          "bit_identical_to_single_gpu": bool(ok), "merge_ms_rank0": [round(1e3 * t, 3) for t in times],
                          "records_received_rank0": sent, "share_sizes": [int(len(s[3])) for s in gathered]}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
