"""Multi-GPU check of the voxel-map merge, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_voxel_check.py

Several merges in a row, each on fresh clouds: every rank voxel-hashes its own synthetic cloud, VoxelExchange.merge
routes the records to their owners through CUDA-IPC peer memory (da3s_voxel_send) and the owners wait for the arrival
flags on the device — there is no host synchronisation and no barrier between the steps, and one rank per step is held
back on the host so that its peers run a step ahead of it.  Rank 0 then compares, step by step, the union of all shares
with the grid it builds alone from ALL points: keys, counts, positions and colours must be bit-identical.  One JSON line.
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


def cloud(dev, r, it, n):
    g = torch.Generator(device=dev)
    g.manual_seed(100 + 17 * it + r)
    xyz = torch.randn((n, 3), device=dev, generator=g) * 0.8
    rgb = torch.randint(0, 256, (n, 3), device=dev, generator=g, dtype=torch.uint8)
    return xyz, rgb


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    n, voxel, slots = int(os.environ.get("POINTS", 2_000_000)), 0.02, 1 << 23
    steps = int(os.environ.get("STEPS", 5))
    grid = ops.VoxelGrid(dev, slots, slots, True)
    cap = slots // max(1, world) + (1 << 16)
    ex = VoxelExchange(dev, world, rank, cap)
    mine = [cloud(dev, rank, it, n) for it in range(steps)]
    torch.cuda.synchronize()
    dist.barrier()
    # `steps` merges back to back: NO host synchronisation and NO process-group barrier between them (the arrival
    # flags live on the devices); one rank per step is held back on the host, so its peers run ahead and their
    # next-step records arrive while it is still busy with the previous step (the double-buffered inboxes' job)
    shares, t0 = [], time.perf_counter()
    for it in range(steps):
        if rank == it % world:
            time.sleep(0.05)
        grid.begin()
        grid.insert(mine[it][0], mine[it][1], None, voxel)
        ex.merge(grid, voxel)
        nv = grid.nv.clone()
        shares.append((grid.xyz.clone(), grid.rgb.clone(), grid.count.clone(), grid.key.clone(), nv))   # stream-ordered copies
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    out = []
    for xyz, rgb, cnt, key, nv in shares:
        m, dropped = (int(v) for v in nv.cpu())
        assert dropped == 0
        order = torch.argsort(key[:m])
        out.append([t[:m][order].cpu().numpy() for t in (xyz, rgb, cnt, key)])
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    ok, n_vox = True, []
    if rank == 0:
        for it in range(steps):
            grid.begin()
            for r in range(world):
                xyz_r, rgb_r = cloud(dev, r, it, n)
                grid.insert(xyz_r, rgb_r, None, voxel)
            grid.finish(voxel)
            e_xyz, e_rgb, e_cnt, e_key = [t.cpu().numpy() for t in grid.read(sort=True)]
            parts = [g_[it] for g_ in gathered]
            key = np.concatenate([s_[3] for s_ in parts]); order = np.argsort(key)
            ok = ok and (np.array_equal(key[order], e_key) and np.array_equal(np.concatenate([s_[2] for s_ in parts])[order], e_cnt)
                         and np.array_equal(np.concatenate([s_[0] for s_ in parts])[order], e_xyz)
                         and np.array_equal(np.concatenate([s_[1] for s_ in parts])[order], e_rgb))
            n_vox.append(int(len(e_key)))
        print(json.dumps({"check": "multi_gpu_voxel_merge", "world": world, "points_per_rank": n, "steps": steps, "voxels": n_vox,
                          "bit_identical_to_single_gpu": bool(ok), "ms_per_step_incl_insert_and_host_delays": round(1e3 * elapsed / steps, 3),
                          "host_sync_between_steps": False,
                          "share_sizes_last_step": [int(len(g_[-1][3])) for g_ in gathered]}))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
