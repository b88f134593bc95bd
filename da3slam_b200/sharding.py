"""Multi-GPU sharding of independent submap pairs.

One process per GPU (torchrun); pairs — consecutive submap pairs and loop-closure candidates
alike — are independent units, so each rank aligns a contiguous block of the pair list and the
ONLY exchange is one all_gather of the [n_local, 16] float64 Sim(3) rows (128 B per pair,
latency bound; NCCL over NVLink on GPUs, gloo on CPU for the tests).  The chain accumulation
(utils/geometry.py:73-119) then runs on every rank on the gathered table.

Global map export: every rank voxel-hashes its own submaps, then
`VoxelExchange.merge` routes each voxel record to the rank that owns its key.  The transfer is done
by the compaction kernel itself (da3s_voxel_send: peer stores into CUDA-IPC-mapped inboxes over
NVLink); torch.distributed only carries the IPC handles once and a barrier per merge.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ROW_LEN = 16


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous block partition: the first (n_units % world) ranks get one extra unit."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_units: int, world: int):
    return [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]


def gather_rows(rows_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather the per-rank row blocks into the full [n_total, 16] table (same on every rank).
    Blocks are padded to the largest shard so that a single fixed-size collective is used."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert rows_local.shape[0] == n_total
        return rows_local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    pad = max(sizes)
    buf = torch.zeros((pad, ROW_LEN), dtype=torch.float64, device=rows_local.device)
    buf[: rows_local.shape[0]] = rows_local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


class RowExchange:
    """The alignment path's only exchange between ranks: every rank's block of [n_local, 16] float64 Sim(3) rows
    all_gathered into the global [n_total, 16] table (128 B per pair; NCCL over NVLink on GPUs, gloo in the CPU tests).
    Buffers and the un-padding index are built once, so calling it inside a step only enqueues a copy, the collective
    and one gather — no allocation, no host synchronisation.  sizes[r] = rows rank r contributes (default: the
    contiguous block partition of shard_range); blocks are concatenated in rank order."""

    def __init__(self, n_total: int, device, group=None, sizes=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.sizes = list(sizes) if sizes is not None else shard_sizes(n_total, self.world)
        assert len(self.sizes) == self.world and sum(self.sizes) == n_total
        self.n_total = n_total
        self.pad = max(1, max(self.sizes))
        dev = torch.device(device)
        self.send = torch.zeros((self.pad, ROW_LEN), dtype=torch.float64, device=dev)
        self.recv = torch.zeros((self.world * self.pad, ROW_LEN), dtype=torch.float64, device=dev)
        idx = [r * self.pad + i for r in range(self.world) for i in range(self.sizes[r])]
        self.index = torch.tensor(idx, dtype=torch.int64, device=dev)
        self.out = torch.empty((n_total, ROW_LEN), dtype=torch.float64, device=dev)
        self.into_tensor = self.active and dist.get_backend(group) == "nccl"

    def __call__(self, rows_local: torch.Tensor) -> torch.Tensor:
        n_local = self.sizes[self.rank]
        assert rows_local.shape[0] == n_local
        if not self.active:
            return rows_local
        if n_local:
            self.send[:n_local].copy_(rows_local)
        if self.into_tensor:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            dist.all_gather(list(self.recv.view(self.world, self.pad, ROW_LEN).unbind(0)), self.send, group=self.group)
        torch.index_select(self.recv, 0, self.index, out=self.out)
        return self.out


def shard_sequence(n_submaps: int, rank: int, world: int):
    """Ownership of ONE sequence of n_submaps spread over the ranks (global map export):
    rank r owns submaps [a, b) (exports them) and the pairs (k, k + 1) for k in [a, b) with k + 1 < n_submaps, for which
    it also needs submap b's overlap frames (a halo copy).  Returns dict(a, b, halo, pair_sizes) — pair_sizes[r] for
    RowExchange (rank order == pair order)."""
    a, b = shard_range(n_submaps, rank, world)
    sizes = []
    for r in range(world):
        ra, rb = shard_range(n_submaps, r, world)
        sizes.append(max(0, min(rb, n_submaps - 1) - ra))
    return dict(a=a, b=b, halo=(b < n_submaps and b > a), pair_sizes=sizes)


class VoxelExchange:
    """Inboxes for the multi-GPU voxel merge.  One per rank; built once (collective call).

    cap = records any one rank may send to any other one (>= voxels_of_a_rank / world, with slack).
    local=True: every rank lives in THIS process (several grids on one device — the single-GPU test);
    call attach_local() with all of them, no process group is used.  Otherwise the inbox storages are
    shared through CUDA IPC and mapped into every rank of `group`.

    No host synchronisation in a merge: the send kernel of rank s raises flags[d][s] = step in rank d's memory after a
    system-scope fence over its peer stores, and rank d's merge waits for its `world` flags ON THE DEVICE.  Inbox, counts
    and flags are double-buffered by step parity, so the records a fast peer sends for step k + 1 can never land in the
    buffer a slow owner is still merging for step k (csrc/voxel.cu explains why one spare buffer is enough)."""

    def __init__(self, device, world: int, rank: int, cap: int, group=None, local: bool = False):
        self.device, self.world, self.rank, self.cap, self.group = torch.device(device), world, rank, int(cap), group
        # [parity][...]: one allocation each, so that ONE IPC handle per array covers both halves
        self.inbox = torch.zeros((2, world, self.cap, 6), dtype=torch.int64, device=self.device)
        self.counts = torch.zeros((2, world), dtype=torch.int64, device=self.device)
        self.flags = torch.zeros((2, world), dtype=torch.int64, device=self.device)
        self.peer_inbox, self.peer_counts, self.peer_flags = [None] * world, [None] * world, [None] * world
        self._keep = []
        self._peer_devices, self._enabled_for = set(), None
        self.local = bool(local)
        self.step = 0
        if world == 1:
            self.peer_inbox, self.peer_counts, self.peer_flags = [self.inbox], [self.counts], [self.flags]
        elif not self.local:
            self._map_ipc()

    def attach_local(self, peers):
        """peers: the VoxelExchange of every rank, all in this process."""
        self.peer_inbox = [p.inbox for p in peers]
        self.peer_counts = [p.counts for p in peers]
        self.peer_flags = [p.flags for p in peers]

    def _map_ipc(self):
        mine = tuple(t.untyped_storage()._share_cuda_() for t in (self.inbox, self.counts, self.flags))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        shapes = ((2, self.world, self.cap, 6), (2, self.world), (2, self.world))
        for r, handles in enumerate(everyone):
            if r == self.rank:
                self.peer_inbox[r], self.peer_counts[r], self.peer_flags[r] = self.inbox, self.counts, self.flags
                continue
            views = []
            for h, shp in zip(handles, shapes):
                # open the handles with OUR device current (first field of the tuple): the mapping is then made for the
                # device whose kernels will store through it (cudaIpcOpenMemHandle + lazy peer access over NVLink)
                st = torch.UntypedStorage._new_shared_cuda(self.device.index, *h[1:])
                self._keep.append(st)
                self._peer_devices.add(h[0])
                # the mapping belongs to the peer's device; only its address is used (by kernels running on OUR device)
                views.append(torch.empty(0, dtype=torch.int64, device=st.device).set_(st).view(*shp))
            self.peer_inbox[r], self.peer_counts[r], self.peer_flags[r] = views
        dist.barrier(group=self.group)                               # nobody frees a storage before it is mapped

    def send(self, grid):
        """Starts merge number self.step + 1: enqueue this rank's send (compaction + peer stores + arrival flags)."""
        if self._peer_devices and self._enabled_for is not grid.ctx:
            from . import _lib
            for d in sorted(self._peer_devices):
                _lib.check(grid.ctx.lib.da3s_enable_peer_access(grid.ctx.h, d), "da3s_enable_peer_access")
            self._enabled_for = grid.ctx
        self.step += 1
        p = self.step & 1
        grid.send(self.world, self.rank, [t[p] for t in self.peer_inbox], [t[p] for t in self.peer_counts],
                  [t[p] for t in self.peer_flags], self.step, self.cap)

    def fold(self, grid):
        """Enqueue the owner side of the current merge: device-side wait for every rank's arrival flag, then the fold."""
        p = self.step & 1
        grid.merge_inbox(self.inbox[p], self.counts[p], self.world, self.cap, flags=self.flags[p], step=self.step)

    def merge(self, grid, voxel: float):
        """Collective (every rank calls it once per step), enqueue-only: after it, `grid.read()` returns this rank's share
        of the global map (the voxels whose key it owns)."""
        if self.local and self.world > 1:
            raise RuntimeError("local exchanges are driven rank by rank: send() all, then fold() + finish() each")
        self.send(grid)
        self.fold(grid)
        grid.finish(voxel)
