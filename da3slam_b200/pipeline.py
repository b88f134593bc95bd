"""Submap-level host logic over the kernels: device-resident submaps, batched pair
alignment of a sequence / of loop candidates, Sim(3) chaining, map export.

Mirrors the shape of the reference's intended long-sequence pipeline
(utils/da3_streaming.py:523-695: pairwise align -> accumulate -> apply -> export) with
every per-pixel stage on the GPU and no host synchronisation until the Sim(3) rows are
read back.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from . import ops


def _field(pred, key):
    return pred[key] if isinstance(pred, dict) else getattr(pred, key)


@dataclass
class DeviceSubmap:
    """A DA3 prediction for one chunk of frames, resident in HBM.
    depth/conf [F,H,W] f32, cams = da3s_cam table [F,200] u8, images [F,H,W,3] u8 or None."""
    depth: torch.Tensor
    conf: torch.Tensor
    cams: torch.Tensor
    intrinsics: torch.Tensor
    extrinsics: torch.Tensor
    images: torch.Tensor | None = None

    @property
    def shape(self):
        return tuple(self.depth.shape)

    @staticmethod
    def from_prediction(pred, device="cuda", general_inverse=False, non_blocking=True) -> "DeviceSubmap":
        def up(x, dtype):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            return t.to(device=device, dtype=dtype, non_blocking=non_blocking).contiguous()
        depth = up(_field(pred, "depth"), torch.float32)
        conf = up(_field(pred, "conf"), torch.float32)
        K = up(_field(pred, "intrinsics"), torch.float32)
        E = up(_field(pred, "extrinsics"), torch.float32)
        img = None
        has_img = ("processed_images" in pred) if isinstance(pred, dict) else hasattr(pred, "processed_images")
        if has_img:
            img = up(_field(pred, "processed_images"), torch.uint8)
        F, H, W = depth.shape
        if (H * W) % 4:
            raise L.Da3sError(L.EALIGN, "DeviceSubmap", "H*W must be a multiple of 4 (16-byte frame alignment)")
        return DeviceSubmap(depth, conf, ops.build_cams(K, E, general_inverse), K, E, img)


def pair_entry(prev: DeviceSubmap, cur: DeviceSubmap, overlap: int):
    """prev's last `overlap` frames (target, A) vs cur's first `overlap` frames (source, B):
    utils/align.py:323-335, align_geometry.py:271-285."""
    F = prev.depth.shape[0]
    return (prev.depth[F - overlap:], prev.conf[F - overlap:], cur.depth[:overlap], cur.conf[:overlap],
            prev.cams[F - overlap:], cur.cams[:overlap])


def rows_to_sim3(rows: np.ndarray):
    """[n,16] rows -> list of (s, R, t) exactly as the reference returns them
    (python float, (3,3) float64, (3,) float64)."""
    out = []
    for r in np.asarray(rows):
        out.append((float(r[0]), r[1:10].reshape(3, 3).copy(), r[10:13].copy()))
    return out


def align_submap_pairs(pairs_list, overlap=1, sample_idx=None, want_aux=False, want_counts=False, **opt_kw):
    """pairs_list: list of (prev DeviceSubmap, cur DeviceSubmap).  One batched launch
    sequence for all of them.  Returns CUDA tensors (rows, aux, counts)."""
    if not pairs_list:
        raise ValueError("no pairs")
    dev = pairs_list[0][0].depth.device
    _, H, W = pairs_list[0][0].depth.shape
    entries = [pair_entry(a, b, overlap) for a, b in pairs_list]
    table = ops.make_pairs(entries, dev)
    opts = L.default_opts(**opt_kw)
    return ops.align_pairs(table, len(entries), overlap, H, W, opts, sample_idx, want_aux, want_counts)


def align_sequence(submaps, overlap=1, **kw):
    """Consecutive pairs (k, k+1) of a sequence; rows[k] maps submap k+1 into submap k."""
    return align_submap_pairs([(submaps[k], submaps[k + 1]) for k in range(len(submaps) - 1)], overlap, **kw)


def loop_constraints(items, overlap=None, **opt_kw):
    """Loop-closure constraints from jointly predicted loop chunks (utils/da3_streaming.py:366-481).

    items: list of (a, b, frames_a, loop_a, frames_b, loop_b): chunk indices a, b and four DeviceSubmaps —
    `frames_a` / `frames_b` the chunks' own predictions of the frames that also went through the network together
    as ONE loop chunk, `loop_a` / `loop_b` the loop chunk's predictions of the same frames.  Both halves of every
    item are aligned on the GPU in one batch (loop -> a and loop -> b, every frame an overlap frame), and
    T_ab = T_a o T_b^-1 maps chunk b into chunk a (compute_sim3_ab upstream).  Returns [(a, b, (s, R, t))] for
    posegraph.optimize."""
    from . import posegraph
    if not items:
        return []
    overlap = overlap or items[0][2].depth.shape[0]        # every shared frame is an overlap frame
    pairs = []
    for a, b, fa, la, fb, lb in items:                      # target = the chunk's own frames, source = the loop chunk's
        pairs += [(fa, la), (fb, lb)]
    rows, _, _ = align_submap_pairs(pairs, overlap=overlap, **opt_kw)
    sims = rows_to_sim3(rows.cpu().numpy())
    out = []
    for i, (a, b, *_rest) in enumerate(items):
        T_a, T_b = sims[2 * i], sims[2 * i + 1]
        out.append((a, b, posegraph._compose(posegraph._as_sim3(T_a), posegraph._inverse(posegraph._as_sim3(T_b)))))
    return out


def close_loops(sequential, loops, max_iterations=30, lambda_init=1e-6):
    """Sim(3) pose graph over the chunk chain with the loop constraints (posegraph.optimize): returns the corrected
    sequential list, ready for accumulate_sim3 (upstream: Sim3LoopOptimizer.optimize, utils/da3_streaming.py:617)."""
    from . import posegraph
    return posegraph.optimize(sequential, loops, max_iterations=max_iterations, lambda_init=lambda_init)


def accumulate_sim3(chain):
    """utils/geometry.py:73-119: identity first, then left-to-right composition
    (output length = len(chain) + 1).  Host float64: it is len(chain) 3x3 products."""
    if not chain:
        return []
    acc = [(1.0, np.eye(3), np.zeros(3)), (chain[0][0], chain[0][1], chain[0][2])]
    for i in range(1, len(chain)):
        s_n, R_n, t_n = chain[i]
        s_p, R_p, t_p = acc[i]
        acc.append((s_p * s_n, R_p @ R_n, s_p * (R_p @ t_n) + t_p))
    return acc


def export_map(submaps, cumulative, voxel, conf_percentile=None, conf_thr=None, mode="fast", table_slots=None,
               max_voxels=None, skip_overlap=0, sort=True):
    """Unproject every submap to its world frame, move it into submap 0's frame with its
    cumulative Sim(3) (fused into the unprojection kernel), keep confident pixels and
    voxel-downsample the union.  conf_percentile follows viewer.py:333-336 per submap
    (percentile of the positive confidences, keep >=); conf_thr is a fixed '>' threshold."""
    dev = submaps[0].depth.device
    clouds = []
    for k, sm in enumerate(submaps):
        s, R, t = cumulative[k]
        row = ops.sim3_row(s, R, t, dev)
        first = skip_overlap if k > 0 else 0
        depth, conf, cams = sm.depth[first:], sm.conf[first:], sm.cams[first:]
        kw = dict(mode=mode, world=True, sim3=row, depth_eps=1e-6, want_count=False)
        if conf_percentile is not None:
            sel = ops.select([dict(a=conf, kind=L.SEL_POSITIVE, stat=L.SEL_PERCENTILE, percent=float(min(conf_percentile, 99.9)))], dev)
            kw.update(conf_cmp=">=", conf_thr=float(sel["value"][0]), conf_floor=0.0)
        elif conf_thr is not None:
            kw.update(conf_cmp=">", conf_thr=float(conf_thr))
        xyz, mask, _ = ops.unproject_filter(depth, conf, cams, **kw)
        rgb = sm.images[first:].reshape(-1, 3) if sm.images is not None else None
        clouds.append((xyz.view(-1, 3), rgb, mask.reshape(-1)))
    return ops.voxel_downsample(clouds, voxel, table_slots, max_voxels, sort=sort)


class SequencePlan:
    """Everything that is static for one submap sequence resident in HBM — pair table,
    selection segments, output buffers, the voxel grid — so that run() ONLY enqueues kernels
    on the current stream (no allocation, no host<->device copy, no synchronisation):

        pair thresholds (exact medians) -> [RANSAC] -> IRLS Umeyama -> rows
        -> Sim(3) chain (device) -> per-submap confidence percentile (exact)
        -> unproject + Sim(3) + filter per submap -> voxel grid insert -> compaction.

    read() synchronises and returns host/trimmed results.  This is one "step" of bench.py."""

    def __init__(self, submaps, overlap=1, voxel=0.02, conf_percentile=65.0, unproject_mode="fast", table_slots=None,
                 max_voxels=None, sample_idx=None, export=True, skip_overlap=True, fuse_export=True, exchange=None,
                 overlap_percentile=True, pairs=None, export_submaps=None, chain_index=None, n_chain=None,
                 rows_hook=None, chain=True, **opt_kw):
        """pairs: list of (i, j) indices into `submaps` (target, source); default = consecutive (k, k + 1).
        The remaining keywords describe a SHARD of a longer sequence (sharding.shard_sequence): `export_submaps` = the
        local submaps this rank exports (default all), `chain_index[i]` = position of local submap i in the global
        chain of `n_chain` submaps (default i), `rows_hook(rows_local) -> rows_global` = the exchange of the Sim(3)
        rows between ranks (default identity), chain=False skips the accumulation (loop candidates)."""
        self.submaps = submaps
        self.n = len(submaps)
        self.dev = submaps[0].depth.device
        self.F, self.H, self.W = submaps[0].depth.shape
        self.overlap = overlap
        self.voxel = float(voxel)
        self.mode = unproject_mode
        self.export = export
        self.side = None
        # exchange: a sharding.VoxelExchange -> the step ends with the multi-GPU merge of the rank-local grids
        # (each rank keeps the voxels whose key it owns) instead of a local compaction
        self.exchange = exchange
        # fuse_export: the exported points go straight from the depth maps into the voxel grid (one kernel);
        # False keeps the per-point arrays (self.xyz / self.mask) for callers that want the full cloud
        self.fuse_export = bool(fuse_export) and unproject_mode == "fast"
        self.opts = L.default_opts(**opt_kw)
        self.pairs = [(k, k + 1) for k in range(self.n - 1)] if pairs is None else [(int(i), int(j)) for i, j in pairs]
        self.n_pairs = len(self.pairs)
        self.rows_hook = rows_hook
        self.chain = bool(chain)
        self.chain_index = list(range(self.n)) if chain_index is None else [int(c) for c in chain_index]
        self.n_chain = int(n_chain) if n_chain is not None else self.n
        export_ids = list(range(self.n)) if export_submaps is None else [int(i) for i in export_submaps]
        self.export_ids = export_ids
        self.entries = [pair_entry(submaps[i], submaps[j], overlap) for i, j in self.pairs]
        self.pair_table = ops.make_pairs(self.entries, self.dev) if self.entries else None
        self.sample_idx = None
        if self.opts.n_hyp > 0:
            if sample_idx is None:
                raise ValueError("SequencePlan: n_hyp > 0 needs sample_idx [n_pairs, n_hyp, 3] (pixel indices drawn by the caller)")
            if tuple(sample_idx.shape) != (self.n_pairs, self.opts.n_hyp, 3):
                raise ValueError(f"SequencePlan: sample_idx must be [{self.n_pairs}, {self.opts.n_hyp}, 3], got {tuple(sample_idx.shape)}")
            self.sample_idx = sample_idx.to(self.dev, torch.int32).contiguous()
        self.rows = None
        self.rows_local = torch.empty((max(self.n_pairs, 1), L.ROW_LEN), dtype=torch.float64, device=self.dev)[:self.n_pairs]
        self.cum = torch.empty((self.n_chain, 13), dtype=torch.float64, device=self.dev)
        if export:
            # frames re-observed by the next submap are exported once (solver.py:100-114 adds them twice)
            self.first = [(overlap if (skip_overlap and self.chain_index[i] > 0) else 0) for i in range(self.n)]
            segs = [dict(a=submaps[i].conf[self.first[i]:], kind=L.SEL_POSITIVE, stat=L.SEL_PERCENTILE,
                         percent=float(min(conf_percentile, 99.9))) for i in export_ids]
            # the export thresholds depend on the confidence maps only, not on the alignment: with overlap_percentile
            # the pairs are aligned on a high-priority side stream (its short dependent kernels are scheduled first)
            # while the thresholds are selected on the caller's stream with their own da3s_ctx (= own scratch)
            self.side = torch.cuda.Stream(self.dev, priority=-1) if (overlap_percentile and self.n_pairs > 0) else None
            self.fork, self.join = torch.cuda.Event(), torch.cuda.Event()
            self.percentiles = ops.SelectPlan(segs, self.dev, private_ctx=overlap_percentile)
            total = sum((self.F - self.first[i]) * self.H * self.W for i in export_ids)
            if table_slots is None:
                table_slots = 1 << max(12, int(np.ceil(np.log2(max(total // 4, 4096)))))
                table_slots = min(table_slots, 1 << 27)
            if max_voxels is None:
                max_voxels = table_slots
            self.grid = ops.VoxelGrid(self.dev, table_slots, max_voxels, submaps[0].images is not None)
            if self.fuse_export:
                frames = []
                for e, i in enumerate(export_ids):
                    sm, f0 = submaps[i], self.first[i]
                    for f in range(f0, self.F):
                        frames.append(dict(depth=sm.depth[f], conf=sm.conf[f], cam=sm.cams[f], sim3=self.cum[self.chain_index[i]],
                                           conf_thr=self.percentiles.value_ptr_tensor(e),
                                           rgb=sm.images[f] if sm.images is not None else None))
                self.export_jobs = self.grid.make_export_jobs(frames)
            else:
                self.xyz = [torch.empty((self.F - self.first[i], self.H, self.W, 3), dtype=torch.float32, device=self.dev) for i in export_ids]
                self.mask = [torch.empty((self.F - self.first[i], self.H, self.W), dtype=torch.uint8, device=self.dev) for i in export_ids]
                # one job per exported frame: static pointers into the submaps, the cumulative Sim(3) table,
                # the percentile records and the output buffers -> the whole export is ONE unprojection launch
                jobs = []
                for e, i in enumerate(export_ids):
                    sm, f0 = submaps[i], self.first[i]
                    for f in range(f0, self.F):
                        jobs.append(dict(depth=sm.depth[f], conf=sm.conf[f], cam=sm.cams[f], sim3=self.cum[self.chain_index[i]],
                                         conf_thr=self.percentiles.value_ptr_tensor(e), xyz=self.xyz[e][f - f0],
                                         mask=self.mask[e][f - f0]))
                self.n_jobs = len(jobs)
                self.job_table = ops.make_frame_jobs(jobs, self.dev)
                # every submap's cloud goes into the grid in one launch
                self.rgb_views = [submaps[i].images[self.first[i]:] if submaps[i].images is not None else None for i in export_ids]
                self.voxel_jobs = self.grid.make_jobs(list(zip(self.xyz, self.rgb_views, self.mask)))
            self.points_per_step = total
        else:
            self.points_per_step = 0
        ops.context(self.dev)            # make sure the context (and its workspace) exists before the first timed run

    def _align(self):
        if self.n_pairs:
            ops.align_pairs(self.pair_table, self.n_pairs, self.overlap, self.H, self.W, self.opts, self.sample_idx,
                            rows_out=self.rows_local)

    def run(self, mark=None, after_align=None):
        """Enqueue one step.  `mark(name)` is called between stages (bench.py records CUDA events);
        `after_align()` right after the Sim(3) rows have been enqueued."""
        mark = mark or (lambda name: None)
        main = torch.cuda.current_stream(self.dev)
        if self.export and self.side is not None:
            self.fork.record(main)                                   # inputs are ready where the caller's stream is now
            self.side.wait_event(self.fork)
            with torch.cuda.stream(self.side), ops.nvtx_range("da3s.align"):
                self._align()
            self.join.record(self.side)
            with ops.nvtx_range("da3s.export_percentiles"):
                self.percentiles.run()
            main.wait_event(self.join)
        else:
            with ops.nvtx_range("da3s.align"):
                self._align()
        mark("align")
        # the only data that ever crosses NVLink on the alignment path: the [n,16] rows (sharding.RowExchange)
        self.rows = self.rows_hook(self.rows_local) if self.rows_hook is not None else self.rows_local
        if self.rows_hook is not None:
            mark("row_exchange")
        if after_align is not None:
            after_align()
        if self.chain:
            with ops.nvtx_range("da3s.chain"):
                ops.accumulate_sim3(self.rows, out=self.cum)
        if not self.export:
            mark("chain")
            return
        if self.side is None:
            with ops.nvtx_range("da3s.export_percentiles"):
                self.percentiles.run()
        mark("percentile")
        self.grid.begin()
        mark("voxel_clear")
        with ops.nvtx_range("da3s.export"):
            if self.fuse_export:
                self.grid.insert_frames(self.export_jobs, self.H, self.W, self.voxel, world=True, conf_cmp=">=", conf_floor=0.0,
                                        depth_eps=1e-6)
                mark("export_fused")
            else:
                ops.unproject_filter_jobs(self.job_table, self.n_jobs, self.H, self.W, mode=self.mode, world=True, conf_cmp=">=",
                                          conf_floor=0.0, depth_eps=1e-6)
                mark("unproject")
                self.grid.insert_jobs(self.voxel_jobs, self.voxel, width=self.W)
                mark("voxel_insert")
        with ops.nvtx_range("da3s.compact"):
            if self.exchange is not None:
                self.exchange.merge(self.grid, self.voxel)
                mark("voxel_merge")
            else:
                self.grid.finish(self.voxel)
                mark("voxel_compact")

    def contexts(self):
        """Every da3s_ctx this plan enqueues on: the device context the alignment and the grid share, plus the private one
        of the export percentiles."""
        ctxs = {id(c): c for c in (ops.context(self.dev), getattr(getattr(self, "grid", None), "ctx", None),
                                    getattr(getattr(self, "percentiles", None), "ctx", None)) if c is not None}
        return list(ctxs.values())

    @property
    def launches(self) -> int:
        """Kernel launches so far by every da3s_ctx of contexts()."""
        return sum(c.launches for c in self.contexts())

    def read(self, sort=False):
        out = {"rows": self.rows.cpu().numpy(), "cum": self.cum.cpu().numpy()}
        if self.export:
            xyz, rgb, cnt, key = self.grid.read(sort=sort)
            out.update(voxel_xyz=xyz, voxel_rgb=rgb, voxel_count=cnt, voxel_key=key)
        return out


class SequenceStream:
    """Serving loop for sequences whose predictions arrive in (pinned) host memory: a few slots, each with its
    own device buffers and SequencePlan, three streams.  While sequence k is aligned and exported, sequence
    k+1 is uploaded (host->device) and the results of sequence k-1 are read back (device->host): the PCIe
    link is full duplex, so a step costs max(upload, compute, download) instead of their sum.

        stream = SequenceStream(example_host_submaps, device, overlap=1, voxel=0.02, ...)
        for result in stream.process(iterable_of_host_submap_lists):
            result["rows"], result["voxel_xyz"], ...        # pinned views, valid until the slot is reused

    Host submaps: dicts with depth, conf, intrinsics, extrinsics (+ processed_images), shapes as in the example."""

    class _Slot:
        pass

    def __init__(self, example, device="cuda", slots=2, general_inverse=False, with_keys=False, upload_streams=2, **plan_kw):
        self.dev = torch.device(device)
        self.with_keys = with_keys                                  # also download the voxel keys (canonical order = ascending key)
        self.general_inverse = general_inverse
        self.s_in, self.s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        # several upload streams: concurrent copies keep the host->device link busier than one queue of copies
        self.s_up = [self.s_in] + [torch.cuda.Stream(self.dev) for _ in range(max(0, int(upload_streams) - 1))]
        self.slots = []
        for _ in range(slots):
            sl = SequenceStream._Slot()
            sl.subs = [DeviceSubmap.from_prediction(p, self.dev, general_inverse) for p in example]
            sl.plan = SequencePlan(sl.subs, **plan_kw)
            n_pairs = sl.plan.n_pairs
            sl.rows = torch.empty((n_pairs, 16), dtype=torch.float64, pin_memory=True)
            sl.cum = torch.empty((sl.plan.n, 13), dtype=torch.float64, pin_memory=True)
            sl.nv = torch.zeros((2,), dtype=torch.int64, pin_memory=True)
            sl.vox = None
            if sl.plan.export:
                g = sl.plan.grid
                sl.vox = dict(xyz=torch.empty((g.max_voxels, 3), dtype=torch.float32, pin_memory=True),
                              rgb=torch.empty((g.max_voxels, 3), dtype=torch.uint8, pin_memory=True) if g.rgb is not None else None,
                              count=torch.empty((g.max_voxels,), dtype=torch.int32, pin_memory=True),
                              key=torch.empty((g.max_voxels,), dtype=torch.int64, pin_memory=True) if with_keys else None)
            sl.h2d_done = sl.compute_done = sl.d2h_done = None
            self.slots.append(sl)
        torch.cuda.synchronize(self.dev)
        self.h2d_bytes = self.d2h_bytes = 0

    def _upload(self, sl, host_subs):
        streams = self.s_up
        for st in streams:
            if sl.compute_done is not None:
                st.wait_event(sl.compute_done)                      # the previous sequence in this slot has been consumed
        n = 0
        for k, (sm, hp) in enumerate(zip(sl.subs, host_subs)):
            st = streams[k % len(streams)]
            with torch.cuda.stream(st):
                for name, dst in (("depth", sm.depth), ("conf", sm.conf), ("intrinsics", sm.intrinsics), ("extrinsics", sm.extrinsics),
                                  ("processed_images", sm.images)):
                    if dst is None:
                        continue
                    src = hp[name] if isinstance(hp[name], torch.Tensor) else torch.from_numpy(hp[name])
                    dst.copy_(src, non_blocking=True)
                    n += dst.numel() * dst.element_size()
                ops.build_cams(sm.intrinsics, sm.extrinsics, self.general_inverse, out=sm.cams)
        for st in streams[1:]:
            join = torch.cuda.Event()
            join.record(st)
            self.s_in.wait_event(join)
        sl.h2d_done = torch.cuda.Event()
        sl.h2d_done.record(self.s_in)
        self.h2d_bytes = n

    def _compute(self, sl):
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(sl.h2d_done)
        if sl.d2h_done is not None:
            cur.wait_event(sl.d2h_done)                             # the slot's previous results have left the device
        sl.plan.run()
        sl.rows.copy_(sl.plan.rows, non_blocking=True)
        sl.cum.copy_(sl.plan.cum, non_blocking=True)
        if sl.plan.export:
            sl.nv.copy_(sl.plan.grid.nv, non_blocking=True)
        sl.compute_done = torch.cuda.Event()
        sl.compute_done.record(cur)

    def _download(self, sl):
        sl.compute_done.synchronize()                               # host: the result size is known
        out = {"rows": sl.rows.numpy(), "cum": sl.cum.numpy()}
        moved = sl.rows.numel() * 8 + sl.cum.numel() * 8
        if sl.plan.export:
            nv, dropped = int(sl.nv[0]), int(sl.nv[1])
            if dropped:
                raise L.Da3sError(L.ENOMEM, "SequenceStream", f"voxel table full: {dropped} points dropped")
            g = sl.plan.grid
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(sl.compute_done)
                sl.vox["xyz"][:nv].copy_(g.xyz[:nv], non_blocking=True)
                if g.rgb is not None:
                    sl.vox["rgb"][:nv].copy_(g.rgb[:nv], non_blocking=True)
                sl.vox["count"][:nv].copy_(g.count[:nv], non_blocking=True)
                if self.with_keys:
                    sl.vox["key"][:nv].copy_(g.key[:nv], non_blocking=True)
                sl.d2h_done = torch.cuda.Event()
                sl.d2h_done.record(self.s_out)
            out.update(voxel_xyz=sl.vox["xyz"][:nv], voxel_rgb=sl.vox["rgb"][:nv] if g.rgb is not None else None,
                       voxel_count=sl.vox["count"][:nv], voxel_key=sl.vox["key"][:nv] if self.with_keys else None, n_voxels=nv)
            moved += nv * (12 + 4 + (3 if g.rgb is not None else 0) + (8 if self.with_keys else 0)) + 16
        self.d2h_bytes = moved
        return out, sl

    def process(self, sequences):
        """Generator: yields one result dict per input sequence, in order.  A result's pinned arrays are complete
        when it is yielded and stay valid until `slots` more sequences have been submitted."""
        it = iter(sequences)
        pending = []                                                # slots whose compute has been enqueued
        k = 0
        nxt = next(it, None)
        if nxt is not None:
            self._upload(self.slots[0], nxt)
        while nxt is not None:
            sl = self.slots[k % len(self.slots)]
            following = next(it, None)
            if following is not None:
                self._upload(self.slots[(k + 1) % len(self.slots)], following)   # overlaps with the compute below
            self._compute(sl)
            pending.append(sl)
            if len(pending) >= len(self.slots):                     # read back the oldest one while the newest computes
                out, done = self._download(pending.pop(0))
                if done.d2h_done is not None:
                    done.d2h_done.synchronize()
                yield out
            nxt, k = following, k + 1
        while pending:
            out, done = self._download(pending.pop(0))
            if done.d2h_done is not None:
                done.d2h_done.synchronize()
            yield out
