#!/usr/bin/env python
"""bench.py — submap-alignment hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--workload seq2000|c3vd300|loop512|hires] [--no-sections]
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores), same config keys

One STEP = one pass of the whole hot path over one synthetic submap sequence already resident in HBM
(da3slam_b200.pipeline.SequencePlan.run): exact-median thresholds -> [RANSAC] -> IRLS Umeyama per consecutive
submap pair -> Sim(3) chain -> per-submap confidence percentile -> unproject + Sim(3) + filter + voxel-grid insert
(one fused kernel) -> compaction.  Metric: submap pairs aligned per second (whole job, all ranks); points/s next to it.

Default workload = `seq2000`, BASELINE.json configs[2], the largest single-GPU configuration (2000 frames, 32-frame
submaps at 518 x 518, RANSAC 1024 hypotheses).  N > 1 (torchrun): WEAK scaling — every rank owns its OWN sequence
(per-rank seed, so scenes, voxel counts and iteration counts differ between ranks); the only data-path exchange is one
all_gather of the [pairs, 16] float64 Sim(3) rows over NCCL inside the timed step.

Two further measurements ride on the same JSON line (`sections`; skipped with --no-sections):
  loop512_sharded    configs[3]: 512 independent loop-candidate pairs STRONG-sharded over the N ranks
                     (sharding.shard_range, per-pair seeds so the content does not depend on N, RowExchange in the timed
                     step, gathered table checked bit for bit against rank 0 aligning all 512 pairs alone)
  hires_global_map   configs[4]: ONE sequence of 1036 x 1036, 64-frame submaps spread over the N ranks, one global voxel map:
                     rows all_gather -> chain -> every rank exports its own submaps -> voxel records routed to their owner
                     rank through NVLink peer memory (sharding.VoxelExchange) -> each rank emits its share
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: 300 frames, 16-frame submaps (solver.py deque semantics -> 19 submaps / 18 pairs)
    "c3vd300": dict(n_submaps=19, frames=16, H=518, W=518, overlap=1, n_hyp=0, outlier=0.0, export=True,
                    table_slots=1 << 25, max_voxels=1 << 24, voxel=0.02, desc="300 frames, 19 submaps x 16 x 518x518, 18 pairs"),
    # configs[2]: 2000 frames, 32-frame submaps, RANSAC 1024 hypotheses (make_image_chunks -> 65 submaps / 64 pairs)
    "seq2000": dict(n_submaps=65, frames=32, H=518, W=518, overlap=1, n_hyp=1024, outlier=0.3, export=True,
                    table_slots=1 << 27, max_voxels=1 << 26, voxel=0.02,
                    desc="2000 frames, 65 submaps x 32 x 518x518, 64 pairs, RANSAC 1024"),
    # configs[3]: 512 independent loop-candidate pairs (2-frame submaps, so every pair reads distinct frames)
    "loop512": dict(n_submaps=513, frames=2, H=518, W=518, overlap=1, n_hyp=0, outlier=0.0, export=False,
                    table_slots=0, max_voxels=0, voxel=0.02, desc="512 submap pairs, 518x518, 1 overlap frame, alignment only"),
    # configs[4]: 1036x1036, 64-frame submaps + global voxel map
    "hires": dict(n_submaps=8, frames=64, H=1036, W=1036, overlap=1, n_hyp=0, outlier=0.0, export=True,
                  table_slots=1 << 25, max_voxels=1 << 24, voxel=0.02, desc="8 submaps x 64 x 1036x1036, 7 pairs, voxel map"),
    # tiny case for CI / smoke runs of this script
    "tiny": dict(n_submaps=4, frames=4, H=64, W=80, overlap=1, n_hyp=16, outlier=0.1, export=True,
                 table_slots=1 << 16, max_voxels=1 << 16, voxel=0.05, desc="4 submaps x 4 x 64x80"),
}
RANSAC_THR = 0.02
CONF_PERCENTILE = 65.0          # viewer.py:86-88 default slider value
SUBMAP_SCALE = (0.8, 1.25)      # every synthetic submap's own metric scale (synth.make_sequence_device abs_scale)
METRIC, UNIT = "submap_pairs_aligned_per_sec", "submap-pairs/s"


def config_of(name, w, world, extra=None):
    """The `config` object — identical keys and values on both arms (ours / reference)."""
    M = w["overlap"] * w["H"] * w["W"]
    inputs_mb = w["n_submaps"] * w["frames"] * w["H"] * w["W"] * 8 / 1e6
    cfg = {"workload": name, "desc": w["desc"], "n_submaps_per_gpu": w["n_submaps"], "frames_per_submap": w["frames"],
           "H": w["H"], "W": w["W"], "overlap": w["overlap"], "pairs_per_gpu": w["n_submaps"] - 1,
           "correspondences_per_pair": M, "n_hyp": w["n_hyp"], "ransac_thr": RANSAC_THR if w["n_hyp"] else None,
           "outlier_ratio": w["outlier"], "voxel": w["voxel"] if w["export"] else None,
           "conf_percentile": CONF_PERCENTILE if w["export"] else None,
           "irls": "huber delta=1.0, <=20 it, tol 1e-6 (utils/align.py defaults)",
           "seeds": "1234 + 1000 * rank (every rank its own scenes)",
           "submap_scale": f"every submap at its own scale, uniform in [{SUBMAP_SCALE[0]}, {SUBMAP_SCALE[1]}]",
           "l2": f"inputs {inputs_mb:.0f} MB (depth + conf) per GPU vs 126 MB L2; no explicit flush",
           "parallelism": f"{world} rank(s), one sequence per rank (weak scaling), Sim(3) rows all_gather only"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy, reference algorithm) on the host cores
# --------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker_init():
    os.environ["OMP_NUM_THREADS"] = "1"             # one unit per process: the threads used == the worker count
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(1)
    except Exception:
        pass


def _cpu_pair(k):
    """One submap pair through the reference's algorithm: thresholds (utils/align.py:140-146), fused-equivalent
    unprojection, [RANSAC: oracle/SPEC.md 4], IRLS (utils/align.py:169-211) with weighted Umeyama (:14-40)."""
    from oracle import spec_port as sp
    (prev, cur), w = _CPU["pairs"][k % len(_CPU["pairs"])], _CPU["w"]
    t0 = time.perf_counter()
    if w["n_hyp"] == 0:
        out = sp.align_pair(prev, cur, overlap=w["overlap"], world=True)
        t_ransac = 0.0
    else:
        rng = np.random.default_rng(1000 + k)
        si = rng.integers(0, w["H"] * w["W"] * w["overlap"], size=(w["n_hyp"], 3))
        corr = sp.pair_correspondences(prev, cur, w["overlap"], True)
        t1 = time.perf_counter()
        xs, ys = sp.ransac_points(corr, True)
        A, T, ok, _ = sp.ransac_hypotheses(xs, ys, corr["mask"], si)
        counts = sp.ransac_score_c(A, T, ok, xs, ys, corr["mask"], RANSAC_THR)
        best, _ = sp.ransac_best(counts, ok)
        mask = sp.ransac_inlier_mask_c(A, T, best, xs, ys, corr["mask"], RANSAC_THR)
        t_ransac = time.perf_counter() - t1
        s, R, t, info = sp.irls_dense(corr["x"], corr["y"], corr["c"], mask)
        out = dict(s=s)
    return ("pair", time.perf_counter() - t0, t_ransac, float(out["s"]))


def _cpu_submap(k):
    """Reference-style export of one submap, the SAME frames the GPU arm exports (overlap frames once): float64
    unprojection (utils/geometry.py:4-40), Sim(3) (utils/geometry.py:43-70), percentile-of-positive-confidence filter
    (viewer.py:333-336), voxel grid (oracle/SPEC.md 5 — the reference has none)."""
    from oracle import ref_port as rp
    from oracle import spec_port as sp
    sub, w = _CPU["sub"], _CPU["w"]
    f0 = w["overlap"]
    t0 = time.perf_counter()
    world = rp.unproject_world_f64(sub["depth"][f0:], sub["intrinsics"][f0:], sub["extrinsics"][f0:])
    world = rp.apply_sim3(world, 1.1, np.eye(3), np.array([0.1, 0.2, 0.3]))
    conf = sub["conf"][f0:].reshape(-1)
    mask, _ = rp.viewer_conf_mask(conf, CONF_PERCENTILE)
    mask &= (sub["depth"][f0:].reshape(-1) > 1e-6)
    t1 = time.perf_counter()
    sp.voxel_downsample(world.reshape(-1, 3).astype(np.float32), w["voxel"], None, mask)
    t2 = time.perf_counter()
    return ("submap", t2 - t0, t2 - t1, int(mask.sum()))


def _cpu_unit(arg):
    kind, k = arg
    return _cpu_pair(k) if kind == "pair" else _cpu_submap(k)


class CpuArm:
    """The reference's CPU path on a BOUNDED sample of the workload with EVERY host core busy: `workers` single-thread
    processes run at the same time, some on one submap pair each and (where the workload exports a map) the others on
    one submap export each.  Per-unit latencies are therefore measured under full load; the machine's rates are
    workers / latency, and the step time of the whole workload is n_pairs / pair_rate + n_submaps / export_rate."""

    def __init__(self, w, max_workers=None):
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        from da3slam_b200 import synth
        self.w = w
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.host_cores = os.cpu_count() or cores
        self.workers = max(1, min(max_workers or cores, cores))
        # pairs only read the overlap frames: 2-frame submaps of the workload's frame size are the same work
        subs, _ = synth.make_sequence(3, max(2, w["overlap"] + 1), w["H"], w["W"], w["overlap"], seed=4321, outlier_ratio=w["outlier"])
        _CPU["pairs"] = [(subs[0], subs[1]), (subs[1], subs[2])]
        _CPU["sub"] = None
        if w["export"]:
            full, _ = synth.make_sequence(1, w["frames"], w["H"], w["W"], w["overlap"], seed=4322)
            _CPU["sub"] = full[0]
        _CPU["w"] = w
        # a 31-frame float64 submap is ~1 GB of temporaries, a 63-frame 1036^2 one ~8 GB: cap the exports in flight
        self.sub_cap = self.workers if w["H"] <= 600 else max(1, min(self.workers, 4))
        self.pool = ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context("fork"), initializer=_cpu_worker_init)

    def step(self):
        """One bounded sample, every worker busy.  Returns rates and the workload's step time derived from them."""
        w = self.w
        n_pairs_total, n_sub_total = w["n_submaps"] - 1, w["n_submaps"]
        n_sub = min(self.sub_cap, max(1, self.workers // 2)) if w["export"] else 0
        n_pair = max(1, self.workers - n_sub)
        units = [("submap", k) for k in range(n_sub)] + [("pair", k) for k in range(n_pair)]     # long units first
        t0 = time.perf_counter()
        res = list(self.pool.map(_cpu_unit, units))
        wall = time.perf_counter() - t0
        pair_res = [r for r in res if r[0] == "pair"]
        sub_res = [r for r in res if r[0] == "submap"]
        lat_pair = float(np.mean([r[1] for r in pair_res]))
        lat_sub = float(np.mean([r[1] for r in sub_res])) if sub_res else None
        pair_rate = self.workers / lat_pair
        sub_rate = self.sub_cap / lat_sub if sub_res else None
        step_s = n_pairs_total / pair_rate + (n_sub_total / sub_rate if sub_rate else 0.0)
        return dict(step_s=step_s, pairs_per_s=n_pairs_total / step_s, cores=self.workers, host_cores=self.host_cores,
                    align_only_pairs_per_s=pair_rate, export_submaps_per_s=sub_rate, sample_wall_s=wall,
                    lat_pair_s=lat_pair, lat_pair_ransac_s=float(np.mean([r[2] for r in pair_res])),
                    lat_submap_s=lat_sub, lat_submap_voxel_s=float(np.mean([r[2] for r in sub_res])) if sub_res else None,
                    sample=f"{len(pair_res)} pairs" + (f" + {len(sub_res)} submap exports ({w['frames'] - w['overlap']} frames each)" if sub_res else "")
                           + f" at the same time on {self.workers} single-thread workers ({wall:.1f} s); rates = workers / per-unit latency "
                           + f"under that load; step = {n_pairs_total} pairs / pair rate" + (f" + {n_sub_total} submaps / export rate" if sub_res else ""))

    def close(self):
        self.pool.shutdown(wait=True, cancel_futures=True)


def cpu_baseline_entry(info):
    return {"value": info["pairs_per_s"], "unit": UNIT, "cores": info["cores"], "kind": "port",
            "sample": info["sample"], "host_cores": info["host_cores"], "step_s": info["step_s"],
            "note": "oracle port of the reference's functions (oracle/ref_port.py pinned by tests/golden; RANSAC and the voxel "
                    "grid are the builder's own spec, oracle/SPEC.md — the reference has neither); a reported baseline, not a target",
            "parts": {"align_only_pairs_per_s": info["align_only_pairs_per_s"], "export_submaps_per_s": info["export_submaps_per_s"],
                      "per_pair_latency_s": info["lat_pair_s"], "of_which_ransac_s": info["lat_pair_ransac_s"],
                      "per_submap_export_latency_s": info["lat_submap_s"], "of_which_voxel_grid_s": info["lat_submap_voxel_s"]}}


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.3)                 # the first sample takes a moment: be sampling before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, torch copy, burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_numa_node(local):
    try:
        import torch
        prop = torch.cuda.get_device_properties(local)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            return int(fh.read().strip())
    except Exception:
        return None


def bind_to_gpu_numa(local):
    """Run this process (and place its pinned host buffers, first-touch) on the NUMA node the GPU hangs off: with
    several ranks uploading GBs per step each, host->device copies that cross the socket interconnect halve the
    end-to-end rate.  Returns the node, or None when the topology cannot be read (then nothing is changed)."""
    node = gpu_numa_node(local)
    try:
        if node is None or node < 0:
            return node
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return node


class Timer:
    """K steps bracketed by barrier + synchronize, timed with CUDA events on the launching stream, MAX over ranks."""

    def __init__(self, dev, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def run(self, step, steps, warmup, want_events=False):
        torch = self.torch
        for _ in range(warmup):
            step(None)
        self.barrier()
        all_events = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            ev = [] if want_events else None
            step(ev)
            if want_events:
                all_events.append(ev)
        t1.record()
        self.barrier()
        ms = t0.elapsed_time(t1)
        if self.world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
            ms = float(tt.item())
        stages = {}
        for ev in all_events:
            for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
                stages[n1] = stages.get(n1, 0.0) + e0.elapsed_time(e1) / steps
        return ms / steps, stages


def make_mark(events):
    import torch

    def mark(name):
        if events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events.append((name, e))
    return mark


def load_traffic(workload, stage):
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as fh:
                ent = json.load(fh).get(workload, {}).get(stage)
            if ent:
                return float(ent["bytes"]), f"{ent['kernel']}: {ent['source']}"
        except (OSError, ValueError):
            pass
    return None, None


def main_section(args, w, rank, world, dev, local, timer, numa_node):
    """The headline measurement: one full sequence per rank, resident in HBM (value) and from pinned host memory (e2e)."""
    import torch
    import torch.distributed as dist
    from da3slam_b200 import _lib as L
    from da3slam_b200 import ops, synth
    from da3slam_b200.pipeline import DeviceSubmap, SequencePlan, SequenceStream
    from da3slam_b200.sharding import RowExchange

    seed = 1234 + 1000 * (rank if args.scene_of_rank is None else args.scene_of_rank)
    subs, gt = synth.make_sequence_device(w["n_submaps"], w["frames"], w["H"], w["W"], w["overlap"], seed=seed,
                                          outlier_ratio=w["outlier"], with_images=w["export"], device=dev, abs_scale=SUBMAP_SCALE)
    dsubs = [DeviceSubmap.from_prediction(s, dev) for s in subs]
    n_pairs = w["n_submaps"] - 1
    M = w["overlap"] * w["H"] * w["W"]
    sample_idx = None
    opt = dict(world=1)
    if w["n_hyp"] > 0:
        rng = np.random.default_rng(99 + rank)
        sample_idx = torch.from_numpy(rng.integers(0, M, size=(n_pairs, w["n_hyp"], 3)).astype(np.int32))
        opt.update(n_hyp=w["n_hyp"], ransac_thr=RANSAC_THR)
    plan_kw = dict(overlap=w["overlap"], voxel=w["voxel"], conf_percentile=CONF_PERCENTILE, table_slots=w["table_slots"] or None,
                   max_voxels=w["max_voxels"] or None, sample_idx=sample_idx, export=w["export"], **opt)
    plan = SequencePlan(dsubs, **plan_kw)
    # the only exchange between ranks: every rank's rows, all_gathered (each rank then holds every sequence's Sim(3) table)
    gather = RowExchange(world * n_pairs, dev, sizes=[n_pairs] * world) if world > 1 else None
    stage_rows = {}

    def run_step(events):
        mark = make_mark(events)
        mark("start")
        plan.run(mark)
        if gather is not None:
            stage_rows["all"] = gather(plan.rows)
            mark("allgather")

    for _ in range(args.warmup):
        run_step(None)
    timer.barrier()
    check = plan.read()                                 # raises if the voxel table overflowed
    rows0 = check["rows"]
    ok = rows0[:, 15] == 0
    err_s = float(np.max([abs(rows0[k, 0] - gt[k][0]) / gt[k][0] for k in range(n_pairs) if ok[k]])) if ok.any() else None
    iters = rows0[:, 14]
    n_vox = int(check["voxel_key"].shape[0]) if w["export"] else 0
    del check

    clocks = ClockSampler(local) if rank == 0 else None
    for c in plan.contexts():
        c.kernel_timers(True)                           # event pairs around the four big kernels, read after the timed region
    launches0 = plan.launches
    ms_per_step, stages = timer.run(run_step, args.steps, 0, want_events=True)
    launches = plan.launches - launches0
    kernel_ms = {}
    for c in plan.contexts():
        for which in L.TIMED_NAMES:
            t_ms, n_t, n_l, wk = c.kernel_time(which)
            if n_l:
                a_ms, a_t, a_n, a_w = kernel_ms.get(which, (0.0, 0, 0, 0.0))
                kernel_ms[which] = (a_ms + t_ms, a_t + n_t, a_n + n_l, a_w + wk)
        c.kernel_timers(False)
    clock_info = clocks.stop() if clocks else None
    per_rank = None
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"stages_ms": {k: round(v, 4) for k, v in stages.items()}, "voxels": n_vox,
                                          "irls_passes": float(np.sum(iters)), "numa_node": numa_node})

    # ---- end-to-end: pinned host predictions in, rows + voxel map out, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        del plan
        torch.cuda.empty_cache()
        host_seq = []
        for s_ in subs:
            hp = {}
            for k in ("depth", "conf", "processed_images", "intrinsics", "extrinsics"):
                if k in s_:
                    hp[k] = torch.empty(s_[k].shape, dtype=s_[k].dtype, pin_memory=True)
                    hp[k].copy_(s_[k])
            host_seq.append(hp)
        torch.cuda.synchronize()
        stream = SequenceStream(host_seq, dev, slots=2, upload_streams=int(os.environ.get("DA3S_UPLOAD_STREAMS", "2")), **plan_kw)
        k_stream = max(4, min(args.steps, 10))
        for _ in stream.process([host_seq] * 3):                 # warm-up (also fills the pipeline once)
            pass
        timer.barrier()
        t0 = time.perf_counter()
        n_out = 0
        for res in stream.process([host_seq] * k_stream):
            n_out += 1
            last_rows = res["rows"]
        timer.barrier()
        e2e_s = (time.perf_counter() - t0) / k_stream
        assert n_out == k_stream and np.array_equal(last_rows, rows0)       # same inputs -> same rows as the resident run
        h2d, d2h = stream.h2d_bytes, stream.d2h_bytes
        per_rank_e2e = None
        if world > 1:
            mine = {"ms": e2e_s * 1e3, "h2d_GBps": h2d / e2e_s / 1e9, "numa_node": numa_node}
            per_rank_e2e = [None] * world
            dist.all_gather_object(per_rank_e2e, mine)
            tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt[0].item())
        e2e = {"value": world * n_pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": k_stream,
               "h2d_GBps_per_rank": h2d / e2e_s / 1e9, "per_rank": per_rank_e2e,
               "api": "SequenceStream.process on pinned host predictions: 2 slots, upload k+1 | compute k | download k-1 on "
                      "three streams; every step uploads its inputs and downloads rows + voxel map inside the timed region"}
        del stream
        torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- rooflines: every big kernel against the resource that bounds it; `roofline` = the one with the largest share ----
    peak, peak_src = measured_peaks()
    px_export = sum((w["frames"] - (w["overlap"] if k > 0 else 0)) for k in range(w["n_submaps"])) * w["H"] * w["W"] if w["export"] else 0
    passes = float(np.sum(iters))
    kept = 1.0 - CONF_PERCENTILE / 100.0
    # algorithmic work PER STEP of each timed kernel: bytes for the HBM-bound ones, flops for the scoring
    work = {
        L.TIMED_IRLS: ("hbm", passes * 16.0 * M, "16 B per correspondence (depth + conf of both overlap frames) x executed IRLS passes, summed over the pairs"),
        L.TIMED_EXPORT_VOXEL: ("hbm", (8.0 + 3.0 * kept) * px_export,
                               "8 B/pixel read (depth + conf) + 3 B rgb per kept point; hash-table traffic is not algorithmic (see traffic)"),
        L.TIMED_VOXEL_EMIT: ("hbm", (64.0 + 27.0) * n_vox, "per voxel: 64 B record read, 27 B written (xyz 12, rgb 3, count 4, key 8)"),
        # oracle/SPEC.md 4: per (correspondence, hypothesis) 12 FMA + 3 add = 27 flop, over EVERY hypothesis and correspondence
        # the specification scores; the kernel skips invalid hypotheses and masked correspondences (they score 0 by definition)
        L.TIMED_RANSAC_SCORE: ("fp32", 27.0 * M * n_pairs * w["n_hyp"],
                               "27 flop per (hypothesis, correspondence) evaluation x the evaluations the kernel EXECUTED (valid hypotheses x "
                               "correspondences that pass the joint mask, counted on the device); invalid hypotheses and masked "
                               "correspondences score 0 by definition (oracle/SPEC.md 4) and are not algorithmic work"),
    }
    fp32_peak = ops.fp32_peak_tflops(dev) if w["n_hyp"] > 0 else None
    kernels = {}
    for which, (bound, amount, note) in work.items():
        sum_ms, n_t, n_l, executed = kernel_ms.get(which, (0.0, 0, 0, 0.0))
        if n_t == 0 or amount <= 0:
            continue
        if which == L.TIMED_RANSAC_SCORE:
            spec_flops, amount = amount, 27.0 * executed / args.steps       # what the kernel evaluated, counted on the device
        per_step = n_l / float(args.steps)
        avg_ms = sum_ms / n_t
        name = L.TIMED_NAMES[which]
        traffic, traffic_src = load_traffic(args.workload, name)
        if bound == "hbm":
            ach, pk, unit, pk_src = amount / per_step / (avg_ms * 1e-3) / 1e9, peak, "GB/s", peak_src
        else:
            ach, pk, unit = amount / per_step / (avg_ms * 1e-3) / 1e12, fp32_peak, "TFLOP/s"
            pk_src = "measured live: da3s_measure_fp32_peak (packed FFMA2 chains on all SMs); CUDA-core float32 FMA — the bit-exact fmaf specification has no tensor-core form"
        kernels[name] = {"bound": bound, "kernel": name, "achieved": ach, "peak": pk, "unit": unit, "frac": ach / pk,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_src,
                         ("algorithmic_bytes_per_launch" if bound == "hbm" else "algorithmic_flops_per_launch"): amount / per_step,
                         "launches_per_step": per_step, "avg_launch_ms": avg_ms, "timing": "CUDA event pair around each launch on the "
                         "launching stream, inside the timed region (da3s_kernel_time), averaged over the timed steps",
                         "share_of_step": avg_ms * per_step / ms_per_step, "note": note}
        if which == L.TIMED_RANSAC_SCORE:
            pipe, pipe_src = load_traffic(args.workload, "ransac_fp32_pipe")
            kernels[name].update(fma_pipe_busy_frac_ncu=pipe, fma_pipe_source=pipe_src, spec_flops_per_launch=spec_flops / per_step,
                                 executed_over_spec=amount / spec_flops,
                                 instruction_mix_ceiling="54 flop per 16 packed FMA-pipe instructions (12 FFMA2/FMUL2 + 4 FADD2) vs 64 for pure "
                                                         "FFMA2: at most 0.84 of the peak (profiles/r2_fp32_pipes.txt: 52.8 TFLOP/s measured for this mix)")
    roofline = max(kernels.values(), key=lambda k: k["share_of_step"]) if kernels else None
    hbm_only = [k for k in kernels.values() if k["bound"] == "hbm"]
    roofline_hbm = max(hbm_only, key=lambda k: k["share_of_step"]) if hbm_only else None
    stage_bytes = {"align": (passes * 16.0 + 3 * 8.0) * M + (w["n_hyp"] > 0) * n_pairs * 16.0 * M}
    if w["export"]:
        stage_bytes["export_fused"] = work[L.TIMED_EXPORT_VOXEL][1]
        stage_bytes["voxel_compact"] = (8.0 + 64.0 + 27.0) * n_vox
    per_stage = {k: {"ms": stages[k], "GB/s": b / (stages[k] * 1e-3) / 1e9, "frac_of_peak": b / (stages[k] * 1e-3) / 1e9 / peak}
                 for k, b in stage_bytes.items() if stages.get(k, 0) > 0.02}

    return {
        "metric": METRIC, "value": world * n_pairs / (ms_per_step * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": config_of(args.workload, w, world), "numa_node_rank0": numa_node,
        "points_per_sec": world * px_export / (ms_per_step * 1e-3) if w["export"] else None,
        "pixels_per_step_per_gpu": px_export, "voxels_out": n_vox,
        "timed_region_s": ms_per_step * args.steps * 1e-3,
        "stages_ms": stages, "per_rank": per_rank, "stage_bandwidth": per_stage,
        "accuracy": {"max_rel_scale_error_vs_ground_truth": err_s, "irls_iterations_mean": float(np.mean(iters)),
                     "pairs_ok": int(ok.sum())},
        "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "roofline_hbm": roofline_hbm,
        "kernels": kernels,
    }


def loop512_section(args, rank, world, dev, timer):
    """configs[3]: 512 independent loop-candidate pairs, STRONG-sharded over the ranks."""
    import torch
    from da3slam_b200 import synth
    from da3slam_b200.pipeline import DeviceSubmap, SequencePlan
    from da3slam_b200.sharding import RowExchange, shard_range
    n_total, H, W = (512, 518, 518) if args.workload != "tiny" else (12, 64, 80)

    def make_pairs(lo, hi):
        subs = []
        for p in range(lo, hi):                                      # per-pair seed: the content does not depend on the sharding
            pair, _ = synth.make_sequence_device(2, 1, H, W, 1, seed=77000 + p, with_images=False, device=dev)
            subs += [DeviceSubmap.from_prediction(s_, dev) for s_ in pair]
        return subs

    lo, hi = shard_range(n_total, rank, world)
    subs = make_pairs(lo, hi)
    gather = RowExchange(n_total, dev)
    plan = SequencePlan(subs, overlap=1, export=False, chain=False, world=1, pairs=[(2 * i, 2 * i + 1) for i in range(hi - lo)],
                        rows_hook=gather if world > 1 else None)

    def step(events):
        mark = make_mark(events)
        mark("start")
        plan.run(mark)

    steps = max(args.steps, 20)
    ms, stages = timer.run(step, steps, max(3, args.warmup), want_events=True)
    table = plan.rows.cpu().numpy()
    digest = hashlib.sha256(np.ascontiguousarray(table).tobytes()).hexdigest()
    identical = None
    if world > 1 and rank == 0:                                      # rank 0 aligns all 512 pairs alone: must be the same bits
        del plan
        full = make_pairs(0, n_total)
        ref = SequencePlan(full, overlap=1, export=False, chain=False, world=1, pairs=[(2 * i, 2 * i + 1) for i in range(n_total)])
        ref.run()
        identical = bool(np.array_equal(ref.rows.cpu().numpy(), table))
    timer.barrier()
    passes = float(table[:, 14].sum())
    peak, _ = measured_peaks()
    M = H * W
    return {"workload": "loop512 (BASELINE configs[3])", "scaling": "strong", "pairs_total": n_total, "pairs_per_rank": hi - lo,
            "value": n_total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "stages_ms": stages,
            "irls_passes_total": passes, "rows_sha256": digest, "bit_identical_to_single_rank": identical,
            "align_stage_GBps_per_gpu": ((passes / n_total * 16.0 + 16.0) * M * (hi - lo)) / (stages.get("align", ms) * 1e-3) / 1e9,
            "align_stage_frac_of_hbm_peak": ((passes / n_total * 16.0 + 16.0) * M * (hi - lo)) / (stages.get("align", ms) * 1e-3) / 1e9 / peak,
            "exchange": "RowExchange: one all_gather of [pairs/N, 16] float64 per step"}


def global_map_section(args, rank, world, dev, timer):
    """configs[4]: ONE sequence of 1036 x 1036, 64-frame submaps spread over the ranks, one global voxel map."""
    import torch
    import torch.distributed as dist
    from da3slam_b200 import synth
    from da3slam_b200.pipeline import DeviceSubmap, SequencePlan
    from da3slam_b200.sharding import RowExchange, VoxelExchange, shard_sequence
    w = WORKLOADS["hires"] if args.workload != "tiny" else dict(WORKLOADS["tiny"], n_hyp=0, outlier=0.0)
    n = w["n_submaps"]
    sh = shard_sequence(n, rank, world)
    a, b = sh["a"], sh["b"]
    # every rank draws the whole sequence (same seed) and keeps its own submaps + the halo submap's overlap frames
    local = []
    subs, _ = synth.make_sequence_device(n, w["frames"], w["H"], w["W"], w["overlap"], seed=555, with_images=True, device=dev,
                                         abs_scale=SUBMAP_SCALE)
    for k in range(a, b + (1 if sh["halo"] else 0)):
        local.append(DeviceSubmap.from_prediction(subs[k], dev))
    del subs
    torch.cuda.empty_cache()
    n_own = b - a
    n_pairs_local = sh["pair_sizes"][rank]
    gather = RowExchange(n - 1, dev, sizes=sh["pair_sizes"]) if world > 1 else None
    slots = w["table_slots"]
    exchange = VoxelExchange(dev, world, rank, cap=max(1 << 16, slots // max(1, world))) if world > 1 else None
    plan = SequencePlan(local, overlap=w["overlap"], voxel=w["voxel"], conf_percentile=CONF_PERCENTILE, table_slots=slots,
                        max_voxels=w["max_voxels"], export=True, exchange=exchange, world=1,
                        pairs=[(i, i + 1) for i in range(n_pairs_local)], export_submaps=list(range(n_own)),
                        chain_index=[a + i for i in range(len(local))], n_chain=n, rows_hook=gather)

    def step(events):
        mark = make_mark(events)
        mark("start")
        plan.run(mark)

    steps = max(args.steps, 20)
    ms, stages = timer.run(step, steps, max(3, args.warmup), want_events=True)
    out = plan.read()
    nv = int(out["voxel_key"].shape[0])
    px = plan.points_per_step
    tot = torch.tensor([nv, px], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        dist.all_reduce(tot)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"stages_ms": {k: round(v, 4) for k, v in stages.items()}, "voxels_owned": nv, "submaps": [a, b]})
    return {"workload": "hires (BASELINE configs[4])", "scaling": "strong", "desc": w["desc"], "submaps_total": n,
            "submaps_this_rank": [a, b], "ms_per_step": ms, "steps": steps, "stages_ms": stages, "per_rank": per_rank,
            "pixels_per_step_total": float(tot[1].item()), "points_per_sec": float(tot[1].item()) / (ms * 1e-3),
            "global_voxels": int(tot[0].item()), "value": (n - 1) / (ms * 1e-3), "unit": UNIT,
            "merge": ("voxel records routed to the rank owning their key through NVLink peer stores (VoxelExchange), "
                      "device-side arrival flags, no host synchronisation" if world > 1 else "single rank: local compaction")}


def gpu_arm(args, w, rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs CUDA: the alignment path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    # pinned host buffers are allocated after this: first touch places them next to the GPU
    numa_node = bind_to_gpu_numa(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                        # keeps NCCL's version banner off stdout (ONE JSON line)
        dist.init_process_group("nccl", device_id=dev)
    timer = Timer(dev, world)
    out = main_section(args, w, rank, world, dev, local, timer, numa_node)
    sections = {}
    if not args.no_sections:
        torch.cuda.empty_cache()
        for name, fn in (("loop512_sharded", loop512_section), ("hires_global_map", global_map_section)):
            try:
                sections[name] = fn(args, rank, world, dev, timer)
            except Exception as err:                                  # a section never takes the headline down with it
                sections[name] = {"error": f"{type(err).__name__}: {err}"}
                if world > 1:
                    raise
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()
    if out is not None:
        out["sections"] = sections
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="seq2000", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sections", action="store_true", help="skip the loop512_sharded / hires_global_map measurements")
    ap.add_argument("--cpu-workers", type=int, default=None)
    ap.add_argument("--scene-of-rank", type=int, default=None, help="N=1 only: run rank R's scenes (per-rank load check)")
    ap.add_argument("--table-slots-log2", type=int, default=None, help="override the workload's voxel table size (tuning)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.table_slots_log2 is not None and w["table_slots"]:
        w = dict(w, table_slots=1 << args.table_slots_log2, max_voxels=1 << (args.table_slots_log2 - 1))
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        if rank != 0:
            return 0
        arm = CpuArm(w, max_workers=args.cpu_workers)
        vals, info = [], None
        for _ in range(max(0, args.warmup)):
            arm.step()
        for _ in range(max(1, args.steps)):
            info = arm.step()
            vals.append(info["pairs_per_s"])
        arm.close()
        # the reference has no multi-GPU path: N ranks = N independent copies of the host's work, so the CPU arm's
        # whole-job value is the one host's rate regardless of --gpus
        v = float(np.mean(vals))
        base = cpu_baseline_entry(info)
        base["value"] = v
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * info["step_s"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": config_of(args.workload, w, world),
                "cpu_baseline": base,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(w, max_workers=args.cpu_workers)             # before CUDA is initialised (fork-safe)
        arm.step()                                                # warm-up: imports, page faults, the C oracle's dlopen
        cpu = arm.step()
        arm.close()
    out = gpu_arm(args, w, rank, world)
    if out is None:
        return 0
    out["cpu_baseline"] = cpu_baseline_entry(cpu) if cpu is not None else None
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
