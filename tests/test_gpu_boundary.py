"""The documented drop-in boundary, exercised the way INTEGRATION.md tells a maintainer to bind it:
`da3s_align_pairs_host` (include/da3s.h) — HOST pointers in, [n,16] float64 rows out — through
`da3slam_b200.host.align_prediction_pairs` / `align_pairs_host_arrays`.  The rows must equal, bit for bit, what the
device-pointer entry (`da3s_align_pairs`) returns for the same predictions: with pageable and with pinned host
buffers, with RANSAC indices, and while a voxel table is active in the same context (the table owns the tail of the
workspace; the host entry stages its inputs at the head)."""
import numpy as np
import pytest
import torch

from da3slam_b200 import _lib as L
from da3slam_b200 import host, ops, synth
from da3slam_b200.pipeline import DeviceSubmap, pair_entry
from oracle import spec_port as sp

pytestmark = pytest.mark.gpu


def device_rows(subs, cuda, overlap, sample_idx=None, **kw):
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    entries = [pair_entry(dsubs[k], dsubs[k + 1], overlap) for k in range(len(dsubs) - 1)]
    si = None if sample_idx is None else torch.from_numpy(sample_idx).to(cuda)
    rows, _, _ = ops.align_pairs(ops.make_pairs(entries, cuda), len(entries), overlap, *dsubs[0].depth.shape[1:], L.default_opts(**kw), si)
    return rows.cpu().numpy()


def stacked(subs, overlap):
    o = overlap
    prev, cur = subs[:-1], subs[1:]
    A = [np.stack([p[k][-o:] for p in prev]).astype(np.float32) for k in ("depth", "conf", "intrinsics", "extrinsics")]
    B = [np.stack([c[k][:o] for c in cur]).astype(np.float32) for k in ("depth", "conf", "intrinsics", "extrinsics")]
    return A + B


@pytest.mark.parametrize("overlap", [1, 2])
def test_host_entry_equals_device_entry(cuda, overlap):
    H, W = 64, 80
    subs, _ = synth.make_sequence(5, 3, H, W, overlap=overlap, seed=60 + overlap)
    want = device_rows(subs, cuda, overlap, world=1)
    pairs = [(subs[k], subs[k + 1]) for k in range(4)]
    got = host.align_prediction_pairs(pairs, overlap=overlap, world=1)                      # pageable numpy
    assert got.dtype == np.float64 and got.shape == (4, 16) and np.array_equal(got, want)
    o = sp.align_pair(subs[0], subs[1], overlap=overlap, world=True)                        # and it is the oracle's answer
    assert int(got[0, 13]) == o["n_valid"] and abs(got[0, 0] - o["s"]) <= 1e-6 * o["s"]
    pinned = [torch.from_numpy(a).pin_memory() for a in stacked(subs, overlap)]             # pinned host buffers
    assert all(t.is_pinned() for t in pinned)
    got_p = host.align_pairs_host_arrays(*[t.numpy() for t in pinned], world=1)
    assert np.array_equal(got_p, want)
    # predictions given as objects with attributes (main_align.py) instead of dicts (solver.py)
    import types
    objs = [(types.SimpleNamespace(**a), types.SimpleNamespace(**b)) for a, b in pairs]
    assert np.array_equal(host.align_prediction_pairs(objs, overlap=overlap, world=1), want)


def test_host_entry_with_ransac_and_an_active_voxel_table(cuda):
    H, W, n_hyp = 48, 64, 64
    subs, _ = synth.make_sequence(4, 2, H, W, overlap=1, seed=71, outlier_ratio=0.3)
    rng = np.random.default_rng(5)
    si = rng.integers(0, H * W, size=(3, n_hyp, 3)).astype(np.int32)
    kw = dict(world=1, n_hyp=n_hyp, ransac_thr=0.02)
    want = device_rows(subs, cuda, 1, si, **kw)
    pairs = [(subs[k], subs[k + 1]) for k in range(3)]
    assert np.array_equal(host.align_prediction_pairs(pairs, overlap=1, sample_idx=si, **kw), want)
    # a grid that is being filled in the SAME context: begin + insert, align through the host entry, then finish
    pts = rng.normal(0, 0.5, (30000, 3)).astype(np.float32)
    grid = ops.VoxelGrid(cuda, 1 << 16, 1 << 16, False)
    assert grid.ctx is ops.context(cuda)
    grid.begin()
    grid.insert(torch.from_numpy(pts[:15000]).to(cuda), None, None, 0.05)
    got = host.align_prediction_pairs(pairs, overlap=1, sample_idx=si, **kw)
    assert grid.ctx is ops.context(cuda)                                                     # same context, same workspace
    grid.insert(torch.from_numpy(pts[15000:]).to(cuda), None, None, 0.05)
    grid.finish(0.05)
    xyz, _, cnt, key = grid.read(sort=True)
    e_xyz, _, e_cnt, e_key = sp.voxel_downsample(pts, 0.05, None, None)
    assert np.array_equal(got, want)
    assert np.array_equal(key.cpu().numpy(), e_key) and np.array_equal(cnt.cpu().numpy(), e_cnt)
    assert np.array_equal(xyz.cpu().numpy(), e_xyz)
    with pytest.raises(ValueError):
        host.align_prediction_pairs(pairs, overlap=1, **kw)                                  # n_hyp > 0 without indices


def test_export_keeps_a_submap_without_positive_confidence(cuda):
    """viewer.py:333-338: when no confidence is positive the percentile is undefined and the reference keeps every
    point; the export path must do the same instead of comparing against NaN (which would drop the submap silently)."""
    from da3slam_b200.pipeline import SequencePlan
    H, W, F = 32, 40, 2
    subs, _ = synth.make_sequence(3, F, H, W, overlap=1, seed=81, with_images=True)
    subs[1]["conf"][:] = 0.0                                                                 # nothing positive in submap 1
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    for fuse in (True, False):
        plan = SequencePlan(dsubs, overlap=1, voxel=0.05, conf_percentile=65.0, table_slots=1 << 15, world=1, fuse_export=fuse)
        plan.run()
        out = plan.read(sort=True)
        kept1 = (F - 1) * H * W                                                              # every pixel of the exported frames
        from oracle import ref_port as rp
        tot = 0
        for k in (0, 2):
            f0 = plan.first[k]
            m, _ = rp.viewer_conf_mask(subs[k]["conf"][f0:].reshape(-1), 65.0)
            tot += int((m & (subs[k]["depth"][f0:].reshape(-1) > np.float32(1e-6))).sum())
        assert int(out["voxel_count"].sum().item()) == tot + kept1, fuse


def test_sequence_plan_validates_its_arguments(cuda):
    from da3slam_b200.pipeline import SequencePlan
    subs, _ = synth.make_sequence(3, 2, 32, 40, overlap=1, seed=82)
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    with pytest.raises(ValueError):
        SequencePlan(dsubs, overlap=1, export=False, world=1, n_hyp=16)                     # RANSAC without sample_idx
    with pytest.raises(ValueError):
        SequencePlan(dsubs, overlap=1, export=False, world=1, n_hyp=16, sample_idx=torch.zeros((2, 8, 3), dtype=torch.int32))


def test_sharded_sequence_plans_equal_the_whole(cuda):
    """One sequence spread over two (virtual) ranks the way bench.py's hires_global_map section does it
    (sharding.shard_sequence + SequencePlan(pairs=, export_submaps=, chain_index=, n_chain=, rows_hook=)): every rank's rows are
    bit-identical to the same rows of the single-rank plan, the chain is identical, and the ranks' exports partition the
    whole export (every kept point lands in exactly one rank's grid; per key the counts add up to the whole map's)."""
    from da3slam_b200.pipeline import SequencePlan
    from da3slam_b200.sharding import shard_sequence
    H, W, F, n, world = 48, 64, 3, 5, 2
    subs, _ = synth.make_sequence(n, F, H, W, overlap=1, seed=91, with_images=True)
    dsubs = [DeviceSubmap.from_prediction(s, cuda) for s in subs]
    kw = dict(overlap=1, voxel=0.05, conf_percentile=65.0, table_slots=1 << 16, world=1)
    whole = SequencePlan(dsubs, **kw)
    whole.run()
    ref = whole.read(sort=True)
    rows_full = torch.from_numpy(ref["rows"]).to(cuda)
    merged = {}
    for r in range(world):
        sh = shard_sequence(n, r, world)
        a, b = sh["a"], sh["b"]
        local = dsubs[a:b + (1 if sh["halo"] else 0)]
        n_local_pairs = sh["pair_sizes"][r]
        first_pair = sum(sh["pair_sizes"][:r])
        seen = {}

        def hook(rows_local, first_pair=first_pair, n_local_pairs=n_local_pairs, seen=seen):
            seen["rows"] = rows_local.clone()
            return rows_full                                   # what RowExchange would hand back on every rank
        plan = SequencePlan(local, pairs=[(i, i + 1) for i in range(n_local_pairs)], export_submaps=list(range(b - a)),
                            chain_index=[a + i for i in range(len(local))], n_chain=n, rows_hook=hook, **kw)
        plan.run()
        out = plan.read(sort=True)
        assert np.array_equal(seen["rows"].cpu().numpy(), ref["rows"][first_pair:first_pair + n_local_pairs])
        assert np.array_equal(out["cum"], ref["cum"])
        for k_, c_ in zip(out["voxel_key"].cpu().numpy(), out["voxel_count"].cpu().numpy()):
            merged[int(k_)] = merged.get(int(k_), 0) + int(c_)
    keys = np.array(sorted(merged))
    assert np.array_equal(keys, ref["voxel_key"].cpu().numpy())
    assert np.array_equal(np.array([merged[int(k_)] for k_ in keys]), ref["voxel_count"].cpu().numpy())
