"""CPU checks (no GPU): the C-ABI library loads and exports every symbol include/da3s.h
declares; host-side logic of the reference-named shims; loud failure without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "da3s.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(da3s_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from da3slam_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/da3s.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures and header disagree"
    assert lib.da3s_version() == 100
    assert lib.da3s_strerror(-2).decode() == "pointer not 16-byte aligned"


def test_struct_layouts_match_the_header():
    import ctypes as C
    from da3slam_b200 import _lib as L
    assert C.sizeof(L.Pair) == 48 and C.sizeof(L.SelectSeg) == 64 and C.sizeof(L.SelectOut) == 24
    assert L.CAM_BYTES == 200 and L.SelectOut.value.offset == 16
    o = L.default_opts()
    assert (o.world, o.huber, o.max_iterations, o.min_points, o.n_hyp) == (1, 1, 20, 100, 0)
    assert o.huber_delta == 1.0 and o.tol == 1e-6 and abs(o.depth_conf_th - 0.2) < 1e-7 and np.isnan(o.conf_thr_override)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import align_geometry
    import utils.align as ua
    import utils.geometry as ug
    from da3slam_b200 import ops
    with pytest.raises(RuntimeError):
        ops.context()
    with pytest.raises(RuntimeError):
        ug.apply_sim3_transform(np.zeros((4, 3)), 1.0, np.eye(3), np.zeros(3))
    with pytest.raises(RuntimeError):
        ua.weighted_umeyama_alignment(np.zeros((4, 3)), np.zeros((4, 3)), np.ones(4))
    with pytest.raises(RuntimeError):
        align_geometry.depth_to_point_cloud_vectorized(np.ones((1, 4, 4), np.float32), np.eye(3)[None], np.eye(4)[None, :3])


def test_shim_host_logic_matches_oracle(golden):
    import align_geometry as ag
    import utils
    import utils.align as ua
    import utils.align_geometry_single as ags
    import utils.geometry as ug
    from oracle import ref_port as rp
    g = golden("chunks")
    for key in g.files:
        if key.startswith("chunks_"):
            n, c, o = (int(x) for x in key.split("_")[1:])
            ch = ag.make_image_chunks(list(range(n)), c, o)
            assert [x[0] for x in ch] == list(g[key])
    g = golden("sim3_chain")
    chain = [(float(a), b, c) for a, b, c in zip(g["chain_s"], g["chain_R"], g["chain_t"])]
    acc = ug.accumulate_sim3_transforms(chain)
    assert np.array_equal(np.stack([a[1] for a in acc]), g["acc_R"]) and np.array_equal(np.stack([a[2] for a in acc]), g["acc_t"])
    assert ug.accumulate_sim3_transforms([]) == []
    assert np.array_equal(ug.transform_camara_extrinsics(g["E_local"][2], float(g["s"]), g["R"], g["t"]), g["rebase"])
    E = g["E_local"].astype(np.float64)
    assert np.array_equal(ag.compute_aligned_chunk_extrinsics_from_prev_overlap(E[4], E, g["T"]), g["chain_overlap"])
    assert ua.huber_weight(0.3) == 1.0 and ua.huber_weight(2.0) == 0.5 and ua.huber_weight(-4.0, 0.5) == 0.125
    assert np.array_equal(ags.to4x4(E[0])[:3], E[0]) and ags._get({"a": 1}, "a") == 1
    img = np.arange(2 * 3 * 4 * 3, dtype=np.uint8).reshape(2, 3, 4, 3)
    assert np.array_equal(ag.images_to_chw01(img), img.transpose(0, 3, 1, 2) / 255.0)
    assert np.array_equal(ags.image_to_chw01({"processed_images": img}, 1), img[1].transpose(2, 0, 1) / 255.0)
    # utils package re-exports the helpers of the reference's utils.py
    assert utils.extract_keyframe(list("abcdef"), 2) == ["a", "c", "e"] and utils.extract_keyframe(["a"], 0) == ["a"]
    assert utils.get_distinct_color(9) == utils.get_distinct_color(1)
    out = utils.apply_chunk_color_to_images_batch(np.zeros((2, 3, 4, 5)), 1)
    assert out.shape == (2, 3, 4, 5) and np.all(out[:, 1] == 1.0) and np.all(out[:, 0] == 0.0)
    with pytest.raises(ValueError):
        ags.get_aligned_chunk_extrinsics_single_overlap(None, {}, {})


def test_load_image_sorts_numerically(tmp_path):
    import utils
    for name in ("img10.png", "img2.png", "img1.jpg", "notes.txt"):
        (tmp_path / name).write_bytes(b"x")
    got = [os.path.basename(p) for p in utils.load_image(str(tmp_path))]
    assert got == ["img1.jpg", "img2.png", "img10.png"]
    assert utils.load_image(str(tmp_path / "missing")) == []


def test_shard_ranges_cover_everything():
    from da3slam_b200.sharding import shard_range, shard_sizes
    for n in (0, 1, 7, 18, 64, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1
