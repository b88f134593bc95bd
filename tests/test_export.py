"""Host-side export writers: PLY round trip, 3DGS initialisation layout, chained camera
poses against a direct restatement of utils/da3_streaming.py:733-774."""
import numpy as np

from da3slam_b200 import export
from oracle import ref_port as rp


def test_ply_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.normal(0, 1, (1000, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (1000, 3), dtype=np.uint8)
    p = tmp_path / "map.ply"
    assert export.write_ply(p, xyz, rgb) == 1000
    got = export.read_ply(p)
    assert np.array_equal(np.stack([got["x"], got["y"], got["z"]], 1), xyz)
    assert np.array_equal(np.stack([got["red"], got["green"], got["blue"]], 1), rgb)
    assert export.write_ply(tmp_path / "empty.ply", np.zeros((0, 3), np.float32)) == 0
    assert len(export.read_ply(tmp_path / "empty.ply")["x"]) == 0


def test_3dgs_init_layout(tmp_path):
    xyz = np.array([[0, 0, 1], [1, 2, 3]], np.float32)
    rgb = np.array([[255, 0, 128], [0, 255, 0]], np.uint8)
    p = tmp_path / "init.ply"
    assert export.write_3dgs_init(p, xyz, rgb, voxel=0.02) == 2
    g = export.read_ply(p)
    assert list(g)[:6] == ["x", "y", "z", "nx", "ny", "nz"] and len(g) == 17
    assert np.allclose(g["scale_0"], np.log(0.01)) and np.allclose(g["rot_0"], 1.0) and np.allclose(g["rot_1"], 0.0)
    back = np.stack([g["f_dc_0"], g["f_dc_1"], g["f_dc_2"]], 1) * export.SH_C0 + 0.5
    assert np.allclose(back, rgb / 255.0, atol=1e-6)
    assert np.allclose(1 / (1 + np.exp(-g["opacity"])), 0.1, atol=1e-6)


def test_chunk_camera_poses_follow_the_chain(tmp_path):
    rng = np.random.default_rng(3)
    from da3slam_b200 import synth
    F, overlap, n_chunks = 5, 1, 3
    Es = [synth.trajectory_w2c(rng, F) for _ in range(n_chunks)]
    rel = [synth.random_sim3(rng) for _ in range(n_chunks - 1)]
    cum = rp.accumulate_sim3(rel)                       # identity first (utils/geometry.py:100)
    poses = export.chunk_camera_poses(Es, cum, overlap)
    assert poses.shape == (n_chunks * F - (n_chunks - 1) * overlap, 4, 4)
    # direct restatement, frame by frame
    want = []
    for k, E in enumerate(Es):
        s, R, t = cum[k]
        lo = 0                                          # overlap_s = 0 (utils/da3_streaming.py:138)
        hi = F - overlap if k < n_chunks - 1 else F     # overlap_e = overlap (:139, :741, :761)
        for i in range(lo, hi):
            w2c = np.eye(4); w2c[:3] = E[i]
            c2w = np.linalg.inv(w2c)
            if k > 0:
                S = np.eye(4); S[:3, :3] = s * R; S[:3, 3] = t
                c2w = S @ c2w
                c2w[:3, :3] /= s
            want.append(c2w)
    assert np.allclose(poses, np.stack(want), atol=1e-12)
    for p in poses:                                     # rotation blocks stay orthonormal after the 1/s normalisation
        assert np.allclose(p[:3, :3] @ p[:3, :3].T, np.eye(3), atol=1e-6)
    export.write_camera_poses(tmp_path, poses, np.stack([np.eye(3)] * len(poses)))
    txt = np.loadtxt(tmp_path / "camera_poses.txt")
    assert txt.shape == (len(poses), 16) and np.allclose(txt.reshape(-1, 4, 4), poses)
    assert (tmp_path / "camera_poses.ply").read_text().splitlines()[2] == f"element vertex {len(poses)}"
