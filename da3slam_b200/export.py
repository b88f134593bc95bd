"""Export of the aligned map: host-side file writers for what the GPU path produces —
the voxel-downsampled cloud (`VoxelGrid.read`) and the chained cameras.  Plain numpy IO, no device code.

* write_ply / read_ply        binary little-endian PLY, float32 x y z (+ uchar red green blue): the format
                              upstream's un-vendored `save_confident_pointcloud_batch` is called to produce
                              (utils/da3_streaming.py:665-690).
* chunk_camera_poses          global camera-to-world matrices of every frame from the per-chunk local w2c
                              extrinsics and the cumulative Sim(3) chain — utils/da3_streaming.py:733-774
                              (first chunk as is; later chunks S @ c2w with the rotation block divided by s;
                              overlap frames are taken from the later chunk).
* write_camera_poses          camera_poses.txt (16 numbers per line), intrinsic.txt (fx fy cx cy),
                              camera_poses.ply (ascii, one vertex per camera) — :776-817.
* write_3dgs_init             3D Gaussian Splatting initialisation PLY (the INRIA layout: x y z nx ny nz
                              f_dc_0..2 opacity scale_0..2 rot_0..3) from the voxel map: one isotropic
                              Gaussian per occupied voxel, scale = log(voxel / 2), colour as SH degree 0,
                              opacity logit(0.1).  `main_3dgs.py` is an empty stub in the reference, so the
                              format is this project's choice.
"""
from __future__ import annotations

import os

import numpy as np

SH_C0 = 0.28209479177387814


def _as_np(x):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def write_ply(path, xyz, rgb=None):
    xyz = np.ascontiguousarray(_as_np(xyz), dtype=np.float32).reshape(-1, 3)
    rgb = _as_np(rgb)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if rgb is not None:
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
        if len(rgb) != len(xyz):
            raise ValueError("rgb and xyz differ in length")
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.empty(len(xyz), dtype=fields)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if rgb is not None:
        rec["red"], rec["green"], rec["blue"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {len(xyz)}",
              "property float x", "property float y", "property float z"]
    if rgb is not None:
        header += ["property uchar red", "property uchar green", "property uchar blue"]
    header += ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())
    return len(xyz)


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "uint8": "u1", "int": "<i4", "uint": "<u4"}


def read_ply(path):
    """Reader for the binary little-endian PLY files written here.  Returns a dict of property arrays."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("not a PLY file")
        n, fields = 0, []
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format" and tok[1] != "binary_little_endian":
                raise ValueError("only binary_little_endian PLY is supported")
            if tok[0] == "element":
                if tok[1] != "vertex":
                    raise ValueError("only vertex elements are supported")
                n = int(tok[2])
            if tok[0] == "property":
                fields.append((tok[2], _PLY_TYPES[tok[1]]))
        rec = np.frombuffer(f.read(), dtype=fields, count=n)
    return {name: rec[name].copy() for name, _ in fields}


def chunk_camera_poses(chunk_extrinsics, cumulative_sim3, overlap=1):
    """chunk_extrinsics: list over chunks of [F,3,4] local world-to-camera matrices; cumulative_sim3: list of
    (s, R, t) per chunk (identity first — accumulate_sim3_transforms).  Returns [n_frames,4,4] float64 global
    camera-to-world matrices; every chunk but the last drops its last `overlap` frames, which the NEXT chunk
    re-observes (overlap_s = 0, overlap_e = overlap: utils/da3_streaming.py:138-139, :733-774)."""
    n_chunks = len(chunk_extrinsics)
    out = []
    for k, E in enumerate(chunk_extrinsics):
        E = np.asarray(E, np.float64)
        s, R, t = cumulative_sim3[k]
        S = np.eye(4)
        S[:3, :3] = float(s) * np.asarray(R, np.float64)
        S[:3, 3] = np.asarray(t, np.float64)
        first = 0                                  # overlap_s = 0, overlap_e = overlap (:138-139)
        last = E.shape[0] - overlap if k < n_chunks - 1 else E.shape[0]
        for i in range(first, last):
            w2c = np.eye(4)
            w2c[:3, :] = E[i]
            c2w = np.linalg.inv(w2c)
            if k > 0:
                c2w = S @ c2w                      # left multiplication (:769)
                c2w[:3, :3] /= float(s)            # normalise the rotation block (:770)
            out.append(c2w)
    return np.stack(out) if out else np.zeros((0, 4, 4))


def write_camera_poses(out_dir, poses_c2w, intrinsics=None, color=(255, 0, 0)):
    os.makedirs(out_dir, exist_ok=True)
    poses = np.asarray(poses_c2w, np.float64)
    with open(os.path.join(out_dir, "camera_poses.txt"), "w") as f:
        for p in poses:
            f.write(" ".join(str(x) for x in p.flatten()) + "\n")
    if intrinsics is not None:
        with open(os.path.join(out_dir, "intrinsic.txt"), "w") as f:
            for K in np.asarray(intrinsics):
                f.write(f"{K[0, 0]} {K[1, 1]} {K[0, 2]} {K[1, 2]}\n")
    with open(os.path.join(out_dir, "camera_poses.ply"), "w") as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write(f"element vertex {len(poses)}\n")
        f.write("property float x\nproperty float y\nproperty float z\n")
        f.write("property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
        for p in poses:
            f.write(f"{p[0, 3]} {p[1, 3]} {p[2, 3]} {color[0]} {color[1]} {color[2]}\n")


def write_3dgs_init(path, xyz, rgb, voxel, opacity=0.1):
    xyz = np.ascontiguousarray(_as_np(xyz), dtype=np.float32).reshape(-1, 3)
    n = len(xyz)
    rgb = _as_np(rgb)
    col = np.full((n, 3), 0.5, np.float32) if rgb is None else np.asarray(rgb, np.float32).reshape(-1, 3) / 255.0
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
             "rot_0", "rot_1", "rot_2", "rot_3"]
    rec = np.zeros(n, dtype=[(nm, "<f4") for nm in names])
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    for c in range(3):
        rec[f"f_dc_{c}"] = (col[:, c] - 0.5) / SH_C0
        rec[f"scale_{c}"] = np.float32(np.log(float(voxel) / 2.0))
    rec["opacity"] = np.float32(np.log(opacity / (1.0 - opacity)))
    rec["rot_0"] = 1.0
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + [f"property float {nm}" for nm in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(rec.tobytes())
    return n
