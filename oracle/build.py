"""Build the oracle's C restatement (oracle/c/oracle.c) into oracle/_build/liboracle.so.
TEST INFRASTRUCTURE: building the checker is not using it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle C build failed:\n" + r.stdout)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        P, I, L, F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
        lib.oracle_ransac_score.argtypes = [P, P, P, I, P, P, P, L, F, P]
        lib.oracle_ransac_score.restype = None
        lib.oracle_ransac_inlier_mask.argtypes = [P, P, P, P, P, L, F, P]
        lib.oracle_ransac_inlier_mask.restype = None
        lib.oracle_cam_fast_f32.argtypes = [P, I, I, I, P, P]
        lib.oracle_cam_fast_f32.restype = None
        lib.oracle_voxel_key.argtypes = [P, F, P, P]
        lib.oracle_voxel_key.restype = I
        _lib = lib
    return _lib


if __name__ == "__main__":
    print(build(force=True))
