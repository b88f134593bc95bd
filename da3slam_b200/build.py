"""Build libda3s.so (sm_100a) in-tree with nvcc.  No GPU needed: nvcc cross-compiles.

    python -m da3slam_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libda3s.so")
SOURCES = ["api.cu", "unproject.cu", "select.cu", "pair_align.cu", "voxel.cu", "icp.cu"]
HEADERS = ["common.cuh", "sim3_math.cuh", "unproject_frame.cuh", os.path.join("..", "..", "include", "da3s.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libda3s.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """defines / out: build a VARIANT of the library (extra -D flags) next to the product, e.g. for A/B timing on the
    GPU box; select it at run time with the environment variable DA3S_LIB (da3slam_b200/_lib.py)."""
    if out is None and not defines and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(HERE, "_build" if out is None else "_build_" + os.path.basename(out).replace(".so", ""))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for s in SOURCES:
        obj = os.path.join(build_dir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, obj, p in procs:
        text, _ = p.communicate()
        log.append(f"==== {s} ====\n{text}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(obj)
    with open(os.path.join(build_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    target = LIB if out is None else out
    cmd = [nvcc, "-shared", "-o", target, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    if verbose:
        print("\n".join(log))
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None)
    print(path)
