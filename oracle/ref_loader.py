"""Import the real reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE.  Used by ``tests/golden/make_golden.py`` to write the
golden fixtures and by the optional "live reference" cross-check tests.  The GPU
box has no /root/reference, so everything here is gated on ``available()``.

The reference imports ``open3d`` at module scope (align_geometry.py:4,
utils/align_geometry_single.py:5) although only the ICP / KD-tree functions use
it; an empty stub module is enough for every other function.  ``utils/align.py``
imports its sibling as a top-level module called ``geometry`` (utils/align.py:8).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("DA3S_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "align.py"))


def _load(name: str, rel: str):
    path = os.path.join(REF_ROOT, rel)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load():
    """Return a namespace with the reference modules:
    ``ag`` (align_geometry), ``ags`` (utils/align_geometry_single),
    ``geo`` (utils/geometry), ``al`` (utils/align), ``vg`` (src/vggt/utils/geometry).
    """
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    saved_geometry = sys.modules.get("geometry")
    saved_o3d = sys.modules.get("open3d")
    sys.modules["open3d"] = types.ModuleType("open3d")
    try:
        ns = types.SimpleNamespace()
        ns.ag = _load("_da3ref_align_geometry", "align_geometry.py")
        ns.ags = _load("_da3ref_align_geometry_single", "utils/align_geometry_single.py")
        ns.geo = _load("geometry", "utils/geometry.py")
        ns.al = _load("_da3ref_align", "utils/align.py")
        # vendored VGGT geometry needs the 'src' package on the path
        sys.path.insert(0, REF_ROOT)
        try:
            ns.vg = importlib.import_module("src.vggt.utils.geometry")
        finally:
            sys.path.remove(REF_ROOT)
    finally:
        if saved_geometry is not None:
            sys.modules["geometry"] = saved_geometry
        else:
            sys.modules.pop("geometry", None)
        if saved_o3d is not None:
            sys.modules["open3d"] = saved_o3d
        else:
            sys.modules.pop("open3d", None)
    _cache["ns"] = ns
    return ns
