"""Drop-in for the reference's ``config.py``: YAML config with recursive ``inherit_from``
(config.py:4-33) and a recursive dict merge (config.py:36-50).  Host-only glue, out of the
hot path; only ``Model.chunk_size / overlap_size /
keyframe_interval / sleep_between_chunk / port`` and ``Weights.DA3`` are ever read by the
solver.  New knobs of this implementation live under ``Model.Align`` (empty in the
reference's configs/config1.yaml:13)."""
from __future__ import annotations

import yaml


def update_recursive(dict1, dict2):
    """Merge dict2 into dict1 in place: nested dicts are merged key by key, everything else
    is overwritten (config.py:36-50)."""
    for key, value in dict2.items():
        if isinstance(value, dict):
            if not isinstance(dict1.get(key), dict):
                dict1[key] = dict()
            update_recursive(dict1[key], value)
        else:
            dict1[key] = value


def load_config(path, default_path=None):
    """Load `path`; if it names ``inherit_from`` load that first (recursively), else start
    from `default_path` when given; the file's own entries win (config.py:4-33)."""
    with open(path) as f:
        own = yaml.full_load(f) or {}
    parent = own.get("inherit_from")
    if parent is not None:
        cfg = load_config(parent, default_path)
    elif default_path is not None:
        with open(default_path) as f:
            cfg = yaml.full_load(f) or {}
    else:
        cfg = dict()
    update_recursive(cfg, own)
    return cfg
