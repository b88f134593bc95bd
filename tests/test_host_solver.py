"""CPU tests of the host-side orchestration shims (no GPU): config loading, the solver's
deque / chunk formation."""
import sys
import types

import numpy as np


def test_load_config_inheritance(tmp_path):
    import config
    (tmp_path / "base.yaml").write_text("Model:\n  chunk_size: 15\n  port: 8080\n  IRLS:\n    delta: 0.1\nWeights:\n  DA3: a\n")
    (tmp_path / "child.yaml").write_text(f"inherit_from: {tmp_path / 'base.yaml'}\nModel:\n  chunk_size: 16\n  IRLS:\n    tol: 1.0e-9\n")
    cfg = config.load_config(str(tmp_path / "child.yaml"))
    assert cfg["Model"]["chunk_size"] == 16 and cfg["Model"]["port"] == 8080
    assert cfg["Model"]["IRLS"] == {"delta": 0.1, "tol": 1e-9} and cfg["Weights"]["DA3"] == "a"
    cfg2 = config.load_config(str(tmp_path / "base.yaml"), default_path=str(tmp_path / "child.yaml"))
    assert cfg2["Model"]["chunk_size"] == 15
    d = {"a": {"b": 1}}
    config.update_recursive(d, {"a": {"c": 2}, "e": 3})
    assert d == {"a": {"b": 1, "c": 2}, "e": 3}


def test_solver_chunk_formation(monkeypatch, tmp_path):
    """solver.py deque semantics: 300 frames, chunk 16, overlap 1 -> 19 chunks starting every 15 frames."""
    fake = types.ModuleType("depth_anything_3"); api = types.ModuleType("depth_anything_3.api")

    class Model:
        @classmethod
        def from_pretrained(cls, p): return cls()
        def to(self, d): return self
        def eval(self): return self
    api.DepthAnything3 = Model; fake.api = api
    monkeypatch.setitem(sys.modules, "depth_anything_3", fake)
    monkeypatch.setitem(sys.modules, "depth_anything_3.api", api)
    import solver
    from oracle import ref_port as rp
    monkeypatch.setattr(solver.time, "sleep", lambda s: None)
    monkeypatch.setattr(solver.SLAMSolver, "init_viewer", lambda self: setattr(self, "viewer", None))
    seen = []
    monkeypatch.setattr(solver.SLAMSolver, "run_single_chunk_prediction",
                        lambda self, paths: seen.append(paths) or {"image_paths": paths, "extrinsics": np.zeros((len(paths), 3, 4))})
    monkeypatch.setattr(solver.SLAMSolver, "process_chunk_alignment", lambda self, a, b: (1.0, np.eye(3), np.zeros(3)))
    for i in range(300):
        (tmp_path / f"f{i}.jpg").write_bytes(b"x")
    cfg = {"Model": {"chunk_size": 16, "overlap_size": 1, "keyframe_interval": 1, "sleep_between_chunk": 0, "port": 1}, "Weights": {"DA3": "x"}}
    s = solver.SLAMSolver(str(tmp_path), cfg)
    s.run()
    starts = [int(p[0].split("f")[-1].split(".")[0]) for p in seen]
    assert starts == rp.solver_chunk_starts(300, 16, 1) and len(starts) == 19 and all(len(p) == 16 for p in seen)
    assert seen[1][0] == seen[0][-1]                       # consecutive chunks share exactly the overlap frame
