"""CPU oracle for the DA3-SLAM submap-alignment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product (``da3slam_b200`` and the root-level shim
modules) never imports from here and fails loudly if ``libda3s.so`` is missing.

Contents
--------
``ref_port``   numpy/torch-CPU restatement of every reference function on the
               path.  PINNED: checked
               bit-for-bit / to 1e-12 against the real reference imported from
               /root/reference (``tests/golden/make_golden.py`` wrote the
               fixtures in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
               replays them without the reference present).
``spec_port``  CPU statement of the pieces the reference does NOT contain
               (fused f32 unprojection, dense joint-mask IRLS, RANSAC scoring,
               voxel grid, exact-selection thresholds).  PARITY UNPINNED: the
               reference holds nothing for these; the spec is ``oracle/SPEC.md``.
``c/``         plain-C restatement of the two loops that numpy cannot state
               exactly or quickly (fp32 FMA residual scoring, fixed-point voxel
               accumulation); built by ``oracle/build.py`` into ``oracle/_build``.
``ref_loader`` imports the real reference from /root/reference (this
               container only) for fixture generation and cross-checks.
"""
