"""CPU oracle for the numgrids hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU arm.  The product package ``numgrids_b200`` never imports it and
fails loudly when its CUDA library is missing.

Contents
--------
findiff_port.py   restatement of the third-party ``findiff.FinDiff`` (uniform
                  grid) that the reference calls at numgrids/diff.py:10,88,259.
                  findiff is pinned only as ``findiff>=0.10`` in the reference's
                  pyproject.toml:23 and is NOT available offline, so FD parity
                  is pinned against this restatement (which passes the
                  reference's own tests), not against the real package:
                  **FD parity vs. real findiff is unpinned.**
pencil.py         bit-faithful NumPy restatements of the four per-axis applies
                  and of the curvilinear composites that never build a
                  Kronecker matrix or a meshgrid (usable at config sizes).
reference_loader.py  imports the real reference from /root/reference (build
                  container only) with ``findiff`` -> findiff_port injected.
pencil_c.c        C restatement of pencil.py (OpenMP), for full-size checks
                  and the multi-threaded CPU baseline.
"""
