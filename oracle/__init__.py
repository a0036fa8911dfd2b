"""CPU oracle for the Frenetix-Occlusion per-planning-step assessment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``frenetix_occlusion_b200/`` (the product) imports
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may.  The product path has no CPU fallback.

Contents
--------
* ``geometry``       float64 convex-polygon ``distance`` / ``intersects`` (stands in for GEOS).
* ``metric_oracle``  "oracle B": vectorised float64 numpy restatement of the dense metric core
                     (reference ``frenetix_occlusion/metrics/*``), SURVEY.md §8(a) rows M0-M9.
* ``ref_shims`` + ``ref_runner``  "oracle A": imports the reference's own ``metrics`` package
                     *unmodified* from ``/root/reference`` over stand-ins for its un-installable
                     third-party leaves.  Only usable in the build container (the reference
                     tree does not travel); it generates ``tests/golden/*.json``.
* ``prediction_oracle``, ``visibility_oracle``  restatements of stage 2 / stage 1.

Parity pinning status: the reference ships NO tests or golden vectors, and its
third-party arithmetic (GEOS, scipy mvnun, commonroad-io, frenetix) is not installed.  The dense
core is pinned by running the reference's metric modules verbatim over shims (oracle A ->
``tests/golden``); the visibility and vehicle-rollout stages are "parity unpinned" (restated from
the reference source).
"""
