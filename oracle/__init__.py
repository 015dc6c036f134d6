"""CPU oracle for the LiDAR hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package, and only as the checker (or as the timed CPU baseline).  The
product package `lidar_ai_recommendation_software_b200` never imports it and has no CPU fallback.

What is here
  np_semantics.py   the numpy / scikit-learn arithmetic the reference delegates to, restated from
                    their published algorithms (numpy 2.3.5, scikit-learn 1.9.0 — un-pinned upstream,
                    these are the versions of this image): arange fill rule, histogramdd binning,
                    linear percentile, DBSCAN labelling in closed form.
  ref_path.py       restatement of the reference's own hot-path functions (REF rows of SURVEY.md
                    §8a), each citing the reference file:line it follows.
  new_ops.py        frozen definitions of the ops the north star names but the reference lacks
                    (NEW rows: voxel downsample, ROI crop, frame flow, FPS, ball query, grouping,
                    shared MLP) — SURVEY.md Appendix B.

Pinning
  REF rows: pinned against outputs of the UNMODIFIED reference imported from /root/reference in
  the build container; the vectors and the script that made them are in tests/golden/.
  (The reference ships no tests, golden vectors or fixtures of its own.)
  NEW rows: the reference has nothing to pin against — "parity unpinned"; the definitions are ours,
  frozen in SURVEY.md Appendix B, and PointNet++-CUDA semantics where applicable.
"""
