"""Generate tests/golden/matcher_*.npz from the REFERENCE's own chargrid.cpp compiled verbatim
(oracle/_ref/libref_chargrid.so, recipe in oracle/Makefile). Runs only in the build container
(needs /root/reference at build time); the fixtures it writes are committed and travel.

    python tools/make_golden_matcher.py

Each fixture holds the inputs (map points, scan points, regions, parameters) and the reference's
outputs: sha256 of the rasterised grid, subsampled points, the full ordered result list.
"""
import hashlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import bindings  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_case(ref, orc, name, cfg, pair, regions, theta_res, max_score, subsample=True,
               levels=0):
    stamp = orc.make_stamp(cfg["res"], cfg["kernel_range"])  # scan_matcher.cpp needs g2o: restated
    g = ref.grid(cfg["ll"], cfg["ur"], cfg["res"])
    g.fill(int(cfg["kernel_range"] * 128))
    g.raster(pair["map_pts"], stamp)
    grid = g.download()
    pts = pair["cur_pts"]
    sub = ref.subsample(pts, 0.1) if subsample else pts
    if levels:
        res = g.hierarchical_search(sub, regions, theta_res, max_score, cases.BINS, levels)
    else:
        res = g.greedy_search_res(sub, regions, theta_res, max_score, cases.BINS)
    np.savez_compressed(
        os.path.join(OUT, "matcher_%s.npz" % name),
        ll=np.array(cfg["ll"]), ur=np.array(cfg["ur"]), res=cfg["res"],
        kernel_range=cfg["kernel_range"], map_pts=pair["map_pts"], cur_pts=pts,
        subsample=subsample, sub_pts=sub, regions=regions, theta_res=theta_res,
        max_score=max_score, bins=np.array(cases.BINS), levels=levels, stamp=stamp,
        grid_shape=np.array(grid.shape), grid_sha256=hashlib.sha256(grid.tobytes()).hexdigest(),
        grid_sum=int(grid.astype(np.int64).sum()), results=res)
    print("%-22s grid %s  pts %d -> %d  regions %d  results %d  best %s" %
          (name, grid.shape, len(pts), len(sub), len(regions), len(res),
           res[0] if len(res) else None))


def main():
    os.makedirs(OUT, exist_ok=True)
    bindings.build()
    ref = bindings.MatcherLib("reference")
    orc = bindings.MatcherLib("oracle")
    p361 = cases.scan_pair(1)
    reg, th = cases.close_window()
    build_case(ref, orc, "close_361", cases.CLOSE, p361, reg, th, 0.15)
    cent = [(0, 0, 0), (0.3, -0.2, 0.1), (1.0, 0.5, -0.3), (-0.4, 0.1, 0.05), (0.2, 0.2, 0.0),
            (2, 1, 1), (0.1, 0, 0)]
    reg, th = cases.lc_regions(cent)
    build_case(ref, orc, "lc_7regions", cases.LC, p361, reg, th, 0.3)
    reg, th = cases.lc_regions(cent, flip=True)
    build_case(ref, orc, "lc_7regions_flip", cases.LC, p361, reg, th, 0.45)
    p1081 = cases.scan_pair(2, 1081, 1.5 * math.pi, (2.0, 2.0, 0.5))
    reg, th = cases.global_window()
    build_case(ref, orc, "global_4levels", cases.LC, p1081, reg, th, 0.25, levels=4)
    build_case(ref, orc, "global_4levels_ties", cases.LC, p1081, reg, th, 0.3, levels=4)
    reg, th = cases.bench_window()
    build_case(ref, orc, "bench_1081_raw", cases.LC, p1081, reg, th, 0.15, subsample=False)
    p3 = cases.scan_pair(3, 1081, 1.5 * math.pi, (0.2, 0.2, 0.1))
    reg, th = cases.close_window()
    build_case(ref, orc, "close_1081", cases.CLOSE, p3, reg, th, 0.15)


if __name__ == "__main__":
    main()
