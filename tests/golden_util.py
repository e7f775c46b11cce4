"""Helpers to replay tests/golden/matcher_*.npz through any matcher implementation."""
import glob
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def matcher_fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN, "matcher_*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["name"] = os.path.basename(path)[len("matcher_"):-len(".npz")]
    return d


def cfg_of(fx):
    return dict(ll=tuple(float(v) for v in fx["ll"]), ur=tuple(float(v) for v in fx["ur"]),
                res=float(fx["res"]), kernel_range=float(fx["kernel_range"]))


def sha(grid):
    return hashlib.sha256(np.ascontiguousarray(grid).tobytes()).hexdigest()


def replay_product(fx, make_matcher, subsample_fn):
    """Run one fixture through the product API (a cg_mrslam_b200.matcher.Matcher factory).
    Returns (grid sha256, subsampled points, results)."""
    cfg = cfg_of(fx)
    m = make_matcher(cfg)
    try:
        assert np.array_equal(m.stamp(), fx["stamp"])
        m.raster_batch([fx["map_pts"]])
        grid = m.download()
        pts = fx["cur_pts"]
        sub = subsample_fn(pts, 0.1) if bool(fx["subsample"]) else pts
        bins = tuple(float(v) for v in fx["bins"])
        if int(fx["levels"]):
            res = m.hierarchical_search(sub, fx["regions"], float(fx["theta_res"]),
                                        float(fx["max_score"]), bins, int(fx["levels"]))
        else:
            res = m.greedy_search_res(sub, fx["regions"], float(fx["theta_res"]),
                                      float(fx["max_score"]), bins)
    finally:
        m.close()
    return sha(grid), sub, res


def replay_cpu(fx, lib, stamp):
    """Run one fixture through a CPU matcher (oracle.bindings.MatcherLib)."""
    cfg = cfg_of(fx)
    g = lib.grid(cfg["ll"], cfg["ur"], cfg["res"])
    g.fill(int(cfg["kernel_range"] * 128))
    g.raster(fx["map_pts"], stamp)
    grid = g.download()
    pts = fx["cur_pts"]
    sub = lib.subsample(pts, 0.1) if bool(fx["subsample"]) else pts
    bins = tuple(float(v) for v in fx["bins"])
    if int(fx["levels"]):
        res = g.hierarchical_search(sub, fx["regions"], float(fx["theta_res"]),
                                    float(fx["max_score"]), bins, int(fx["levels"]))
    else:
        res = g.greedy_search_res(sub, fx["regions"], float(fx["theta_res"]),
                                  float(fx["max_score"]), bins)
    return sha(grid), sub, res
