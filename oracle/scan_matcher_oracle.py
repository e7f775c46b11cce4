"""TEST INFRASTRUCTURE ONLY: restatement of the g2o-facing front half of the reference's
ScanMatcher (src/matcher/scan_matcher.cpp:78-110 transformPointsFromVSet / applyTransfToScan,
:112-160 closeScanMatching, :201-294 scanMatchingLC, :366-401 globalMatching) on top of a CPU
matcher from oracle/bindings.py (the verbatim reference chargrid.cpp when oracle/_ref exists, else
the restatement). scan_matcher.cpp itself needs g2o and cannot be compiled here;
SE2 follows oracle/pgo_oracle.py (appendix C1). Vertices are dicts: id, pose (x, y, th),
ranges, first_angle, step, max_range, laser_pose."""
import math

import numpy as np

from . import pgo_oracle as po

CLOSE = dict(ll=(-15.0, -15.0), ur=(15.0, 15.0), res=0.025, kernel_range=0.2)
LC = dict(ll=(-35.0, -35.0), ur=(35.0, 35.0), res=0.1, kernel_range=0.5)


def cartesian(v):
    """RawLaser::cartesian (appendix C13): beams with r < maxRange, laser frame."""
    r = np.asarray(v["ranges"], dtype=np.float64)
    i = np.nonzero(r < v["max_range"])[0]
    a = v["first_angle"] + i * v["step"]
    return np.stack([np.cos(a) * r[i], np.sin(a) * r[i]], axis=1)


def apply_transf(t, pts):
    """applyTransfToScan: (transf * SE2(p, 0)).translation()."""
    c, s = math.cos(t[2]), math.sin(t[2])
    return np.stack([t[0] + (c * pts[:, 0] - s * pts[:, 1]),
                     t[1] + (s * pts[:, 0] + c * pts[:, 1])], axis=1)


def points_from_vset(vset, ref):
    """transformPointsFromVSet with the vertex set iterated in ascending id."""
    out = []
    for v in sorted(vset, key=lambda q: q["id"]):
        if v.get("ranges") is None:
            continue
        pts = cartesian(v)
        trl = np.asarray(v["laser_pose"], dtype=np.float64)
        if v["id"] == ref["id"]:
            t = trl
        else:
            trel = po.se2_mul(po.se2_inv(ref["pose"]), v["pose"])[0]
            t = po.se2_mul(trel, trl)[0]
        out.append(apply_transf(t, pts))
    return np.concatenate(out) if out else np.zeros((0, 2))


def _grid(lib, orc, cfg, map_pts):
    g = lib.grid(cfg["ll"], cfg["ur"], cfg["res"])
    g.fill(int(cfg["kernel_range"] * 128))
    g.raster(map_pts, orc.make_stamp(cfg["res"], cfg["kernel_range"]))
    return g


def close_scan_matching(lib, orc, vset, origin, current, max_score):
    g = _grid(lib, orc, CLOSE, points_from_vset(vset, origin))
    red = lib.subsample(cartesian(current), 0.1)
    pts = apply_transf(np.asarray(current["laser_pose"], dtype=np.float64), red)
    d = po.se2_mul(po.se2_inv(origin["pose"]), current["pose"])[0]
    lo = np.array([-.3 + d[0], -.3 + d[1], -0.2 + d[2]], dtype=np.float32)
    hi = np.array([.3 + d[0], .3 + d[1], 0.2 + d[2]], dtype=np.float32)
    res = g.greedy_search_res(pts, np.concatenate([lo, hi])[None, :], 0.0125 * .5, max_score,
                              (0.5, 0.5, 0.2))
    return (True, res[0, :3]) if len(res) else (False, None)


def scan_matching_lc(lib, orc, ref_vset, ref, current, max_score):
    g = _grid(lib, orc, LC, points_from_vset(ref_vset, ref))
    red = lib.subsample(points_from_vset([current], current), 0.1)
    regions, regions_pi = [], []
    for v in sorted(ref_vset, key=lambda q: q["id"]):
        rel = np.zeros(3)
        if v["id"] != ref["id"]:
            rel = po.se2_mul(po.se2_inv(ref["pose"]), v["pose"])[0]
        lo = np.array([-.5 + rel[0], -1.5 + rel[1], -0.8 + rel[2]], dtype=np.float32)
        hi = np.array([.5 + rel[0], 1.5 + rel[1], 0.8 + rel[2]], dtype=np.float32)
        regions.append(np.concatenate([lo, hi]))
        lo2, hi2 = lo.copy(), hi.copy()
        lo2[2] = np.float32(float(lo[2]) + math.pi)   # float += double, rounded to float
        hi2[2] = np.float32(float(hi[2]) + math.pi)
        regions_pi.append(np.concatenate([lo2, hi2]))
    bins = (0.5, 0.5, 0.2)
    found = {}
    for regs in (regions, regions_pi):
        res = g.greedy_search_res(red, np.array(regs, dtype=np.float32), 0.025, max_score, bins)
        if len(res):
            best = res[0].copy()
            best[2] = po.normalize_theta(np.array([best[2]]))[0]
            key = (float(int(best[0] / bins[0])), float(int(best[1] / bins[1])),
                   float(int(best[2] / bins[2])))
            if key not in found or found[key][3] > best[3]:
                found[key] = best
    return [found[k][:3] for k in sorted(found)]


def global_matching(lib, orc, ref_vset, ref, current, max_score):
    g = _grid(lib, orc, LC, points_from_vset(ref_vset, ref))
    red = lib.subsample(points_from_vset([current], current), 0.1)
    lo = np.array([-10, -5, -math.pi], dtype=np.float32)
    hi = np.array([10, 5, math.pi], dtype=np.float32)
    res = g.hierarchical_search(red, np.concatenate([lo, hi])[None, :], 0.025, max_score,
                                (0.5, 0.5, 0.2), 4)
    return (True, res[0, :3]) if len(res) else (False, None)
