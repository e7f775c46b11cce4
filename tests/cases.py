"""Seeded matcher test cases shared by the CPU and GPU suites.

Grid configurations are the reference's own (src/slam/graph_slam.cpp:60-62):
  close: +-15 m, resolution 0.025, kernel range 0.2  -> 1200 x 1200 cells, 17 x 17 stamp
  lc:    +-35 m, resolution 0.1,   kernel range 0.5  ->  700 x  700 cells, 11 x 11 stamp
Windows are the ones ScanMatcher issues (scan_matcher.cpp:148-151, 230-246, 384-391).
"""
import math

import numpy as np

from cg_mrslam_b200 import synth

CLOSE = dict(ll=(-15.0, -15.0), ur=(15.0, 15.0), res=0.025, kernel_range=0.2)
LC = dict(ll=(-35.0, -35.0), ur=(35.0, 35.0), res=0.1, kernel_range=0.5)
BINS = (0.5, 0.5, 0.2)


def se2_apply(pose, pts):
    c, s = math.cos(pose[2]), math.sin(pose[2])
    return np.stack([pose[0] + c * pts[:, 0] - s * pts[:, 1],
                     pose[1] + s * pts[:, 0] + c * pts[:, 1]], axis=1)


def scan_pair(seed, n_beams=361, fov=math.pi, max_delta=(0.25, 0.25, 0.15)):
    return synth.make_scan_pair(seed, n_beams=n_beams, fov=fov, max_delta=max_delta)


def close_window(guess=(0.0, 0.0, 0.0)):
    """closeScanMatching: +-0.3 m, +-0.2 rad around the guess, float bounds."""
    lo = np.array([guess[0] - 0.3, guess[1] - 0.3, guess[2] - 0.2], dtype=np.float32)
    hi = np.array([guess[0] + 0.3, guess[1] + 0.3, guess[2] + 0.2], dtype=np.float32)
    return np.concatenate([lo, hi])[None, :], 0.00625


def lc_regions(centres, flip=False):
    """scanMatchingLC: one region of half-widths (0.5, 1.5, 0.8) per reference vertex."""
    out = []
    for c in centres:
        th = c[2] + (math.pi if flip else 0.0)
        lo = np.array([c[0] - 0.5, c[1] - 1.5, th - 0.8], dtype=np.float32)
        hi = np.array([c[0] + 0.5, c[1] + 1.5, th + 0.8], dtype=np.float32)
        out.append(np.concatenate([lo, hi]))
    return np.array(out, dtype=np.float32), 0.025


def bench_window():
    """BASELINE cfg 5: x, y in [-5, 5), theta in [-1.25, 1.25), step 0.1 / 0.025."""
    lo = np.array([-5.0, -5.0, -1.25], dtype=np.float32)
    hi = np.array([5.0, 5.0, 1.25], dtype=np.float32)
    return np.concatenate([lo, hi])[None, :], 0.025


def global_window():
    """globalMatching: +-10 m x +-5 m x +-pi, 4 hierarchical levels."""
    lo = np.array([-10.0, -5.0, -math.pi], dtype=np.float32)
    hi = np.array([10.0, 5.0, math.pi], dtype=np.float32)
    return np.concatenate([lo, hi])[None, :], 0.025


def same(a, b):
    """Bit-for-bit equality of two result lists (rows x, y, theta, score)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def same_multiset(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    if a.shape != b.shape:
        return False
    ka = np.sort(a.view(np.uint64).copy().view([("f", np.uint64, 4)]).ravel(), order="f")
    kb = np.sort(b.view(np.uint64).copy().view([("f", np.uint64, 4)]).ravel(), order="f")
    return np.array_equal(ka, kb)
