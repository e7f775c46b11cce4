"""Seeded synthetic inputs for the two hot paths (there are no datasets in the container).

* ``make_scan_pair``  -- SURVEY.md section 8(d) "Scan-match synthetic input (cfg 5)": a random
  rectilinear room, a 1081-beam 270 degree reference scan taken at pose A and a current scan taken
  at pose B = A (+) delta, both with N(0, 0.01^2) range noise and an 8 m cut-off.
* ``make_pose_graph`` -- section 8(d) "GN synthetic input (cfg 4)": a Manhattan-style lattice walk
  with odometry edges and proximity loop closures, information matrices as the reference sets them
  (src/slam/graph_slam.cpp:72-76).

Pure numpy; used by bench.py, tests/ and tools/. Not part of the timed path.
"""
import math

import numpy as np

TWO_PI = 2.0 * math.pi


# ----------------------------------------------------------------------------------------------
# scans
# ----------------------------------------------------------------------------------------------
def _room_segments(rng):
    """Axis-aligned wall segments of a w x h room with 0-3 interior boxes.
    Returns (vertical [n,3] = x, y0, y1 ; horizontal [m,3] = y, x0, x1 ; (w, h) ; boxes)."""
    w = rng.uniform(6.0, 14.0)
    h = rng.uniform(4.0, 10.0)
    vert = [(0.0, 0.0, h), (w, 0.0, h)]
    horiz = [(0.0, 0.0, w), (h, 0.0, w)]
    boxes = []
    for _ in range(int(rng.integers(0, 4))):
        bw, bh = rng.uniform(0.4, 1.5), rng.uniform(0.4, 1.5)
        bx, by = rng.uniform(0.3, w - bw - 0.3), rng.uniform(0.3, h - bh - 0.3)
        boxes.append((bx, by, bx + bw, by + bh))
        vert += [(bx, by, by + bh), (bx + bw, by, by + bh)]
        horiz += [(by, bx, bx + bw), (by + bh, bx, bx + bw)]
    return np.array(vert), np.array(horiz), (w, h), boxes


def _free_pose(rng, size, boxes, margin=0.4):
    w, h = size
    for _ in range(1000):
        x, y = rng.uniform(margin, w - margin), rng.uniform(margin, h - margin)
        if all(not (bx0 - margin < x < bx1 + margin and by0 - margin < y < by1 + margin)
               for bx0, by0, bx1, by1 in boxes):
            return x, y
    return margin, margin


def cast_scan(vert, horiz, pose, n_beams, first_angle, step, max_range, sigma, rng):
    """Ranges of an ideal 2-D laser at ``pose`` = (x, y, theta) against axis-aligned segments.
    Beams that hit nothing or exceed ``max_range`` read exactly ``max_range`` (as the bags do)."""
    x, y, th = pose
    ang = th + first_angle + step * np.arange(n_beams)
    c, s = np.cos(ang), np.sin(ang)
    best = np.full(n_beams, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        # vertical walls: x = X, y in [y0, y1]
        t = (vert[:, 0][None, :] - x) / c[:, None]
        yy = y + t * s[:, None]
        ok = (t > 1e-9) & (yy >= vert[:, 1][None, :]) & (yy <= vert[:, 2][None, :])
        best = np.minimum(best, np.where(ok, t, np.inf).min(axis=1))
        t = (horiz[:, 0][None, :] - y) / s[:, None]
        xx = x + t * c[:, None]
        ok = (t > 1e-9) & (xx >= horiz[:, 1][None, :]) & (xx <= horiz[:, 2][None, :])
        best = np.minimum(best, np.where(ok, t, np.inf).min(axis=1))
    r = best + rng.normal(0.0, sigma, n_beams)
    r = np.where(np.isfinite(r) & (r < max_range), r, max_range)
    return r


def scan_to_points(ranges, first_angle, step, max_range):
    """RawLaser::cartesian (g2o, recalled): keep beams with r < maxRange."""
    i = np.nonzero(ranges < max_range)[0]
    a = first_angle + step * i
    return np.stack([ranges[i] * np.cos(a), ranges[i] * np.sin(a)], axis=1)


def make_scan_pair(seed, n_beams=1081, fov=1.5 * math.pi, max_range=8.0, sigma=0.01,
                   max_delta=(4.0, 4.0, 1.0)):
    """One (reference scan, current scan) pair. Returns a dict with
    ``map_pts`` (reference scan, frame A), ``cur_pts`` (current scan, frame B) and ``delta``
    (pose of B in A: the transformation the matcher should recover)."""
    rng = np.random.default_rng(1234 + seed)
    vert, horiz, size, boxes = _room_segments(rng)
    first = -fov / 2.0
    step = fov / (n_beams - 1)
    ax, ay = _free_pose(rng, size, boxes)
    ath = rng.uniform(-math.pi, math.pi)
    for _ in range(1000):
        d = rng.uniform(-1.0, 1.0, 3) * np.array(max_delta)
        bx = ax + math.cos(ath) * d[0] - math.sin(ath) * d[1]
        by = ay + math.sin(ath) * d[0] + math.cos(ath) * d[1]
        inside = 0.4 < bx < size[0] - 0.4 and 0.4 < by < size[1] - 0.4
        if inside and all(not (b[0] - 0.4 < bx < b[2] + 0.4 and b[1] - 0.4 < by < b[3] + 0.4)
                          for b in boxes):
            break
    else:
        d = np.zeros(3)
        bx, by = ax, ay
    bth = ath + d[2]
    ra = cast_scan(vert, horiz, (ax, ay, ath), n_beams, first, step, max_range, sigma, rng)
    rb = cast_scan(vert, horiz, (bx, by, bth), n_beams, first, step, max_range, sigma, rng)
    return {
        "map_pts": scan_to_points(ra, first, step, max_range),
        "cur_pts": scan_to_points(rb, first, step, max_range),
        "delta": d,
        "ranges_a": ra,
        "ranges_b": rb,
        "first_angle": first,
        "angular_step": step,
        "max_range": max_range,
    }


# ----------------------------------------------------------------------------------------------
# SE(2) pose graphs
# ----------------------------------------------------------------------------------------------
def wrap(a):
    """Map angles to [-pi, pi)."""
    return (np.asarray(a) + math.pi) % TWO_PI - math.pi


def se2_mul(a, b):
    """Rows (x, y, th): a * b."""
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    c, s = np.cos(a[:, 2]), np.sin(a[:, 2])
    out = np.empty((max(len(a), len(b)), 3))
    out[:, 0] = a[:, 0] + c * b[:, 0] - s * b[:, 1]
    out[:, 1] = a[:, 1] + s * b[:, 0] + c * b[:, 1]
    out[:, 2] = wrap(a[:, 2] + b[:, 2])
    return out


def se2_inv(a):
    a = np.atleast_2d(a)
    th = -a[:, 2]
    c, s = np.cos(th), np.sin(th)
    out = np.empty_like(a)
    out[:, 0] = -(c * a[:, 0] - s * a[:, 1])
    out[:, 1] = -(s * a[:, 0] + c * a[:, 1])
    out[:, 2] = wrap(th)
    return out


INFO_ODOM = np.diag([100.0, 100.0, 1000.0])        # graph_slam.cpp:72-73
INFO_CLOSURE = np.diag([1000.0, 1000.0, 10000.0])  # graph_slam.cpp:75-76


def make_pose_graph(n_vertices=50000, n_edges=200000, seed=42, box=250.0, radius=2.5,
                    min_gap=10, max_per_pose=6, noise_scale=1.0, init="odometry", start_noise=1.0):
    """Manhattan-style walk. Returns dict(poses0 [V,3] initial guess, truth [V,3], edge_ij [E,2],
    meas [E,3], info [E,6] upper triangle row-major, fixed = [0]).

    ``noise_scale`` scales the measurement noise standard deviation (1.0 = Sigma = Omega^-1 as in
    SURVEY 8(d)). ``init`` = "odometry" (dead reckoning along the chain) or "truth_noisy" (truth +
    N(0, (0.05 m, 0.05 m, 0.01 rad) x ``start_noise``): a start Gauss-Newton converges from)."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    truth = np.zeros((n_vertices, 3))
    x = y = box / 2.0
    heading = 0
    dirs = [(1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0)]
    turns = rng.choice([0, 1, -1], size=n_vertices, p=[0.8, 0.1, 0.1])
    truth[0] = (x, y, 0.0)
    for k in range(1, n_vertices):
        heading = (heading + int(turns[k])) % 4
        dx, dy = dirs[heading]
        if not (0.0 <= x + dx <= box and 0.0 <= y + dy <= box):
            heading = (heading + 2) % 4  # bounce off the wall
            dx, dy = dirs[heading]
        x += dx
        y += dy
        truth[k] = (x, y, heading * math.pi / 2.0)
    truth[:, 2] = wrap(truth[:, 2])

    ei = [np.arange(n_vertices - 1)]
    ej = [np.arange(1, n_vertices)]
    n_closure = max(0, n_edges - (n_vertices - 1))
    if n_closure:
        tree = cKDTree(truth[:, :2])
        pairs = tree.query_pairs(radius, output_type="ndarray")
        pairs = pairs[np.abs(pairs[:, 0] - pairs[:, 1]) > min_gap]
        lo = np.minimum(pairs[:, 0], pairs[:, 1])
        hi = np.maximum(pairs[:, 0], pairs[:, 1])
        order = np.lexsort((lo, hi))  # group by the later pose, earlier partners ascending
        lo, hi = lo[order], hi[order]
        # at most max_per_pose closures per (later) pose
        start = np.r_[0, np.nonzero(np.diff(hi))[0] + 1]
        rank = np.arange(len(hi)) - np.repeat(start, np.diff(np.r_[start, len(hi)]))
        keep = rank < max_per_pose
        lo, hi = lo[keep], hi[keep]
        if len(lo) > n_closure:
            sel = np.sort(rng.choice(len(lo), size=n_closure, replace=False))
            lo, hi = lo[sel], hi[sel]
        ei.append(lo)
        ej.append(hi)
    ei = np.concatenate(ei).astype(np.int32)
    ej = np.concatenate(ej).astype(np.int32)
    n_odo = n_vertices - 1
    is_odo = np.arange(len(ei)) < n_odo

    rel = se2_mul(se2_inv(truth[ei]), truth[ej])
    sig = np.where(is_odo[:, None],
                   1.0 / np.sqrt(np.diag(INFO_ODOM))[None, :],
                   1.0 / np.sqrt(np.diag(INFO_CLOSURE))[None, :])
    noise = rng.normal(size=rel.shape) * sig * noise_scale
    meas = se2_mul(rel, noise)
    info = np.zeros((len(ei), 6))
    for m, sel in ((INFO_ODOM, is_odo), (INFO_CLOSURE, ~is_odo)):
        info[sel] = (m[0, 0], m[0, 1], m[0, 2], m[1, 1], m[1, 2], m[2, 2])

    if init == "odometry":
        poses0 = np.zeros_like(truth)
        poses0[0] = truth[0]
        cur = truth[0].copy()
        for k in range(n_odo):
            cur = se2_mul(cur, meas[k])[0]
            poses0[k + 1] = cur
    elif init == "truth_noisy":
        poses0 = truth + rng.normal(size=truth.shape) * np.array([0.05, 0.05, 0.01]) * start_noise
        poses0[:, 2] = wrap(poses0[:, 2])
        poses0[0] = truth[0]
    else:
        raise ValueError(init)
    return {
        "poses0": poses0,
        "truth": truth,
        "edge_ij": np.stack([ei, ej], axis=1),
        "meas": meas,
        "info": info,
        "fixed": np.array([0], dtype=np.int32),
        "ids": np.arange(n_vertices, dtype=np.int32),
    }
