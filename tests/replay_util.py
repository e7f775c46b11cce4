"""Shared by tests/test_ref_frontend.py, bench.py --workload bag and tools/make_golden_replay.py:
keyframe files for the replay drivers (tests/cpp/ref_replay.cpp) and parsing of their result lines."""
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LASER_POSE = (0.05, 0.0, 0.0)      # base_link -> base_laser_link in the bags' /tf


def write_keyframes(path, fx, n, min_inliers=7):
    with open(path, "w") as f:
        f.write("%d %.17g %.17g %.17g %.17g %.17g %.17g %d\n" % (
            fx["ranges"].shape[1], float(fx["first_angle"]), float(fx["angular_step"]),
            float(fx["max_range"]), LASER_POSE[0], LASER_POSE[1], LASER_POSE[2], min_inliers))
        for k in range(min(n, len(fx["odom"]))):
            f.write("%.17g %.17g %.17g " % tuple(fx["odom"][k]))
            f.write(" ".join("%.9g" % r for r in fx["ranges"][k]) + "\n")


def parse(lines):
    """-> (frames: list of dict(est, edges [(from, to, xyz, w)], cands [(from, to, xyz)]), poses {id: xyz},
    times dict or None)."""
    frames, poses, times = [], {}, None
    pending = None
    pending_t = None
    for ln in lines:
        tok = ln.split()
        if not tok:
            continue
        if tok[0] == "K":
            frames.append({"id": int(tok[2]), "est": np.array([float(x) for x in tok[3:6]]), "edges": [], "cands": [],
                           "follow_diff": pending, "ms": pending_t})
            pending = None
            pending_t = None
        elif tok[0] == "E":
            frames[-1]["edges"].append((int(tok[1]), int(tok[2]), np.array([float(x) for x in tok[3:6]]), float(tok[6])))
        elif tok[0] == "C":
            frames[-1]["cands"].append((int(tok[1]), int(tok[2]), np.array([float(x) for x in tok[3:6]])))
        elif tok[0] == "P":
            poses[int(tok[1])] = np.array([float(x) for x in tok[2:5]])
        elif tok[0] == "T":
            pending_t = [float(x) for x in tok[2:5]]
        elif tok[0] == "D":
            pending = float(tok[2])      # printed before the K line of its keyframe
        elif tok[0] == "TIMES_MS":
            times = dict(times or {}, **{tok[i]: float(tok[i + 1]) for i in range(1, len(tok) - 1, 2)})
        elif tok[0] == "SOLVER_MS":  # wall clock inside the solver library by entry point: (ms, calls)
            times = dict(times or {}, solver_ms={tok[i]: (float(tok[i + 1]), int(float(tok[i + 2]))) for i in range(1, len(tok) - 2, 3)})
        elif tok[0] == "TRACE":      # per kind of solver call: how many, worst difference follower vs leader
            times = dict(times or {}, trace={tok[i]: (int(tok[i + 1]), float(tok[i + 2])) for i in range(1, len(tok) - 2, 3)})
    return frames, poses, times


def load_golden(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", "replay_%s_cpu.txt.gz" % name), "rt") as f:
        lines = f.read().splitlines()
    return lines[lines.index("BEGIN") + 1:lines.index("END")] if "BEGIN" in lines else lines


def angle_diff(a, b):
    return (a - b + np.pi) % (2 * np.pi) - np.pi


def compare(got, want, n, tol, allow_tie=False):
    """Frames [0, n) of two parsed replays: identical vertex indices everywhere, estimates and
    measurements within tol. Returns (kinds dict, worst difference).

    allow_tie: the one kind of decision that two solvers agreeing to 1e-13 cannot be asked to share.
    When the robot turns in place, consecutive keyframes coincide; VerticesFinder::findClosestVertex
    (vertices_finder.cpp:95-112) then picks among vertices whose distances differ at rounding level,
    and graph_slam.cpp:417 drops the match if the pick happens to be the previous keyframe. If the two
    runs differ by exactly such an edge (its endpoints less than 5 cm apart), the comparison stops
    there (stats["tie_at"] = keyframe): from then on the graphs differ."""
    worst = 0.0
    stats = {"edges": 0, "closures": 0, "cands": 0, "tie_at": None}
    where = {}

    def d3(a, b):
        d = a - b
        d[2] = angle_diff(a[2], b[2])
        return float(np.abs(d).max())

    for k in range(n):
        g, w = got[k], want[k]
        assert g["id"] == w["id"], (k, g["id"], w["id"])
        where[g["id"]] = g["est"][:2]
        ge, we = [(e[0], e[1]) for e in g["edges"]], [(e[0], e[1]) for e in w["edges"]]
        if allow_tie and ge != we:
            odd = set(ge) ^ set(we)
            assert odd and all(a in where and b in where and np.hypot(*(where[a] - where[b])) < 0.05 for a, b in odd), \
                (k, g["edges"], w["edges"])
            stats["tie_at"] = k
            return stats, worst
        assert [(e[0], e[1]) for e in g["edges"]] == [(e[0], e[1]) for e in w["edges"]], (k, g["edges"], w["edges"])
        assert [(e[0], e[1]) for e in g["cands"]] == [(e[0], e[1]) for e in w["cands"]], (k, g["cands"], w["cands"])
        worst = max(worst, d3(g["est"], w["est"]))
        for eg, ew in zip(g["edges"], w["edges"]):
            worst = max(worst, d3(eg[2], ew[2]))
            assert eg[3] == ew[3], (k, eg, ew)       # odometry vs scan-matched information
            stats["edges"] += 1
            stats["closures"] += abs(eg[0] - eg[1]) > 1
        for cg, cw in zip(g["cands"], w["cands"]):
            worst = max(worst, d3(cg[2], cw[2]))
            stats["cands"] += 1
        assert worst < tol, (k, worst)
    return stats, worst
