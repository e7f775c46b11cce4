"""Evidence run for BASELINE cfg 3 at full length (every keyframe of the 4-robot bag, the reference's
default inter-robot quorum): compares a GPU leader's output (single process, or one robot per rank /
GPU) with the CPU follower's (tests/mr_replay.py --follow), with the checks of
tests/test_mr_replay.py::test_four_robots_gpu_lockstep.

    python tools/mr_full_compare.py <leader.npz | leader.rank%d.npz stem> <follower.npz>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_mr_replay import _same_graph  # noqa: E402

lead, foll = sys.argv[1], sys.argv[2]
b = np.load(foll)
if os.path.exists(lead):
    leaders = {r: np.load(lead) for r in range(4)}
    how = "one process, one GPU"
else:
    leaders = {r: np.load(lead.replace(".npz", ".rank%d.npz" % r)) for r in range(4)}
    how = "one robot per rank / GPU, datagrams by NCCL all-gather"
print("leader: reference MRGraphSLAM over libcgmrslam_b200.so (%s); follower: the same sources over the"
      " reference's chargrid.cpp + the CPU oracle solver" % how)
worst_all = 0.0
TOL = 1e-6


def tie_at(a, b, r):
    """First keyframe at which the two runs' edge counts differ, after checking that the cause is the
    one decision two solvers that agree to 1e-13 cannot be asked to share (tests/replay_util.py:
    compare, allow_tie): a robot turning in place produces coincident keyframes,
    VerticesFinder::findClosestVertex (vertices_finder.cpp:95-112) picks among vertices whose distances
    differ at rounding level, and graph_slam.cpp:417 drops the match when the pick is the previous
    keyframe. Returns None when the runs never part."""
    na, nb = a["n_edges%d" % r], b["n_edges%d" % r]
    d = np.nonzero(na != nb)[0]
    if not len(d):
        return None
    k = int(d[0])
    vid = 10000 * r + k
    pos = {int(v[0]): v[1:3] for v in a["vertices%d" % r]}
    ea = set(map(tuple, a["edges%d" % r][:, [0, 1]].astype(int)))
    eb = set(map(tuple, b["edges%d" % r][:, [0, 1]].astype(int)))
    mine = [e for e in (ea ^ eb) if max(e) == vid and min(e) // 10000 == r]
    assert mine, (r, k, "edge counts part but no edge of that keyframe differs")
    for i, j in mine:
        assert np.hypot(*(pos[i] - pos[j])) < 0.05, (r, k, i, j, "not a coincident-keyframe tie")
    return k


for r in range(4):
    a = leaders[r]
    same_traffic = bool(np.array_equal(a["msgs"], b["msgs"]))
    k = tie_at(a, b, r)
    e = a["edges%d" % r]
    inter = int(((e[:, 0] // 10000 != r) | (e[:, 1] // 10000 != r)).sum())
    n = len(a["est%d" % r])
    if k is None:
        worst = _same_graph(a, b, r, TOL)
        fd = float(b["follow_diff%d" % r].max())
        span = "all %d keyframes" % n
    else:   # per-keyframe records up to the tie
        d = a["est%d" % r][:k] - b["est%d" % r][:k]
        d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
        worst = float(np.abs(d).max())
        fd = float(b["follow_diff%d" % r][:k].max())
        span = "keyframes 0..%d of %d (coincident-keyframe tie at %d, graphs differ from there)" % (k - 1, n, k)
    print("robot %d: %s; %d vertices, %d edges (%d inter-robot, %d condensed star edges) at the end; max |difference| %.2e,"
          " follower's own solve differs by at most %.2e"
          % (r, span, len(a["vertices%d" % r]), len(e), inter, int((e[:, 6] > 0).sum()), worst, fd))
    assert same_traffic and worst < TOL and fd < TOL
    worst_all = max(worst_all, worst, fd)
print("datagrams %d (identical in both runs), sizes %d..%d bytes; worst difference %.2e (tolerance 1e-6)"
      % (len(b["msgs"]), int(b["msgs"][:, 3].min()), int(b["msgs"][:, 3].max()), worst_all))
