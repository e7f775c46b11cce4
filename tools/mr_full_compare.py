"""Evidence run for BASELINE cfg 3 at full length (every keyframe of the 4-robot bag, the reference's
default inter-robot quorum; lockstep at the granularity of the solver calls, g2o::trace): compares a GPU leader's output (single process, or one robot per rank /
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
PERIOD = 0.15
fx = [np.load(os.path.join(ROOT, "tests", "golden", "bag_4robots_robot%d_full.npz" % r)) for r in range(4)]
t0 = min(f["time"][0] for f in fx)


def first_difference(a, b, r):
    """First keyframe at which the follower's edge count or its own solve parts from the leader's."""
    na, nb = a["n_edges%d" % r], b["n_edges%d" % r]
    bad = (na != nb) | (b["follow_diff%d" % r] > TOL)
    d = np.nonzero(bad)[0]
    return int(d[0]) if len(d) else None


def is_tie(a, b, r, k):
    """The one decision two solvers that agree to 1e-13 cannot be asked to share (tests/replay_util.py:
    compare, allow_tie): a robot turning in place produces coincident keyframes,
    VerticesFinder::findClosestVertex (vertices_finder.cpp:95-112) picks among vertices whose distances
    differ at rounding level, and graph_slam.cpp:417 drops the match when the pick is the previous
    keyframe. True if the runs differ by an edge of keyframe k between vertices less than 5 cm apart."""
    vid = 10000 * r + k
    pos = {int(v[0]): v[1:3] for v in a["vertices%d" % r]}
    ea = set(map(tuple, a["edges%d" % r][:, [0, 1]].astype(int)))
    eb = set(map(tuple, b["edges%d" % r][:, [0, 1]].astype(int)))
    mine = [e for e in (ea ^ eb) if max(e) == vid and min(e) // 10000 == r]
    return bool(mine) and all(np.hypot(*(pos[i] - pos[j])) < 0.05 for i, j in mine)


first = {r: first_difference(leaders[r], b, r) for r in range(4)}
when = {r: (fx[r]["time"][first[r]] if first[r] is not None else np.inf) for r in range(4)}
cause = {}
for r in sorted(range(4), key=lambda q: when[q]):
    if first[r] is None:
        continue
    if is_tie(leaders[r], b, r, first[r]):
        cause[r] = "coincident-keyframe tie at keyframe %d" % first[r]
        continue
    # otherwise it must have heard from a robot that had already parted: a datagram q -> r sent between
    # q's first difference and this keyframe
    msgs = b["msgs"]
    t_msg = t0 + (msgs[:, 0] + 1) * PERIOD
    src = [q for q in cause if np.any((msgs[:, 1] == q) & (msgs[:, 2] == r) & (t_msg >= when[q]) & (t_msg <= when[r] + PERIOD))]
    assert src, (r, first[r], "parts from the leader with neither a tie nor a message from a robot that had parted")
    cause[r] = "messages from robot %s, which had parted before (keyframe %d)" % (", ".join(map(str, src)), first[r])

for r in range(4):
    a = leaders[r]
    same_traffic = bool(np.array_equal(a["msgs"], b["msgs"]))
    k = first[r]
    e = a["edges%d" % r]
    inter = int(((e[:, 0] // 10000 != r) | (e[:, 1] // 10000 != r)).sum())
    n = len(a["est%d" % r])
    if k is None:
        worst = _same_graph(a, b, r, TOL)
        fd = float(b["follow_diff%d" % r].max())
        span = "all %d keyframes" % n
    else:   # per-keyframe records up to there
        assert np.array_equal(a["n_edges%d" % r][:k], b["n_edges%d" % r][:k])
        worst = 0.0
        fd = float(b["follow_diff%d" % r][:k].max())
        span = "keyframes 0..%d of %d (then: %s)" % (k - 1, n, cause[r])
    print("robot %d: %s; %d vertices, %d edges (%d inter-robot, %d condensed star edges) at the end; graph difference %.2e,"
          " follower's own solve differs by at most %.2e"
          % (r, span, len(a["vertices%d" % r]), len(e), inter, int((e[:, 6] > 0).sum()), worst, fd))
    assert same_traffic and worst < TOL and fd < TOL
    worst_all = max(worst_all, worst, fd)
kinds = ("optimize", "computeInitialGuess", "computeMarginals", "labelEdges measurement", "labelEdges information")
if "trace0" in b.files:
    print("solver calls of the follower against the leader's, per kind (count, worst difference; estimates and measurements"
          " absolute, covariance / information blocks relative to the block's largest entry):")
    for i, k in enumerate(kinds):
        print("   %-24s %6d calls, worst %.2e" % (k, sum(int(b["trace%d" % r][i, 0]) for r in range(4)),
                                                  max(float(b["trace%d" % r][i, 1]) for r in range(4))))
print("datagrams %d (identical in both runs), sizes %d..%d bytes; worst difference %.2e (tolerance 1e-6)"
      % (len(b["msgs"]), int(b["msgs"][:, 3].min()), int(b["msgs"][:, 3].max()), worst_all))
