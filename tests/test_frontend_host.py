"""Candidate selection and closure voting: the C++ mirrors in
include/cgm/slam_frontend.hpp (VerticesFinder, ClosureBuffer, LoopClosureChecker,
addNeighboringVertices) against the restated oracle, through the reference-shaped graph API.
Host logic only: runs on a GPU-less box (the covariance gate, which needs marginals from the
solver, is in test_cpp_compat.py under the gpu marker)."""
import os
import subprocess

import numpy as np
import pytest

from cg_mrslam_b200 import synth
from oracle import frontend_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cg_mrslam_b200", "lib")


@pytest.fixture(scope="module", params=["gpu", "cpu"])
def driver(request):
    """The reference's own VerticesFinder / ClosureBuffer / LoopClosureChecker / MRClosureBuffer /
    GraphSLAM::addNeighboringVertices over the compat layer (gpu build; no compute call is made, so
    it runs on a GPU-less box) and over the CPU oracle build."""
    import ref_frontend
    return ref_frontend.driver_path("compat_driver", request.param)


def graph(n=400, e=1500, seed=12, box=22.0, base=10000):
    g = synth.make_pose_graph(n, e, seed=seed, box=box, init="truth_noisy")
    ids = [base + k for k in range(n)]
    poses = {ids[k]: g["poses0"][k] for k in range(n)}
    edges = [(ids[a], ids[b]) for a, b in g["edge_ij"]]
    return g, ids, poses, edges


def write(path, g, ids, commands):
    with open(path, "w") as f:
        for k, vid in enumerate(ids):
            p = g["poses0"][k]
            f.write("V %d %.17g %.17g %.17g %d 0 0 0 8\n" % (vid, p[0], p[1], p[2], 1 if k == 0 else 0))
        for (a, b), z, w in zip(g["edge_ij"], g["meas"], g["info"]):
            f.write("E %d %d %.17g %.17g %.17g %s\n" % (ids[a], ids[b], z[0], z[1], z[2],
                                                      " ".join("%.17g" % x for x in w)))
        for c in commands:
            f.write(c + "\n")


def run(exe, path):
    import ref_frontend
    return ref_frontend.run_driver(exe, [path], timeout=300)[0]


def test_candidate_selection(driver, tmp_path):
    g, ids, poses, edges = graph()
    curs = [ids[-1], ids[200], ids[37]]
    cmds = []
    for c in curs:
        cmds += ["FINDSM %d" % c, "SETS %d" % c]
    path = str(tmp_path / "s.txt")
    write(path, g, ids, cmds)
    lines = run(driver, path)
    at = 0
    for c in curs:
        want = fo.find_vertices_scan_matching(poses, edges, c)
        got = [int(x) for x in lines[at].split()[2:]]
        assert lines[at].startswith("FINDSM") and got == want and len(want) > 5
        at += 1
        groups = fo.find_sets_of_vertices(edges, want)
        assert lines[at] == "SETS %d" % len(groups)
        for k, grp in enumerate(groups):
            tok = lines[at + 1 + k].split()
            assert tok[0] == "G" and [int(x) for x in tok[3:]] == grp
            assert int(tok[1]) == fo.find_closest_vertex(poses, grp, c)     # the closure vertex index
        at += 1 + len(groups)


def test_add_neighboring_vertices(driver, tmp_path):
    g, ids, poses, edges = graph(120, 300, seed=3, box=12.0)
    cur = ids[-1]
    seeds = [ids[5], ids[6], ids[40], ids[110], ids[118]]
    path = str(tmp_path / "s.txt")
    write(path, g, ids, ["NEIGH %d 8 %d %s" % (cur, len(seeds), " ".join(map(str, seeds)))])
    got = [int(x) for x in run(driver, path)[0].split()[2:]]
    assert got == fo.add_neighboring_vertices(poses, seeds, cur, 8) and cur not in got


def test_closure_voting(driver, tmp_path):
    g, ids, poses, edges = graph(200, 600, seed=8, box=15.0)
    rng = np.random.default_rng(1)
    window = ids[180:190]
    info = np.diag([1000.0, 1000.0, 10000.0])
    from oracle import pgo_oracle as po
    cands = []
    for k, w in enumerate(window):
        other = ids[20 + 3 * k]
        rel = po.se2_mul(po.se2_inv(poses[other]), poses[w])[0]
        noise = [0.01, 0.01, 0.003] if k % 4 else [0.8, 0.8, 0.3]        # every 4th is an outlier
        z = po.se2_mul(rel, rng.normal(0, noise))[0]
        cands.append((other, w, tuple(z)))
    cmd = "CLOSURES 2.0 %d %s %d %s" % (len(window), " ".join(map(str, window)), len(cands),
                                        " ".join("%d %d %.17g %.17g %.17g" % (a, b, z[0], z[1], z[2])
                                                 for a, b, z in cands))
    path = str(tmp_path / "s.txt")
    write(path, g, ids, [cmd])
    lines = run(driver, path)
    inl, chi2, per = fo.closure_check(poses, set(window), cands, 2.0, info)
    tok = lines[0].split()
    assert int(tok[1]) == inl and 0 < inl < len(cands)
    assert abs(float(tok[2]) - chi2) <= 1e-9 * max(1.0, abs(chi2))
    got = np.array([float(ln.split()[1]) for ln in lines[1:1 + len(cands)]])
    assert np.allclose(got, per, rtol=1e-9, atol=1e-12)
    assert [c < 2.0 for c in got] == [c < 2.0 for c in per]              # the accepted closures


def test_closure_buffer_window(driver, tmp_path):
    g, ids, poses, edges = graph(60, 120, seed=2, box=9.0)
    steps = [(ids[10], 2), (ids[11], 0), (ids[12], 1), (ids[13], 3), (ids[14], 0), (ids[15], 2),
             (ids[16], 0), (ids[17], 0), (ids[18], 1), (ids[19], 0), (ids[20], 0), (ids[21], 0)]
    cmd = "BUFSIM 4 %d %s" % (len(steps), " ".join("%d %d" % s for s in steps))
    path = str(tmp_path / "s.txt")
    write(path, g, ids, [cmd])
    lines = run(driver, path)
    win = fo.ClosureWindow(4)
    for ln, (v, ne) in zip(lines, steps):
        vote, n_edges, n_vertices, vl = win.step(v, ne)
        tok = ln.split()
        assert tok[0] == "B" and int(tok[1]) == int(vote)
        assert int(tok[2]) == n_edges and int(tok[3]) == n_vertices
        assert [tuple(int(x) for x in t.split(":")) for t in tok[4:]] == vl
    assert any(ln.split()[1] == "1" for ln in lines)


def test_mr_closure_buffer(driver, tmp_path):
    """MRClosureBuffer (mr_closure_buffer.cpp:30-118): per-peer windows under insert / remove /
    update, incl. a vertex inserted twice, removal of the last vertex of a peer, and ageing out."""
    g, ids, poses, edges = graph(60, 120, seed=4, box=9.0)
    ops = [("I", 1, ids[5], 1), ("I", 2, ids[7], 0), ("I", 1, ids[6], 2), ("U", 0, 0, 0),
           ("I", 1, ids[5], 0), ("I", 3, ids[9], 1), ("R", 2, ids[7], 0), ("U", 0, 0, 0),
           ("R", 1, ids[6], 0), ("U", 0, 0, 0), ("I", 2, ids[8], 3), ("U", 0, 0, 0), ("U", 0, 0, 0),
           ("R", 4, ids[5], 0), ("U", 0, 0, 0), ("U", 0, 0, 0), ("U", 0, 0, 0)]
    cmd = "MRBUF 3 %d %s" % (len(ops), " ".join("%s %d %d %d" % o for o in ops))
    path = str(tmp_path / "s.txt")
    write(path, g, ids, [cmd])
    lines = run(driver, path)
    mr = fo.MRClosureWindows()
    seen_sizes = set()
    for ln, (op, robot, v, ne) in zip(lines, ops):
        if op == "I":
            mr.insert(robot, v, ne)
        elif op == "R":
            mr.remove(robot, v)
        else:
            mr.update(3)
        want = mr.state()
        parts = ln.split(" | ")
        assert parts[0] == "MR %d" % len(want), (ln, want)
        for part, (r, n_edges, n_vertices, vl) in zip(parts[1:], want):
            tok = part.split()
            assert [int(tok[0]), int(tok[1]), int(tok[2])] == [r, n_edges, n_vertices], (ln, want)
            assert [tuple(int(x) for x in t.split(":")) for t in tok[3:]] == vl, (ln, want)
        seen_sizes.add(len(want))
    assert seen_sizes >= {0, 1, 2, 3}

