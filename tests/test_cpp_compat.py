"""The drop-in boundary, exercised by the REFERENCE'S OWN host sources (compiled verbatim from
/root/reference, oracle/Makefile target `frontend`): ScanMatcher::closeScanMatching / scanMatchingLC /
globalMatching over include/cgm/chargrid.hpp, SparseOptimizer::optimize / computeMarginals /
EdgeLabeler::labelEdges / GraphManipulator / CondensedGraphBuffer over include/g2o_compat -- on the
GPU (gpu marker) they must reproduce the oracles; the same sources over the reference's CPU matcher
and the CPU oracle solver (no GPU needed) pin the Python oracles to the reference's code."""
import math
import os
import subprocess

import numpy as np
import pytest

from cg_mrslam_b200 import synth
from oracle import pgo_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cg_mrslam_b200", "lib")
LASER_POSE = (0.05, 0.0, 0.0)


@pytest.fixture(scope="module", params=[pytest.param("gpu", marks=pytest.mark.gpu), "cpu"])
def driver(request):
    """compat_driver_gpu: the reference's host classes over include/g2o_compat + include/cgm/chargrid.hpp
    + the CUDA library; compat_driver_cpu: the same reference sources over the reference's own
    chargrid.cpp and the CPU oracle solver -- run here without a GPU, it pins the Python oracles
    (scan_matcher_oracle, frontend_oracle, pgo_oracle) to the reference's code."""
    import ref_frontend
    return ref_frontend.driver_path("compat_driver", request.param)


@pytest.fixture(scope="module")
def gpu_driver():
    import ref_frontend
    return ref_frontend.driver_path("compat_driver", "gpu")


def _scenario(n=14, seed=4):
    """Keyframes along a short path in a synthetic room, with scans and a small pose graph."""
    rng = np.random.default_rng(seed)
    vert, horiz, size, boxes = synth._room_segments(rng)
    x0, y0 = synth._free_pose(rng, size, boxes, margin=1.5)
    truth = []
    th = rng.uniform(-math.pi, math.pi)
    x, y = x0, y0
    for k in range(n):
        truth.append((x, y, th))
        th += rng.uniform(-0.25, 0.25)
        nx, ny = x + 0.25 * math.cos(th), y + 0.25 * math.sin(th)
        if 1.0 < nx < size[0] - 1.0 and 1.0 < ny < size[1] - 1.0 and all(
                not (b[0] - 0.6 < nx < b[2] + 0.6 and b[1] - 0.6 < ny < b[3] + 0.6) for b in boxes):
            x, y = nx, ny
        else:
            th += math.pi / 2
    truth = np.array(truth)
    truth[:, 2] = po.normalize_theta(truth[:, 2])
    first, step, max_range, nb = -math.pi / 2, math.pi / 360, 8.0, 361
    verts = []
    for k in range(n):
        lp = po.se2_mul(truth[k], np.array(LASER_POSE))[0]
        r = synth.cast_scan(vert, horiz, lp, nb, first, step, max_range, 0.01, rng)
        pose = truth[k] + rng.normal(0, [0.03, 0.03, 0.01])
        pose[2] = po.normalize_theta(np.array([pose[2]]))[0]
        verts.append(dict(id=100 + 3 * k, pose=pose if k else truth[0].copy(), ranges=r,
                          first_angle=first, step=step, max_range=max_range,
                          laser_pose=LASER_POSE, fixed=(k == 0)))
    edges = []
    pairs = [(k, k + 1) for k in range(n - 1)] + [q for q in [(0, 5), (2, 9), (4, 12), (1, 13)] if q[1] < n]
    for a, b in pairs:
        rel = po.se2_mul(po.se2_inv(truth[a]), truth[b])[0]
        z = po.se2_mul(rel, rng.normal(0, [0.01, 0.01, 0.003]))[0]
        w = (1000.0, 0.0, 0.0, 1000.0, 0.0, 10000.0) if b != a + 1 else (100.0, 0, 0, 100.0, 0, 1000.0)
        edges.append((a, b, z, w))
    return verts, edges


def _write(path, verts, edges, commands):
    with open(path, "w") as f:
        for v in verts:
            f.write("V %d %.17g %.17g %.17g %d %d %.17g %.17g %.17g %s\n" % (
                v["id"], v["pose"][0], v["pose"][1], v["pose"][2], 1 if v["fixed"] else 0,
                len(v["ranges"]), v["first_angle"], v["step"], v["max_range"],
                " ".join("%.17g" % r for r in v["ranges"])))
        for a, b, z, w in edges:
            f.write("E %d %d %.17g %.17g %.17g %s\n" % (
                verts[a]["id"], verts[b]["id"], z[0], z[1], z[2], " ".join("%.17g" % x for x in w)))
        for c in commands:
            f.write(c + "\n")


def _run(exe, path):
    import ref_frontend
    return ref_frontend.run_driver(exe, [path], timeout=300)


def test_graph_bookkeeping_and_text_format(gpu_driver, tmp_path):
    driver = gpu_driver
    verts, edges = _scenario()
    g2o_path = str(tmp_path / "out.g2o")
    path = str(tmp_path / "s.txt")
    _write(path, verts, edges, ["POSES", "SAVE " + g2o_path])
    lines, _ = _run(driver, path)
    poses = np.array([[float(x) for x in ln.split()[2:]] for ln in lines if ln.startswith("P ")])
    assert np.array_equal(poses, np.array([v["pose"] for v in verts]))   # id order == input order
    assert "SAVE 1" in lines
    g = po.read_g2o(g2o_path)
    assert list(g["ids"]) == [v["id"] for v in verts] and list(g["fixed"]) == [0]
    assert np.array_equal(g["poses"], np.array([v["pose"] for v in verts]))
    assert np.array_equal(g["meas"], np.array([e[2] for e in edges]))
    assert np.array_equal(g["edge_ij"], np.array([[e[0], e[1]] for e in edges]))


def test_compute_calls_fail_loudly_without_gpu(gpu_driver, tmp_path):
    import torch
    driver = gpu_driver
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    verts, edges = _scenario(6)
    path = str(tmp_path / "s.txt")
    _write(path, verts, edges, ["OPT 3"])
    lines, err = _run(driver, path)
    assert "OPT 0" in lines and "SparseOptimizer" in err   # reported, not silently computed


def test_cpp_api_matches_oracles(driver, tmp_path, oracle_lib):
    from oracle import bindings
    from oracle import scan_matcher_oracle as smo
    lib = bindings.MatcherLib("reference") if bindings.have_reference() else oracle_lib
    verts, edges = _scenario()
    ids = [v["id"] for v in verts]
    cmds = ["CLOSE %d %d 0.2 6 %s" % (ids[5], ids[6], " ".join(str(i) for i in ids[0:6])),
            "LC %d %d 0.3 5 %s" % (ids[2], ids[11], " ".join(str(i) for i in ids[0:5])),
            "GLOBAL %d %d 0.3 4 %s" % (ids[1], ids[10], " ".join(str(i) for i in ids[0:4])),
            "OPT 5", "POSES",
            "MARG 3 %d %d %d" % (ids[3], ids[7], ids[13]),
            "STAR %d 4 %d %d %d %d" % (ids[6], ids[2], ids[6], ids[9], ids[13]),
            "POSES"]
    path = str(tmp_path / "s.txt")
    _write(path, verts, edges, cmds)
    lines, _ = _run(driver, path)

    # --- matcher front half + GPU search vs the CPU matcher -------------------------------------
    ok, want = smo.close_scan_matching(lib, oracle_lib, verts[0:6], verts[5], verts[6], 0.2)
    got = [ln for ln in lines if ln.startswith("CLOSE")][0].split()
    assert int(got[1]) == int(ok)
    if ok:
        assert np.array_equal(np.array([float(x) for x in got[2:]]), want)
    want_lc = smo.scan_matching_lc(lib, oracle_lib, verts[0:5], verts[2], verts[11], 0.3)
    k = [i for i, ln in enumerate(lines) if ln.startswith("LC ")][0]
    n_lc = int(lines[k].split()[2])
    got_lc = [np.array([float(x) for x in lines[k + 1 + t].split()[1:]]) for t in range(n_lc)]
    assert n_lc == len(want_lc)
    for a, b in zip(got_lc, want_lc):
        assert np.array_equal(a, b)
    okg, want_g = smo.global_matching(lib, oracle_lib, verts[0:4], verts[1], verts[10], 0.3)
    got = [ln for ln in lines if ln.startswith("GLOBAL")][0].split()
    assert int(got[1]) == int(okg)
    if okg:
        assert np.array_equal(np.array([float(x) for x in got[2:]]), want_g)

    # --- optimize(5), marginals, condensed star vs the pose-graph oracle -------------------------
    poses0 = np.array([v["pose"] for v in verts])
    e_ij = np.array([[a, b] for a, b, _, _ in edges], dtype=np.int64)
    meas = np.array([z for _, _, z, _ in edges])
    info = np.array([w for _, _, _, w in edges], dtype=np.float64)
    ref = po.gauss_newton(poses0, e_ij, meas, info, [0], 5)
    assert "OPT 5" in lines
    blocks = [i for i, ln in enumerate(lines) if ln.startswith("P ")]
    first = np.array([[float(x) for x in lines[i].split()[2:]] for i in blocks[:len(verts)]])
    d = first - ref.poses
    d[:, 2] = po.normalize_theta(d[:, 2])
    assert np.abs(d).max() < 1e-6
    hidx = po.hessian_index(len(verts), [0])
    want_m = po.marginals(ref, hidx, [(3, 3), (7, 7), (13, 13)])
    got_m = np.array([[float(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("M ")])
    assert np.abs(got_m.reshape(-1, 3, 3) - want_m).max() < 1e-9 * max(1.0, np.abs(want_m).max())
    z, om, vs = po.condensed_star(ref.poses, e_ij, meas, info, 6, [2, 6, 9, 13])
    got_s = np.array([[float(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("S ")])
    assert [int(r[0]) for r in got_s] == [ids[v] for v in vs]
    dz = got_s[:, 1:4] - z
    dz[:, 2] = po.normalize_theta(dz[:, 2])
    assert np.abs(dz).max() < 1e-6
    assert np.abs(got_s[:, 4:].reshape(-1, 3, 3) - om).max() < 1e-6 * np.abs(om).max()
    # push/pop around the condensed-graph computation restores the estimates
    second = np.array([[float(x) for x in lines[i].split()[2:]] for i in blocks[len(verts):]])
    assert np.array_equal(first, second)


def test_covariance_gate_matches_oracle(driver, tmp_path):
    """GraphSLAM::checkCovariance (graph_slam.cpp:311-354) through CovarianceEstimator: the whole
    graph re-optimised once with the current vertex as the gauge, marginal blocks from the GPU
    solver, chi2 gate on the xy offset. The surviving candidate ids must equal the oracle's."""
    from oracle import frontend_oracle as fo
    g = synth.make_pose_graph(300, 1100, seed=17, box=19.0, init="truth_noisy")
    n = len(g["poses0"])
    ids = [20000 + k for k in range(n)]
    cur = n - 1
    d = np.hypot(*(g["poses0"][:n - 1, :2] - g["poses0"][cur, :2]).T)
    cand = sorted(set(np.argsort(d)[:30].tolist()) | set(range(1, n - 1, 23)))   # near ones + far ones
    path = str(tmp_path / "s.txt")
    with open(path, "w") as f:
        for k in range(n):
            p = g["poses0"][k]
            f.write("V %d %.17g %.17g %.17g %d 0 0 0 8\n" % (ids[k], p[0], p[1], p[2], 1 if k == 0 else 0))
        for (a, b), z, w in zip(g["edge_ij"], g["meas"], g["info"]):
            f.write("E %d %d %.17g %.17g %.17g %s\n" % (ids[a], ids[b], z[0], z[1], z[2],
                                                      " ".join("%.17g" % x for x in w)))
        f.write("COVGATE %d %d %s\nPOSES\n" % (ids[cur], len(cand), " ".join(str(ids[k]) for k in cand)))
    lines, _ = _run(driver, path)
    got = [int(x) for x in lines[0].split()[2:]]
    # oracle: gauge = current vertex, spanning-tree guess, one GN iteration, marginals of that H
    guess = po.initial_guess(g["poses0"], g["edge_ij"], g["meas"], [cur])
    res = po.gauss_newton(guess, g["edge_ij"], g["meas"], g["info"], [cur], 1)
    hidx = po.hessian_index(n, [cur])
    covs = po.marginals(res, hidx, [(k, k) for k in cand])
    poses = {ids[k]: g["poses0"][k] for k in range(n)}       # the gate uses the LIVE estimates (popState)
    want = fo.check_covariance(poses, [ids[k] for k in cand], ids[cur], {ids[k]: c for k, c in zip(cand, covs)})
    assert got == want and 0 < len(want) < len(cand)
    # GraphManipulator::popState left the estimates untouched
    after = np.array([[float(x) for x in ln.split()[2:]] for ln in lines if ln.startswith("P ")])
    assert np.array_equal(after, g["poses0"])


def test_condensed_graph_buffer_matches_oracle(driver, tmp_path):
    """CondensedGraphBuffer (condensed_graph_buffer.cpp:94-510) + CondensedGraphCreator
    (condensed_graph_creator.cpp:33-66) through their C++ mirrors: the star a peer robot gets over
    the vertices it asked about -- centroid gauge and 'optimal' gauge (argmin of sum det(Omega^-1)
    over all candidate gauges) --, its replacement on the next request, the bookkeeping of a star
    received from a peer, and that neither kind of star leaks into 'my own edges'."""
    g = synth.make_pose_graph(260, 950, seed=23, box=17.0, init="truth_noisy")
    n = len(g["poses0"])
    ids = [10000 + k for k in range(n)]
    asked = [12, 57, 58, 131, 190, 244]
    more = [77, 131, 203]                       # a later request: 131 is already listed
    peer_star = [(57, 190, (0.4, -0.2, 0.05)), (57, 12, (-1.0, 0.3, -0.4))]
    path = str(tmp_path / "cgb.txt")
    with open(path, "w") as f:
        for k in range(n):
            p = g["poses0"][k]
            f.write("V %d %.17g %.17g %.17g %d 0 0 0 8\n" % (ids[k], p[0], p[1], p[2], 1 if k == 0 else 0))
        for (a, b), z, w in zip(g["edge_ij"], g["meas"], g["info"]):
            f.write("E %d %d %.17g %.17g %.17g %s\n" % (ids[a], ids[b], z[0], z[1], z[2],
                                                      " ".join("%.17g" % x for x in w)))
        f.write("CGB 1 0 %d %s\n" % (len(asked), " ".join(str(ids[k]) for k in asked)))
        f.write("CGIN 1 %d %s\n" % (len(peer_star), " ".join(
            "%d %d %.17g %.17g %.17g 500 0 0 500 0 5000" % (ids[a], ids[b], z[0], z[1], z[2])
            for a, b, z in peer_star)))
        f.write("CGB 1 0 %d %s\n" % (len(more), " ".join(str(ids[k]) for k in more)))
        f.write("CGB 2 1 %d %s\n" % (len(asked[:4]), " ".join(str(ids[k]) for k in asked[:4])))
        f.write("POSES\n")
    lines, _ = _run(driver, path)
    heads = [i for i, ln in enumerate(lines) if ln.startswith("CGB ")]
    n_e = len(g["edge_ij"])

    def star_of(i):
        head = [int(x) for x in lines[heads[i]].split()[1:]]
        rows = []
        for ln in lines[heads[i] + 1:]:
            if not ln.startswith("C "):
                break
            rows.append([float(x) for x in ln.split()[1:]])
        return head, np.array(rows)

    def check(rows, seps, gauge, edge_ij, meas, info):
        z, om, vs = po.condensed_star(g["poses0"], edge_ij, meas, info, gauge, sorted(seps))
        assert [int(r[0]) for r in rows] == [ids[v] for v in vs]
        dz = rows[:, 1:4] - z
        dz[:, 2] = po.normalize_theta(dz[:, 2])
        assert np.abs(dz).max() < 1e-6
        assert np.abs(rows[:, 4:].reshape(-1, 3, 3) - om).max() < 1e-6 * np.abs(om).max()
        return om

    # 1. first request of robot 1: centroid gauge, star over the asked vertices, level-2 edges
    head, rows = star_of(0)
    gauge = po.select_gauge_centroid(g["poses0"], sorted(asked))
    assert head == [ids[gauge], len(asked) - 1, n_e + len(asked) - 1, n_e, len(asked) - 1, len(asked)]
    check(rows, asked, gauge, g["edge_ij"], g["meas"], g["info"])
    # 2. the peer's star arrives: level-0 edges in the graph, but not among "my own edges"
    assert [int(x) for x in lines[[i for i, ln in enumerate(lines) if ln.startswith("CGIN ")][0]].split()[1:]] == \
        [len(peer_star), n_e + len(asked) - 1 + len(peer_star), n_e]
    # 3. a later request adds two vertices; the old star is replaced; the star is computed from
    #    this robot's own edges only (the peer's star does not enter)
    head, rows = star_of(1)
    union = sorted(set(asked) | set(more))
    gauge = po.select_gauge_centroid(g["poses0"], union)
    assert head == [ids[gauge], len(union) - 1, n_e + len(union) - 1 + len(peer_star), n_e, len(union) - 1,
                    len(union)]
    check(rows, union, gauge, g["edge_ij"], g["meas"], g["info"])
    # 4. robot 2, optimal gauge: every candidate's star is computed, the least uncertain one wins
    head, rows = star_of(2)
    cand = sorted(asked[:4])
    cost = []
    for c in cand:
        _, om, _ = po.condensed_star(g["poses0"], g["edge_ij"], g["meas"], g["info"], c, cand)
        cost.append(sum(np.linalg.det(np.linalg.inv(o)) for o in om))
    best = cand[int(np.argmin(cost))]
    assert head[0] == ids[best] and head[1] == len(cand) - 1 and head[4] == len(cand) - 1
    check(rows, cand, best, g["edge_ij"], g["meas"], g["info"])
    # push / pop inside the creator left every estimate untouched
    after = np.array([[float(x) for x in ln.split()[2:]] for ln in lines if ln.startswith("P ")])
    assert np.array_equal(after, g["poses0"])

