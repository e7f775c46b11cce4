"""Two MRGraphSLAM mirrors (include/cgm/mr_graph_slam.hpp) exchange condensed graphs through the
float32 wire format, as cg_mrslam.cpp:236-259 / graph_comm.cpp:126-154 do over UDP: B asks A about
the A-vertices it closed loops with, A computes the star on the GPU and answers, B optimises with
the star. Checked against the pose-graph oracle: datagram sizes, the star (1e-6, and exactly
float32-representable: it crossed the wire), B's estimates after optimize(5) (1e-6)."""
import os
import subprocess

import numpy as np
import pytest

from cg_mrslam_b200 import synth
from oracle import pgo_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cg_mrslam_b200", "lib")


@pytest.fixture(scope="module", params=[pytest.param("gpu", marks=pytest.mark.gpu), "cpu"])
def driver(request):
    """mr_exchange_gpu: the reference's own MRGraphSLAM (compiled verbatim) over the CUDA library;
    mr_exchange_cpu: the same sources over the reference's CPU matcher and the CPU oracle solver."""
    import ref_frontend
    return ref_frontend.driver_path("mr_exchange", request.param)


def scenario(path):
    rng = np.random.default_rng(31)
    ga = synth.make_pose_graph(200, 720, seed=41, box=15.0, init="truth_noisy")
    gb = synth.make_pose_graph(180, 640, seed=42, box=15.0, init="truth_noisy")
    na, nb = len(ga["poses0"]), len(gb["poses0"])
    ida = list(range(na))                       # robot 0: ids 0 .. (id / 10000 == 0)
    idb = [10000 + k for k in range(nb)]        # robot 1
    asked = [23, 88, 141, 190]                  # A's vertices B closed loops with
    partners = [15, 60, 99, 170]                # ... from these B vertices
    # B's copies of A's vertices: A's estimate seen with some error; inter-robot closure edges
    copies = ga["poses0"][asked] + rng.normal(size=(len(asked), 3)) * [0.05, 0.05, 0.01]
    ir_meas = []
    for p, c in zip(partners, copies):
        rel = po.se2_mul(po.se2_inv(gb["poses0"][p]), c)[0]
        ir_meas.append(po.se2_mul(rel, rng.normal(0, [0.02, 0.02, 0.005]))[0])
    ir_info = [100.0, 0.0, 0.0, 100.0, 0.0, 1000.0]
    with open(path, "w") as f:
        for r, g, ids in ((0, ga, ida), (1, gb, idb)):
            for k, vid in enumerate(ids):
                p = g["poses0"][k]
                f.write("V %d %d %.17g %.17g %.17g %d\n" % (r, vid, p[0], p[1], p[2], 1 if k == 0 else 0))
            for (a, b), z, w in zip(g["edge_ij"], g["meas"], g["info"]):
                f.write("E %d %d %d %.17g %.17g %.17g %s\n" % (r, ids[a], ids[b], z[0], z[1], z[2],
                                                             " ".join("%.17g" % x for x in w)))
        for a, c in zip(asked, copies):
            f.write("V 1 %d %.17g %.17g %.17g 0\n" % (ida[a], c[0], c[1], c[2]))
        for p, a, z in zip(partners, asked, ir_meas):
            f.write("E 1 %d %d %.17g %.17g %.17g %s\n" % (idb[p], ida[a], z[0], z[1], z[2],
                                                        " ".join("%.17g" % x for x in ir_info)))
        f.write("WANT 1 0 %d %s\n" % (len(asked), " ".join(str(ida[a]) for a in asked)))
    return ga, gb, ida, idb, asked, partners, copies, ir_meas, ir_info


def test_condensed_graph_exchange(driver, tmp_path):
    path = str(tmp_path / "mr.txt")
    ga, gb, ida, idb, asked, partners, copies, ir_meas, ir_info = scenario(path)
    na, nb = len(ga["poses0"]), len(gb["poses0"])
    import ref_frontend
    lines, _ = ref_frontend.run_driver(driver, [path])
    msgs = [ln.split() for ln in lines if ln.startswith("MSG ")]
    n_star = len(asked) - 1
    # B -> A: header 8 + edge count 8 + closure count 8 + 4 ids; A -> B: the star, no requests
    assert msgs[0] == ["MSG", "1", "0", str(8 + 8 + 8 + 4 * len(asked)), "closures", str(len(asked)), "edges", "0"]
    assert msgs[1] == ["MSG", "0", "1", str(8 + 8 + n_star * (8 + 9 * 4) + 8), "closures", "0", "edges", str(n_star)]
    # the star A computed: centroid gauge, its own edges only
    gauge = po.select_gauge_centroid(ga["poses0"], sorted(asked))
    z, om, vs = po.condensed_star(ga["poses0"], ga["edge_ij"], ga["meas"], ga["info"], gauge, sorted(asked))
    rows = np.array([[float(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("C ")])
    assert [ln.split()[:3] for ln in lines if ln.startswith("STAR ")][0] == \
        ["STAR", str(n_star), str(len(gb["edge_ij"]) + len(asked) + n_star)]
    assert [int(r[0]) for r in rows] == [ida[gauge]] * n_star and [int(r[1]) for r in rows] == [ida[v] for v in vs]
    got_z, got_om = rows[:, 2:5], rows[:, 5:].reshape(-1, 3, 3)
    dz = got_z - z
    dz[:, 2] = po.normalize_theta(dz[:, 2])
    assert np.abs(dz).max() < 1e-6
    assert np.abs(got_om - om).max() < 1e-6 * np.abs(om).max()
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    assert np.array_equal(got_z, f32(got_z)) and np.array_equal(got_om, f32(got_om))   # it crossed the wire
    assert np.array_equal(got_om, np.transpose(got_om, (0, 2, 1)))
    # B's optimisation: its own graph + the closures + the star as received
    idx = {vid: k for k, vid in enumerate(idb)}
    for j, a in enumerate(asked):
        idx[ida[a]] = nb + j
    poses0 = np.vstack([gb["poses0"], copies])
    e_ij = [list(e) for e in gb["edge_ij"]] + [[idx[idb[p]], idx[ida[a]]] for p, a in zip(partners, asked)] + \
        [[idx[int(r[0])], idx[int(r[1])]] for r in rows]
    meas = np.vstack([gb["meas"], np.array(ir_meas), got_z])
    info = np.vstack([gb["info"], np.tile(ir_info, (len(asked), 1)),
                      np.array([[o[0, 0], o[0, 1], o[0, 2], o[1, 1], o[1, 2], o[2, 2]] for o in got_om])])
    ref = po.gauss_newton(poses0, np.array(e_ij, dtype=np.int64), meas, info, [0], 5)
    got_p = {int(ln.split()[1]): [float(x) for x in ln.split()[2:]] for ln in lines if ln.startswith("P ")}
    got = np.array([got_p[vid] for vid in sorted(idx, key=idx.get)])
    d = got - ref.poses
    d[:, 2] = po.normalize_theta(d[:, 2])
    assert np.abs(d).max() < 1e-6
    assert np.abs(got - poses0).max() > 1e-3          # the optimisation moved something
    # second round: the star is replaced on both sides, nothing is duplicated
    assert [ln.split()[1:] for ln in lines if ln.startswith("AGAIN ")][0] == \
        [str(n_star), str(len(gb["edge_ij"]) + len(asked) + n_star), str(len(ga["edge_ij"]) + n_star)]


def test_whole_graph_exchange(driver, tmp_path):
    """The GraphMessage alternative (mr_graph_slam.cpp:397-483, 672-739): A answers B's request with
    its whole own graph; B creates A's vertices (float32 estimates), refreshes the copies it had,
    takes A's edges as level-0 edges and optimises the union."""
    path = str(tmp_path / "mr.txt")
    ga, gb, ida, idb, asked, partners, copies, ir_meas, ir_info = scenario(path)
    na, nb = len(ga["poses0"]), len(gb["poses0"])
    import ref_frontend
    lines, _ = ref_frontend.run_driver(driver, [path, "graph"])
    msgs = [ln.split() for ln in lines if ln.startswith("MSG ")]
    n_ea = len(ga["edge_ij"])
    # B -> A: only the request (A has not been asked before: no vertices, no edges)
    assert msgs[0] == ["MSG", "1", "0", str(8 + 8 + 8 + 8 + 4 * len(asked)), "closures", str(len(asked)),
                       "edges", "0", "vertices", "0"]
    # A -> B: every vertex and every own edge of A (its star for B, level 2, is not among them)
    assert msgs[1] == ["MSG", "0", "1", str(8 + 8 + na * 16 + 8 + n_ea * 44 + 8), "closures", "0",
                       "edges", str(n_ea), "vertices", str(na)]
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)
    # B's union graph: its own vertices, then A's (all of them now), estimates as they crossed the wire
    idx = {vid: k for k, vid in enumerate(idb)}
    for k, vid in enumerate(ida):
        idx[vid] = nb + k
    poses0 = np.vstack([gb["poses0"], f32(ga["poses0"])])
    info_a = f32(ga["info"])
    e_ij = [list(e) for e in gb["edge_ij"]] + [[idx[idb[p]], idx[ida[a]]] for p, a in zip(partners, asked)] + \
        [[nb + a, nb + b] for a, b in ga["edge_ij"]]
    meas = np.vstack([gb["meas"], np.array(ir_meas), f32(ga["meas"])])
    info = np.vstack([gb["info"], np.tile(ir_info, (len(asked), 1)), info_a])
    ref = po.gauss_newton(poses0, np.array(e_ij, dtype=np.int64), meas, info, [0], 5)
    got_p = {int(ln.split()[1]): [float(x) for x in ln.split()[2:]] for ln in lines if ln.startswith("P ")}
    assert sorted(got_p) == sorted(idx)
    got = np.array([got_p[vid] for vid in sorted(idx, key=idx.get)])
    d = got - ref.poses
    d[:, 2] = po.normalize_theta(d[:, 2])
    assert np.abs(d).max() < 1e-6
    star = [ln.split() for ln in lines if ln.startswith("STAR ")][0]
    assert star[1:] == [str(n_ea), str(len(gb["edge_ij"]) + len(asked) + n_ea)]
    # second round: A's edges replace the ones B already had; nothing is duplicated
    assert [ln.split()[1:] for ln in lines if ln.startswith("AGAIN ")][0] == \
        [str(n_ea), str(len(gb["edge_ij"]) + len(asked) + n_ea), str(n_ea + len(asked) - 1)]


def test_two_robots_2k_nodes_8k_edges(driver, tmp_path):
    """BASELINE cfg 2 in synthetic form: two robots with 1000 poses / 4000 edges each, six
    inter-robot closures either way; each asks the other about the vertices it closed loops with,
    each computes the other's star on the GPU from its own edges (closures to the peer included),
    the stars cross the wire, both optimise. Stars and both robots' estimates against the oracle."""
    rng = np.random.default_rng(77)
    g = [synth.make_pose_graph(1000, 4000, seed=51 + r, box=35.0, init="truth_noisy") for r in range(2)]
    ids = [[10000 * r + k for k in range(1000)] for r in range(2)]
    # robot r closed loops between its vertices mine[r] and the peer's vertices theirs[r]
    mine = [[40, 300, 301, 640, 777, 910], [15, 220, 480, 481, 700, 950]]
    theirs = [[100, 250, 400, 555, 800, 990], [60, 333, 334, 500, 720, 880]]
    ir_info = [100.0, 0.0, 0.0, 100.0, 0.0, 1000.0]
    copies, ir_meas = [], []
    for r in range(2):
        peer = 1 - r
        c = g[peer]["poses0"][theirs[r]] + rng.normal(size=(len(theirs[r]), 3)) * [0.05, 0.05, 0.01]
        copies.append(c)
        zs = []
        for p, cc in zip(mine[r], c):
            rel = po.se2_mul(po.se2_inv(g[r]["poses0"][p]), cc)[0]
            zs.append(po.se2_mul(rel, rng.normal(0, [0.02, 0.02, 0.005]))[0])
        ir_meas.append(np.array(zs))
    path = str(tmp_path / "mr2k.txt")
    with open(path, "w") as f:
        for r in range(2):
            for k, vid in enumerate(ids[r]):
                p = g[r]["poses0"][k]
                f.write("V %d %d %.17g %.17g %.17g %d\n" % (r, vid, p[0], p[1], p[2], 1 if k == 0 else 0))
            for (a, b), z, w in zip(g[r]["edge_ij"], g[r]["meas"], g[r]["info"]):
                f.write("E %d %d %d %.17g %.17g %.17g %s\n" % (r, ids[r][a], ids[r][b], z[0], z[1], z[2],
                                                             " ".join("%.17g" % x for x in w)))
        for r in range(2):
            peer = 1 - r
            for a, c in zip(theirs[r], copies[r]):
                f.write("V %d %d %.17g %.17g %.17g 0\n" % (r, ids[peer][a], c[0], c[1], c[2]))
            for p, a, z in zip(mine[r], theirs[r], ir_meas[r]):
                f.write("E %d %d %d %.17g %.17g %.17g %s\n" % (r, ids[r][p], ids[peer][a], z[0], z[1], z[2],
                                                             " ".join("%.17g" % x for x in ir_info)))
            f.write("WANT %d %d %d %s\n" % (r, peer, len(theirs[r]), " ".join(str(ids[peer][a]) for a in theirs[r])))
    import ref_frontend
    lines, _ = ref_frontend.run_driver(driver, [path])

    def local_graph(r):
        """robot r's graph as the oracle sees it: own vertices, then the peer copies; own edges + closures"""
        peer = 1 - r
        idx = {vid: k for k, vid in enumerate(ids[r])}
        for j, a in enumerate(theirs[r]):
            idx[ids[peer][a]] = 1000 + j
        poses = np.vstack([g[r]["poses0"], copies[r]])
        e_ij = np.array([list(e) for e in g[r]["edge_ij"]] +
                        [[idx[ids[r][p]], idx[ids[peer][a]]] for p, a in zip(mine[r], theirs[r])], dtype=np.int64)
        meas = np.vstack([g[r]["meas"], ir_meas[r]])
        info = np.vstack([g[r]["info"], np.tile(ir_info, (len(mine[r]), 1))])
        return idx, poses, e_ij, meas, info

    def check_star(rows, r, asked_by_peer):
        """the star robot r computed over its vertices the peer asked about, as the peer received it"""
        idx, poses, e_ij, meas, info = local_graph(r)
        seps = sorted(asked_by_peer)
        gauge = po.select_gauge_centroid(poses, seps)
        z, om, vs = po.condensed_star(poses, e_ij, meas, info, gauge, seps)
        assert [int(x[0]) for x in rows] == [ids[r][gauge]] * len(vs)
        assert [int(x[1]) for x in rows] == [ids[r][v] for v in vs]
        dz = rows[:, 2:5] - z
        dz[:, 2] = po.normalize_theta(dz[:, 2])
        assert np.abs(dz).max() < 1e-6
        got_om = rows[:, 5:].reshape(-1, 3, 3)
        assert np.abs(got_om - om).max() < 1e-6 * np.abs(om).max()
        return rows[:, 2:5], got_om

    def check_poses(tag, r, star_rows):
        """robot r after optimize(5): its local graph + the peer's star between the peer copies"""
        idx, poses, e_ij, meas, info = local_graph(r)
        z, om = star_rows[:, 2:5], star_rows[:, 5:].reshape(-1, 3, 3)
        e2 = np.vstack([e_ij, np.array([[idx[int(x[0])], idx[int(x[1])]] for x in star_rows], dtype=np.int64)])
        m2 = np.vstack([meas, z])
        i2 = np.vstack([info, np.array([[o[0, 0], o[0, 1], o[0, 2], o[1, 1], o[1, 2], o[2, 2]] for o in om])])
        ref = po.gauss_newton(poses, e2, m2, i2, [0], 5)
        got_p = {int(ln.split()[1]): [float(x) for x in ln.split()[2:]] for ln in lines if ln.startswith(tag + " ")}
        got = np.array([got_p[vid] for vid in sorted(idx, key=idx.get)])
        d = got - ref.poses
        d[:, 2] = po.normalize_theta(d[:, 2])
        assert np.abs(d).max() < 1e-6, np.abs(d).max()

    rows_c = np.array([[float(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("C ")])
    rows_d = np.array([[float(x) for x in ln.split()[1:]] for ln in lines if ln.startswith("D ")])
    assert len(rows_c) == 5 and len(rows_d) == 5
    check_star(rows_c, 0, theirs[1])       # A's star over the A-vertices B asked about, as B got it
    check_star(rows_d, 1, theirs[0])       # B's star over the B-vertices A asked about, as A got it
    check_poses("P", 1, rows_c)
    check_poses("Q", 0, rows_d)

