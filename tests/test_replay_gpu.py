"""BASELINE config 0/1 as a parity case: keyframes of robot_0 and of robot_1 from the reference's
own 2robots-hospital bag (tests/golden/bag_2robots_robot{0,1}_kf160.npz, made by
tools/extract_bag_keyframes.py) replayed through the GPU-backed GraphSLAM mirror
(include/cgm/graph_slam.hpp: addDataSM -> findConstraints -> optimize(5), the loop of
src/srslam.cpp:190-221 with its default parameters) and through the CPU oracle pipeline
(oracle/graph_slam_oracle.py). Every decision must agree -- which edges are scan-matched, which
vertex pairs become loop-closure candidates, which closures the vote accepts (closure vertex
indices: exact) -- and measurements / final poses must agree to 1e-6 (north_star)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "cg_mrslam_b200", "lib")
FIXTURES = [os.path.join(ROOT, "tests", "golden", "bag_2robots_robot%d_kf160.npz" % r) for r in range(2)]
LASER_POSE = (0.05, 0.0, 0.0)      # base_link -> base_laser_link in the bag's /tf
TOL = 1e-6


@pytest.fixture(scope="module")
def replay_exe(tmp_path_factory):
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(LIBDIR, "libcgmrslam_b200.so")):
        g.build()
    exe = str(tmp_path_factory.mktemp("cpp") / "srslam_replay")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror",
                           "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "srslam_replay.cpp"), "-o", exe,
                           "-L" + LIBDIR, "-lcgmrslam_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def run_mirror(exe, path, save, robot):
    out = subprocess.run([exe, path, save, str(robot)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = out.stdout.splitlines()
    assert "BEGIN" in lines and lines[-1] == "END", out.stdout[-2000:] + out.stderr[-2000:]
    frames, poses = [], {}
    for ln in lines[lines.index("BEGIN") + 1:-1]:
        tok = ln.split()
        if tok[0] == "K":
            frames.append([])
        elif tok[0] == "EV":
            frames[-1].append((tok[1], int(tok[2]), int(tok[3]), np.array([float(x) for x in tok[4:7]])))
        elif tok[0] == "P":
            poses[int(tok[1])] = np.array([float(x) for x in tok[2:5]])
    return frames, poses


def run_oracle(fx, n, min_inliers, oracle_lib, robot):
    from oracle import bindings
    from oracle.graph_slam_oracle import GraphSlamOracle
    lib = bindings.MatcherLib("reference") if bindings.have_reference() else oracle_lib
    geom = (float(fx["first_angle"]), float(fx["angular_step"]), float(fx["max_range"]))
    gs = GraphSlamOracle(lib, oracle_lib, geom, LASER_POSE, min_inliers=min_inliers, id_robot=robot)
    frames = []
    for k in range(n):
        ranges = fx["ranges"][k].astype(np.float64)
        if k == 0:
            gs.set_initial_data(fx["odom"][k], ranges)
        else:
            gs.add_data_sm(fx["odom"][k], ranges)
            gs.find_constraints()
            gs.optimize(5)
        frames.append(gs.take_events())
    return frames, {v: d["pose"] for v, d in gs.vertices.items()}


def angle_diff(a, b):
    return (a - b + np.pi) % (2 * np.pi) - np.pi


@pytest.mark.gpu
@pytest.mark.parametrize("robot,n,min_inliers", [(0, 60, 7), (0, 160, 7), (1, 160, 7), (1, 160, 4)])
def test_bag_replay_matches_oracle(replay_exe, oracle_lib, tmp_path, robot, n, min_inliers):
    """robot_0, 60 keyframes: the outbound leg (odometry refinement + close edges). 160 keyframes:
    through the first revisit (keyframe 118 closes on keyframe 56); the vote accepts its first seven
    closures at keyframe 127 with the reference's default quorum of 7 inliers. robot_1 (vertex ids
    10000 + k): 59 loop-closure candidates from keyframe 110 on; the default quorum accepts none of
    them within 160 keyframes, a quorum of 4 accepts its first four at keyframe 127."""
    fx = np.load(FIXTURES[robot])
    path = str(tmp_path / "kf.txt")
    with open(path, "w") as f:
        f.write("%d %.17g %.17g %.17g %.17g %.17g %.17g %d\n" % (
            fx["ranges"].shape[1], float(fx["first_angle"]), float(fx["angular_step"]),
            float(fx["max_range"]), LASER_POSE[0], LASER_POSE[1], LASER_POSE[2], min_inliers))
        for k in range(n):
            f.write("%.17g %.17g %.17g " % tuple(fx["odom"][k]))
            f.write(" ".join("%.9g" % r for r in fx["ranges"][k]) + "\n")
    g2o_path = str(tmp_path / ("robot-%d.g2o" % robot))
    got_frames, got_poses = run_mirror(replay_exe, path, g2o_path, robot)
    want_frames, want_poses = run_oracle(fx, n, min_inliers, oracle_lib, robot)
    assert len(got_frames) == len(want_frames) == n
    kinds = {}
    for k, (g, w) in enumerate(zip(got_frames, want_frames)):
        assert [(e[0], e[1], e[2]) for e in g] == [(e[0], e[1], e[2]) for e in w], (k, g, w)
        for eg, ew in zip(g, w):
            d = eg[3] - ew[3]
            d[2] = angle_diff(eg[3][2], ew[3][2])
            assert np.abs(d).max() < TOL, (k, eg, ew)
            kinds[eg[0]] = kinds.get(eg[0], 0) + 1
    assert sorted(got_poses) == sorted(want_poses)
    worst = 0.0
    for v in got_poses:
        d = got_poses[v] - want_poses[v]
        d[2] = angle_diff(got_poses[v][2], want_poses[v][2])
        worst = max(worst, float(np.abs(d).max()))
    assert worst < TOL, worst
    assert kinds.get("S", 0) > n // 2            # odometry edges refined by the matcher
    if n > 120:
        assert kinds.get("L", 0) > 0, kinds                               # loop-closure candidates
        assert (kinds.get("A", 0) > 0) == ((robot, min_inliers) != (1, 7)), kinds   # accepted closures
    print("replay robot", robot, n, "keyframes:", kinds, "max |pose - oracle| = %.2e" % worst)
    if n > 120 and robot == 0:
        condensed_graph_on_replayed_graph(g2o_path, got_poses)


def condensed_graph_on_replayed_graph(g2o_path, mirror_poses):
    """BASELINE cfg 1, the per-robot half: the graph the replay saved (g2o text, as saveGraph
    writes it) is reduced to a condensed star on a set of separator vertices, as
    CondensedGraphBuffer::computeCondensedGraph does for a peer robot
    (condensed_graph_buffer.cpp:437-485): gauge = separator nearest the centroid, optimise with the
    gauge fixed, one labelled EdgeSE2 gauge -> v per separator. GPU solver vs oracle."""
    from cg_mrslam_b200 import pgo
    from oracle import pgo_oracle as po
    g = po.read_g2o(g2o_path)
    ids = list(g["ids"])
    assert sorted(ids) == sorted(mirror_poses)
    for k, v in enumerate(ids):                       # the text round trip keeps 17 digits
        assert np.abs(g["poses"][k] - mirror_poses[v]).max() < 1e-12
    seps = [k for k in range(5, len(ids), 17)]        # vertices a peer robot would have matched
    gauge = po.select_gauge_centroid(g["poses"], seps)
    zr, omr, vs_r = po.condensed_star(g["poses"], g["edge_ij"], g["meas"], g["info"], gauge, seps)
    z, om, vs = pgo.condensed_star(pgo.Solver, g["poses"], g["edge_ij"], g["meas"], g["info"], gauge, seps)
    assert vs == vs_r and len(vs) == len(seps) - 1
    dz = z - zr
    dz[:, 2] = (dz[:, 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(dz).max() < TOL
    assert np.allclose(om, omr, rtol=1e-6, atol=1e-6 * np.abs(omr).max())
