"""BASELINE configs 0/1 as parity cases, with the REFERENCE'S OWN front-end in the loop: keyframes of
robot_0 and robot_1 from the reference's 2robots-hospital bag (tests/golden/bag_2robots_robot*_full.npz,
made by tools/extract_bag_keyframes.py: all ~885 keyframes per robot) are replayed through the
reference's GraphSLAM -- src/slam/*.cpp + src/matcher/scan_matcher.cpp compiled verbatim, driven like
src/srslam.cpp:190-221 drives them (tests/cpp/ref_replay.cpp) --

  * over include/cgm/chargrid.hpp + include/g2o_compat + libcgmrslam_b200.so  (ref_replay_gpu), and
  * over the reference's own chargrid.cpp and the CPU oracle solver           (ref_replay_cpu; its
    free-running output is committed as tests/golden/replay_*_cpu.txt.gz, tools/make_golden_replay.py).

The GPU test runs the two in lockstep (the pipeline is chaotic at the float-ulp level, see
tests/cpp/ref_replay.cpp). Every decision must agree -- which edges enter the graph (odometry refined by the matcher or not,
closures the vote accepts), which loop-closure candidates are buffered: vertex indices exact -- and
every measurement / estimate to 1e-6 (north_star)."""
import os
import shutil

import numpy as np
import pytest

import ref_frontend
import replay_util

TOL = 1e-6
FIXTURE = os.path.join(replay_util.ROOT, "tests", "golden", "bag_2robots_robot%d_full.npz")


def _replay(kind, robot, n, tmp_path, save=None):
    exe = ref_frontend.driver_path("ref_replay", kind)
    fx = np.load(FIXTURE % robot)
    path = str(tmp_path / ("kf%d.txt" % robot))
    replay_util.write_keyframes(path, fx, n)
    lines, _ = ref_frontend.run_driver(exe, [path, save or "-", robot, n], timeout=3000)
    return replay_util.parse(lines)


@pytest.mark.parametrize("robot", [0, 1])
def test_cpu_replay_reproduces_golden(robot, tmp_path):
    """The committed golden is what the reference front-end over the CPU oracles produces (first 140
    keyframes re-run here: robot_0 meets its first loop-closure candidates at keyframe 118)."""
    frames, _, _ = _replay("cpu", robot, 140, tmp_path)
    want, _, _ = replay_util.parse(replay_util.load_golden("2robots_robot%d" % robot))
    stats, worst = replay_util.compare(frames, want, 140, 1e-12)
    assert len(frames) == 140 and stats["edges"] >= 139
    if robot == 0:
        assert stats["cands"] > 0


def test_call_trace_mechanism(tmp_path):
    """The lockstep instrument itself, without a GPU: the CPU build leading the CPU build follows
    with zero difference through every solver call; a follower fed DIFFERENT scans makes the same
    sequence of calls on different numbers, and the trace reports the difference."""
    exe = ref_frontend.driver_path("ref_replay", "cpu")
    fx = np.load(FIXTURE % 0)
    path, other, states = str(tmp_path / "kf.txt"), str(tmp_path / "other.txt"), str(tmp_path / "trace.bin")
    replay_util.write_keyframes(path, fx, 60)
    ref_frontend.run_driver(exe, [path, "-", 0, 60, "--dump", states], timeout=600)
    lines, _ = ref_frontend.run_driver(exe, [path, "-", 0, 60, "--follow", states], timeout=600)
    frames, _, times = replay_util.parse(lines)
    assert max(f["follow_diff"] for f in frames[1:]) == 0.0
    assert times["trace"]["optimize"] == (3 * 59, 0.0) and times["trace"]["marginals"][0] == 59
    fy = np.load(FIXTURE % 1)                            # another robot's scans against robot 0's trace
    replay_util.write_keyframes(other, fy, 60)
    lines, _ = ref_frontend.run_driver(exe, [other, "-", 0, 60, "--follow", states], timeout=600)
    frames, _, times = replay_util.parse(lines)
    assert max(f["follow_diff"] for f in frames[1:]) > 1e-3   # reported, not hidden
    assert times["trace"]["optimize"][1] > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("robot", [0, 1])
def test_gpu_replay_matches_reference_pipeline(robot, tmp_path):
    """The whole bag in LOCKSTEP at the granularity of the solver calls (tests/cpp/ref_replay.cpp,
    g2o::trace in include/g2o_compat/g2o_compat.hpp): ~885 keyframes per robot, default quorum
    (7 inliers). The GPU build leads and records the result of every optimize / computeInitialGuess /
    computeMarginals; the CPU build (reference matcher + CPU oracle solver) computes every one of
    these calls from the same inputs and carries on with the leader's numbers. Every decision of every
    keyframe must be identical (vertex indices of new graph edges and of buffered closure
    candidates, to the end of the bag), measurements identical, and every solver result of the
    follower within 1e-6 of the leader's (estimates absolute, covariance blocks relative)."""
    g2o_path = str(tmp_path / ("robot-%d.g2o" % robot))
    states = str(tmp_path / "states.bin")
    fx = np.load(FIXTURE % robot)
    path = str(tmp_path / ("kf%d.txt" % robot))
    replay_util.write_keyframes(path, fx, 1 << 20)
    n = len(fx["odom"])
    lines, _ = ref_frontend.run_driver(ref_frontend.driver_path("ref_replay", "gpu"),
                                       [path, g2o_path, robot, n, "--dump", states], timeout=3000)
    frames, poses, times = replay_util.parse(lines)
    lines, _ = ref_frontend.run_driver(ref_frontend.driver_path("ref_replay", "cpu"),
                                       [path, "-", robot, n, "--follow", states], timeout=3000)
    want, want_poses, cpu_times = replay_util.parse(lines)
    assert len(frames) == len(want) == n > 800
    stats, worst = replay_util.compare(frames, want, n, TOL)
    follow = max(f["follow_diff"] for f in want[1:])
    assert 0.0 <= follow < TOL, follow                  # -1 = the runs parted
    trace = cpu_times["trace"]
    assert trace["optimize"][0] >= 3 * (n - 1) and trace["marginals"][0] > 500, trace
    assert all(w < TOL for _, w in trace.values()), trace
    assert sorted(poses) == sorted(want_poses)
    for v in poses:                                     # the final graphs: the same estimates
        d = poses[v] - want_poses[v]
        assert abs(d[0]) < TOL and abs(d[1]) < TOL and abs(replay_util.angle_diff(poses[v][2], want_poses[v][2])) < TOL
    assert stats["closures"] > 0 and stats["cands"] > 50, stats
    print("replay robot", robot, n, "keyframes:", stats, "max |measurement diff| = %.2e," % worst,
          "solver calls (count, worst difference):", trace, "GPU ms", {k: v for k, v in times.items() if k != "trace"},
          "CPU ms", {k: v for k, v in cpu_times.items() if k != "trace"})   # solver_ms: (ms, calls) inside the solver library
    if robot == 0:
        condensed_graph_on_replayed_graph(g2o_path, poses)


def condensed_graph_on_replayed_graph(g2o_path, replay_poses):
    """BASELINE cfg 1, the per-robot half: the graph the replay saved (g2o text, as saveGraph
    writes it) is reduced to a condensed star on a set of separator vertices, as
    CondensedGraphBuffer::computeCondensedGraph does for a peer robot
    (condensed_graph_buffer.cpp:437-485): gauge = separator nearest the centroid, optimise with the
    gauge fixed, one labelled EdgeSE2 gauge -> v per separator. GPU solver vs oracle."""
    from cg_mrslam_b200 import pgo
    from oracle import pgo_oracle as po
    g = po.read_g2o(g2o_path)
    ids = list(g["ids"])
    assert sorted(ids) == sorted(replay_poses)
    for k, v in enumerate(ids):                       # the text round trip keeps 17 digits
        assert np.abs(g["poses"][k] - replay_poses[v]).max() < 1e-12
    seps = [k for k in range(5, len(ids), 37)]        # vertices a peer robot would have matched
    gauge = po.select_gauge_centroid(g["poses"], seps)
    zr, omr, vs_r = po.condensed_star(g["poses"], g["edge_ij"], g["meas"], g["info"], gauge, seps)
    z, om, vs = pgo.condensed_star(pgo.Solver, g["poses"], g["edge_ij"], g["meas"], g["info"], gauge, seps)
    assert vs == vs_r and len(vs) == len(seps) - 1
    dz = z - zr
    dz[:, 2] = (dz[:, 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(dz).max() < TOL
    assert np.allclose(om, omr, rtol=1e-6, atol=1e-6 * np.abs(omr).max())


def test_reference_sources_compile_verbatim():
    """'Link unchanged': every reference source of the two hot paths' callers compiles from
    /root/reference with zero edits against include/ref_names + include/g2o_compat +
    include/cgm/chargrid.hpp (the recipe is oracle/Makefile, target frontend)."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("needs /root/reference (build-time check; the GPU box runs the prebuilt drivers)")
    import subprocess
    root = replay_util.ROOT
    srcs = ["slam/graph_slam.cpp", "slam/graph_manipulator.cpp", "slam/vertices_finder.cpp",
            "slam/closure_checker.cpp", "slam/closure_buffer.cpp", "matcher/scan_matcher.cpp",
            "mrslam/mr_graph_slam.cpp", "mrslam/mr_closure_buffer.cpp", "mrslam/msg_factory.cpp",
            "mrslam/condensed_graph/condensed_graph_buffer.cpp", "mrslam/condensed_graph/condensed_graph_creator.cpp"]
    for s in srcs:
        subprocess.check_call(["/usr/bin/g++", "-std=gnu++14", "-fsyntax-only", "-w", "-Werror=return-type",
                               "-I" + os.path.join(root, "include", "ref_names"), "-I/root/reference/src",
                               "-include", os.path.join(root, "include", "cgm", "chargrid.hpp"),
                               "/root/reference/src/" + s])
    assert shutil.which("nm")
    out = subprocess.check_output(["nm", "-D", "--undefined-only", ref_frontend.driver_path("ref_replay", "gpu")], text=True)
    assert "cgm_matcher_search" in out and "pgo_iterate" in out     # the reference code calls the C ABI
