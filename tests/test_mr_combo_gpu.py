"""A peer's keyframe (ComboMessage) through the MRGraphSLAM mirror: wire format -> globalMatching on
the GPU against the receiver's map -> per-peer window -> vote -> the accepted inter-robot closure in
the graph -> the request for a condensed graph (mr_graph_slam.cpp:60-329, 564-670). The matcher
decisions and transforms are compared with the CPU matcher oracle fed with what crossed the wire
(float32 estimates, ranges and laser parameters; max range forced to 8 m; no laser offset)."""
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
    """mr_combo_gpu: the reference's own MRGraphSLAM (compiled verbatim) over the CUDA library;
    mr_combo_cpu: the same sources over the reference's CPU matcher and the CPU oracle solver."""
    import ref_frontend
    return ref_frontend.driver_path("mr_combo", request.param)


def two_robots(seed, n=(26, 9)):
    """Two short trajectories in ONE synthetic room, with scans; ids consecutive per robot."""
    rng = np.random.default_rng(seed)
    vert, horiz, size, boxes = synth._room_segments(rng)
    first, step, max_range, nb = -math.pi / 2, math.pi / 360, 8.0, 361
    robots = []
    for r, count in enumerate(n):
        x, y = synth._free_pose(rng, size, boxes, margin=1.5)
        th = rng.uniform(-math.pi, math.pi)
        verts = []
        for k in range(count):
            truth = np.array([x, y, po.normalize_theta(np.array([th]))[0]])
            lp = po.se2_mul(truth, np.array(LASER_POSE))[0]
            ranges = synth.cast_scan(vert, horiz, lp, nb, first, step, max_range, 0.01, rng)
            pose = truth + rng.normal(0, [0.02, 0.02, 0.005])
            verts.append(dict(id=10000 * r + k, pose=pose, ranges=ranges, first_angle=first, step=step,
                              max_range=max_range, laser_pose=LASER_POSE, fixed=(k == 0)))
            th += rng.uniform(-0.25, 0.25)
            nx, ny = x + 0.25 * math.cos(th), y + 0.25 * math.sin(th)
            if 1.0 < nx < size[0] - 1.0 and 1.0 < ny < size[1] - 1.0 and all(
                    not (b[0] - 0.6 < nx < b[2] + 0.6 and b[1] - 0.6 < ny < b[3] + 0.6) for b in boxes):
                x, y = nx, ny
            else:
                th += math.pi / 2
        robots.append(verts)
    return robots


def f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("seed,ref,max_score", [(5, 14, 0.3), (8, 3, 0.3), (11, 20, 0.12), (5, 14, 0.01)])
def test_combo_message_to_accepted_closure(driver, tmp_path, oracle_lib, seed, ref, max_score):
    from oracle import bindings
    from oracle import scan_matcher_oracle as smo
    lib = bindings.MatcherLib("reference") if bindings.have_reference() else oracle_lib
    a, b = two_robots(seed)
    path = str(tmp_path / "combo.txt")
    with open(path, "w") as f:
        for r, verts in enumerate((a, b)):
            for v in verts:
                f.write("V %d %d %.17g %.17g %.17g %d %d %.17g %.17g %.17g %s\n" % (
                    r, v["id"], v["pose"][0], v["pose"][1], v["pose"][2], 1 if v["fixed"] else 0,
                    len(v["ranges"]), v["first_angle"], v["step"], v["max_range"],
                    " ".join("%.17g" % x for x in v["ranges"])))
        f.write("RUN %d 2 1 %.17g\n" % (a[ref]["id"], max_score))
    import ref_frontend
    lines, _ = ref_frontend.run_driver(driver, [path])
    last = b[-1]
    nv, nr = 5, len(last["ranges"])
    assert lines[0].split() == ["MSG", str(8 + 8 + nv * 16 + 4 + 8 + nr * 4 + 16), "vertices", str(nv),
                                "readings", str(nr), "node", str(last["id"])]
    # what robot 0 sees of the peer's keyframe
    peer = dict(id=last["id"], pose=f32(last["pose"]), ranges=f32(last["ranges"]),
                first_angle=float(f32(last["first_angle"])), step=float(f32(last["step"])), max_range=8.0,
                laser_pose=(0.0, 0.0, 0.0), fixed=False)

    def state(tag):
        k = [i for i, ln in enumerate(lines) if ln.startswith(tag)]
        out = []
        for i in k:
            t = lines[i].split()
            cands = []
            for ln in lines[i + 1:]:
                if not ln.startswith("CAND "):
                    break
                c = ln.split()
                cands.append((int(c[1]), int(c[2]), np.array([float(x) for x in c[3:]])))
            out.append(dict(closures=int(t[2]), waiting=int(t[4]), edges=int(t[6]), peers=int(t[8]), cands=cands))
        return out

    arrival = state("ARRIVAL")[0]
    rounds = state("ROUND")
    refset1 = [v for v in a if abs(v["id"] - a[ref]["id"]) <= 10]
    ok1, t1 = smo.global_matching(lib, oracle_lib, refset1, a[ref], peer, max_score)
    matched_from = None
    if ok1:
        assert (arrival["closures"], arrival["waiting"]) == (1, 0)
        frm, to, t = arrival["cands"][0]
        assert (frm, to) == (a[ref]["id"], last["id"]) and np.array_equal(t, t1)
        matched_from = a[ref]["id"]
    else:
        assert (arrival["closures"], arrival["waiting"], arrival["cands"]) == (0, 1, [])
        refset2 = a[-21:]
        ok2, t2 = smo.global_matching(lib, oracle_lib, refset2, a[-1], peer, max_score)
        if ok2:
            # matched by findInterRobotConstraints; by the time the state is printed the window has aged once
            assert (rounds[0]["closures"], rounds[0]["waiting"]) == (1, 0)
            frm, to, t = rounds[0]["cands"][0]
            assert (frm, to) == (a[-1]["id"], last["id"]) and np.array_equal(t, t2)
            matched_from = a[-1]["id"]
    # (seeds 5 and 8 match on arrival, seed 11 only on the retry, max_score 0.01 never)
    ask = lines[[i for i, ln in enumerate(lines) if ln.startswith("ASK ")][0]].split()
    repeat = state("REPEAT")[0]
    if matched_from is not None:
        # round 0: no vote yet (age 0 -> 1); round 1: the window votes, the closure enters the graph
        # and the window drops the vertex; robot 0 then asks robot 1 about it
        assert (rounds[0]["edges"], rounds[0]["peers"]) == (0, 0)
        assert (rounds[1]["closures"], rounds[1]["edges"], rounds[1]["peers"]) == (0, 1, 1)
        assert (rounds[2]["closures"], rounds[2]["edges"], rounds[2]["peers"]) == (0, 1, 1)
        assert ask == ["ASK", "1", str(last["id"])]
        # the same keyframe again: its vertex is in the graph now, nothing is matched twice
        assert (repeat["closures"], repeat["waiting"], repeat["edges"]) == (0, 0, 1)
    else:
        # never matched: it waits two keyframes and is dropped; the repeat starts over
        assert [r["waiting"] for r in rounds] == [1, 0, 0] and all(r["edges"] == 0 for r in rounds)
        assert ask == ["ASK", "0"]
        assert (repeat["closures"], repeat["waiting"]) == (0, 1)
