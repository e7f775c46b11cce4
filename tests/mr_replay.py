"""Multi-robot bag replay (BASELINE cfgs 2-3) over the REFERENCE'S OWN MRGraphSLAM
(oracle/_ref/libref_robot_{gpu,cpu}.so = tests/cpp/ref_robot.cpp + the reference sources compiled
verbatim): N robots of a hospital bag, each driven like src/cg_mrslam.cpp:206-226, exchanging
ComboMessages and CondensedGraphMessages on the cadence of src/mrslam/graph_comm.cpp:126-154.

The schedule, restated deterministically (the reference runs three threads per robot on wall-clock
time): bag time is cut into periods of 150 ms (graph_comm.cpp:152). Within a period every robot
processes the keyframes that fall into it. At the end of the period robot r may send to robot q iff
their ground-truth positions (as of their latest keyframes) are closer than 5 m (the `sim`
modality, graph_comm.h:48, graph_comm.cpp:66-68): first EVERY robot builds its datagrams
(a ComboMessage if its last vertex changed since it last sent one, then a CondensedGraphMessage per
peer that is owed one), then all datagrams are delivered, ordered by sender. A receiver handles a
datagram at once, with its current last vertex as the reference vertex (graph_comm.cpp:178-211).
Nothing else couples the robots, so the same schedule runs in ONE process (all robots) or with ONE
ROBOT PER RANK under torch.distributed, the datagrams crossing ranks in an all-gather (NCCL on the
GPU box, gloo in the CPU tests) in place of the reference's UDP sockets.

    python tests/mr_replay.py --kind cpu --bag 4robots --robots 4 --keyframes 120 --out /tmp/a.npz
    torchrun --nproc-per-node 4 tests/mr_replay.py --kind gpu --dist ...      # one robot per GPU

Lockstep (see tests/cpp/ref_replay.cpp for why) at the granularity of the solver calls: --dump STEM
records every robot's solver results (STEM.r<robot>.bin, g2o::trace in g2o_compat.hpp), --follow STEM
makes this run compute every call itself, note the difference and carry on with the recorded result.
Test / bench tooling; nothing in the package imports it."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LASER_POSE = (0.05, 0.0, 0.0)
PERIOD = 0.15          # graph_comm.cpp:152
COMM_RANGE = 5.0       # graph_comm.h:48
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def fixture(bag, robot):
    return np.load(os.path.join(ROOT, "tests", "golden", "bag_%s_robot%d_full.npz" % (bag, robot)))


class RobotLib:
    def __init__(self, kind):
        path = os.environ.get("CGM_ROBOT_LIB") or os.path.join(ROOT, "oracle", "_ref", "libref_robot_%s.so" % kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (make -C oracle frontend; needs /root/reference)")
        self.lib = lib = C.CDLL(path)
        lib.robot_create.restype = C.c_void_p
        lib.robot_create.argtypes = [C.c_int] + [C.c_double] * 6 + [C.c_int] + [C.c_double] * 6 + [_dp, C.c_int, C.c_double,
                                                                                                 C.c_int, C.c_int]
        lib.robot_destroy.argtypes = [C.c_void_p]
        lib.robot_keyframe.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _dp]
        lib.robot_last_vertex.restype = C.c_int
        lib.robot_last_vertex.argtypes = [C.c_void_p]
        lib.robot_outgoing.restype = C.c_int
        lib.robot_outgoing.argtypes = [C.c_void_p, _ip, C.c_int, C.c_char_p, C.c_int]
        lib.robot_deliver.restype = C.c_int
        lib.robot_deliver.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.robot_vertices.restype = C.c_int
        lib.robot_vertices.argtypes = [C.c_void_p, _dp, C.c_int]
        lib.robot_edges.restype = C.c_int
        lib.robot_edges.argtypes = [C.c_void_p, _dp, C.c_int]
        lib.robot_trace_open.restype = C.c_int
        lib.robot_trace_open.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.robot_trace_stats.argtypes = [C.c_void_p, _dp]
        lib.robot_trace_close.argtypes = [C.c_void_p]


class Robot:
    def __init__(self, rl, rid, fx, min_inliers, mr):
        self.rl, self.id, self.fx = rl, rid, fx
        gt = fx["gt"]
        first = gt[np.isfinite(gt[:, 0])][0]            # sim modality: start at the ground truth
        r0 = np.ascontiguousarray(fx["ranges"][0], dtype=np.float64)
        od = fx["odom"][0]
        self.h = rl.lib.robot_create(rid, first[0], first[1], first[2], od[0], od[1], od[2], fx["ranges"].shape[1],
                                     float(fx["first_angle"]), float(fx["angular_step"]), float(fx["max_range"]),
                                     LASER_POSE[0], LASER_POSE[1], LASER_POSE[2], r0.ctypes.data_as(_dp), min_inliers,
                                     mr[0], mr[1], mr[2])
        self.k = 1                                        # next keyframe
        self.buf = C.create_string_buffer(1 << 22)

    def keyframe(self):
        od = self.fx["odom"][self.k]
        r = np.ascontiguousarray(self.fx["ranges"][self.k], dtype=np.float64)
        self.rl.lib.robot_keyframe(self.h, od[0], od[1], od[2], r.ctypes.data_as(_dp))
        self.k += 1

    def vertices(self):
        out = np.empty((4096, 4))
        n = self.rl.lib.robot_vertices(self.h, out.ctypes.data_as(_dp), len(out))
        assert n >= 0
        return out[:n].copy()

    def edges(self):
        out = np.empty((16384, 7))
        n = self.rl.lib.robot_edges(self.h, out.ctypes.data_as(_dp), len(out))
        assert n >= 0
        e = out[:n]
        return e[np.lexsort((e[:, 4], e[:, 3], e[:, 2], e[:, 1], e[:, 0]))].copy()

    def outgoing(self, peers):
        p = np.asarray(peers, dtype=np.int32)
        n = self.rl.lib.robot_outgoing(self.h, p.ctypes.data_as(_ip), len(p), self.buf, len(self.buf))
        assert n >= 0, "datagram buffer too small"
        return self.buf.raw[:n]

    def deliver(self, data):
        assert self.rl.lib.robot_deliver(self.h, data, len(data)) == 0

    def trace_open(self, stem, follow):
        path = "%s.r%d.bin" % (stem, self.id)
        assert self.rl.lib.robot_trace_open(self.h, path.encode(), 1 if follow else 0) == 0, path

    def trace_stats(self):
        """(parted_at or -1, {kind: (calls, worst difference)})"""
        out = np.zeros(11)
        self.rl.lib.robot_trace_stats(self.h, out.ctypes.data_as(_dp))
        kinds = ("optimize", "initial_guess", "marginals", "star_meas", "star_info")
        return int(out[0]), {k: (int(out[2 * i + 1]), float(out[2 * i + 2])) for i, k in enumerate(kinds)}


def split_datagrams(blob):
    """[int32 peer][int32 size][bytes]... -> [(peer, bytes)]"""
    out, pos = [], 0
    while pos < len(blob):
        peer, n = np.frombuffer(blob, dtype=np.int32, count=2, offset=pos)
        out.append((int(peer), blob[pos + 8:pos + 8 + int(n)]))
        pos += 8 + int(n)
    return out


def run(kind, bag, n_robots, n_keyframes, min_inliers=7, mr=(0.15, 5, 10), dist=None, dump=None, follow=None,
        log=None):
    """Returns a dict of per-robot records. dist = (rank, world, all_gather_bytes) for one robot per
    rank; dump / follow = file stem of the per-robot solver-call traces (lockstep)."""
    rl = RobotLib(kind)
    fxs = [fixture(bag, r) for r in range(n_robots)]
    mine = range(n_robots) if dist is None else [dist[0]]
    robots = {r: Robot(rl, r, fxs[r], min_inliers, mr) for r in mine}
    for rb in robots.values():
        if dump or follow:
            rb.trace_open(dump or follow, follow is not None)
    n_kf = [min(n_keyframes, len(fx["odom"])) for fx in fxs]
    t0 = min(fx["time"][0] for fx in fxs)
    t1 = max(fx["time"][n - 1] for fx, n in zip(fxs, n_kf))
    done = [1] * n_robots                                  # keyframes processed per robot (all ranks track all)
    rec = {r: {"est": [fxs[r]["gt"][np.isfinite(fxs[r]["gt"][:, 0])][0]], "follow_diff": [0.0], "n_edges": [0]}
           for r in mine}
    msgs = []                                              # (period, sender, receiver, bytes)
    period = 0
    while t0 + period * PERIOD <= t1 + PERIOD:
        t_end = t0 + (period + 1) * PERIOD
        for r in range(n_robots):
            while done[r] < n_kf[r] and fxs[r]["time"][done[r]] < t_end:
                if r in robots:
                    rb = robots[r]
                    rb.keyframe()
                    if follow is not None:                # worst difference of any solver call so far
                        parted, kinds = rb.trace_stats()
                        assert parted < 0, ("the follower's solver calls part from the leader's", r, done[r], parted)
                        rec[r]["follow_diff"].append(max(w for _, w in kinds.values()))
                    v = rb.vertices()
                    rec[r]["est"].append(v[v[:, 0] == rl.lib.robot_last_vertex(rb.h)][0, 1:])
                    rec[r]["n_edges"].append(len(rb.edges()))
                done[r] += 1
        # who may talk to whom at the end of this period (ground truth at the latest keyframes)
        pos = [fxs[r]["gt"][done[r] - 1, :2] for r in range(n_robots)]
        peers = {r: [q for q in range(n_robots) if q != r and np.all(np.isfinite(pos[r])) and
                     np.all(np.isfinite(pos[q])) and np.hypot(*(pos[r] - pos[q])) < COMM_RANGE]
                 for r in range(n_robots)}
        if any(peers.values()):
            out = {r: (robots[r].outgoing(peers[r]) if peers[r] else b"") for r in mine}
            if dist is not None:
                blobs = dist[2](out[dist[0]])              # every rank's blob, by rank
            else:
                blobs = [out[r] for r in range(n_robots)]
            for sender in range(n_robots):
                for peer, data in split_datagrams(blobs[sender]):
                    msgs.append((period, sender, peer, len(data)))
                    if peer in robots:
                        robots[peer].deliver(data)
        period += 1
    result = {"msgs": np.array(msgs, dtype=np.int64).reshape(-1, 4)}
    for r in mine:
        result["est%d" % r] = np.array(rec[r]["est"])
        result["follow_diff%d" % r] = np.array(rec[r]["follow_diff"])
        result["n_edges%d" % r] = np.array(rec[r]["n_edges"])
        result["vertices%d" % r] = robots[r].vertices()
        result["edges%d" % r] = robots[r].edges()
        if dump or follow:
            parted, kinds = robots[r].trace_stats()
            result["trace%d" % r] = np.array([[c, w] for c, w in kinds.values()])   # rows: optimize, initial guess,
            rl.lib.robot_trace_close(robots[r].h)                                    # marginals, star meas, star info
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="cpu", choices=["cpu", "gpu"])
    ap.add_argument("--bag", default="4robots")
    ap.add_argument("--robots", type=int, default=4)
    ap.add_argument("--keyframes", type=int, default=120)
    ap.add_argument("--min-inliers-mr", type=int, default=5)
    ap.add_argument("--out", required=True)
    ap.add_argument("--dump")
    ap.add_argument("--follow")
    ap.add_argument("--dist", action="store_true", help="one robot per rank (torchrun)")
    args = ap.parse_args()
    dist = None
    if args.dist:
        import torch
        import torch.distributed as td
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        assert world == args.robots, "one robot per rank"
        use_cuda = args.kind == "gpu"
        if use_cuda:
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            os.environ["CGM_DEVICE"] = str(local)
        td.init_process_group("nccl" if use_cuda else "gloo")
        dev = torch.device("cuda") if use_cuda else torch.device("cpu")

        def all_gather_bytes(blob):
            n = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
            sizes = [torch.zeros_like(n) for _ in range(world)]
            td.all_gather(sizes, n)
            cap = max(int(s) for s in sizes)
            if cap == 0:
                return [b""] * world
            mine = torch.zeros(cap, dtype=torch.uint8)
            if blob:
                mine[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
            mine = mine.to(dev)
            parts = [torch.empty_like(mine) for _ in range(world)]
            td.all_gather(parts, mine)
            return [bytes(p[:int(s)].cpu().numpy().tobytes()) for p, s in zip(parts, sizes)]

        dist = (rank, world, all_gather_bytes)
    res = run(args.kind, args.bag, args.robots, args.keyframes, mr=(0.15, args.min_inliers_mr, 10), dist=dist,
              dump=args.dump, follow=args.follow)
    out = args.out if dist is None else args.out.replace(".npz", ".rank%d.npz" % dist[0])
    np.savez_compressed(out, **res)
    if dist is not None:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    print("mr_replay", args.kind, "messages", len(res["msgs"]),
          {k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items() if k.startswith("edges")})


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
