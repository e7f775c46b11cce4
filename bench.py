#!/usr/bin/env python
"""Benchmark of the two hot paths on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

One JSON line on stdout (rank 0). A "step" is one pass of the scan-match hot path over one batch
of synthetic scan pairs (BASELINE cfg 5: 1081-beam scans, one 100 x 100 x 101 = 1.01 M candidate
window per pair): reset + rasterise every pair's map grid, rotate/quantise/score every candidate,
compact the per-bin minima. `value` is device throughput with inputs resident in HBM; `e2e` is the
same work through the C-ABI call with host buffers (plan + H2D + kernels + D2H + result assembly).
The GN solve numbers ride in the same line under "gn".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "scan-match scores/sec"
UNIT = "scores/s"
WINDOW_LO = (-5.0, -5.0, -1.25)
WINDOW_HI = (5.0, 5.0, 1.25)
THETA_RES = 0.025
MAX_SCORE = 0.15
BINS = (0.5, 0.5, 0.2)
LC = dict(ll=(-35.0, -35.0), ur=(35.0, 35.0), res=0.1, kernel_range=0.5)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=10000, help="scan pairs per GPU per step")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the CPU sample (0 = auto)")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 global, 2 tiled")
    ap.add_argument("--no-gn", action="store_true")
    ap.add_argument("--workload", default="both", choices=["both", "gn"],
                    help="gn: the GN metric alone as the top-level line (both arms), so that a driver "
                         "can form value(ours) / value(reference) for it")
    ap.add_argument("--e2e-chunks", type=int, default=4,
                    help="handles / host threads the end-to-end step is pipelined over")
    ap.add_argument("--gn-batch", type=int, default=128, help="graph instances per GPU in the GN arm")
    ap.add_argument("--no-bag", action="store_true", help="skip the real-data rows (bag replay)")
    ap.add_argument("--bag-cpu-keyframes", type=int, default=502,
                    help="keyframes of the bag the CPU arm of the real-data rows replays (bounded sample)")
    return ap.parse_args()


def make_pairs(n, rank):
    from cg_mrslam_b200 import synth
    maps, curs = [], []
    for i in range(n):
        p = synth.make_scan_pair(rank * 1000003 + i)
        maps.append(p["map_pts"])
        curs.append(p["cur_pts"])
    return maps, curs


def window_regions(n):
    reg = np.array(list(WINDOW_LO) + list(WINDOW_HI), dtype=np.float32)[None, :]
    return np.repeat(reg, n, axis=0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(key):
    """DRAM traffic and on-chip figures of a kernel from the committed ncu captures
    (profiles/ncu_traffic.json names the capture and the kernel revision); {} if absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key, {})
    except Exception:
        return {}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU legs (the reference's own chargrid.cpp compiled verbatim when oracle/_ref exists, else the
# restated oracle). One pair per host thread: the reference itself uses min(#regions, 4) OpenMP
# threads, i.e. one thread for this one-region window (chargrid.cpp:223-227).
# ------------------------------------------------------------------------------------------------
def cpu_run(n_pairs, n_threads, repeat=1):
    from oracle import bindings
    kind = "reference" if bindings.have_reference() else "port"
    lib = bindings.MatcherLib("reference" if kind == "reference" else "oracle")
    orc = bindings.MatcherLib("oracle")
    stamp = orc.make_stamp(LC["res"], LC["kernel_range"])
    maps, curs = make_pairs(n_pairs, 0)
    grids = []
    for mp in maps:
        g = lib.grid(LC["ll"], LC["ur"], LC["res"])
        g.fill(64)
        g.raster(mp, stamp)
        grids.append(g)
    reg = window_regions(1)
    counts = [0] * n_pairs

    def work(tid):
        for i in range(tid, n_pairs, n_threads):
            g = grids[i]
            g.fill(64)
            g.raster(maps[i], stamp)
            r = g.greedy_search(curs[i], reg, (0.1, 0.1, THETA_RES), MAX_SCORE, BINS)
            counts[i] = len(r)

    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        ts = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        times.append(time.perf_counter() - t0)
    cand = 100 * 100 * 101 * n_pairs
    return kind, cand, times


WORKLOAD = ("cfg5 scan matcher: %d pairs per GPU per step, 1081-beam scans, one 100x100x101 = "
            "1.01M-candidate window per pair, LC grid 700x700 res 0.1, raster + score + compact")


def gn_reference_line(args):
    blk, _, _ = gn_cpu()
    return {"impl": "reference", "metric": "GN iters/sec (50k-node SE2 graph)", "value": blk["value"],
            "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / blk["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg4 synthetic Manhattan graph, %d vertices / %d edges, seed 42" % (GN_V, GN_E)},
            "cpu_baseline": blk,
            "e2e": {"value": blk["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "gn":
        print(json.dumps(gn_reference_line(args)))
        return
    cores = os.cpu_count() or 1
    n_threads = min(cores, 64)
    n_pairs = args.cpu_pairs or n_threads * 2
    total = args.steps + args.warmup
    kind, cand, times = cpu_run(n_pairs, n_threads, repeat=total)
    timed = times[args.warmup:]
    sec = sum(timed)
    value = cand * len(timed) / sec
    sample = "%d pairs x 1.01M candidates per step, one pair per thread" % n_pairs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(timed),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32",
        "data": "synthetic",
        # the GPU arm's workload; each CPU step is a bounded sample of it (sample_pairs_per_step)
        "config": {"workload": WORKLOAD % args.pairs, "pairs_per_gpu": args.pairs,
                   "sample_pairs_per_step": n_pairs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_gn:
        blk, _, _ = gn_cpu()
        line["gn"] = {"impl": "reference", "metric": "GN iters/sec (50k-node SE2 graph)", **blk}
    if not args.no_bag:
        rd = real_data(args, with_gpu=False)
        if rd is not None:
            line["real_data"] = rd
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GN solve (BASELINE cfg 4: 50 k vertices / 200 k edges), reported under "gn" in the same line
# ------------------------------------------------------------------------------------------------
GN_V, GN_E = 50000, 200000


def gn_graph(n_vertices=GN_V, n_edges=GN_E):
    from cg_mrslam_b200 import synth
    # truth + noise as the starting estimate: plain Gauss-Newton (no damping, graph_slam.cpp:54)
    # does not survive 50 k steps of dead-reckoning drift, and the timing does not depend on it
    return synth.make_pose_graph(n_vertices, n_edges, seed=42, box=250.0 * math.sqrt(n_vertices / 50000.0),
                                 init="truth_noisy")


def gn_bytes(st):
    """Algorithmic HBM bytes of one GN iteration (direct form): linearise + assemble
    152 E + 120 V, factorisation 2 x nnz(L) x 8 B, the two triangular solves 2 x nnz(L) x 8 B,
    with nnz(L) = 9 scalars per 3x3 factor block."""
    E, V, nnzb = st["n_edges"], st["n_vertices"], st["factor_blocks"]
    return 152 * E + 120 * V + 4 * 72 * nnzb


def gn_ours(args, local, world, barrier):
    import torch
    from cg_mrslam_b200 import pgo
    g = gn_graph()
    nv = len(g["poses0"])
    B = max(1, args.gn_batch)
    # B instances of the graph: same structure, measurements re-drawn around the first instance's
    rng = np.random.default_rng(4242 + local)
    sig = np.array([0.01, 0.01, 0.002])
    inst_meas = [g["meas"]] + [g["meas"] + rng.normal(size=g["meas"].shape) * sig for _ in range(B - 1)]
    poses0 = torch.from_numpy(g["poses0"]).pin_memory().numpy()
    info = torch.from_numpy(g["info"]).pin_memory().numpy()
    inst_meas = [torch.from_numpy(np.ascontiguousarray(m)).pin_memory().numpy() for m in inst_meas]

    # ---- single graph: the latency of one optimize() call ------------------------------------------
    s1 = pgo.Solver(device=local)
    t0 = time.perf_counter()
    s1.set_graph(nv, g["edge_ij"], g["fixed"])
    analyse_s = time.perf_counter() - t0
    # parity scenario: PARITY_ITERS iterations from the far start, checked against the CPU arm below
    far = gn_far_start()
    s1.upload(far["poses0"], far["meas"], far["info"])
    par_done, par_chi2, par_poses = s1.optimize(PARITY_ITERS)
    s1.upload(poses0, inst_meas[0], info)
    _, chi2_warm, _ = s1.optimize(args.warmup, want_poses=False)
    barrier()
    done1, chi2_1, _ = s1.optimize(args.steps, want_poses=False)
    st1 = s1.stats()
    single_ms = st1["last_iterate_ms"] / max(done1, 1)
    s1.close()

    # ---- batch: B instances per GPU served by the same kernels (the throughput figure) ----------------
    s = pgo.Solver(device=local, batch=B)
    s.set_graph(nv, g["edge_ij"], g["fixed"])
    for b in range(B):
        s.upload_instance(b, poses0, inst_meas[b], info)
    s.optimize_batch(args.warmup)
    barrier()
    done, chi2 = s.optimize_batch(args.steps)
    st = s.stats()
    dev_ms = st["last_iterate_ms"]
    iters = int(done.sum())
    # end to end: every step uploads estimates + measurements of every instance, runs one iteration
    # and reads all estimates back. The instances are served by K solver handles (B / K instances,
    # one stream and one host thread each): the uploads of handle k+1 overlap the kernels of handle k.
    s.close()
    K = max(1, min(args.e2e_chunks, B))
    bounds = [B * k // K for k in range(K + 1)]
    parts = []
    for k in range(K):
        sk = pgo.Solver(device=local, batch=bounds[k + 1] - bounds[k])
        sk.set_graph(nv, g["edge_ij"], g["fixed"])
        parts.append(sk)
    outs = [torch.empty((nv, 3), dtype=torch.float64).pin_memory().numpy() for _ in range(B)]
    up_lock = threading.Lock()
    e2e_done = np.zeros(B, dtype=np.int32)

    def part_step(k):
        sk = parts[k]
        with up_lock:
            for b in range(bounds[k], bounds[k + 1]):
                sk.upload_instance(b - bounds[k], poses0, inst_meas[b], info)
        d, _ = sk.optimize_batch(1)
        e2e_done[bounds[k]:bounds[k + 1]] = d
        for b in range(bounds[k], bounds[k + 1]):
            sk.poses_of(b - bounds[k], out=outs[b])

    times = []
    for it in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=part_step, args=(k,)) for k in range(K)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        t1 = time.perf_counter()
        if it >= args.warmup:
            times.append(t1 - t0)
    assert int(e2e_done.sum()) == B, e2e_done
    for sk in parts:
        sk.close()
    t = torch.tensor([dev_ms, sum(times) * 1e3, single_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pk, pk_kind = peaks()
    nbytes = gn_bytes(st)
    per_iter_ms = float(t[0]) / max(iters, 1)     # per instance-iteration
    achieved = nbytes / (per_iter_ms * 1e-3) / 1e9
    launches_per_iter = (st["kernel_launches"]) // max(1, 2 * args.warmup + 2 * args.steps)
    out = {
        "metric": "GN iters/sec (50k-node SE2 graph)", "unit": "iters/s",
        "value": world * iters / (float(t[0]) * 1e-3),
        "ms_per_iter": per_iter_ms, "iters_done": iters,
        "scaling": "weak (%d graph instances per GPU)" % B, "dtype": "f64",
        "single_graph": {"iters_per_s": 1e3 / float(t[2]), "ms_per_iter": float(t[2]),
                         "stage_ms_last_iter": st1["stage_ms"],
                         "note": "one graph alone: about a hundred dependent panel levels, latency-bound"},
        "config": {"workload": "cfg4 synthetic Manhattan graph, %d vertices / %d edges, seed 42, "
                               "truth+noise start, vertex 0 fixed; %d instances per GPU with one "
                               "structure and re-drawn measurements, every kernel of an iteration "
                               "serves all instances" % (st["n_vertices"], st["n_edges"], B),
                   "batch": B, "factor_blocks": st["factor_blocks"], "elimination_tree_height": st["n_levels"],
                   "supernodes": st["n_supernodes"], "panels": st["n_panels"],
                   "panel_levels": st["n_panel_levels"], "supernode_levels": st["n_supernode_levels"],
                   "analyse_seconds": analyse_s, "stage_ms_last_iter": st["stage_ms"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"],
                     # DRAM bytes per instance-iteration summed over the kernels of one iteration
                     "traffic": ncu_traffic("gn_iteration").get("dram_bytes_per_instance_iteration"),
                     "traffic_source": ncu_traffic("gn_iteration").get("source"), "peak_kind": pk_kind,
                     "kernel": "iteration graph (sn_k_factor / sn_k_update / sn_k_bwd_* + linearise)",
                     "algorithmic_bytes": nbytes,
                     "fp64": {"flops_per_iteration": st.get("factor_flops", 0.0),
                              "achieved_tflops": st.get("factor_flops", 0.0) / (per_iter_ms * 1e-3) / 1e12,
                              "peak_tflops": 37.1,
                              "peak_kind": "measured on this part: profiles/r01_ubench_fp64_rate.txt"},
                     "note": "bytes per instance-iteration = 152 E + 120 V + 4 x 72 B x factor blocks "
                             "(SURVEY 8d direct form); the factorisation itself is bound by fp64 "
                             "issue and dependent-level latency, not by HBM"},
        "e2e": {"value": world * args.steps * B / (float(t[1]) * 1e-3), "unit": "iters/s",
                "h2d_bytes_per_step": int(B * (poses0.nbytes + inst_meas[0].nbytes + info.nbytes)),
                "d2h_bytes_per_step": int(B * poses0.nbytes), "chunks": K},
        "chi2_first_last": [float(chi2[0, 0]), float(chi2[0, -1])] if chi2.size else None,
        "gpu_launches": int(launches_per_iter * args.steps),
    }
    out["_parity"] = (par_done, par_chi2, par_poses)
    if world > 1:
        out["dd"] = gn_domain_decomposed(args, g, local, world, barrier, poses0, inst_meas[0], info,
                                         float(chi2_warm[0]) if len(chi2_warm) else None)
    return out


def gn_domain_decomposed(args, g, local, world, barrier, poses0, meas, info, chi2_single):
    """ONE cfg-4 graph cut over all ranks: per iteration each rank eliminates its interior, one
    NCCL all-reduce sums the separator Schur blocks, the separators are solved replicated."""
    import torch
    import torch.distributed as dist
    from cg_mrslam_b200 import pgo
    rank = dist.get_rank()
    s = pgo.Solver(device=local)

    def bcast(b):
        t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())

    # the solver's own NCCL communicator: the loop (local stage, ncclAllReduce of the separator
    # blocks, shared stage) runs inside the C ABI (pgo_dd_iterate), no Python between the stages
    s.comm_init(rank, world, bcast)
    s.set_graph(len(g["poses0"]), g["edge_ij"], g["fixed"])
    res = {}
    for phase, iters in (("warmup", args.warmup), ("timed", args.steps)):
        s.upload(poses0, meas, info)
        barrier()
        done, chi2 = s.optimize_dd(iters)
        barrier()
        res[phase] = (done, chi2, s.stats()["last_iterate_ms"])
    done, chi2, ms = res["timed"]
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    import ctypes as C
    ptr, n = C.c_void_p(), C.c_int64()
    s.lib.pgo_dd_exchange_buffer(s.h, C.byref(ptr), C.byref(n))
    owner, st = pgo.analyse_partition(len(g["poses0"]), g["edge_ij"], g["fixed"], world)
    out = {"what": "one cfg-4 graph over %d GPUs, one NCCL all-reduce per iteration" % world,
           "scaling": "strong", "iters_done": done, "ms_per_iter": float(t[0]) / max(done, 1),
           "value": done / (float(t[0]) * 1e-3), "unit": "iters/s",
           "allreduce_bytes_per_iter": int(n.value) * 8, "shared_vertices": st["shared_vertices"],
           "rank_updates": st["rank_updates"], "shared_updates": st["shared_updates"],
           "chi2_first": float(chi2[0]) if len(chi2) else None,
           "chi2_first_single_gpu": chi2_single}
    s.close()
    return out


PARITY_ITERS = 2


def gn_far_start():
    """The cfg-4 graph from a NON-converged start (5x the start noise): the parity scenario, and
    what the CPU arm times (the cost of an iteration does not depend on the estimate)."""
    from cg_mrslam_b200 import synth
    return synth.make_pose_graph(GN_V, GN_E, seed=42, box=250.0, init="truth_noisy", start_noise=5.0)


def gn_cpu(n_iters=PARITY_ITERS):
    """The CPU arm of the GN metric: oracle/pgo_oracle_c.cpp, a compiled single-thread restatement
    of g2o's stack for this path (per-edge linearisation, scalar CCS, fill-reducing ordering,
    up-looking sparse Cholesky as in CSparse, two triangular solves). g2o itself cannot be built
    here. Returns (json block, poses after n_iters, chi2)."""
    from oracle import bindings
    g = gn_far_start()
    r = bindings.CGaussNewton().gauss_newton(g["poses0"], g["edge_ij"], g["meas"], g["info"], g["fixed"],
                                             n_iters)
    t = r["times"]
    per_iter = (t["linearise"] + t["factor"] + t["solve"]) / max(r["iterations"], 1)
    blk = {"value": 1.0 / per_iter, "unit": "iters/s", "cores": 1, "kind": "port-c++",
           "what": "compiled restatement of g2o + LinearSolverCSparse (not g2o itself): minimum-degree "
                   "ordering, up-looking scalar sparse Cholesky, one thread",
           "sample": "%d full GN iterations of the same 50k/200k graph: linearise %.2f s, factorise "
                     "%.2f s, solve %.2f s per iteration; ordering + symbolic analysis %.2f s once, "
                     "excluded here as on the GPU side (analyse_seconds)" %
                     (r["iterations"], t["linearise"] / n_iters, t["factor"] / n_iters,
                      t["solve"] / n_iters, t["analyse"]),
           "analyse_seconds": t["analyse"], "nnz_l_scalars": r["nnz_l"],
           "value_with_analysis": r["iterations"] / (per_iter * r["iterations"] + t["analyse"])}
    return blk, r["poses"], r["chi2"]


# ------------------------------------------------------------------------------------------------
# Real data (BASELINE cfgs 0-2): what the product actually calls, per keyframe
# ------------------------------------------------------------------------------------------------
def bag_rows(kind, n_keyframes, device=0):
    """Replays robot_0 of the reference's 2robots-hospital bag through the REFERENCE'S OWN GraphSLAM
    (src/slam + src/matcher/scan_matcher.cpp compiled verbatim, tests/cpp/ref_replay.cpp;
    srslam.cpp:190-221): kind "gpu" = over include/cgm/chargrid.hpp + include/g2o_compat +
    libcgmrslam_b200.so, kind "cpu" = over the reference's chargrid.cpp and the CPU oracle solver.
    Returns per-keyframe latencies in ms, averaged over windows of the keyframe index."""
    import tempfile
    import replay_util
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_replay_" + kind)
    fixture = os.path.join(ROOT, "tests", "golden", "bag_2robots_robot0_full.npz")
    if not (os.path.exists(exe) and os.path.exists(fixture)):
        return None
    fx = np.load(fixture)
    with tempfile.TemporaryDirectory() as d:
        kf, res = os.path.join(d, "kf.txt"), os.path.join(d, "out.txt")
        replay_util.write_keyframes(kf, fx, n_keyframes)
        t0 = time.perf_counter()
        subprocess.check_call([exe, kf, "-", "0", str(n_keyframes)], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, env=dict(os.environ, CGM_OUT=res, CGM_DEVICE=str(device)))
        wall = time.perf_counter() - t0
        frames, _, _ = replay_util.parse(open(res).read().splitlines())
    ms = np.array([f["ms"] for f in frames[1:]])          # keyframe 0 only initialises
    edges = np.cumsum([len(f["edges"]) for f in frames])[1:]
    rows = []
    for lo, hi in ((100, 200), (400, 500), (780, 880)):
        if hi <= len(ms):
            w = ms[lo:hi]
            rows.append({"keyframes": [lo, hi], "graph_vertices": [lo + 1, hi], "graph_edges": int(edges[hi - 1]),
                         "closeScanMatching_ms": float(w[:, 0].mean()),
                         "findConstraints_ms": float(w[:, 1].mean()),
                         "optimize5_ms": float(w[:, 2].mean()),
                         "worst_keyframe_ms": float(w.sum(axis=1).max())})
    return {"keyframes": len(frames), "wall_s": wall, "total_ms": {"closeScanMatching": float(ms[:, 0].sum()),
            "findConstraints": float(ms[:, 1].sum()), "optimize5": float(ms[:, 2].sum())}, "rows": rows}


def real_data(args, device=0, with_gpu=True):
    out = {"what": "per-keyframe latency of the reference's own keyframe loop (srslam.cpp:190-221) on robot_0 of "
                   "2robots-hospital.bag, 361-beam scans: closeScanMatching = addDataSM (24x24x65 candidates on "
                   "the 1200^2 / 0.025 m grid, scan_matcher.cpp:148-151), findConstraints = optimize(1) + "
                   "covariance gate (marginals) + loop-closure sweeps (10x30x65 x regions x 2, "
                   "scan_matcher.cpp:230-254) + vote, optimize5 = initializeOptimization + optimize(5)",
           "unit": "ms per keyframe (mean over the keyframe window)"}
    if with_gpu:
        out["gpu"] = bag_rows("gpu", 1 << 20, device)
    out["cpu"] = bag_rows("cpu", args.bag_cpu_keyframes)
    if out["cpu"] is not None:
        out["cpu"]["what"] = ("the same reference sources over the reference's CPU matcher (chargrid.cpp, up to 4 "
                              "OpenMP threads) and the compiled CPU solver oracle (not g2o); first %d keyframes"
                              % args.bag_cpu_keyframes)
    if out.get("gpu") is None and out["cpu"] is None:
        return None
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist

    from cg_mrslam_b200 import matcher

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "gn":
        gn = gn_ours(args, local, world, barrier)
        if rank == 0:
            par = gn.pop("_parity")
            line = {"metric": gn["metric"], "value": gn["value"], "unit": gn["unit"], "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": gn["ms_per_iter"] * gn["config"]["batch"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic", "config": gn["config"], "roofline": gn["roofline"], "e2e": gn["e2e"],
                    "gpu_launches": gn["gpu_launches"], "single_graph": gn["single_graph"]}
            if world == 1:
                blk, cpu_poses, _ = gn_cpu()
                line["cpu_baseline"] = blk
                d = par[2] - cpu_poses
                d[:, 2] = (d[:, 2] + math.pi) % (2.0 * math.pi) - math.pi
                line["parity_max_abs"] = float(np.abs(d).max())
            if "dd" in gn:
                line["dd"] = gn["dd"]
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    n = args.pairs
    maps, curs = make_pairs(n, rank)
    map_counts = np.array([len(p) for p in maps], dtype=np.int32)
    cur_counts = np.array([len(p) for p in curs], dtype=np.int32)
    # pinned host buffers (the e2e path copies from these every step)
    map_host = torch.from_numpy(np.concatenate(maps)).pin_memory()
    cur_host = torch.from_numpy(np.concatenate(curs)).pin_memory()
    regions = window_regions(n)
    reg_counts = np.ones(n, dtype=np.int32)
    map_np, cur_np = map_host.numpy(), cur_host.numpy()
    step = (0.1, 0.1, THETA_RES)

    stream = torch.cuda.Stream()
    m = matcher.Matcher(LC["ll"], LC["ur"], LC["res"], LC["kernel_range"], n_slots=n,
                        device=local, stream=stream.cuda_stream)
    m.set_kernel(args.kernel)

    # ---- device-resident arm: stage once, time launches ----------------------------------------
    m.batch_stage(cur_np, cur_counts, regions, reg_counts, step, MAX_SCORE, BINS,
                  map_pts=map_np, map_counts=map_counts)
    launches0 = m.launch_count()
    for _ in range(args.warmup):
        m.batch_launch()
    m.batch_collect(cap=4)  # also validates the launch (error flag) and drains the survivors
    stats = m.batch_stats()
    kms = []
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches1 = m.launch_count()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            m.batch_launch()
        ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches2 = m.launch_count()
    kms.append(m.kernel_ms())  # last step's per-kernel split (events on the same stream)
    res, n_out = m.batch_collect(cap=4)
    found = int((n_out > 0).sum())

    # ---- strong scaling of the same job: BASELINE cfg 5 is 10 k pairs IN TOTAL, so at N GPUs each
    # rank takes n / N of them (device-resident, same kernels; launch and planning overheads do
    # not shrink with the share, which is what this line shows) ------------------------------------
    strong_ms = None
    if world > 1:
        ns = max(1, n // world)
        m.batch_stage(cur_np[:int(np.sum(cur_counts[:ns]))], cur_counts[:ns], regions[:ns], reg_counts[:ns], step,
                      MAX_SCORE, BINS, map_pts=map_np[:int(np.sum(map_counts[:ns]))], map_counts=map_counts[:ns])
        for _ in range(args.warmup):
            m.batch_launch()
        m.batch_collect(cap=4)
        barrier()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(args.steps):
                m.batch_launch()
            ev1.record(stream)
        barrier()
        strong_ms = ev0.elapsed_time(ev1)
        strong_cand = m.batch_stats()["candidates"]
        m.batch_collect(cap=4)

    # ---- end-to-end arm: host buffers in, host results out, every step -------------------------
    # The step's pairs are cut into K chunks, one cgm_matcher handle + stream + host thread each
    # (the reference's own search runs on up to 4 OpenMP threads, chargrid.cpp:223-227): chunk
    # k+1's host planning and H2D copies overlap chunk k's kernels. Every step still stages every
    # pair from the pinned host buffers and collects every result list on the host.
    K = max(1, min(args.e2e_chunks, n))
    bounds = [n * k // K for k in range(K + 1)]
    map_off = np.concatenate([[0], np.cumsum(map_counts)])
    cur_off = np.concatenate([[0], np.cumsum(cur_counts)])
    m.close()
    chunk_m, chunk_streams = [], []
    for k in range(K):
        cs = torch.cuda.Stream()
        cm = matcher.Matcher(LC["ll"], LC["ur"], LC["res"], LC["kernel_range"],
                             n_slots=bounds[k + 1] - bounds[k], device=local, stream=cs.cuda_stream)
        cm.set_kernel(args.kernel)
        chunk_m.append(cm)
        chunk_streams.append(cs)
    n_out = np.zeros(n, dtype=np.int32)
    stage_lock = threading.Lock()

    def chunk_step(k):
        a, b = bounds[k], bounds[k + 1]
        cm = chunk_m[k]
        with stage_lock:   # one chunk at a time on the copy engine
            cm.batch_stage(cur_np[cur_off[a]:cur_off[b]], cur_counts[a:b], regions[a:b], reg_counts[a:b],
                           step, MAX_SCORE, BINS, map_pts=map_np[map_off[a]:map_off[b]],
                           map_counts=map_counts[a:b])
            cm.batch_launch()
        _, no = cm.batch_collect(cap=4)
        n_out[a:b] = no

    e2e_t = []
    h2d = (map_np.nbytes + cur_np.nbytes + regions.nbytes + 2 * map_counts.nbytes)
    for it in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        if K == 1:
            chunk_step(0)
        else:
            ths = [threading.Thread(target=chunk_step, args=(k,)) for k in range(K)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if it >= args.warmup:
            e2e_t.append(t1 - t0)
    e2e_found = int((n_out > 0).sum())
    assert e2e_found == found, (e2e_found, found)   # same answers as the one-handle run
    clocks = sampler.stop() if rank == 0 else None
    d2h = int(n_out.sum()) * 16 + 32
    e2e_sec = sum(e2e_t)

    for cm in chunk_m:
        cm.close()
    gn = None if args.no_gn else gn_ours(args, local, world, barrier)

    cand = stats["candidates"]
    t = torch.tensor([dev_ms, e2e_sec * 1e3, strong_ms or 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    value = cand * world * args.steps / (dev_ms_max * 1e-3)
    e2e_value = cand * world * args.steps / (e2e_ms_max * 1e-3)

    if rank == 0:
        pk, pk_kind = peaks()
        alg_bytes = stats["cell_reads"] + 4 * cand
        score_ms = kms[-1]["score"]
        tr = ncu_traffic("score_tiled")
        achieved = alg_bytes / (score_ms * 1e-3) / 1e9 if score_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD % n,
                "pairs_per_gpu": n, "candidates_per_step_per_gpu": cand,
                "mean_k": stats["cell_reads"] / max(cand, 1),
                "l2": "inputs larger than L2 (%.1f GB of grids per GPU)" % (n * 700 * 704 / 1e9),
                "pairs_with_match": found, "kernel": args.kernel},
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"],
                # ncu dram__bytes_read + write of the same kernel, per pair (profiles/ncu_traffic.json)
                "traffic": (tr.get("dram_bytes_per_pair") * n
                            if args.kernel != 1 and tr.get("dram_bytes_per_pair") else None),
                "on_chip": {"alu_pipe_pct": tr.get("alu_pipe_pct"),
                            "shared_wavefronts_pct": tr.get("shared_wavefronts_pct"),
                            "issue_active_pct": tr.get("issue_active_pct"),
                            "source": tr.get("source"),
                            "note": "the limiter is the shared-memory gather + integer pipe, not HBM"},
                "peak_kind": pk_kind,
                "kernel": "score_tiled" if args.kernel != 1 else "score_global",
                "kernel_ms": score_ms, "algorithmic_bytes": alg_bytes,
                "note": "algorithmic bytes = sum over candidates of k_theta x 1 B + 4 B per score "
                        "(SURVEY 8d); the gathers hit shared memory, not DRAM, so frac may exceed 1",
                "step_split_ms": kms[-1]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / args.steps,
                    "chunks": K,
                    "note": "%d chunks of pairs, one matcher handle + stream + host thread each: host "
                            "planning and H2D of a chunk overlap the kernels of the previous one" % K},
            "gpu_launches": int(launches2 - launches1),
            "clocks": clocks,
        }
        if world > 1:
            line["strong_scaling"] = {
                "what": "the same job with %d pairs IN TOTAL (BASELINE cfg 5), %d per GPU, device-resident" %
                        (n // world * world, n // world),
                "value": strong_cand * world * args.steps / (float(t[2]) * 1e-3), "unit": UNIT,
                "ms_per_step": float(t[2]) / args.steps}
        # CPU baseline on a bounded sample, rank 0 at N=1 only
        if world == 1:
            cores = os.cpu_count() or 1
            n_threads = min(cores, 64)
            kind, ccand, times = cpu_run(args.cpu_pairs or n_threads, n_threads, repeat=1)
            line["cpu_baseline"] = {
                "value": ccand / times[0], "unit": UNIT, "cores": n_threads, "kind": kind,
                "sample": "%d pairs x 1.01M candidates, one pair per thread (%.1f s)" %
                          (args.cpu_pairs or n_threads, times[0])}
            if gn is not None:
                blk, cpu_poses, cpu_chi2 = gn_cpu()
                gn["cpu_baseline"] = blk
                par_done, par_chi2, par_poses = gn.pop("_parity")
                d = par_poses - cpu_poses
                d[:, 2] = (d[:, 2] + math.pi) % (2.0 * math.pi) - math.pi
                gn["parity_max_abs"] = float(np.abs(d).max())
                gn["parity"] = {"what": "%d GN iterations of the cfg-4 graph from a non-converged start "
                                        "(5x start noise), GPU vs the CPU arm; tolerance 1e-6 (north_star)"
                                        % PARITY_ITERS,
                                "iters": [int(par_done), PARITY_ITERS],
                                "chi2_gpu": [float(x) for x in par_chi2],
                                "chi2_cpu": [float(x) for x in cpu_chi2]}
                assert par_done == PARITY_ITERS and gn["parity_max_abs"] <= 1e-6, gn["parity_max_abs"]
            if not args.no_bag:
                rd = real_data(args, local)
                if rd is not None:
                    line["real_data"] = rd
        if gn is not None:
            gn.pop("_parity", None)
        if gn is not None:
            line["gn"] = gn
            line["gpu_launches"] += gn["gpu_launches"] * (args.steps)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner
    # at communicator creation, for one) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
