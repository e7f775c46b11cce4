"""Tuning aid (GPU): device throughput of the benchmark graph with the instances spread over K
solver handles (K streams, K host threads) instead of one handle. Arguments: total_batch K [iters]."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cg_mrslam_b200 import pgo, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = synth.make_pose_graph(50000, 200000, seed=42, box=250.0, init="truth_noisy")
parts = []
for k in range(K):
    s = pgo.Solver(batch=B // K)
    s.set_graph(50000, g["edge_ij"], g["fixed"])
    s.upload(g["poses0"], g["meas"], g["info"])
    s.optimize_batch(2)
    parts.append(s)


def run(s):
    s.optimize_batch(iters)


best = 1e9
for rep in range(3):
    ths = [threading.Thread(target=run, args=(s,)) for s in parts]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    best = min(best, time.perf_counter() - t0)
print("B %d over %d handles: %.3f ms per instance-iteration, %.1f iters/s" %
      (B, K, best * 1e3 / (iters * B), iters * B / best))
for s in parts:
    s.close()
