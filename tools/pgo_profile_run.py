"""Runs a few Gauss-Newton iterations of the benchmark graph (GPU needed); meant to be wrapped in
ncu, e.g. the per-launch list of one iteration (every phase is a kernel of its own):

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gn_launches.csv \
        python tools/pgo_profile_run.py 1 16

Arguments: iterations [batch n_vertices n_edges box]. Not a benchmark."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cg_mrslam_b200 import pgo, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nv = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
ne = int(sys.argv[4]) if len(sys.argv) > 4 else 200000
box = float(sys.argv[5]) if len(sys.argv) > 5 else 250.0
g = synth.make_pose_graph(nv, ne, seed=42, box=box, init="truth_noisy")
s = pgo.Solver(batch=batch)
s.set_graph(nv, g["edge_ij"], g["fixed"])
s.upload(g["poses0"], g["meas"], g["info"])
done, chi2 = s.optimize_batch(iters)
st = s.stats()
print("iters", list(done), "chi2[0]", list(chi2[0]), "ms per instance-iteration",
      st["last_iterate_ms"] / max(int(np.sum(done)), 1))
print("stage_ms", st["stage_ms"], "launches", st["kernel_launches"])
s.close()
