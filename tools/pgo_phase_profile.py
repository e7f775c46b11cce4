"""Per-barrier timing of one Gauss-Newton iteration of the supernodal solver (GPU needed).

    python tools/pgo_phase_profile.py [n_vertices n_edges box]

Prints, for every grid-wide barrier of the last iteration, the time since the previous one:
kind 1 = panel factorisation, 2 = outer products, 3 / 4 = forward (triangular / rows),
5 / 6 = backward (rows / triangular). Used to decide what to optimise next; not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cg_mrslam_b200 import pgo, synth  # noqa: E402

pos = [a for a in sys.argv[1:] if not a.startswith("-")]
nv = int(pos[0]) if len(pos) > 0 else 50000
ne = int(pos[1]) if len(pos) > 1 else 200000
box = float(pos[2]) if len(pos) > 2 else 250.0
g = synth.make_pose_graph(nv, ne, seed=42, box=box, init="truth_noisy")
s = pgo.Solver()
s.set_graph(nv, g["edge_ij"], g["fixed"])
s.upload(g["poses0"], g["meas"], g["info"])
s.optimize(3, want_poses=False)
done, chi2, _ = s.optimize(2, want_poses=False)
st = s.stats()
print("iters", done, "chi2", list(chi2), "ms/iter", st["last_iterate_ms"] / max(done, 1))
print("stage_ms", st["stage_ms"])
ticks = s.phase_ticks()
tot = {}
for kind, level, us in ticks:
    tot.setdefault(kind, [0, 0.0])
    tot[kind][0] += 1
    tot[kind][1] += us
names = {1: "factor panels", 2: "outer products", 3: "fwd tri", 4: "fwd rows", 5: "bwd rows", 6: "bwd tri"}
for k in sorted(tot):
    print("%-15s barriers %4d  total %9.1f us  mean %7.2f us" % (names.get(k, k), tot[k][0], tot[k][1],
                                                                tot[k][1] / tot[k][0]))
if "-v" in sys.argv:
    for kind, level, us in ticks:
        print(kind, level, "%.2f" % us)
s.close()
