import sys, numpy as np
sys.path.insert(0, ".")
from cg_mrslam_b200 import pgo, synth
from oracle import pgo_oracle as po
g = synth.make_pose_graph(1500, 6000, seed=3, box=43.0)
s = pgo.Solver(batch=2)
s.set_graph(1500, g["edge_ij"], g["fixed"])
s.upload(g["poses0"], g["meas"], g["info"])
done, chi2 = s.optimize_batch(2)
ref = po.gauss_newton(g["poses0"], g["edge_ij"], g["meas"], g["info"], g["fixed"], 2)
print("done", list(done), "err", float(np.abs(s.poses_of(1) - ref.poses).max()))
c = s.marginals([(5, 5), (700, 700), (1400, 1399)])
print("marg", c[0, 0, 0])
s.close()
# wide supernodes: a dense-ish graph
n = 160
e = [(i, j) for i in range(n) for j in range(i + 1, min(n, i + 90))]
rng = np.random.default_rng(0)
poses = np.cumsum(rng.normal(size=(n, 3)) * [0.3, 0.3, 0.05], axis=0)
ei = np.array(e)
rel = po.se2_mul(po.se2_inv(poses[ei[:, 0]]), poses[ei[:, 1]])
info = np.tile([100.0, 0, 0, 100.0, 0, 1000.0], (len(e), 1))
s = pgo.Solver()
s.set_graph(n, ei, [0])
s.upload(poses + rng.normal(size=poses.shape) * 0.01, rel, info)
done, chi2, out = s.optimize(3)
print("dense done", done, chi2[-1])
s.close()
