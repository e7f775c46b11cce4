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
# scan matcher: band rasteriser (bulk shared->global stores), tiled and global scorers, hierarchy
from cg_mrslam_b200 import matcher
rng = np.random.default_rng(1)
for ll, ur, kr in (((-15.0, -15.0), (15.0, 15.0), 0.2), ((-35.0, -35.0), (35.0, 35.0), 0.5)):
    res = 0.025 if kr < 0.3 else 0.1
    m = matcher.Matcher(ll, ur, res, kr, n_slots=3)
    pts = np.column_stack([rng.uniform(ll[0] * 1.05, ur[0] * 1.05, 700), rng.uniform(ll[1] * 1.05, ur[1] * 1.05, 700)])
    m.reset(0); m.raster(pts, 0)
    m.raster_batch([pts[:300], pts[300:], pts[:0]], 0)
    cells = m.download(0)
    scan = pts[:181] * 0.3
    reg = np.array([[-0.3, -0.3, -0.2, 0.3, 0.3, 0.2], [1.0, 1.0, 0.0, 1.4, 1.4, 0.1]], dtype=np.float32)
    r1 = m.greedy_search_res(scan, reg, 0.025, 200.0, (0.5, 0.5, 0.5), slot=0)
    r2 = m.greedy_search(scan, reg, (res, res, 0.025), 200.0, (0.5, 0.5, 0.5), slot=1)
    r3 = m.hierarchical_search(scan, reg, 0.025, 200.0, (0.5, 0.5, 0.5), 3, slot=1)
    outs = m.search_batch([scan, scan * 0.5, scan[:7]], [reg, reg[:1], reg], (res, res, 0.025), 200.0, (0.5, 0.5, 0.5))
    print("matcher", m.rows, m.cols, int(cells.min()), len(r1), len(r2), len(r3), [len(o) for o in outs],
          m.count_points(ll, ur, 0), len(m.search_non_matched(scan, 50.0, 0)))
    m.close()
