"""Self-checks of the pose-graph oracle (oracle/pgo_oracle.py). g2o is absent, so these pin the
restatement against independent mathematics, not against g2o's bits ("parity unpinned")."""
import math

import numpy as np
import pytest

from cg_mrslam_b200 import synth
from oracle import pgo_oracle as po


def _graph(n=120, seed=5, box=12.0):
    return synth.make_pose_graph(n, 4 * n, seed=seed, box=box)


def test_se2_algebra():
    rng = np.random.default_rng(0)
    a = rng.uniform(-3, 3, (50, 3))
    b = rng.uniform(-3, 3, (50, 3))
    ident = po.se2_mul(a, po.se2_inv(a))
    assert np.abs(ident[:, :2]).max() < 1e-12
    assert np.abs(po.normalize_theta(ident[:, 2])).max() < 1e-12
    c = rng.uniform(-3, 3, (50, 3))
    lhs = po.se2_mul(po.se2_mul(a, b), c)
    rhs = po.se2_mul(a, po.se2_mul(b, c))
    d = lhs - rhs
    d[:, 2] = po.normalize_theta(d[:, 2])
    assert np.abs(d).max() < 1e-12
    assert po.normalize_theta(np.array([math.pi]))[0] == -math.pi          # [-pi, pi)
    assert po.normalize_theta(np.array([-math.pi]))[0] == -math.pi
    assert abs(po.normalize_theta(np.array([7.0]))[0] - (7.0 - 2 * math.pi)) < 1e-15


def test_jacobians_match_central_differences():
    g = _graph()
    poses = g["poses0"] + np.random.default_rng(1).normal(0, 0.05, g["poses0"].shape)
    e_ij = g["edge_ij"].astype(np.int64)
    ji, jj = po.edge_jacobians(poses, e_ij, g["meas"])
    h = 1e-6
    for k in range(3):
        for which, jac in ((0, ji), (1, jj)):
            d = np.zeros_like(poses)
            vp, vm = poses.copy(), poses.copy()
            idx = e_ij[:, which]
            # perturb edge by edge to avoid shared vertices interfering
            num = np.zeros((len(e_ij), 3))
            for e in range(0, len(e_ij), 7):
                vp[:] = poses
                vm[:] = poses
                vp[idx[e], k] += h
                vm[idx[e], k] -= h
                ep = po.edge_errors(vp, e_ij[e:e + 1], g["meas"][e:e + 1])[0]
                em = po.edge_errors(vm, e_ij[e:e + 1], g["meas"][e:e + 1])[0]
                de = ep - em
                de[2] = po.normalize_theta(np.array([de[2]]))[0]
                num[e] = de / (2 * h)
                assert np.abs(num[e] - jac[e, :, k]).max() < 1e-6
            del d


def test_gn_against_dense_solve_and_fixed_point():
    g = _graph(150, seed=7)
    e_ij = g["edge_ij"].astype(np.int64)
    hidx = po.hessian_index(len(g["poses0"]), g["fixed"])
    H, b, _ = po.build_system(g["poses0"], e_ij, g["meas"], g["info"], hidx)
    Hd = H.toarray()
    assert np.abs(Hd - Hd.T).max() < 1e-9 * np.abs(Hd).max()
    assert np.linalg.eigvalsh(Hd).min() > 0
    res1 = po.gauss_newton(g["poses0"], e_ij, g["meas"], g["info"], g["fixed"], 1)
    delta = np.linalg.solve(Hd, b).reshape(-1, 3)
    want = g["poses0"].copy()
    want[hidx >= 0] = po.oplus(want[hidx >= 0], delta)
    assert np.abs(res1.poses - want).max() < 1e-9
    # converged graph: b -> 0, chi2 stationary and close to the degrees of freedom
    res = po.gauss_newton(g["poses0"], e_ij, g["meas"], g["info"], g["fixed"], 12)
    assert res.iterations == 12 and res.deltas[-1] < 1e-9
    _, b_end, c_end = po.build_system(res.poses, e_ij, g["meas"], g["info"], hidx)
    assert np.abs(b_end).max() < 1e-6
    dof = 3 * len(e_ij) - 3 * int((hidx >= 0).sum())
    assert 0.6 * dof < c_end < 1.5 * dof
    assert abs(po.chi2(res.poses, e_ij, g["meas"], g["info"]) - c_end) < 1e-9 * c_end


def test_marginals_and_labelling():
    g = _graph(80, seed=9, box=9.0)
    e_ij = g["edge_ij"].astype(np.int64)
    res = po.gauss_newton(g["poses0"], e_ij, g["meas"], g["info"], g["fixed"], 3)
    hidx = po.hessian_index(len(g["poses0"]), g["fixed"])
    cov = np.linalg.inv(res.H.toarray())
    pairs = [(5, 5), (17, 17), (40, 12), (79, 79)]
    got = po.marginals(res, hidx, pairs)
    for (r, c), blk in zip(pairs, got):
        want = cov[3 * hidx[r]:3 * hidx[r] + 3, 3 * hidx[c]:3 * hidx[c] + 3]
        assert np.abs(blk - want).max() < 1e-10 * max(1.0, np.abs(want).max())
    # UT labelling equals first-order propagation (J Sigma J^T)^-1 for these tiny covariances
    pts, wm, wc = po.sample_unscented(got[0])
    assert abs(wm.sum() - 1.0) < 1e-12
    z, om = po.label_star_edge(res.poses[0], res.poses[5], got[0])
    _, jj = po.edge_jacobians(res.poses, np.array([[0, 5]]), z[None, :])
    first_order = np.linalg.inv(jj[0] @ got[0] @ jj[0].T)
    assert np.abs(om - first_order).max() < 1e-4 * np.abs(first_order).max()


def test_initial_guess_and_condensed_star():
    g = _graph(100, seed=11, box=10.0)
    e_ij = g["edge_ij"].astype(np.int64)
    # along the odometry chain the spanning tree from vertex 0 reproduces dead reckoning
    chain = e_ij[:99]
    guess = po.initial_guess(np.zeros_like(g["poses0"]) + g["poses0"][0], chain, g["meas"][:99], [0])
    assert np.abs(guess - g["poses0"]).max() < 1e-9
    seps = [10, 35, 60, 90]
    gauge = po.select_gauge_centroid(g["truth"], seps)
    z, om, vs = po.condensed_star(g["poses0"], e_ij, g["meas"], g["info"], gauge, seps)
    assert len(vs) == 3 and z.shape == (3, 3) and om.shape == (3, 3, 3)
    for m in om:
        assert np.linalg.eigvalsh((m + m.T) / 2).min() > 0


def test_g2o_text_roundtrip(tmp_path):
    g = _graph(40, seed=13, box=6.0)
    path = str(tmp_path / "g.g2o")
    po.write_g2o(path, g["ids"] * 3 + 7, g["poses0"], g["edge_ij"], g["meas"], g["info"], g["fixed"])
    r = po.read_g2o(path)
    assert np.array_equal(r["poses"], g["poses0"]) and np.array_equal(r["meas"], g["meas"])
    assert np.array_equal(r["edge_ij"], g["edge_ij"]) and list(r["fixed"]) == [0]


def test_compiled_oracle_matches_python_oracle():
    """oracle/pgo_oracle_c.cpp (up-looking scalar Cholesky, minimum-degree ordering) against
    oracle/pgo_oracle.py (SuperLU): two independent solvers, same Gauss-Newton trajectory."""
    from cg_mrslam_b200 import synth
    from oracle import bindings
    bindings.build()
    c = bindings.CGaussNewton()
    for n, box, seed in [(30, 5.0, 1), (400, 22.0, 3), (2000, 50.0, 5)]:
        g = synth.make_pose_graph(n, 4 * n, seed=seed, box=box)
        ref = po.gauss_newton(g["poses0"], g["edge_ij"], g["meas"], g["info"], g["fixed"], 5)
        r = c.gauss_newton(g["poses0"], g["edge_ij"], g["meas"], g["info"], g["fixed"], 5)
        d = r["poses"] - ref.poses
        d[:, 2] = po.normalize_theta(d[:, 2])
        assert r["iterations"] == 5 and np.abs(d).max() < 1e-10
        assert np.allclose(r["chi2"], ref.chi2, rtol=1e-10)
    # a singular system (no gauge) stops at the failed factorisation, like optimize() returning 0
    g = synth.make_pose_graph(30, 120, seed=1, box=5.0)
    r = c.gauss_newton(g["poses0"], g["edge_ij"], g["meas"], g["info"], [], 3)
    assert r["iterations"] < 3 or np.all(np.isfinite(r["poses"]))
