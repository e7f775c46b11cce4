"""TEST INFRASTRUCTURE ONLY: CPU restatement (numpy + scipy.sparse) of the pose-graph arithmetic
the reference delegates to g2o -- SE2 algebra, EdgeSE2 error and analytic Jacobians, block-Hessian
assembly, Gauss-Newton iterations with a sparse direct solve, marginal covariance blocks,
spanning-tree initial guess, unscented edge labelling and the condensed-graph star construction.

PARITY UNPINNED. g2o (RainerKuemmerle/g2o, commit 4b9c2f5b68d14ad479457b18c5a2a0bce1541a90 per the
reference's README.md:17-19) is a third-party dependency that is NOT in /root/reference and not in
this container, and the reference ships no tests or golden vectors for this path. The formulas
below restate g2o's published algorithm as recorded in SURVEY.md appendix C (items C1-C11), and
are anchored on the reference's own call sites (cited per function). They are self-checked in
tests/test_oracle_pgo.py (numeric Jacobians, an independent dense solve, fixed-point properties),
which shows the oracle is self-consistent, not that it equals g2o's bits.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The linear solver here is scipy's SuperLU -- deliberately a different algorithm from
the product's block Cholesky, so agreement between the two is meaningful.
"""
import math
from collections import deque

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

TWO_PI = 2.0 * math.pi


# ---------------------------------------------------------------------------------------------
# C1-C3: SE2, normalize_theta, VertexSE2::oplusImpl
# ---------------------------------------------------------------------------------------------
def normalize_theta(t):
    """g2o normalize_theta (appendix C2, older form): identity inside [-pi, pi), else wrap."""
    t = np.asarray(t, dtype=np.float64)
    out = t.copy()
    m = ~((t >= -math.pi) & (t < math.pi))
    if np.any(m):
        w = t[m] - TWO_PI * np.floor(t[m] / TWO_PI)  # [0, 2pi)
        w = np.where(w >= math.pi, w - TWO_PI, w)
        out[m] = w
    return out


def se2_mul(a, b):
    """A*B: t = tA + R(thA) tB, th = wrap(thA + thB). Rows (x, y, th). (C1)"""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    b = np.atleast_2d(np.asarray(b, dtype=np.float64))
    c, s = np.cos(a[:, 2]), np.sin(a[:, 2])
    out = np.empty((max(len(a), len(b)), 3))
    out[:, 0] = a[:, 0] + (c * b[:, 0] - s * b[:, 1])
    out[:, 1] = a[:, 1] + (s * b[:, 0] + c * b[:, 1])
    out[:, 2] = normalize_theta(a[:, 2] + b[:, 2])
    return out


def se2_inv(a):
    """A^-1: th' = wrap(-thA), t' = R(th') (-tA). (C1)"""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    th = normalize_theta(-a[:, 2])
    c, s = np.cos(th), np.sin(th)
    out = np.empty_like(a)
    out[:, 0] = c * (-a[:, 0]) - s * (-a[:, 1])
    out[:, 1] = s * (-a[:, 0]) + c * (-a[:, 1])
    out[:, 2] = th
    return out


def oplus(poses, delta):
    """VertexSE2::oplusImpl: t += d_xy (world frame), th = wrap(th + d_th). (C3)"""
    out = np.array(poses, dtype=np.float64, copy=True)
    out[:, :2] += delta[:, :2]
    out[:, 2] = normalize_theta(out[:, 2] + delta[:, 2])
    return out


def info_full(info6):
    """Upper triangle (I11 I12 I13 I22 I23 I33) -> [E, 3, 3] symmetric."""
    info6 = np.atleast_2d(info6)
    m = np.empty((len(info6), 3, 3))
    m[:, 0, 0], m[:, 0, 1], m[:, 0, 2] = info6[:, 0], info6[:, 1], info6[:, 2]
    m[:, 1, 0], m[:, 1, 1], m[:, 1, 2] = info6[:, 1], info6[:, 3], info6[:, 4]
    m[:, 2, 0], m[:, 2, 1], m[:, 2, 2] = info6[:, 2], info6[:, 4], info6[:, 5]
    return m


# ---------------------------------------------------------------------------------------------
# C4: EdgeSE2 error and Jacobians
# ---------------------------------------------------------------------------------------------
def edge_errors(poses, edge_ij, meas):
    """EdgeSE2::computeError: e = toVector(Z^-1 * (Xi^-1 * Xj)). (C4)"""
    xi, xj = poses[edge_ij[:, 0]], poses[edge_ij[:, 1]]
    return se2_mul(se2_inv(meas), se2_mul(se2_inv(xi), xj))


def edge_jacobians(poses, edge_ij, meas):
    """EdgeSE2::linearizeOplus, analytic (C4). Returns Ji, Jj as [E, 3, 3]."""
    xi, xj = poses[edge_ij[:, 0]], poses[edge_ij[:, 1]]
    ci, si = np.cos(xi[:, 2]), np.sin(xi[:, 2])
    dx, dy = xj[:, 0] - xi[:, 0], xj[:, 1] - xi[:, 1]
    n = len(edge_ij)
    ji = np.zeros((n, 3, 3))
    jj = np.zeros((n, 3, 3))
    ji[:, 0, 0], ji[:, 0, 1], ji[:, 0, 2] = -ci, -si, -si * dx + ci * dy
    ji[:, 1, 0], ji[:, 1, 1], ji[:, 1, 2] = si, -ci, -ci * dx - si * dy
    ji[:, 2, 2] = -1.0
    jj[:, 0, 0], jj[:, 0, 1] = ci, si
    jj[:, 1, 0], jj[:, 1, 1] = -si, ci
    jj[:, 2, 2] = 1.0
    zinv = se2_inv(meas)
    cz, sz = np.cos(zinv[:, 2]), np.sin(zinv[:, 2])
    rz = np.zeros((n, 3, 3))
    rz[:, 0, 0], rz[:, 0, 1], rz[:, 1, 0], rz[:, 1, 1], rz[:, 2, 2] = cz, -sz, sz, cz, 1.0
    return rz @ ji, rz @ jj


def chi2(poses, edge_ij, meas, info6):
    e = edge_errors(poses, edge_ij, meas)
    om = info_full(info6)
    return float(np.einsum("ei,eij,ej->", e, om, e))


# ---------------------------------------------------------------------------------------------
# C5-C6: active set, Hessian indices, system build
# ---------------------------------------------------------------------------------------------
def hessian_index(n_vertices, fixed):
    """Non-fixed vertices get 0..n-1 in vertex (= id) order, fixed ones -1. (C6)"""
    mask = np.ones(n_vertices, dtype=bool)
    mask[np.asarray(fixed, dtype=np.int64)] = False
    hidx = np.full(n_vertices, -1, dtype=np.int64)
    hidx[mask] = np.arange(int(mask.sum()))
    return hidx


def build_system(poses, edge_ij, meas, info6, hidx):
    """constructQuadraticForm + buildSystem (C5): returns (H csc [3n,3n] full symmetric, b [3n],
    chi2)."""
    n = int(hidx.max()) + 1 if len(hidx) else 0
    e = edge_errors(poses, edge_ij, meas)
    ji, jj = edge_jacobians(poses, edge_ij, meas)
    om = info_full(info6)
    hi, hj = hidx[edge_ij[:, 0]], hidx[edge_ij[:, 1]]
    omega_e = -np.einsum("eij,ej->ei", om, e)
    b = np.zeros(3 * n)
    rows, cols, vals = [], [], []
    r3 = np.arange(3)

    def add_block(hr, hc, blk, sel):
        if not np.any(sel):
            return
        rr = (3 * hr[sel])[:, None, None] + r3[None, :, None]
        cc = (3 * hc[sel])[:, None, None] + r3[None, None, :]
        rows.append(np.broadcast_to(rr, blk[sel].shape).ravel())
        cols.append(np.broadcast_to(cc, blk[sel].shape).ravel())
        vals.append(blk[sel].ravel())

    fi, fj = hi >= 0, hj >= 0
    jit_om = np.einsum("eji,ejk->eik", ji, om)  # Ji^T Omega
    jjt_om = np.einsum("eji,ejk->eik", jj, om)
    add_block(hi, hi, jit_om @ ji, fi)
    add_block(hj, hj, jjt_om @ jj, fj)
    both = fi & fj
    hij = jit_om @ jj
    add_block(hi, hj, hij, both)
    add_block(hj, hi, np.transpose(hij, (0, 2, 1)), both)
    np.add.at(b, (3 * hi[fi])[:, None] + r3[None, :], np.einsum("eji,ej->ei", ji, omega_e)[fi])
    np.add.at(b, (3 * hj[fj])[:, None] + r3[None, :], np.einsum("eji,ej->ei", jj, omega_e)[fj])
    if rows:
        H = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(3 * n, 3 * n))
    else:
        H = sp.csc_matrix((3 * n, 3 * n))
    c2 = float(np.einsum("ei,eij,ej->", e, om, e))
    return H, b, c2


# ---------------------------------------------------------------------------------------------
# C7-C8: Gauss-Newton
# ---------------------------------------------------------------------------------------------
class GNResult:
    def __init__(self):
        self.poses = None
        self.chi2 = []      # chi2 at the linearisation point of every iteration
        self.deltas = []    # max |delta| per iteration
        self.H = None       # the system of the LAST iteration (what computeMarginals sees, C9)
        self.lu = None
        self.iterations = 0


def gauss_newton(poses0, edge_ij, meas, info6, fixed, n_iters):
    """SparseOptimizer::optimize(n) with OptimizationAlgorithmGaussNewton (C7): exactly n
    iterations, no damping, no termination test; returns iterations done (0 if the solve fails)."""
    poses = np.array(poses0, dtype=np.float64, copy=True)
    edge_ij = np.asarray(edge_ij, dtype=np.int64)
    hidx = hessian_index(len(poses), fixed)
    free = np.nonzero(hidx >= 0)[0]
    out = GNResult()
    for _ in range(n_iters):
        H, b, c2 = build_system(poses, edge_ij, meas, info6, hidx)
        out.chi2.append(c2)
        try:
            lu = spla.splu(H.tocsc(), permc_spec="COLAMD",
                           options=dict(SymmetricMode=True))
            delta = lu.solve(b)
        except RuntimeError:
            break
        if not np.all(np.isfinite(delta)):
            break
        d = delta.reshape(-1, 3)
        poses[free] = oplus(poses[free], d)
        out.deltas.append(float(np.abs(d).max()) if len(d) else 0.0)
        out.H, out.lu = H, lu
        out.iterations += 1
    out.poses = poses
    return out


# ---------------------------------------------------------------------------------------------
# C9: marginal covariance blocks
# ---------------------------------------------------------------------------------------------
def marginals(res, hidx, block_pairs):
    """computeMarginals(spinv, pairs): blocks (r, c) of H^-1 for the H stored by the last
    iteration (C9). block_pairs are VERTEX index pairs; returns [n, 3, 3]."""
    out = np.zeros((len(block_pairs), 3, 3))
    n3 = res.H.shape[0]
    cache = {}
    for k, (r, c) in enumerate(block_pairs):
        hr, hc = int(hidx[r]), int(hidx[c])
        if hc not in cache:
            rhs = np.zeros((n3, 3))
            rhs[3 * hc + np.arange(3), np.arange(3)] = 1.0
            cache[hc] = res.lu.solve(rhs)
        out[k] = cache[hc][3 * hr:3 * hr + 3, :]
    return out


# ---------------------------------------------------------------------------------------------
# C10: spanning-tree initial guess
# ---------------------------------------------------------------------------------------------
def initial_guess(poses0, edge_ij, meas, fixed):
    """computeInitialGuess (C10): breadth-first from the fixed vertices over the active edges
    (uniform cost), each newly reached vertex set by EdgeSE2::initialEstimate from its parent
    (Xj = Xi*Z or Xi = Xj*Z^-1). g2o breaks ties by pointer order; here the rule does not depend on
    the order of the edge list: the vertices of a level are expanded in ascending index, the edges of
    a vertex in ascending (other end, direction, measurement), the first edge to reach a vertex sets
    it."""
    poses = np.array(poses0, dtype=np.float64, copy=True)
    meas = np.asarray(meas, dtype=np.float64)
    n = len(poses)
    adj = [[] for _ in range(n)]
    for e, (i, j) in enumerate(np.asarray(edge_ij)):
        adj[int(i)].append((int(j), 0, float(meas[e][0]), float(meas[e][1]), float(meas[e][2]), e))
        adj[int(j)].append((int(i), 1, float(meas[e][0]), float(meas[e][1]), float(meas[e][2]), e))
    seen = np.zeros(n, dtype=bool)
    level = sorted(int(v) for v in fixed)
    seen[level] = True
    while level:
        nxt = []
        for v in level:
            for w, backward, _, _, _, e in sorted(adj[v], key=lambda t: t[:5]):
                if seen[w]:
                    continue
                if not backward:   # v is Xi, w is Xj
                    poses[w] = se2_mul(poses[v], meas[e])[0]
                else:              # v is Xj, w is Xi
                    poses[w] = se2_mul(poses[v], se2_inv(meas[e]))[0]
                seen[w] = True
                nxt.append(w)
        level = sorted(nxt)
    return poses


# ---------------------------------------------------------------------------------------------
# C11: unscented edge labelling; condensed graph (condensed_graph_creator.cpp:33-66)
# ---------------------------------------------------------------------------------------------
def sample_unscented(cov, alpha=1e-3, beta=2.0):
    """g2o sampleUnscented for mean 0 (C11): 2*dim+1 points, (w_mean, w_cov)."""
    dim = cov.shape[0]
    lam = alpha * alpha * dim
    L = np.linalg.cholesky(cov * (dim + lam))
    pts = [np.zeros(dim)]
    wi = 1.0 / (2.0 * (dim + lam))
    w0m = lam / (dim + lam)
    w0c = w0m + (1.0 - alpha * alpha + beta)
    wm, wc = [w0m], [w0c]
    for i in range(dim):
        pts.append(L[:, i].copy())
        pts.append(-L[:, i])
        wm += [wi, wi]
        wc += [wi, wi]
    return np.array(pts), np.array(wm), np.array(wc)


def label_star_edge(pose_g, pose_v, cov_vv):
    """EdgeLabeler::labelEdge for a star edge gauge -> v with the gauge fixed (C11): measurement =
    Xg^-1 * Xv (setMeasurementFromState); information = inverse of the unscented covariance of the
    edge error under v's marginal."""
    z = se2_mul(se2_inv(pose_g), pose_v)[0]
    pts, wm, wc = sample_unscented(cov_vv)
    errs = []
    for p in pts:
        pv = oplus(pose_v[None, :], p[None, :])[0]
        e = se2_mul(se2_inv(z), se2_mul(se2_inv(pose_g), pv))[0]
        errs.append(e)
    errs = np.array(errs)
    mean = (wm[:, None] * errs).sum(axis=0)
    d = errs - mean
    cov = np.einsum("k,ki,kj->ij", wc, d, d)
    return z, np.linalg.inv(cov)


def condensed_star(poses0, edge_ij, meas, info6, gauge, separators):
    """CondensedGraphCreator::compute (condensed_graph_creator.cpp:33-66) on the given edge set:
    fix only the gauge, computeInitialGuess, optimize(1), then label one star edge gauge -> v per
    separator vertex v != gauge. Returns (meas [m, 3], info [m, 3, 3], v list). The caller's poses
    are left untouched (push/pop in GraphManipulator, graph_manipulator.cpp:62-88)."""
    guess = initial_guess(poses0, edge_ij, meas, [gauge])
    res = gauss_newton(guess, edge_ij, meas, info6, [gauge], 1)
    if res.iterations != 1:
        raise RuntimeError("condensed_star: linear solve failed")
    hidx = hessian_index(len(poses0), [gauge])
    vs = [int(v) for v in separators if int(v) != int(gauge)]
    covs = marginals(res, hidx, [(v, v) for v in vs])
    zs, oms = [], []
    for v, c in zip(vs, covs):
        z, om = label_star_edge(res.poses[gauge], res.poses[v], c)
        zs.append(z)
        oms.append(om)
    return np.array(zs).reshape(-1, 3), np.array(oms).reshape(-1, 3, 3), vs


def select_gauge_centroid(poses, separators):
    """selectGaugeCentroid (condensed_graph_buffer.cpp:318-345): the separator vertex closest to
    the centroid of the separators' translations."""
    sep = np.asarray(separators, dtype=np.int64)
    c = poses[sep, :2].mean(axis=0)
    d = np.linalg.norm(poses[sep, :2] - c, axis=1)
    return int(sep[int(np.argmin(d))])


# ---------------------------------------------------------------------------------------------
# C14: g2o text format
# ---------------------------------------------------------------------------------------------
def write_g2o(path, ids, poses, edge_ij, meas, info6, fixed):
    with open(path, "w") as f:
        for i, p in zip(ids, poses):
            f.write("VERTEX_SE2 %d %.17g %.17g %.17g\n" % (int(i), p[0], p[1], p[2]))
        for v in fixed:
            f.write("FIX %d\n" % int(ids[int(v)]))
        for (a, b), z, w in zip(edge_ij, meas, info6):
            f.write("EDGE_SE2 %d %d %.17g %.17g %.17g %s\n" %
                    (int(ids[int(a)]), int(ids[int(b)]), z[0], z[1], z[2],
                     " ".join("%.17g" % x for x in w)))


def read_g2o(path):
    ids, poses, fixed_ids, ei, ej, meas, info = [], [], [], [], [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "VERTEX_SE2":
                ids.append(int(t[1]))
                poses.append([float(x) for x in t[2:5]])
            elif t[0] == "FIX":
                fixed_ids += [int(x) for x in t[1:]]
            elif t[0] == "EDGE_SE2":
                ei.append(int(t[1]))
                ej.append(int(t[2]))
                meas.append([float(x) for x in t[3:6]])
                info.append([float(x) for x in t[6:12]])
    order = np.argsort(ids, kind="stable")
    ids = np.asarray(ids, dtype=np.int64)[order]
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 3)[order]
    lut = {int(v): k for k, v in enumerate(ids)}
    edge_ij = np.array([[lut[a], lut[b]] for a, b in zip(ei, ej)], dtype=np.int64).reshape(-1, 2)
    fixed = np.array(sorted(lut[v] for v in fixed_ids), dtype=np.int64)
    return dict(ids=ids, poses=poses, edge_ij=edge_ij,
                meas=np.asarray(meas, dtype=np.float64).reshape(-1, 3),
                info=np.asarray(info, dtype=np.float64).reshape(-1, 6), fixed=fixed)
