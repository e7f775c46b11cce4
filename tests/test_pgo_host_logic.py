"""Supernodal factorisation / substitution logic of the pose-graph solver on a GPU-less box.

tests/hostsim/libpgo_hostsim.so (TEST ONLY) instantiates the product's task functions
(cg_mrslam_b200/csrc/pgo_supernodal.h -- the source the CUDA kernels are built from) with a
one-thread group, driven by the product's structure analysis. The result of H x = b must match an
independent sparse solve (scipy SuperLU) on graphs that exercise every table: single vertices,
chains, wide dense supernodes split into several panels, panels with more rows than one CTA task
takes (scratch-published diagonal parts), and the benchmark's Manhattan graphs.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "hostsim", "libpgo_hostsim.so")


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cg_mrslam_b200", "csrc"), "hostsim"])
    lb = ctypes.CDLL(LIB)
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    lb.pgo_hostsim_solve.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, f64p, f64p, f64p, i32p]
    lb.pgo_hostsim_solve.restype = ctypes.c_int
    return lb


def block_system(n, edges, seed):
    """Random SPD block matrix with the sparsity of the graph: H = sum_e J_e^T J_e + diag."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    diag = np.tile(np.eye(3) * 0.5, (n, 1, 1))
    for (i, j) in edges:
        ji, jj = rng.normal(size=(3, 3)), rng.normal(size=(3, 3))
        diag[i] += ji.T @ ji
        diag[j] += jj.T @ jj
        r, c, blk = (i, j, ji.T @ jj) if i > j else (j, i, jj.T @ ji)
        rows.append(r)
        cols.append(c)
        vals.append(blk)
    for v in range(n):
        rows.append(v)
        cols.append(v)
        vals.append(diag[v])
    return (np.array(rows, np.int32), np.array(cols, np.int32),
            np.ascontiguousarray(np.array(vals).reshape(-1, 9)), rng.normal(size=(n, 3)))


def dense_reference(n, rows, cols, vals, rhs):
    r_idx, c_idx, data = [], [], []
    for r, c, v in zip(rows, cols, vals.reshape(-1, 3, 3)):
        for a in range(3):
            for b in range(3):
                r_idx.append(3 * r + a)
                c_idx.append(3 * c + b)
                data.append(v[a, b])
                if r != c:
                    r_idx.append(3 * c + b)
                    c_idx.append(3 * r + a)
                    data.append(v[a, b])
    h = sp.csc_matrix((data, (r_idx, c_idx)), shape=(3 * n, 3 * n))
    return spla.spsolve(h, rhs.reshape(-1)).reshape(n, 3)


def run(lib, n, edges, seed=0):
    rows, cols, vals, rhs = block_system(n, edges, seed)
    xx = np.zeros((2, n, 3))   # [0]: forward fused into the factorisation; [1]: stand-alone forward
    stats = np.zeros(8, np.int32)
    rc = lib.pgo_hostsim_solve(n, len(rows), rows, cols, vals, np.ascontiguousarray(rhs), xx, stats)
    assert rc == 0
    ref = dense_reference(n, rows, cols, vals, rhs)
    scale = max(1.0, float(np.abs(ref).max()))
    for x in xx:
        assert np.abs(x - ref).max() <= 1e-9 * scale, (np.abs(x - ref).max(), stats)
    return stats


def test_tiny_and_chain(lib):
    run(lib, 1, [])
    run(lib, 2, [(0, 1)])
    run(lib, 3, [])  # three disconnected vertices
    run(lib, 40, [(i, i + 1) for i in range(39)])


def test_dense_cliques(lib):
    # one clique = one wide supernode: several panels, no rows below
    n = 45
    st = run(lib, n, [(i, j) for i in range(n) for j in range(i + 1, n)], seed=1)
    assert st[1] >= 3 and st[0] == 1
    # a clique hanging under many leaves: panels with > 64 rows below them (scratch path)
    n = 130
    edges = [(i, j) for i in range(30, n) for j in range(i + 1, n)]
    edges += [(i, j) for i in range(30) for j in range(30, n)]
    st = run(lib, n, edges, seed=2)
    assert st[7] > 0  # some diagonal part went through the scratch area


def test_grid_graphs(lib):
    for w, h in [(7, 5), (30, 30)]:
        idx = lambda x, y: y * w + x
        edges = [(idx(x, y), idx(x + 1, y)) for y in range(h) for x in range(w - 1)]
        edges += [(idx(x, y), idx(x, y + 1)) for y in range(h - 1) for x in range(w)]
        edges += [(idx(x, y), idx(x + 1, y + 1)) for y in range(h - 1) for x in range(w - 1)]
        run(lib, w * h, edges, seed=w)


def test_random_graphs(lib):
    rng = np.random.default_rng(5)
    for n, e in [(30, 60), (200, 900), (500, 1500)]:
        pairs = set()
        while len(pairs) < e:
            i, j = rng.integers(0, n, 2)
            if i != j:
                pairs.add((min(i, j), max(i, j)))
        run(lib, n, sorted(pairs), seed=n)


def test_manhattan_graph(lib):
    from cg_mrslam_b200 import synth
    g = synth.make_pose_graph(3000, 12000, seed=7, box=60.0)
    e = g["edge_ij"]
    keep = (e[:, 0] != 0) & (e[:, 1] != 0)
    pairs = sorted({(min(a, b) - 1, max(a, b) - 1) for a, b in e[keep]})
    st = run(lib, 2999, pairs, seed=11)
    assert st[2] > 3 and st[5] > 0


def analyse(lib, n, pairs, world=1):
    lib.pgo_hostsim_analyse.restype = ctypes.c_int
    e = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    ei, ej = np.ascontiguousarray(e[:, 0]), np.ascontiguousarray(e[:, 1])
    out = np.zeros(6, dtype=np.int64)
    rc = lib.pgo_hostsim_analyse(n, len(e), ei.ctypes.data_as(ctypes.c_void_p), ej.ctypes.data_as(ctypes.c_void_p),
                                 world, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return dict(nnzb=int(out[0]), mflop=int(out[1]), plevels=int(out[2]), slevels=int(out[3]),
                supernodes=int(out[4]), shared=int(out[5]))


def test_ordering_quality(lib):
    """Fill and work of the dissection ordering stay where the separator refinement put them
    (thresholds = measured values + 10 %). On the regular grid the level-structure
    cuts this solver started with were as good; on the walk-shaped pose graphs they needed 4x the
    flops."""
    # 60 x 60 king's-move grid: the ideal top separator is one row of 60 vertices
    w = h = 60
    idx = lambda x, y: y * w + x
    edges = [(idx(x, y), idx(x + 1, y)) for y in range(h) for x in range(w - 1)]
    edges += [(idx(x, y), idx(x, y + 1)) for y in range(h - 1) for x in range(w)]
    edges += [(idx(x, y), idx(x + 1, y + 1)) for y in range(h - 1) for x in range(w - 1)]
    st = analyse(lib, w * h, edges)
    assert st["nnzb"] <= 95000 and st["mflop"] <= 100, st
    from cg_mrslam_b200 import synth
    g = synth.make_pose_graph(12000, 48000, seed=9, box=122.0)
    e = g["edge_ij"]
    keep = (e[:, 0] != 0) & (e[:, 1] != 0)
    pairs = sorted({(min(a, b) - 1, max(a, b) - 1) for a, b in e[keep]})
    st = analyse(lib, 11999, pairs)
    assert st["nnzb"] <= 295000 and st["mflop"] <= 350 and st["plevels"] <= 40, st
    # the same graph cut for 4 ranks: the analysis' own partition checks pass, the shared
    # separators stay small and the fill does not suffer
    st4 = analyse(lib, 11999, pairs, world=4)
    assert st4["shared"] <= 200 and st4["nnzb"] <= 1.1 * st["nnzb"], (st, st4)

