"""Python face of the pose-graph solver C ABI (include/pgo_solver.h).

Thin ctypes glue for tests/, bench.py and tools/. Method names follow the g2o calls the reference
makes (SparseOptimizer::initializeOptimization / optimize / computeMarginals /
computeInitialGuess, EdgeLabeler::labelEdges; call sites listed in SURVEY.md section 8b).
"""
import ctypes as C

import numpy as np

from . import _lib

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class pgo_stats(C.Structure):
    _fields_ = [("n_vertices", C.c_int32), ("n_edges", C.c_int32), ("n_free", C.c_int32),
                ("n_levels", C.c_int32), ("factor_blocks", C.c_int64), ("update_ops", C.c_int64),
                ("hessian_blocks", C.c_int64), ("analyse_seconds", C.c_double),
                ("last_iterate_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("stage_ms", C.c_double * 5),
                ("factor_flops", C.c_double),
                ("n_supernodes", C.c_int32), ("n_panels", C.c_int32),
                ("n_panel_levels", C.c_int32), ("n_supernode_levels", C.c_int32),
                ("batch", C.c_int32)]


class SolverError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "pgo error %d: %s" % (code, msg))
        self.code = code


_SIGS = {
    "pgo_last_error": (C.c_char_p, []),
    "pgo_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "pgo_destroy": (None, [C.c_void_p]),
    "pgo_set_graph": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _bp]),
    "pgo_upload": (C.c_int, [C.c_void_p, _dp, _dp, _dp]),
    "pgo_set_poses": (C.c_int, [C.c_void_p, _dp]),
    "pgo_get_poses": (C.c_int, [C.c_void_p, _dp]),
    "pgo_iterate": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, C.POINTER(C.c_int)]),
    "pgo_chi2": (C.c_int, [C.c_void_p, _dp]),
    "pgo_marginals": (C.c_int, [C.c_void_p, C.c_int, _ip, _ip, _dp]),
    "pgo_initial_guess": (C.c_int, [C.c_void_p]),
    "pgo_label_star_edges": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _ip, _dp, _dp]),
    "pgo_set_partition": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "pgo_dd_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "pgo_dd_local": (C.c_int, [C.c_void_p]),
    "pgo_dd_exchange_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pgo_dd_shared": (C.c_int, [C.c_void_p]),
    "pgo_dd_end": (C.c_int, [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int)]),
    "pgo_dd_pose_exchange": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "pgo_dd_pose_commit": (C.c_int, [C.c_void_p]),
    "pgo_dd_unique_id": (C.c_int, [C.c_void_p]),
    "pgo_dd_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pgo_dd_set_comm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pgo_dd_iterate": (C.c_int, [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_int)]),
    "pgo_analyse_partition": (C.c_int, [C.c_int, C.c_int, _ip, _ip, _bp, C.c_int, _ip,
                                        C.POINTER(C.c_int64)]),
    "pgo_get_stats": (C.c_int, [C.c_void_p, C.POINTER(pgo_stats)]),
    "pgo_set_batch": (C.c_int, [C.c_void_p, C.c_int]),
    "pgo_upload_instance": (C.c_int, [C.c_void_p, C.c_int, _dp, _dp, _dp]),
    "pgo_get_poses_instance": (C.c_int, [C.c_void_p, C.c_int, _dp]),
    "pgo_iterate_batch": (C.c_int, [C.c_void_p, C.c_int, _dp, _ip]),
    "pgo_stream": (C.c_void_p, [C.c_void_p]),
}

PGO_SYMBOLS = sorted(_SIGS)
PGO_ERR_NUMERIC = -5


def bind(lib):
    for name, (res, args) in _SIGS.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    return lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None and a.size else None


class Solver:
    """One SparseOptimizer-like solver on one GPU."""

    def __init__(self, device=0, stream=None, batch=1):
        self.lib = bind(_lib.load())
        self.h = C.c_void_p()
        self._check(self.lib.pgo_create(C.byref(self.h), device, stream))
        self.n_vertices = self.n_edges = 0
        self.batch = int(batch)
        if self.batch != 1:
            self._check(self.lib.pgo_set_batch(self.h, self.batch))

    def _check(self, rc):
        if rc:
            raise SolverError(rc, self.lib.pgo_last_error().decode())

    def close(self):
        if self.h:
            self.lib.pgo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_partition(self, rank, world):
        """Domain decomposition: this solver is rank `rank` of `world` (before set_graph)."""
        self._check(self.lib.pgo_set_partition(self.h, rank, world))

    def comm_init(self, rank, world, broadcast_bytes):
        """Domain decomposition with the all-reduce inside the library: creates the solver's NCCL
        communicator. ``broadcast_bytes(b)`` must return rank 0's bytes on every rank (e.g. through
        torch.distributed); call before set_graph."""
        uid = C.create_string_buffer(128)
        if rank == 0:
            self._check(self.lib.pgo_dd_unique_id(uid))
        data = broadcast_bytes(uid.raw)
        self._check(self.lib.pgo_dd_comm_init(self.h, C.c_char_p(data), rank, world))

    def optimize_dd(self, n_iters):
        """optimize(n) of ONE graph cut over the ranks, the whole loop in C (pgo_dd_iterate)."""
        chi2 = np.zeros(max(n_iters, 1))
        done = C.c_int()
        rc = self.lib.pgo_dd_iterate(self.h, n_iters, _p(chi2, _dp), C.byref(done))
        if rc and rc != PGO_ERR_NUMERIC:
            self._check(rc)
        return done.value, chi2[:n_iters]

    def set_graph(self, n_vertices, edge_ij, fixed):
        """initializeOptimization: structure of the active graph. ``fixed`` = vertex indices."""
        e = np.asarray(edge_ij).reshape(-1, 2)
        ei, ej = _i(e[:, 0]), _i(e[:, 1])
        mask = np.zeros(n_vertices, dtype=np.uint8)
        mask[np.asarray(fixed, dtype=np.int64)] = 1
        self._check(self.lib.pgo_set_graph(self.h, n_vertices, len(e), _p(ei, _ip), _p(ej, _ip),
                                           _p(mask, _bp)))
        self.n_vertices, self.n_edges = n_vertices, len(e)

    def upload(self, poses, meas, info6):
        poses, meas, info6 = _d(poses), _d(meas), _d(info6)
        assert poses.shape == (self.n_vertices, 3) and meas.shape == (self.n_edges, 3) \
            and info6.shape == (self.n_edges, 6)
        self._check(self.lib.pgo_upload(self.h, _p(poses, _dp), _p(meas, _dp), _p(info6, _dp)))

    def upload_instance(self, inst, poses, meas, info6):
        """Values of one instance of a batch (upload() gives every instance the same values)."""
        poses, meas, info6 = _d(poses), _d(meas), _d(info6)
        assert poses.shape == (self.n_vertices, 3) and meas.shape == (self.n_edges, 3) \
            and info6.shape == (self.n_edges, 6)
        self._check(self.lib.pgo_upload_instance(self.h, int(inst), _p(poses, _dp), _p(meas, _dp),
                                                 _p(info6, _dp)))

    def poses_of(self, inst, out=None):
        """Estimates of one instance; `out` (C-contiguous float64 [n_vertices, 3], e.g. pinned) is
        filled in place when given."""
        if out is None:
            out = np.empty((self.n_vertices, 3))
        assert out.shape == (self.n_vertices, 3) and out.dtype == np.float64 and out.flags.c_contiguous
        self._check(self.lib.pgo_get_poses_instance(self.h, int(inst), _p(out, _dp)))
        return out

    def optimize_batch(self, n_iters):
        """optimize(n) of every instance: (iterations done [batch], chi2 [batch, n_iters])."""
        chi2 = np.zeros((self.batch, max(n_iters, 1)))
        done = np.zeros(self.batch, dtype=np.int32)
        rc = self.lib.pgo_iterate_batch(self.h, n_iters, _p(chi2, _dp), _p(done, _ip))
        if rc and rc != PGO_ERR_NUMERIC:
            self._check(rc)
        return done, chi2[:, :n_iters]

    def set_poses(self, poses):
        poses = _d(poses)
        assert poses.shape == (self.n_vertices, 3)
        self._check(self.lib.pgo_set_poses(self.h, _p(poses, _dp)))

    def poses(self):
        out = np.empty((self.n_vertices, 3))
        self._check(self.lib.pgo_get_poses(self.h, _p(out, _dp)))
        return out

    def optimize(self, n_iters, want_poses=True):
        """SparseOptimizer::optimize(n): returns (iterations done, chi2 per iteration, poses)."""
        chi2 = np.zeros(max(n_iters, 1))
        out = np.empty((self.n_vertices, 3)) if want_poses else None
        done = C.c_int()
        rc = self.lib.pgo_iterate(self.h, n_iters, _p(out, _dp) if want_poses else None,
                                  _p(chi2, _dp), C.byref(done))
        if rc and rc != PGO_ERR_NUMERIC:
            self._check(rc)
        return done.value, chi2[:n_iters], out

    def chi2(self):
        v = C.c_double()
        self._check(self.lib.pgo_chi2(self.h, C.byref(v)))
        return v.value

    def marginals(self, pairs):
        """computeMarginals: blocks (r, c) of H^-1, vertex index pairs -> [n, 3, 3]."""
        pr = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
        r, c = _i(pr[:, 0]), _i(pr[:, 1])
        out = np.empty((len(pr), 3, 3))
        self._check(self.lib.pgo_marginals(self.h, len(pr), _p(r, _ip), _p(c, _ip), _p(out, _dp)))
        return out

    def initial_guess(self):
        self._check(self.lib.pgo_initial_guess(self.h))

    def label_star_edges(self, gauge, vs):
        v = _i(vs)
        meas = np.empty((len(v), 3))
        info = np.empty((len(v), 3, 3))
        self._check(self.lib.pgo_label_star_edges(self.h, int(gauge), len(v), _p(v, _ip),
                                                  _p(meas, _dp), _p(info, _dp)))
        return meas, info

    def stats(self):
        s = pgo_stats()
        self._check(self.lib.pgo_get_stats(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in pgo_stats._fields_}
        d["stage_ms"] = dict(zip(("linearise", "factor", "forward", "backward", "update"),
                                 list(s.stage_ms)))
        return d

    def stream(self):
        return self.lib.pgo_stream(self.h)


class _DeviceArray:
    """A float64 device buffer exposed through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def analyse_partition(n_vertices, edge_ij, fixed, world):
    """The partition pgo_set_partition(rank, world) + pgo_set_graph would use (host only).
    Returns (vertex_owner [n_vertices]: rank, -1 shared, -2 fixed; stats dict)."""
    lib = bind(_lib.load())
    e = np.asarray(edge_ij).reshape(-1, 2)
    ei, ej = _i(e[:, 0]), _i(e[:, 1])
    mask = np.zeros(n_vertices, dtype=np.uint8)
    mask[np.asarray(fixed, dtype=np.int64)] = 1
    owner = np.empty(n_vertices, dtype=np.int32)
    stats = np.zeros(3 + world, dtype=np.int64)
    rc = lib.pgo_analyse_partition(n_vertices, len(e), _p(ei, _ip), _p(ej, _ip), _p(mask, _bp), world,
                                   _p(owner, _ip), stats.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc:
        raise SolverError(rc, lib.pgo_last_error().decode())
    return owner, {"shared_vertices": int(stats[0]), "shared_blocks": int(stats[1]),
                   "rank_updates": [int(x) for x in stats[2:2 + world]],
                   "shared_updates": int(stats[2 + world])}


def optimize_distributed(solvers, n_iters, all_reduce):
    """Domain-decomposed SparseOptimizer::optimize(n) over the ranks' solvers held by THIS process.

    ``solvers``: this process' Solver objects, each with set_partition(rank, world) applied before
    set_graph (one per process under torchrun; several "virtual ranks" on one GPU in the tests).
    ``all_reduce(list_of_tensors)`` sums the given tensors (one per local solver) over ALL ranks in
    place: with one solver per process it is ``lambda ts: dist.all_reduce(ts[0])``.
    Returns (iterations done, chi2 per iteration)."""
    import torch
    for s in solvers:
        s._check(s.lib.pgo_dd_begin(s.h, n_iters))

    def views(fn):
        out = []
        for s in solvers:
            ptr, n = C.c_void_p(), C.c_int64()
            s._check(fn(s.h, C.byref(ptr), C.byref(n)))
            out.append(torch.as_tensor(_DeviceArray(ptr.value, n.value), device="cuda"))
        return out

    for _ in range(n_iters):
        for s in solvers:
            s._check(s.lib.pgo_dd_local(s.h))
        all_reduce(views(lambda h, p, n: solvers[0].lib.pgo_dd_exchange_buffer(h, p, n)))
        for s in solvers:
            s._check(s.lib.pgo_dd_shared(s.h))
    chi2 = np.zeros(max(n_iters, 1))
    done = []
    for s in solvers:
        d = C.c_int()
        rc = s.lib.pgo_dd_end(s.h, n_iters, _p(chi2, _dp), C.byref(d))
        if rc and rc != PGO_ERR_NUMERIC:
            s._check(rc)
        done.append(d.value)
    all_reduce(views(lambda h, p, n: solvers[0].lib.pgo_dd_pose_exchange(h, p, n)))
    for s in solvers:
        s._check(s.lib.pgo_dd_pose_commit(s.h))
    return min(done), chi2[:n_iters]


def condensed_star(solver_factory, poses, edge_ij, meas, info6, gauge, separators):
    """CondensedGraphCreator::compute (condensed_graph_creator.cpp:33-66) on the GPU solver: fix
    only the gauge, computeInitialGuess, optimize(1), label star edges gauge -> v."""
    s = solver_factory()
    try:
        s.set_graph(len(poses), edge_ij, [gauge])
        s.upload(poses, meas, info6)
        s.initial_guess()
        done, _, _ = s.optimize(1, want_poses=False)
        if done != 1:
            raise SolverError(PGO_ERR_NUMERIC, "optimize(1) failed")
        vs = [int(v) for v in separators if int(v) != int(gauge)]
        z, om = s.label_star_edges(gauge, vs)
    finally:
        s.close()
    return z, om, vs
