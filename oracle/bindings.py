"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracles.

Two libraries expose the same matcher surface under different prefixes:

* ``orc_*``  oracle/_build/liboracle_matcher.so -- the restatement (oracle/matcher_oracle.cpp)
* ``ref_*``  oracle/_ref/libref_chargrid.so     -- the reference's chargrid.cpp compiled verbatim
  (oracle/Makefile, target ``ref``; needs /root/reference at BUILD time only)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
may import this module. Nothing under cg_mrslam_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_MATCHER_SO = os.path.join(HERE, "_build", "liboracle_matcher.so")
ORACLE_PGO_SO = os.path.join(HERE, "_build", "liboracle_pgo.so")
REF_CHARGRID_SO = os.path.join(HERE, "_ref", "libref_chargrid.so")

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


def build(reference="/root/reference"):
    """Compile the oracles. The verbatim reference build is attempted only where the reference
    tree exists (this container); the GPU box uses the prebuilt oracle/_ref/*.so that travelled."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir(os.path.join(reference, "src", "matcher")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", "REFERENCE=" + reference])
        # the reference's own host sources + test drivers, linked to the CUDA library and to the CPU
        # oracles (needs cg_mrslam_b200/lib/libcgmrslam_b200.so: __graft_entry__.build() makes it first)
        if os.path.exists(os.path.join(HERE, "..", "cg_mrslam_b200", "lib", "libcgmrslam_b200.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "frontend", "REFERENCE=" + reference])


def _as_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class MatcherLib:
    """One of the two CPU matchers behind a common Python face."""

    def __init__(self, kind):
        assert kind in ("oracle", "reference")
        self.kind = kind
        self.prefix = "orc_" if kind == "oracle" else "ref_"
        path = ORACLE_MATCHER_SO if kind == "oracle" else REF_CHARGRID_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        self.lib = lib = C.CDLL(path)
        p = self.prefix
        f = getattr(lib, p + "grid_create")
        f.restype = C.c_void_p
        f.argtypes = [C.c_float] * 5 + [C.c_int]
        getattr(lib, p + "grid_destroy").argtypes = [C.c_void_p]
        getattr(lib, p + "grid_size").argtypes = [C.c_void_p, _ip, _ip]
        getattr(lib, p + "grid_fill").argtypes = [C.c_void_p, C.c_int]
        getattr(lib, p + "grid_raster").argtypes = [C.c_void_p, _dp, C.c_int, _up, C.c_int]
        getattr(lib, p + "grid_download").argtypes = [C.c_void_p, _up]
        getattr(lib, p + "grid_upload").argtypes = [C.c_void_p, _up]
        getattr(lib, p + "world2grid").argtypes = [C.c_void_p, C.c_float, C.c_float, _ip, _ip]
        getattr(lib, p + "grid2world").argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp]
        f = getattr(lib, p + "subsample")
        f.restype = C.c_int
        f.argtypes = [_dp, C.c_int, C.c_double, _dp]
        f = getattr(lib, p + "greedy_search")
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, _dp, C.c_int, _fp, C.c_int] + [C.c_double] * 7 + [_dp, C.c_int]
        f = getattr(lib, p + "greedy_search_res")
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, _dp, C.c_int, _fp, C.c_int] + [C.c_double] * 5 + [_dp, C.c_int]
        f = getattr(lib, p + "hierarchical_search")
        f.restype = C.c_int
        f.argtypes = ([C.c_void_p, _dp, C.c_int, _fp, C.c_int] + [C.c_double] * 5 +
                      [C.c_int, _dp, C.c_int])
        f = getattr(lib, p + "count_points")
        f.restype = C.c_double
        f.argtypes = [C.c_void_p] + [C.c_float] * 4
        f = getattr(lib, p + "search_non_matched")
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, _dp, C.c_int, C.c_double, _dp]
        if kind == "oracle":
            f = lib.orc_make_stamp
            f.restype = C.c_int
            f.argtypes = [C.c_double, C.c_double, C.c_int, _up, C.c_int]

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def grid(self, ll, ur, res, kscale=128):
        return CpuGrid(self, ll, ur, res, kscale)

    def subsample(self, pts, res):
        pts = _as_d(pts).reshape(-1, 2)
        out = np.empty_like(pts)
        n = self.fn("subsample")(pts.ctypes.data_as(_dp), len(pts), float(res),
                                 out.ctypes.data_as(_dp))
        return out[:n].copy()

    def make_stamp(self, resolution, kernel_range, kscale=128):
        """ScanMatcher::initializeKernel (scan_matcher.cpp:38-61), restated (oracle only: the
        reference's scan_matcher.cpp needs g2o and cannot be compiled here)."""
        assert self.kind == "oracle"
        buf = np.zeros(4096, dtype=np.uint8)
        dim = self.lib.orc_make_stamp(float(resolution), float(kernel_range), int(kscale),
                                      buf.ctypes.data_as(_up), buf.size)
        assert dim > 0
        return buf[: dim * dim].reshape(dim, dim).copy()  # [col][row] (column-major), symmetric


class CpuGrid:
    def __init__(self, lib, ll, ur, res, kscale):
        self.m = lib
        self.h = C.c_void_p(lib.fn("grid_create")(ll[0], ll[1], ur[0], ur[1], res, kscale))
        r, c = C.c_int(), C.c_int()
        lib.fn("grid_size")(self.h, C.byref(r), C.byref(c))
        self.rows, self.cols = r.value, c.value

    def __del__(self):
        try:
            self.m.fn("grid_destroy")(self.h)
        except Exception:
            pass

    def fill(self, value):
        self.m.fn("grid_fill")(self.h, int(value))

    def raster(self, pts, stamp):
        pts = _as_d(pts).reshape(-1, 2)
        stamp = np.ascontiguousarray(stamp, dtype=np.uint8)
        self.m.fn("grid_raster")(self.h, pts.ctypes.data_as(_dp), len(pts),
                                 stamp.ctypes.data_as(_up), stamp.shape[0])

    def download(self):
        out = np.empty((self.rows, self.cols), dtype=np.uint8)
        self.m.fn("grid_download")(self.h, out.ctypes.data_as(_up))
        return out

    def upload(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        assert cells.shape == (self.rows, self.cols)
        self.m.fn("grid_upload")(self.h, cells.ctypes.data_as(_up))

    def world2grid(self, x, y):
        a, b = C.c_int(), C.c_int()
        self.m.fn("world2grid")(self.h, x, y, C.byref(a), C.byref(b))
        return a.value, b.value

    def grid2world(self, ix, iy):
        a, b = C.c_float(), C.c_float()
        self.m.fn("grid2world")(self.h, ix, iy, C.byref(a), C.byref(b))
        return a.value, b.value

    @staticmethod
    def _regions(regions):
        r = np.ascontiguousarray(regions, dtype=np.float32).reshape(-1, 6)
        return r

    def greedy_search(self, pts, regions, step, max_score, bins, cap=1 << 16):
        """CharGrid::greedySearch(mresvec, points, regions, params). Returns [n][4] doubles."""
        pts = _as_d(pts).reshape(-1, 2)
        r = self._regions(regions)
        out = np.empty((cap, 4), dtype=np.float64)
        n = self.m.fn("greedy_search")(self.h, pts.ctypes.data_as(_dp), len(pts),
                                       r.ctypes.data_as(_fp), len(r), step[0], step[1], step[2],
                                       max_score, bins[0], bins[1], bins[2],
                                       out.ctypes.data_as(_dp), cap)
        assert n <= cap
        return out[:n].copy()

    def greedy_search_res(self, pts, regions, theta_res, max_score, bins, cap=1 << 16):
        pts = _as_d(pts).reshape(-1, 2)
        r = self._regions(regions)
        out = np.empty((cap, 4), dtype=np.float64)
        n = self.m.fn("greedy_search_res")(self.h, pts.ctypes.data_as(_dp), len(pts),
                                           r.ctypes.data_as(_fp), len(r), theta_res, max_score,
                                           bins[0], bins[1], bins[2], out.ctypes.data_as(_dp), cap)
        assert n <= cap
        return out[:n].copy()

    def hierarchical_search(self, pts, regions, theta_res, max_score, bins, n_levels,
                            cap=1 << 16):
        pts = _as_d(pts).reshape(-1, 2)
        r = self._regions(regions)
        out = np.empty((cap, 4), dtype=np.float64)
        n = self.m.fn("hierarchical_search")(self.h, pts.ctypes.data_as(_dp), len(pts),
                                             r.ctypes.data_as(_fp), len(r), theta_res, max_score,
                                             bins[0], bins[1], bins[2], n_levels,
                                             out.ctypes.data_as(_dp), cap)
        assert n <= cap
        return out[:n].copy()

    def count_points(self, ll, ur):
        return self.m.fn("count_points")(self.h, ll[0], ll[1], ur[0], ur[1])

    def search_non_matched(self, pts, max_score):
        pts = _as_d(pts).reshape(-1, 2)
        out = np.empty_like(pts)
        n = self.m.fn("search_non_matched")(self.h, pts.ctypes.data_as(_dp), len(pts), max_score,
                                            out.ctypes.data_as(_dp))
        return out[:n].copy()


def have_reference():
    return os.path.exists(REF_CHARGRID_SO)


class CGaussNewton:
    """The compiled pose-graph oracle (oracle/pgo_oracle_c.cpp): Gauss-Newton with an up-looking
    scalar sparse Cholesky, one thread -- the shape of g2o + LinearSolverCSparse."""

    def __init__(self):
        if not os.path.exists(ORACLE_PGO_SO):
            raise FileNotFoundError(ORACLE_PGO_SO + " (run `make -C oracle`)")
        self.lib = C.CDLL(ORACLE_PGO_SO)
        f = self.lib.pgo_oracle_c_gauss_newton
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _up, _dp, _dp, _dp,
                      C.c_int, _dp, _dp, C.POINTER(C.c_longlong)]

    def gauss_newton(self, poses0, edge_ij, meas, info6, fixed, n_iters):
        """Returns dict(poses, chi2 [iterations done], iterations, times {analyse, linearise,
        factor, solve} in seconds, nnz_l, flops)."""
        poses = np.array(poses0, dtype=np.float64, copy=True, order="C")
        e = np.asarray(edge_ij).reshape(-1, 2)
        ei = np.ascontiguousarray(e[:, 0], dtype=np.int32)
        ej = np.ascontiguousarray(e[:, 1], dtype=np.int32)
        mask = np.zeros(len(poses), dtype=np.uint8)
        mask[np.asarray(fixed, dtype=np.int64)] = 1
        meas, info6 = _as_d(meas), _as_d(info6)
        chi2 = np.zeros(max(n_iters, 1))
        times = np.zeros(4)
        stats = (C.c_longlong * 3)()
        done = self.lib.pgo_oracle_c_gauss_newton(
            len(poses), len(e), ei.ctypes.data_as(C.POINTER(C.c_int32)),
            ej.ctypes.data_as(C.POINTER(C.c_int32)), mask.ctypes.data_as(_up),
            poses.ctypes.data_as(_dp), meas.ctypes.data_as(_dp), info6.ctypes.data_as(_dp), n_iters,
            chi2.ctypes.data_as(_dp), times.ctypes.data_as(_dp), stats)
        return {"poses": poses, "chi2": chi2[:max(done, min(n_iters, done + 1))], "iterations": done,
                "times": dict(zip(("analyse", "linearise", "factor", "solve"), times.tolist())),
                "nnz_l": int(stats[0]), "flops": float(stats[1])}
