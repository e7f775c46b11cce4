"""Python face of the scan-matcher C ABI (include/cgm_matcher.h).

Thin ctypes glue used by tests/, bench.py and tools/: every method forwards to one ``cgm_*`` entry
point. The names follow the reference's CharGrid / ScanMatcher members
(src/matcher/chargrid.h:106-231, src/matcher/scan_matcher.h:41-85).
"""
import ctypes as C

import numpy as np

from . import _lib

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_ubyte)


class cgm_result(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("theta", C.c_double), ("score", C.c_double)]


class MatcherError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "cgm error %d: %s" % (code, msg))
        self.code = code


_SIGS = {
    "cgm_last_error": (C.c_char_p, []),
    "cgm_device_count": (C.c_int, []),
    "cgm_launch_count": (C.c_uint64, []),
    "cgm_matcher_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int] +
                           [C.c_float] * 4 + [C.c_double, C.c_double, C.c_int]),
    "cgm_matcher_destroy": (None, [C.c_void_p]),
    "cgm_matcher_grid_size": (C.c_int, [C.c_void_p, _ip, _ip]),
    "cgm_matcher_stamp": (C.c_int, [C.c_void_p, _up, C.c_int, _ip]),
    "cgm_matcher_world2grid": (C.c_int, [C.c_void_p, C.c_float, C.c_float, _ip, _ip]),
    "cgm_matcher_grid2world": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _fp, _fp]),
    "cgm_matcher_reset": (C.c_int, [C.c_void_p, C.c_int]),
    "cgm_matcher_raster": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int]),
    "cgm_matcher_grid_download": (C.c_int, [C.c_void_p, C.c_int, _up]),
    "cgm_matcher_grid_upload": (C.c_int, [C.c_void_p, C.c_int, _up]),
    "cgm_subsample": (C.c_int, [_dp, C.c_int, C.c_double, _dp, _ip]),
    "cgm_matcher_search": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, _fp, C.c_int] +
                           [C.c_double] * 7 + [C.POINTER(cgm_result), C.c_int, _ip]),
    "cgm_matcher_hierarchical_search": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, _fp, C.c_int] +
                                        [C.c_double] * 5 +
                                        [C.c_int, C.POINTER(cgm_result), C.c_int, _ip]),
    "cgm_matcher_count_points": (C.c_int, [C.c_void_p, C.c_int] + [C.c_float] * 4 + [_dp]),
    "cgm_matcher_search_non_matched": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, C.c_double,
                                                 _dp, _ip]),
    "cgm_matcher_raster_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _ip]),
    "cgm_matcher_hierarchical_search_levels": (C.c_int, [C.c_void_p, C.c_int, _dp, C.c_int, _fp, C.c_int, _dp,
                                                         C.c_int, C.c_void_p, C.c_int, _ip]),
    "cgm_matcher_set_stamp": (C.c_int, [C.c_void_p, _up, C.c_int]),
    "cgm_matcher_fill": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "cgm_matcher_fill_raster": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]),
    "cgm_matcher_copy_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "cgm_matcher_search_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _ip, _fp, _ip] +
                                 [C.c_double] * 7 + [C.POINTER(cgm_result), C.c_int, _ip]),
    "cgm_matcher_batch_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _dp, _ip, _dp, _ip, _fp,
                                          _ip] + [C.c_double] * 7),
    "cgm_matcher_batch_launch": (C.c_int, [C.c_void_p]),
    "cgm_matcher_batch_collect": (C.c_int, [C.c_void_p, C.POINTER(cgm_result), C.c_int, _ip]),
    "cgm_matcher_batch_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_uint64), _ip]),
    "cgm_matcher_batch_kernel_ms": (C.c_int, [C.c_void_p, _fp]),
    "cgm_matcher_stream": (C.c_void_p, [C.c_void_p]),
    "cgm_matcher_set_kernel": (C.c_int, [C.c_void_p, C.c_int]),
}

MATCHER_SYMBOLS = sorted(_SIGS)


def bind(lib):
    for name, (res, args) in _SIGS.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    return lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None and a.size else None


def _results(buf, n):
    out = np.empty((n, 4), dtype=np.float64)
    if n:
        out[:] = np.frombuffer(buf, dtype=np.float64, count=4 * n).reshape(n, 4)
    return out


def subsample(pts, res, lib_path=None):
    """CharGrid::subsample (chargrid.cpp:98-122)."""
    lib = bind(_lib.load(lib_path))
    pts = _d(pts).reshape(-1, 2)
    out = np.empty_like(pts)
    n = C.c_int()
    rc = lib.cgm_subsample(_ptr(pts, _dp), len(pts), float(res), _ptr(out, _dp), C.byref(n))
    if rc:
        raise MatcherError(rc, lib.cgm_last_error().decode())
    return out[: n.value].copy()


class Matcher:
    """n_slots CharGrids of one geometry on one GPU + the stamp of ScanMatcher::initializeKernel.

    ``Matcher(ll, ur, resolution, kernel_range)`` corresponds to ``ScanMatcher::initializeGrid`` +
    ``initializeKernel`` (scan_matcher.cpp:38-66); slot 0 is "the" grid of the reference API.
    """

    def __init__(self, ll, ur, resolution, kernel_range, kscale=128, n_slots=1, device=0,
                 stream=None, lib_path=None):
        self.lib = bind(_lib.load(lib_path))
        self.h = C.c_void_p()
        rc = self.lib.cgm_matcher_create(C.byref(self.h), device, stream, n_slots, ll[0], ll[1],
                                         ur[0], ur[1], resolution, kernel_range, kscale)
        self._check(rc)
        r, c = C.c_int(), C.c_int()
        self._check(self.lib.cgm_matcher_grid_size(self.h, C.byref(r), C.byref(c)))
        self.rows, self.cols, self.n_slots = r.value, c.value, n_slots
        self.resolution, self.kernel_range, self.kscale = resolution, kernel_range, kscale

    def _check(self, rc):
        if rc:
            raise MatcherError(rc, self.lib.cgm_last_error().decode())

    def close(self):
        if self.h:
            self.lib.cgm_matcher_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- geometry ---------------------------------------------------------------------------
    def stamp(self):
        dim = C.c_int()
        self._check(self.lib.cgm_matcher_stamp(self.h, None, 0, C.byref(dim)))
        buf = np.empty(dim.value * dim.value, dtype=np.uint8)
        self._check(self.lib.cgm_matcher_stamp(self.h, _ptr(buf, _up), buf.size, C.byref(dim)))
        return buf.reshape(dim.value, dim.value)

    def world2grid(self, x, y):
        a, b = C.c_int(), C.c_int()
        self._check(self.lib.cgm_matcher_world2grid(self.h, x, y, C.byref(a), C.byref(b)))
        return a.value, b.value

    def grid2world(self, ix, iy):
        a, b = C.c_float(), C.c_float()
        self._check(self.lib.cgm_matcher_grid2world(self.h, ix, iy, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- map building -----------------------------------------------------------------------
    def reset(self, slot=0):
        """ScanMatcher::resetGrid."""
        self._check(self.lib.cgm_matcher_reset(self.h, slot))

    def raster(self, pts, slot=0):
        """CharGrid::addAndConvolvePoints with the matcher's stamp (no reset)."""
        pts = _d(pts).reshape(-1, 2)
        self._check(self.lib.cgm_matcher_raster(self.h, slot, _ptr(pts, _dp), len(pts)))

    def raster_batch(self, pts_list, first_slot=0):
        """resetGrid + addAndConvolvePoints for consecutive slots."""
        counts = np.array([len(p) for p in pts_list], dtype=np.int32)
        packed = _d(np.concatenate([_d(p).reshape(-1, 2) for p in pts_list])
                    if len(pts_list) else np.zeros((0, 2)))
        self._check(self.lib.cgm_matcher_raster_batch(self.h, first_slot, len(pts_list),
                                                      _ptr(packed, _dp), _ptr(counts, _ip)))

    def download(self, slot=0):
        out = np.empty((self.rows, self.cols), dtype=np.uint8)
        self._check(self.lib.cgm_matcher_grid_download(self.h, slot, _ptr(out, _up)))
        return out

    def upload(self, cells, slot=0):
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        assert cells.shape == (self.rows, self.cols)
        self._check(self.lib.cgm_matcher_grid_upload(self.h, slot, _ptr(cells, _up)))

    # -- search -----------------------------------------------------------------------------
    def greedy_search(self, pts, regions, step, max_score, bins, slot=0, cap=1 << 16):
        """CharGrid::greedySearch(mresvec, points, regions, params): rows (x, y, theta, score)."""
        pts = _d(pts).reshape(-1, 2)
        reg = np.ascontiguousarray(regions, dtype=np.float32).reshape(-1, 6)
        buf = (cgm_result * cap)()
        n = C.c_int()
        self._check(self.lib.cgm_matcher_search(self.h, slot, _ptr(pts, _dp), len(pts),
                                                _ptr(reg, _fp), len(reg), step[0], step[1],
                                                step[2], max_score, bins[0], bins[1], bins[2],
                                                buf, cap, C.byref(n)))
        if n.value > cap:
            return self.greedy_search(pts, regions, step, max_score, bins, slot, n.value)
        return _results(buf, n.value)

    def greedy_search_res(self, pts, regions, theta_res, max_score, bins, slot=0, cap=1 << 16):
        """The overloads that search at the grid resolution (chargrid.cpp:182-206)."""
        r = np.float32(self.resolution)
        return self.greedy_search(pts, regions, (float(r), float(r), theta_res), max_score, bins,
                                  slot, cap)

    def hierarchical_search(self, pts, regions, theta_res, max_score, bins, n_levels, slot=0,
                            cap=1 << 16):
        pts = _d(pts).reshape(-1, 2)
        reg = np.ascontiguousarray(regions, dtype=np.float32).reshape(-1, 6)
        buf = (cgm_result * cap)()
        n = C.c_int()
        self._check(self.lib.cgm_matcher_hierarchical_search(
            self.h, slot, _ptr(pts, _dp), len(pts), _ptr(reg, _fp), len(reg), theta_res, max_score,
            bins[0], bins[1], bins[2], n_levels, buf, cap, C.byref(n)))
        if n.value > cap:
            return self.hierarchical_search(pts, regions, theta_res, max_score, bins, n_levels,
                                            slot, n.value)
        return _results(buf, n.value)

    def count_points(self, ll, ur, slot=0):
        s = C.c_double()
        self._check(self.lib.cgm_matcher_count_points(self.h, slot, ll[0], ll[1], ur[0], ur[1],
                                                      C.byref(s)))
        return s.value

    def search_non_matched(self, pts, max_score, slot=0):
        pts = _d(pts).reshape(-1, 2)
        out = np.empty_like(pts)
        n = C.c_int()
        self._check(self.lib.cgm_matcher_search_non_matched(self.h, slot, _ptr(pts, _dp), len(pts),
                                                            max_score, _ptr(out, _dp),
                                                            C.byref(n)))
        return out[: n.value].copy()

    # -- batch ------------------------------------------------------------------------------
    @staticmethod
    def _pack(pts_list, regions_list):
        pc = np.array([len(p) for p in pts_list], dtype=np.int32)
        rc = np.array([len(np.asarray(r).reshape(-1, 6)) for r in regions_list], dtype=np.int32)
        pts = _d(np.concatenate([_d(p).reshape(-1, 2) for p in pts_list])) if len(pts_list) \
            else np.zeros((0, 2))
        reg = np.ascontiguousarray(
            np.concatenate([np.asarray(r, dtype=np.float32).reshape(-1, 6) for r in regions_list])
            if len(regions_list) else np.zeros((0, 6)), dtype=np.float32)
        return pts, pc, reg, rc

    def search_batch(self, pts_list, regions_list, step, max_score, bins, first_slot=0, cap=64):
        """n independent greedySearch problems in one scoring launch; list of [n_i, 4] arrays."""
        pts, pc, reg, rc = self._pack(pts_list, regions_list)
        n = len(pts_list)
        buf = (cgm_result * (cap * max(n, 1)))()
        n_out = np.zeros(max(n, 1), dtype=np.int32)
        self._check(self.lib.cgm_matcher_search_batch(
            self.h, first_slot, n, _ptr(pts, _dp), _ptr(pc, _ip), _ptr(reg, _fp), _ptr(rc, _ip),
            step[0], step[1], step[2], max_score, bins[0], bins[1], bins[2], buf, cap,
            _ptr(n_out, _ip)))
        if n and int(n_out.max()) > cap:
            return self.search_batch(pts_list, regions_list, step, max_score, bins, first_slot,
                                     int(n_out.max()))
        flat = np.frombuffer(buf, dtype=np.float64).reshape(max(n, 1), cap, 4)
        return [flat[i, : n_out[i]].copy() for i in range(n)]

    def batch_stage(self, pts, pts_counts, regions, region_counts, step, max_score, bins,
                    first_slot=0, map_pts=None, map_counts=None):
        """Raw-pointer staging for the benchmark: arrays must stay alive until collect."""
        n = len(pts_counts)
        self._check(self.lib.cgm_matcher_batch_stage(
            self.h, first_slot, n,
            _ptr(map_pts, _dp) if map_pts is not None else None,
            _ptr(map_counts, _ip) if map_counts is not None else None,
            _ptr(pts, _dp), _ptr(pts_counts, _ip), _ptr(regions, _fp), _ptr(region_counts, _ip),
            step[0], step[1], step[2], max_score, bins[0], bins[1], bins[2]))
        self._staged_n = n

    def batch_launch(self):
        self._check(self.lib.cgm_matcher_batch_launch(self.h))

    def batch_collect(self, cap=64):
        n = self._staged_n
        buf = (cgm_result * (cap * max(n, 1)))()
        n_out = np.zeros(max(n, 1), dtype=np.int32)
        self._check(self.lib.cgm_matcher_batch_collect(self.h, buf, cap, _ptr(n_out, _ip)))
        flat = np.frombuffer(buf, dtype=np.float64).reshape(max(n, 1), cap, 4)
        return [flat[i, : min(cap, n_out[i])].copy() for i in range(n)], n_out[:n].copy()

    def batch_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_int()
        self._check(self.lib.cgm_matcher_batch_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"candidates": a.value, "cell_reads": b.value, "score_launches": c.value}

    def kernel_ms(self):
        ms = (C.c_float * 3)()
        self._check(self.lib.cgm_matcher_batch_kernel_ms(self.h, ms))
        return {"raster": ms[0], "score": ms[1], "compact": ms[2]}

    def stream(self):
        return self.lib.cgm_matcher_stream(self.h)

    def set_kernel(self, which):
        self._check(self.lib.cgm_matcher_set_kernel(self.h, which))

    def launch_count(self):
        return self.lib.cgm_launch_count()
