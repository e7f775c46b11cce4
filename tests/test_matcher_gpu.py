"""GPU parity tests of the scan matcher: the product library (CUDA kernels, through the C ABI)
against the committed golden vectors of the verbatim reference and against the CPU oracle on the
same seeded inputs. Bit-exact: grids byte for byte, result lists double for double."""
import math

import numpy as np
import pytest

import cases
import golden_util as gu
from cg_mrslam_b200 import matcher

pytestmark = pytest.mark.gpu

KERNELS = {"auto": 0, "global": 1, "tiled": 2}


def _mk(cfg, n_slots=1, kernel=0):
    m = matcher.Matcher(cfg["ll"], cfg["ur"], cfg["res"], cfg["kernel_range"], n_slots=n_slots)
    m.set_kernel(kernel)
    return m


@pytest.mark.parametrize("kernel", ["auto", "global"])
@pytest.mark.parametrize("path", gu.matcher_fixtures(), ids=lambda p: p.split("matcher_")[-1][:-4])
def test_gpu_reproduces_golden(path, kernel):
    fx = gu.load(path)
    digest, sub, res = gu.replay_product(fx, lambda cfg: _mk(cfg, kernel=KERNELS[kernel]),
                                         matcher.subsample)
    assert digest == str(fx["grid_sha256"])
    assert np.array_equal(sub, fx["sub_pts"])
    assert cases.same(res, fx["results"])


@pytest.mark.parametrize("name", ["lc_7regions", "lc_7regions_flip", "bench_1081_raw"])
def test_tiled_kernel_is_used_and_exact(name):
    """The LC-grid windows must qualify for the shared-memory kernel (kernel=2 fails otherwise)."""
    fx = gu.load([p for p in gu.matcher_fixtures() if p.endswith("matcher_%s.npz" % name)][0])
    _, _, res = gu.replay_product(fx, lambda cfg: _mk(cfg, kernel=2), matcher.subsample)
    assert cases.same(res, fx["results"])


@pytest.mark.parametrize("kernel", ["global", "tiled"])
def test_batch_vs_oracle(oracle_lib, kernel):
    n = 6
    pairs = [cases.scan_pair(40 + i, 1081 if i % 2 else 361, 1.5 * math.pi if i % 2 else math.pi,
                             (1.0, 1.0, 0.4)) for i in range(n)]
    rng = np.random.default_rng(9)
    regs = [cases.lc_regions(rng.uniform(-1.5, 1.5, (k, 3)))[0] for k in (1, 5, 2, 9, 3, 4)]
    m = _mk(cases.LC, n_slots=n + 2, kernel=KERNELS[kernel])
    m.raster_batch([p["map_pts"] for p in pairs], first_slot=2)
    subs = [matcher.subsample(p["cur_pts"], 0.1) if i % 3 else p["cur_pts"]
            for i, p in enumerate(pairs)]
    res = m.search_batch(subs, regs, (0.1, 0.1, 0.025), 0.35, cases.BINS, first_slot=2)
    stamp = oracle_lib.make_stamp(0.1, 0.5)
    total = 0
    for i in range(n):
        g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
        g.fill(64)
        g.raster(pairs[i]["map_pts"], stamp)
        assert np.array_equal(m.download(slot=2 + i), g.download())
        want = g.greedy_search(subs[i], regs[i], (0.1, 0.1, 0.025), 0.35, cases.BINS)
        assert cases.same(res[i], want), "problem %d" % i
        total += len(want)
    assert total > 0
    m.close()


@pytest.mark.parametrize("kernel", ["global", "auto"])
def test_edge_cases_gpu(oracle_lib, kernel):
    m = _mk(cases.LC, kernel=KERNELS[kernel])
    m.raster_batch([np.zeros((0, 2))])
    assert int(m.download().min()) == 64 and int(m.download().max()) == 64
    reg, th = cases.lc_regions([(0, 0, 0)])
    assert len(m.greedy_search_res(np.zeros((0, 2)), reg, th, 0.5, cases.BINS)) == 0
    assert len(m.greedy_search_res(np.ones((3, 2)), np.zeros((0, 6)), th, 0.5, cases.BINS)) == 0
    bad = reg.copy()
    bad[0, 3] = bad[0, 0] - 1.0
    assert len(m.greedy_search_res(np.ones((3, 2)), bad, th, 0.5, cases.BINS)) == 0
    g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
    g.fill(64)
    pts = np.array([[0.0, 0.0], [1.0, 0.5], [2.0, -1.0], [2.0, -1.0], [-3.0, 2.5]])
    # windows hanging over each grid border: out-of-grid cells add 0 but count in k
    for lo, hi in (((33.0, 33.0), (36.5, 36.5)), ((-36.5, -36.0), (-33.5, -33.0)),
                   ((-36.0, 33.5), (-33.0, 36.0)), ((30.0, -40.0), (40.0, -30.0))):
        edge = np.array([[lo[0], lo[1], -0.1, hi[0], hi[1], 0.1]], dtype=np.float32)
        got = m.greedy_search_res(pts, edge, th, 0.6, cases.BINS)
        want = g.greedy_search_res(pts, edge, th, 0.6, cases.BINS)
        assert len(want) > 0 and cases.same(got, want)
    # a very loose threshold accepts everything: every bin reports its first-visited minimum
    pair = cases.scan_pair(55, 361, math.pi)
    m.raster_batch([pair["map_pts"]])
    g.raster(pair["map_pts"], oracle_lib.make_stamp(0.1, 0.5))
    got = m.greedy_search_res(pair["cur_pts"], reg, th, 10.0, cases.BINS)
    want = g.greedy_search_res(pair["cur_pts"], reg, th, 10.0, cases.BINS)
    assert cases.same(got, want)
    with pytest.raises(matcher.MatcherError):
        m.greedy_search(pts, reg, (0.1, 0.1, 0.0), 0.5, cases.BINS)
    with pytest.raises(matcher.MatcherError):
        m.raster(pts, slot=3)
    m.close()


def test_strided_search_and_hierarchy_vs_oracle(oracle_lib):
    """Coarse levels use strides of 8/4/2 cells (global kernel), the last level stride 1 (tiled)."""
    pair = cases.scan_pair(61, 1081, 1.5 * math.pi, (3.0, 2.0, 1.0))
    regions, th = cases.global_window()
    m = _mk(cases.LC)
    m.raster_batch([pair["map_pts"]])
    g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
    g.fill(64)
    g.raster(pair["map_pts"], oracle_lib.make_stamp(0.1, 0.5))
    sub = matcher.subsample(pair["cur_pts"], 0.1)
    for levels, ms in ((4, 0.3), (3, 0.25), (1, 0.3)):
        got = m.hierarchical_search(sub, regions, th, ms, cases.BINS, levels)
        want = g.hierarchical_search(sub, regions, th, ms, cases.BINS, levels)
        assert cases.same(got, want)
    got = m.greedy_search(sub, regions, (0.4, 0.2, 0.05), 0.3, cases.BINS)
    want = g.greedy_search(sub, regions, (0.4, 0.2, 0.05), 0.3, cases.BINS)
    assert len(want) > 0 and cases.same(got, want)
    m.close()


def test_cold_paths_gpu(oracle_lib):
    pair = cases.scan_pair(31, 361, math.pi)
    m = _mk(cases.CLOSE)
    m.raster_batch([pair["map_pts"]])
    g = oracle_lib.grid(cases.CLOSE["ll"], cases.CLOSE["ur"], cases.CLOSE["res"])
    g.fill(25)
    g.raster(pair["map_pts"], oracle_lib.make_stamp(0.025, 0.2))
    for ll, ur in (((-1.0, -1.0), (1.5, 2.0)), ((-16.0, -3.0), (-14.0, 3.0))):
        assert m.count_points(ll, ur) == g.count_points(ll, ur)
    assert np.array_equal(m.search_non_matched(pair["cur_pts"], 0.05),
                          g.search_non_matched(pair["cur_pts"], 0.05))
    m.reset()
    half = len(pair["map_pts"]) // 2
    m.raster(pair["map_pts"][:half])
    m.raster(pair["map_pts"][half:])
    assert np.array_equal(m.download(), g.download())
    m.close()


def test_uploaded_grid_with_large_cells(oracle_lib):
    """Raw uploads may hold any byte: accumulators must not overflow (G = 1 path)."""
    rng = np.random.default_rng(3)
    cells = rng.integers(0, 256, size=(700, 700), dtype=np.uint8)
    m = _mk(cases.LC)
    m.upload(cells)
    g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
    g.upload(cells)
    pts = rng.uniform(-7, 7, (900, 2))
    reg, th = cases.lc_regions([(0.2, -0.3, 0.1), (1.0, 1.0, -0.5)])
    for ms in (0.9, 1.0, 3.0):
        got = m.greedy_search_res(pts, reg, th, ms, cases.BINS)
        want = g.greedy_search_res(pts, reg, th, ms, cases.BINS)
        assert cases.same(got, want)
    m.close()


def test_full_size_properties():
    """BASELINE cfg-5 shape (1.01 M candidates, 1081 beams) at a size the CPU oracle would need
    seconds per pair for: properties that need no oracle.
      * both kernels agree bit for bit on every pair;
      * a scan matched against its own map puts the best pose within one cell of the origin;
      * re-launching the staged batch is idempotent."""
    n = 24
    pairs = [cases.scan_pair(200 + i, 1081, 1.5 * math.pi, (3.0, 3.0, 0.8)) for i in range(n)]
    regions, th = cases.bench_window()
    out = {}
    for kernel in ("global", "tiled"):
        m = _mk(cases.LC, n_slots=n, kernel=KERNELS[kernel])
        m.raster_batch([p["map_pts"] for p in pairs])
        out[kernel] = m.search_batch([p["cur_pts"] for p in pairs], [regions] * n,
                                     (0.1, 0.1, th), 0.15, cases.BINS)
        if kernel == "tiled":
            again = m.search_batch([p["cur_pts"] for p in pairs], [regions] * n,
                                   (0.1, 0.1, th), 0.15, cases.BINS)
            assert all(cases.same(a, b) for a, b in zip(out[kernel], again))
            own = m.search_batch([p["map_pts"] for p in pairs], [regions] * n,
                                 (0.1, 0.1, th), 0.15, cases.BINS)
            for r in own:
                assert len(r) > 0 and abs(r[0, 0]) <= 0.1001 and abs(r[0, 1]) <= 0.1001 \
                    and abs(r[0, 2]) <= 0.0251
        m.close()
    assert all(cases.same(a, b) for a, b in zip(out["global"], out["tiled"]))
    assert sum(len(r) for r in out["tiled"]) > 0


def test_two_geometries_in_one_process(oracle_lib):
    """GraphSLAM holds a close matcher (1200^2, 0.025 m) and a loop-closure matcher (700^2, 0.1 m) and
    uses them alternately (graph_slam.cpp:59-62,244,444): the rasteriser's shared-memory opt-in is a
    property of the kernel, not of the matcher (regression: the second geometry used to lower it
    under the first one's need). Also: a caller-supplied stamp, fill + raster in one launch, and a
    device-to-device grid copy (CharGrid's value semantics)."""
    pair = cases.scan_pair(11, 361, math.pi)
    ms = {name: matcher.Matcher(cfg["ll"], cfg["ur"], cfg["res"], cfg["kernel_range"])
          for name, cfg in (("close", cases.CLOSE), ("lc", cases.LC))}
    for rep in range(3):
        for name, cfg in (("close", cases.CLOSE), ("lc", cases.LC), ("close", cases.CLOSE)):
            m = ms[name]
            m.raster_batch([pair["map_pts"]])
            g = oracle_lib.grid(cfg["ll"], cfg["ur"], cfg["res"])
            k2 = int(cfg["kernel_range"] * 128)
            g.fill(k2)
            g.raster(pair["map_pts"], oracle_lib.make_stamp(cfg["res"], cfg["kernel_range"]))
            assert np.array_equal(m.download(), g.download()), (rep, name)
    # set_stamp + fill_raster + copy_grid through the C ABI
    import ctypes as C
    m = ms["lc"]
    stamp = oracle_lib.make_stamp(0.1, 0.3)                      # a different kernel than the matcher's own
    sb = np.ascontiguousarray(stamp, dtype=np.uint8).ravel()
    assert m.lib.cgm_matcher_set_stamp(m.h, sb.ctypes.data_as(C.POINTER(C.c_ubyte)), stamp.shape[0]) == 0
    pts = np.ascontiguousarray(pair["map_pts"], dtype=np.float64)
    assert m.lib.cgm_matcher_fill_raster(m.h, 0, 38, pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts)) == 0
    g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
    g.fill(38)
    g.raster(pair["map_pts"], stamp)
    assert np.array_equal(m.download(), g.download())
    other = matcher.Matcher(cases.LC["ll"], cases.LC["ur"], cases.LC["res"], cases.LC["kernel_range"])
    assert other.lib.cgm_matcher_copy_grid(other.h, 0, m.h, 0) == 0
    assert np.array_equal(other.download(), g.download())
    sub = matcher.subsample(pair["cur_pts"], 0.1)
    regions, th = cases.lc_regions([(0.0, 0.0, 0.0)])
    assert cases.same(other.greedy_search_res(sub, regions, th, 0.3, cases.BINS),
                      g.greedy_search_res(sub, regions, th, 0.3, cases.BINS))
    for x in list(ms.values()) + [other]:
        x.close()
