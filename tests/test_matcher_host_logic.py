"""CPU tests of the PRODUCT's host logic (matcher_plan.cpp + matcher_api.cpp): the C ABI is
linked against tests/hostsim (a test-only CPU stand-in for the kernels) and must reproduce the
golden vectors and the oracle bit for bit. The kernels themselves are tested in
test_matcher_gpu.py on a B200."""
import math
import os
import subprocess

import numpy as np
import pytest

import cases
import golden_util as gu
from cg_mrslam_b200 import matcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libcgm_hostsim.so")


@pytest.fixture(scope="module")
def hostsim():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cg_mrslam_b200", "csrc"),
                           "hostsim"])
    return HOSTSIM


def _factory(lib_path, n_slots=1):
    return lambda cfg: matcher.Matcher(cfg["ll"], cfg["ur"], cfg["res"], cfg["kernel_range"],
                                       n_slots=n_slots, lib_path=lib_path)


@pytest.mark.parametrize("path", gu.matcher_fixtures(), ids=lambda p: p.split("matcher_")[-1][:-4])
def test_host_logic_reproduces_golden(hostsim, path):
    fx = gu.load(path)
    if fx["name"] == "bench_1081_raw":
        pytest.skip("1.01M candidates: covered on the GPU; too slow for the CPU stand-in")
    digest, sub, res = gu.replay_product(fx, _factory(hostsim),
                                         lambda p, r: matcher.subsample(p, r, lib_path=hostsim))
    assert digest == str(fx["grid_sha256"])
    assert np.array_equal(sub, fx["sub_pts"])
    assert cases.same(res, fx["results"])


def test_batch_equals_single(hostsim, oracle_lib):
    pairs = [cases.scan_pair(20 + i, 361, math.pi, (0.3, 0.3, 0.2)) for i in range(3)]
    rng = np.random.default_rng(5)
    regs = [cases.lc_regions(rng.uniform(-1, 1, (n, 3)))[0] for n in (1, 5, 2)]
    m = _factory(hostsim, n_slots=4)(cases.LC)
    m.raster_batch([p["map_pts"] for p in pairs], first_slot=1)
    subs = [matcher.subsample(p["cur_pts"], 0.1, lib_path=hostsim) for p in pairs]
    res = m.search_batch(subs, regs, (0.1, 0.1, 0.025), 0.3, cases.BINS, first_slot=1)
    stamp = oracle_lib.make_stamp(0.1, 0.5)
    for i in range(3):
        g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
        g.fill(64)
        g.raster(pairs[i]["map_pts"], stamp)
        want = g.greedy_search(subs[i], regs[i], (0.1, 0.1, 0.025), 0.3, cases.BINS)
        assert cases.same(res[i], want)
        assert np.array_equal(m.download(slot=1 + i), g.download())
    st = m.batch_stats()
    assert 0.9 < st["candidates"] / (sum(len(r) for r in regs) * 10 * 30 * 65.0) < 1.1
    m.close()


def test_edge_cases(hostsim, oracle_lib):
    m = _factory(hostsim)(cases.LC)
    m.raster_batch([np.zeros((0, 2))])                      # empty map: all cells = fill
    assert int(m.download().min()) == 64 and int(m.download().max()) == 64
    reg, th = cases.lc_regions([(0, 0, 0)])
    # empty scan: k = 0 => score = maxScore + 1, nothing accepted (chargrid.cpp:276)
    assert len(m.greedy_search_res(np.zeros((0, 2)), reg, th, 0.5, cases.BINS)) == 0
    # no regions: the reference would divide by zero (chargrid.cpp:225); we return nothing
    assert len(m.greedy_search_res(np.ones((3, 2)), np.zeros((0, 6)), th, 0.5, cases.BINS)) == 0
    # inverted region: zero candidates
    bad = reg.copy()
    bad[0, 3] = bad[0, 0] - 1.0
    assert len(m.greedy_search_res(np.ones((3, 2)), bad, th, 0.5, cases.BINS)) == 0
    # window partly outside the grid: out-of-grid cells add 0 but count in k
    pts = np.array([[0.0, 0.0], [1.0, 0.5], [2.0, -1.0], [2.0, -1.0]])
    edge = np.array([[33.0, 33.0, -0.1, 36.5, 36.5, 0.1]], dtype=np.float32)
    got = m.greedy_search_res(pts, edge, th, 0.6, cases.BINS)
    g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
    g.fill(64)
    want = g.greedy_search_res(pts, edge, th, 0.6, cases.BINS)
    assert len(want) > 0 and cases.same(got, want)
    # argument errors
    with pytest.raises(matcher.MatcherError):
        m.greedy_search(pts, edge, (0.1, 0.1, 0.0), 0.5, cases.BINS)
    with pytest.raises(matcher.MatcherError):
        m.raster(pts, slot=3)
    m.close()


def test_cold_paths(hostsim, oracle_lib):
    pair = cases.scan_pair(31, 361, math.pi)
    m = _factory(hostsim)(cases.CLOSE)
    m.raster_batch([pair["map_pts"]])
    g = oracle_lib.grid(cases.CLOSE["ll"], cases.CLOSE["ur"], cases.CLOSE["res"])
    g.fill(25)
    g.raster(pair["map_pts"], oracle_lib.make_stamp(0.025, 0.2))
    for ll, ur in (((-1.0, -1.0), (1.5, 2.0)), ((-16.0, -3.0), (-14.0, 3.0))):
        assert m.count_points(ll, ur) == g.count_points(ll, ur)
    assert np.array_equal(m.search_non_matched(pair["cur_pts"], 0.05),
                          g.search_non_matched(pair["cur_pts"], 0.05))
    assert m.world2grid(0.3, -0.7) == g.world2grid(0.3, -0.7)
    assert m.grid2world(588, 12) == g.grid2world(588, 12)
    m.close()


def test_incremental_raster(hostsim, oracle_lib):
    """addAndConvolvePoints without reset accumulates (min) over calls."""
    pair = cases.scan_pair(32, 361, math.pi)
    m = _factory(hostsim)(cases.LC)
    m.reset()
    half = len(pair["map_pts"]) // 2
    m.raster(pair["map_pts"][:half])
    m.raster(pair["map_pts"][half:])
    g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
    g.fill(64)
    g.raster(pair["map_pts"], oracle_lib.make_stamp(0.1, 0.5))
    assert np.array_equal(m.download(), g.download())
    m.close()
