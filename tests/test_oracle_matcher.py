"""Pins the CPU oracle (oracle/matcher_oracle.cpp): against the committed golden vectors produced
by the reference's chargrid.cpp compiled verbatim (tools/make_golden_matcher.py), and -- where
oracle/_ref is present -- against that library directly on fresh seeded inputs."""
import math

import numpy as np
import pytest

import cases
import golden_util as gu


@pytest.mark.parametrize("path", gu.matcher_fixtures(), ids=lambda p: p.split("matcher_")[-1][:-4])
def test_oracle_reproduces_golden(oracle_lib, path):
    fx = gu.load(path)
    stamp = oracle_lib.make_stamp(float(fx["res"]), float(fx["kernel_range"]))
    assert np.array_equal(stamp, fx["stamp"])
    digest, sub, res = gu.replay_cpu(fx, oracle_lib, stamp)
    assert digest == str(fx["grid_sha256"])
    assert np.array_equal(sub, fx["sub_pts"])
    assert cases.same(res, fx["results"])


def test_golden_fixtures_exist():
    assert len(gu.matcher_fixtures()) >= 7


def test_stamp_known_values(oracle_lib):
    # centre rows of the two stamps, verified by running the reference code.
    close = oracle_lib.make_stamp(0.025, 0.2)
    assert close.shape == (17, 17)
    assert list(close[8]) == [24, 21, 18, 15, 12, 9, 6, 3, 0, 3, 6, 9, 12, 15, 18, 21, 24]
    lc = oracle_lib.make_stamp(0.1, 0.5)
    assert lc.shape == (11, 11)
    assert list(lc[5]) == [60, 48, 36, 24, 12, 0, 12, 24, 36, 48, 60]
    assert int(lc.max()) == 64 and int(close.max()) == 25


def test_grid_geometry(oracle_lib):
    g = oracle_lib.grid((-15.0, -15.0), (15.0, 15.0), 0.025)
    assert (g.rows, g.cols) == (1200, 1200)
    g2 = oracle_lib.grid((-35.0, -35.0), (35.0, 35.0), 0.1)
    assert (g2.rows, g2.cols) == (700, 700)
    # grid2world(588, 588) on the close grid shows the float arithmetic
    x, _ = g.grid2world(588, 588)
    assert abs(x - (-0.300000191)) < 1e-8
    assert g.world2grid(0.0, 0.0) == (600, 600)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_oracle_matches_reference_fresh_inputs(oracle_lib, reference_lib, seed):
    """Fresh (non-golden) inputs: restatement == verbatim reference, bit for bit."""
    rng = np.random.default_rng(seed)
    pair = cases.scan_pair(100 + seed, 361, math.pi, (0.3, 0.3, 0.2))
    for cfg, (regions, th), ms in (
            (cases.CLOSE, cases.close_window(), 0.2),
            (cases.LC, cases.lc_regions(rng.uniform(-1, 1, (5, 3))), 0.35),
            (cases.LC, cases.lc_regions(rng.uniform(-1, 1, (9, 3)), flip=True), 0.5)):
        stamp = oracle_lib.make_stamp(cfg["res"], cfg["kernel_range"])
        outs = []
        for lib in (oracle_lib, reference_lib):
            g = lib.grid(cfg["ll"], cfg["ur"], cfg["res"])
            g.fill(int(cfg["kernel_range"] * 128))
            g.raster(pair["map_pts"], stamp)
            sub = lib.subsample(pair["cur_pts"], 0.1)
            outs.append((g.download(), sub, g.greedy_search_res(sub, regions, th, ms, cases.BINS),
                         g.count_points((-1.0, -1.0), (1.5, 2.0)),
                         g.search_non_matched(pair["cur_pts"], 0.1)))
        assert np.array_equal(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1], outs[1][1])
        assert cases.same(outs[0][2], outs[1][2])
        assert outs[0][3] == outs[1][3]
        assert np.array_equal(outs[0][4], outs[1][4])


def test_oracle_hierarchical_matches_reference(oracle_lib, reference_lib):
    pair = cases.scan_pair(77, 1081, 1.5 * math.pi, (3.0, 2.0, 1.0))
    regions, th = cases.global_window()
    stamp = oracle_lib.make_stamp(cases.LC["res"], cases.LC["kernel_range"])
    outs = []
    for lib in (oracle_lib, reference_lib):
        g = lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
        g.fill(64)
        g.raster(pair["map_pts"], stamp)
        sub = lib.subsample(pair["cur_pts"], 0.1)
        outs.append(g.hierarchical_search(sub, regions, th, 0.3, cases.BINS, 4))
    assert len(outs[0]) > 0
    assert cases.same(outs[0], outs[1])
