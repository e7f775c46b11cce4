"""N > 1 host logic on CPU: two gloo ranks shard a batch of scan-match problems (each rank runs
the product's host logic against the test-only CPU kernel stand-in, tests/hostsim) and gather the
results; the union must equal the single-process answer bit for bit."""
import math
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, math, pickle
import numpy as np
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist
import cases
from cg_mrslam_b200 import matcher, shard
HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libcgm_hostsim.so")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 5
pairs = [cases.scan_pair(300 + i, 361, math.pi, (0.3, 0.3, 0.2)) for i in range(n)]
regs = [cases.lc_regions([(0.0, 0.0, 0.0), (0.2 * i, -0.1, 0.05)])[0] for i in range(n)]
mine = shard.shard_indices(n, rank, world)
m = matcher.Matcher(cases.LC["ll"], cases.LC["ur"], cases.LC["res"], cases.LC["kernel_range"],
                    n_slots=max(len(mine), 1), lib_path=HOSTSIM)
m.raster_batch([pairs[i]["map_pts"] for i in mine])
subs = [matcher.subsample(pairs[i]["cur_pts"], 0.1, lib_path=HOSTSIM) for i in mine]
res = m.search_batch(subs, [regs[i] for i in mine], (0.1, 0.1, 0.025), 0.3, cases.BINS)
full = shard.gather_results(res, mine, n)
if rank == 0:
    pickle.dump(full, open(OUT, "wb"))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_sharding_equals_single_process(tmp_path, oracle_lib):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cg_mrslam_b200", "csrc"), "hostsim"])
    out = str(tmp_path / "full.pkl")
    script = str(tmp_path / "worker.py")
    with open(script, "w") as f:
        f.write("ROOT = %r\nOUT = %r\n" % (ROOT, out) + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                           "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                           "29531", script], env=env, timeout=300)
    import pickle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    full = pickle.load(open(out, "rb"))
    assert len(full) == 5
    stamp = oracle_lib.make_stamp(0.1, 0.5)
    for i in range(5):
        pair = cases.scan_pair(300 + i, 361, math.pi, (0.3, 0.3, 0.2))
        reg = cases.lc_regions([(0.0, 0.0, 0.0), (0.2 * i, -0.1, 0.05)])[0]
        g = oracle_lib.grid(cases.LC["ll"], cases.LC["ur"], cases.LC["res"])
        g.fill(64)
        g.raster(pair["map_pts"], stamp)
        want = g.greedy_search(oracle_lib.subsample(pair["cur_pts"], 0.1), reg,
                               (0.1, 0.1, 0.025), 0.3, cases.BINS)
        assert cases.same(full[i], want)


def test_shard_indices_cover_everything():
    from cg_mrslam_b200 import shard
    for n in (0, 1, 7, 10000):
        for world in (1, 2, 4, 8):
            got = np.sort(np.concatenate([shard.shard_indices(n, r, world) for r in range(world)]))
            assert np.array_equal(got, np.arange(n))
