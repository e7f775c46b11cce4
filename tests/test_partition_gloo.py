"""Host side of the domain-decomposed solve on CPU: the partition is computed independently by two
gloo ranks and must be identical, separate the ranks' interiors, and cover every edge once."""
import os
import subprocess
import sys

import numpy as np

from cg_mrslam_b200 import pgo, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys
import numpy as np
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from cg_mrslam_b200 import pgo, synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = synth.make_pose_graph(1500, 6000, seed=17, box=43.0)
owner, st = pgo.analyse_partition(1500, g["edge_ij"], g["fixed"], world)
mine = torch.from_numpy(owner.astype(np.int64))
both = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(both, mine)
assert all(torch.equal(b, mine) for b in both), "ranks disagree on the partition"
# edges this rank would linearise: interior endpoint owner, else edge index % world
e = g["edge_ij"]
oi, oj = owner[e[:, 0]], owner[e[:, 1]]
edge_owner = np.where(oi >= 0, oi, np.where(oj >= 0, oj, np.arange(len(e)) % world))
count = torch.tensor([int((edge_owner == rank).sum())])
dist.all_reduce(count)
assert int(count) == len(e), (int(count), len(e))
dist.barrier()
dist.destroy_process_group()
'''


def test_partition_is_consistent_across_ranks(tmp_path):
    script = str(tmp_path / "worker.py")
    with open(script, "w") as f:
        f.write("ROOT = %r\n" % ROOT + WORKER)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                           "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                           "29533", script], timeout=300)


def test_partition_properties():
    g = synth.make_pose_graph(4000, 16000, seed=19, box=70.0)
    e = g["edge_ij"]
    for world in (1, 2, 4, 8):
        owner, st = pgo.analyse_partition(4000, e, g["fixed"], world)
        assert owner[0] == -2 and (owner[1:] >= -1).all() and owner.max() < world
        oi, oj = owner[e[:, 0]], owner[e[:, 1]]
        both = (oi >= 0) & (oj >= 0)
        assert (oi[both] == oj[both]).all()          # no edge joins two ranks' interiors
        assert st["shared_vertices"] == int((owner == -1).sum())
        if world > 1:
            sizes = np.bincount(owner[owner >= 0], minlength=world)
            assert sizes.min() > 0 and st["shared_vertices"] < 0.2 * 4000
