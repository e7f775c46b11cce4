"""Real multi-process run of the domain-decomposed solve: one process per GPU under torchrun, the
all-reduce done by NCCL. Needs >= 2 GPUs (skipped on a single-GPU box; the same kernels are
covered there with virtual ranks in test_pgo_gpu.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from cg_mrslam_b200 import pgo, synth
from oracle import pgo_oracle as po
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
g = synth.make_pose_graph(3000, 12000, seed=5, box=61.0)
stream = torch.cuda.Stream()
s = pgo.Solver(device=local, stream=stream.cuda_stream)
s.set_partition(rank, world)
s.set_graph(3000, g["edge_ij"], g["fixed"])
s.upload(g["poses0"], g["meas"], g["info"])
def all_reduce(ts):
    with torch.cuda.stream(stream):
        dist.all_reduce(ts[0])
done, chi2 = pgo.optimize_distributed([s], 4, all_reduce)
ref = po.gauss_newton(g["poses0"], g["edge_ij"], g["meas"], g["info"], g["fixed"], 4)
d = s.poses() - ref.poses
d[:, 2] = po.normalize_theta(d[:, 2])
assert done == 4 and np.abs(d).max() < 1e-6, (done, np.abs(d).max())
assert np.allclose(chi2, ref.chi2, rtol=1e-9)
# the same with the loop and the all-reduce INSIDE the library (pgo_dd_iterate, NCCL bound by dlopen)
def bcast(b):
    t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())
s2 = pgo.Solver(device=local)
s2.comm_init(rank, world, bcast)
s2.set_graph(3000, g["edge_ij"], g["fixed"])
s2.upload(g["poses0"], g["meas"], g["info"])
done2, chi2b = s2.optimize_dd(4)
d2 = s2.poses() - ref.poses
d2[:, 2] = po.normalize_theta(d2[:, 2])
assert done2 == 4 and np.abs(d2).max() < 1e-6, (done2, np.abs(d2).max())
assert np.allclose(chi2b, ref.chi2, rtol=1e-9)
dist.barrier()
if rank == 0:
    print("DD_OK", world, float(np.abs(d).max()), float(np.abs(d2).max()))
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_domain_decomposition_over_nccl(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    script = str(tmp_path / "worker.py")
    with open(script, "w") as f:
        f.write("ROOT = %r\n" % ROOT + WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
                          "--master-port", "29535", script], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DD_OK %d" % world in out.stdout
