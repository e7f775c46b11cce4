"""Multi-GPU sharding of the two paths. One process per GPU.

* scan matching shards embarrassingly: problems (scan pairs, or one robot's loop-closure sweep)
  are dealt round-robin to ranks; there is NO collective on the data path. Only the tiny result
  lists are gathered at the end (`gather_results`), which replaces the reference's UDP exchange of
  closures between robots (src/mrslam/graph_comm.cpp) on one NVSwitch box.
* the pose-graph solve runs as independent replicas (one graph per rank: one robot per GPU, as in
  BASELINE config 3), or as ONE graph cut over the ranks with a single all-reduce of the separator
  blocks per iteration (pgo_set_partition / pgo_dd_* in include/pgo_solver.h, pgo.optimize_distributed).
"""
import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin assignment: item i belongs to rank i % world."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def gather_results(local_results, local_indices, n_items, group=None):
    """Collect per-problem result arrays ([k_i, 4] doubles: x, y, theta, score) from all ranks.

    Every rank passes the results of the problems it owns (`local_indices`, same order). Returns
    the full list of n_items arrays on every rank. Works on any torch.distributed backend (NCCL on
    the GPU box, gloo in the CPU tests): the payload is padded to the largest result count and
    exchanged with all_gather on plain tensors."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        out = [None] * n_items
        for i, r in zip(local_indices, local_results):
            out[int(i)] = np.asarray(r, dtype=np.float64).reshape(-1, 4)
        return out
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    per_rank = (n_items + world - 1) // world
    k_local = max([len(r) for r in local_results], default=0)
    k_max = torch.tensor([k_local], dtype=torch.int64, device=dev)
    dist.all_reduce(k_max, op=dist.ReduceOp.MAX, group=group)
    k = max(int(k_max.item()), 1)
    buf = torch.zeros((per_rank, k, 4), dtype=torch.float64)
    cnt = torch.full((per_rank,), -1, dtype=torch.int64)
    for slot, r in enumerate(local_results):
        r = np.asarray(r, dtype=np.float64).reshape(-1, 4)
        cnt[slot] = len(r)
        if len(r):
            buf[slot, :len(r)] = torch.from_numpy(r)
    buf, cnt = buf.to(dev), cnt.to(dev)
    all_buf = [torch.empty_like(buf) for _ in range(world)]
    all_cnt = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(all_buf, buf, group=group)
    dist.all_gather(all_cnt, cnt, group=group)
    out = [None] * n_items
    for rk in range(world):
        idx = shard_indices(n_items, rk, world)
        b, c = all_buf[rk].cpu().numpy(), all_cnt[rk].cpu().numpy()
        for slot, i in enumerate(idx):
            out[int(i)] = b[slot, :c[slot]].copy()
    return out
