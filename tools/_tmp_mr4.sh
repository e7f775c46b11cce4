mkdir -p gpurun_out /tmp/mr
ARGS="--bag 4robots --robots 4 --keyframes 200 --min-inliers-mr 3"
( time python tests/mr_replay.py --kind gpu --out /tmp/mr/single.npz $ARGS ) > gpurun_out/r02_mr_replay_single.log 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/mr_replay.py --kind gpu --dist --out /tmp/mr/dist.npz $ARGS ) > gpurun_out/r02_mr_replay_4gpu.log 2>&1
python - <<'PY' > gpurun_out/r02_mr_replay_4gpu.txt 2>&1
import sys, numpy as np
sys.path.insert(0, 'tests')
import test_mr_replay as t
a = np.load('/tmp/mr/single.npz')
print("cfg 3: 4 robots of 4robots-hospital.bag, 200 keyframes each, reference MRGraphSLAM over libcgmrslam_b200.so")
print("single process (1 GPU): %d datagrams" % len(a['msgs']))
for r in range(4):
    z = np.load('/tmp/mr/dist.rank%d.npz' % r)
    same_msgs = np.array_equal(z['msgs'], a['msgs'])
    try:
        worst = t._same_graph(z, a, r, 1e-6)
        print("rank %d (GPU %d, NCCL all-gather of datagrams): same traffic %s, same edge set, max |difference| vs single process %.3e" % (r, r, same_msgs, worst))
    except AssertionError as e:
        print("rank %d: graphs differ in structure (a coincident-keyframe tie or an fp64-atomics flip): %r" % (r, e))
    e = z['edges%d' % r]
    print("   vertices %d (peers' %d), edges %d (inter-robot %d, condensed out-stars %d)" % (
        len(z['vertices%d' % r]), int((z['vertices%d' % r][:, 0] // 10000 != r).sum()), len(e),
        int(((e[:, 0] // 10000 != r) | (e[:, 1] // 10000 != r)).sum()), int((e[:, 6] > 0).sum())))
PY
grep real gpurun_out/r02_mr_replay_single.log gpurun_out/r02_mr_replay_4gpu.log >> gpurun_out/r02_mr_replay_4gpu.txt
cat gpurun_out/r02_mr_replay_4gpu.txt
(timeout 600 python -m pytest tests/test_pgo_multigpu.py -m gpu -q 2>&1 | tail -3) > gpurun_out/r02_pytest_multigpu_4.log; tail -2 gpurun_out/r02_pytest_multigpu_4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 4 --steps 5 --warmup 3 --no-bag > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
tail -c 300 gpurun_out/r02_bench_4gpu.err
