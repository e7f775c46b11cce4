set -e
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'tests')
import replay_util
fx=np.load('tests/golden/bag_2robots_robot0_full.npz')
replay_util.write_keyframes('/tmp/kf0.txt', fx, 100000)
PY
CGM_OUT=gpurun_out/replay0_gpu.out ./oracle/_ref/ref_replay_gpu /tmp/kf0.txt - 0 > /tmp/replay0.stdout 2> /tmp/replay0.stderr || true
grep -n "cgm:\|SparseOptimizer\|optimize:\|computeMarginals\|labelEdges" /tmp/replay0.stderr | head -20 > gpurun_out/replay0_errs.txt
wc -l /tmp/replay0.stderr >> gpurun_out/replay0_errs.txt
grep -n "Current vertex: 64[6-9]\|Current vertex: 65[0-2]" /tmp/replay0.stdout | head > gpurun_out/replay0_ctx.txt
awk '/Current vertex: 648/,/Current vertex: 652/' /tmp/replay0.stdout | head -300 >> gpurun_out/replay0_ctx.txt
tail -3 gpurun_out/replay0_gpu.out
