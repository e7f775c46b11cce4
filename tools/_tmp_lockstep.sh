set -e
mkdir -p gpurun_out
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'tests')
import replay_util
fx=np.load('tests/golden/bag_2robots_robot1_full.npz')
replay_util.write_keyframes('/tmp/kf1.txt', fx, 130)
PY
CGM_OUT=gpurun_out/ls1_gpu.out ./oracle/_ref/ref_replay_gpu /tmp/kf1.txt - 1 130 --dump /tmp/st1.bin > /tmp/ls1_gpu.stdout 2>&1
CGM_OUT=gpurun_out/ls1_cpu.out ./oracle/_ref/ref_replay_cpu /tmp/kf1.txt - 1 130 --follow /tmp/st1.bin > /tmp/ls1_cpu.stdout 2>&1
awk '/Current vertex: 10111 /,/Current vertex: 10112 /' /tmp/ls1_gpu.stdout | grep -v "^$" | head -60 > gpurun_out/ls1_gpu_ctx.txt
awk '/Current vertex: 10111 /,/Current vertex: 10112 /' /tmp/ls1_cpu.stdout | grep -v "^$" | head -60 > gpurun_out/ls1_cpu_ctx.txt
grep -c . gpurun_out/ls1_gpu.out gpurun_out/ls1_cpu.out
