#!/bin/bash
# Checking aid (GPU box): the two-robot test drivers under AddressSanitizer, on the scenarios of
# tests/test_mr_exchange_gpu.py and tests/test_mr_combo_gpu.py. Result: profiles/r01_asan_two_robot_drivers.txt
set -e
cd "$(dirname "$0")/.."
L=$PWD/cg_mrslam_b200/lib
for d in mr_exchange mr_combo compat_driver; do
  /usr/bin/g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -Iinclude tests/cpp/$d.cpp -o /tmp/${d}_asan -L$L -lcgmrslam_b200 -Wl,-rpath,$L
done
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0
python - <<'PY'
import sys, subprocess, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import test_mr_exchange_gpu as ex, test_mr_combo_gpu as co
p = '/tmp/asan_mr.txt'
ex.scenario(p)
for mode in ([], ['graph']):
    r = subprocess.run(['/tmp/mr_exchange_asan', p] + mode, capture_output=True, text=True)
    print('mr_exchange', mode, 'rc', r.returncode, 'asan' if 'AddressSanitizer' in r.stderr else 'clean')
    if 'AddressSanitizer' in r.stderr: print(r.stderr[:3000])
a, b = co.two_robots(5)
with open('/tmp/asan_combo.txt', 'w') as f:
    for rr, verts in enumerate((a, b)):
        for v in verts:
            f.write("V %d %d %.17g %.17g %.17g %d %d %.17g %.17g %.17g %s\n" % (rr, v["id"], v["pose"][0], v["pose"][1], v["pose"][2], 1 if v["fixed"] else 0, len(v["ranges"]), v["first_angle"], v["step"], v["max_range"], " ".join("%.17g" % x for x in v["ranges"])))
    f.write("RUN %d 2 1 0.3\n" % a[14]["id"])
r = subprocess.run(['/tmp/mr_combo_asan', '/tmp/asan_combo.txt'], capture_output=True, text=True)
print('mr_combo rc', r.returncode, 'asan' if 'AddressSanitizer' in r.stderr else 'clean')
if 'AddressSanitizer' in r.stderr: print(r.stderr[:3000])
PY
