#!/bin/bash
# Checking aid (no GPU needed): the reference's front-end sources + the test drivers of tests/cpp,
# CPU build (reference chargrid.cpp + CPU oracle solver), under AddressSanitizer: a 60-keyframe bag
# replay as lockstep leader and follower (g2o::trace), and a 4-robot message exchange.
# Needs /root/reference. Result: profiles/r02_asan_frontend_drivers.txt
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD
REF=${REFERENCE:-/root/reference}
OUT=/tmp/asan_frontend
mkdir -p $OUT
make -s -C oracle oracle
FLAGS="-std=gnu++14 -O1 -g -fPIC -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -w -I$ROOT/include/ref_names -I$REF/src -I$ROOT/tests/cpp -fopenmp -I$REF/src/matcher -include $ROOT/oracle/ref_array_allocator_fix.h"
SRCS="slam/graph_slam.cpp slam/graph_manipulator.cpp slam/vertices_finder.cpp slam/closure_checker.cpp slam/closure_buffer.cpp matcher/scan_matcher.cpp mrslam/mr_graph_slam.cpp mrslam/mr_closure_buffer.cpp mrslam/msg_factory.cpp mrslam/condensed_graph/condensed_graph_buffer.cpp mrslam/condensed_graph/condensed_graph_creator.cpp matcher/chargrid.cpp"
OBJS=""
for s in $SRCS; do
  o=$OUT/$(echo $s | tr / _).o
  g++ $FLAGS -c $REF/src/$s -o $o &
  OBJS="$OBJS $o"
done
wait
LINK="-L$ROOT/oracle/_build -loracle_pgo -Wl,-rpath,$ROOT/oracle/_build -lpthread"
g++ $FLAGS tests/cpp/ref_replay.cpp $OBJS -o $OUT/ref_replay_asan $LINK
g++ $FLAGS -shared tests/cpp/ref_robot.cpp $OBJS -o $OUT/libref_robot_asan.so $LINK
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0
python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, "tests")
import numpy as np
import replay_util
out = "/tmp/asan_frontend"
fx = np.load("tests/golden/bag_2robots_robot0_full.npz")
replay_util.write_keyframes(out + "/kf.txt", fx, 150)
for mode in ("--dump", "--follow"):
    r = subprocess.run([out + "/ref_replay_asan", out + "/kf.txt", "-", "0", "150", mode, out + "/trace.bin"],
                       capture_output=True, text=True, env=dict(os.environ, CGM_OUT=out + "/res.txt"))
    print("ref_replay", mode, "150 keyframes: rc", r.returncode, "AddressSanitizer report" if "AddressSanitizer" in r.stderr else "clean")
    if "AddressSanitizer" in r.stderr:
        print(r.stderr[-3000:])
PY
# four robots exchanging Combo / CondensedGraph messages (60 keyframes each, quorum 3), leader and follower
ASAN_LIB=$(g++ -print-file-name=libasan.so)
for mode in --dump --follow; do
  if LD_PRELOAD=$ASAN_LIB CGM_ROBOT_LIB=$OUT/libref_robot_asan.so python tests/mr_replay.py --kind cpu --bag 4robots --robots 4 \
       --keyframes 60 --min-inliers-mr 3 --out $OUT/mr.npz $mode $OUT/mrtrace > $OUT/mr.log 2> $OUT/mr.err; then rc=0; else rc=$?; fi
  if grep -q AddressSanitizer $OUT/mr.err; then echo "mr_replay $mode 4 x 60 keyframes: rc $rc AddressSanitizer report"; grep -A25 AddressSanitizer $OUT/mr.err | head -60;
  else echo "mr_replay $mode 4 x 60 keyframes: rc $rc clean"; fi
done
