"""Golden outputs of the bag replays: the reference's own GraphSLAM (compiled verbatim) over the
reference's CPU matcher and the CPU oracle solver (oracle/_ref/ref_replay_cpu, built by
`make -C oracle frontend`), run HERE on the keyframe fixtures tests/golden/bag_2robots_*_full.npz.

    python tools/make_golden_replay.py            # all fixtures (minutes per robot)

Writes tests/golden/replay_<fixture>_cpu.txt.gz: the driver's result lines (K keyframe estimates,
E new graph edges, C new closure candidates, P final estimates). tests/test_ref_frontend.py compares
the GPU build of the same driver (oracle/_ref/ref_replay_gpu) with them. Test-data tooling only."""
import glob
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import replay_util  # noqa: E402


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_replay_cpu")
    for fx_path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "bag_2robots_*_full.npz"))):
        name = os.path.basename(fx_path)[len("bag_"):-len("_full.npz")]
        robot = int(name.rsplit("robot", 1)[1])
        fx = np.load(fx_path)
        with tempfile.TemporaryDirectory() as d:
            kf = os.path.join(d, "kf.txt")
            replay_util.write_keyframes(kf, fx, len(fx["odom"]), min_inliers=7)
            res = os.path.join(d, "out.txt")
            subprocess.check_call([exe, kf, "-", str(robot)], env=dict(os.environ, CGM_OUT=res),
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            lines = [ln for ln in open(res).read().splitlines() if not ln.startswith(("TIMES_MS", "T ", "TRACE", "SOLVER_MS"))]
        out = os.path.join(ROOT, "tests", "golden", "replay_%s_cpu.txt.gz" % name)
        with gzip.open(out, "wt") as f:
            f.write("\n".join(lines) + "\n")
        print(out, len(lines), "lines")


if __name__ == "__main__":
    main()
