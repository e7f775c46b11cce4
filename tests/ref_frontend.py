"""Drivers built from the REFERENCE'S OWN host sources (oracle/Makefile, target `frontend`):
`<name>_gpu` = reference sources + include/g2o_compat + include/cgm/chargrid.hpp +
libcgmrslam_b200.so; `<name>_cpu` = the same sources + the reference's chargrid.cpp + the CPU
oracle solver. They are compiled where /root/reference exists (this container, by
__graft_entry__.build()) into oracle/_ref/, which travels to the GPU box with the snapshot."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def driver_path(name, kind):
    assert kind in ("gpu", "cpu")
    exe = os.path.join(REF_DIR, "%s_%s" % (name, kind))
    if os.path.isdir("/root/reference/src"):
        import __graft_entry__ as g
        if not os.path.exists(os.path.join(ROOT, "cg_mrslam_b200", "lib", "libcgmrslam_b200.so")):
            g.build()
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle", "frontend"])
    if not os.path.exists(exe):
        pytest.skip("%s was not prebuilt (needs /root/reference at build time)" % exe)
    return exe


def run_driver(exe, args, timeout=900):
    """Runs a driver; returns (result lines between BEGIN and END, stderr). The reference's own
    chatter stays on the process' stdout; the driver's result lines come through $CGM_OUT."""
    import tempfile
    with tempfile.NamedTemporaryFile("r", suffix=".out") as tf:
        env = dict(os.environ, CGM_OUT=tf.name)
        out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout,
                             env=env)
        assert out.returncode == 0, (out.returncode, out.stderr[-3000:])
        lines = tf.read().splitlines()
    assert "BEGIN" in lines and lines[-1] == "END", (lines[-5:], out.stderr[-2000:])
    return lines[lines.index("BEGIN") + 1:-1], out.stderr
