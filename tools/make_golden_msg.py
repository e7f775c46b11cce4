"""Golden datagrams of the inter-robot wire format, produced by the REFERENCE's own msg_factory.cpp
compiled verbatim (oracle/_ref/libref_msg.so; needs /root/reference at build time):

    python tools/make_golden_msg.py      ->  tests/golden/msg_wire.npz

One seeded message of every kind in a small and a large variant; the fixture holds the inputs
(as float64 / int32 arrays) and the reference's bytes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import msg_wire_util as mw  # noqa: E402

ref = mw.Wire(mw.REF_SO)
rng = np.random.default_rng(20261017)
out = {}
n = 0
for big in (False, True):
    for kind in mw.TYPES:
        m = mw.random_message(kind, rng, big=big)
        data = ref.pack(m)
        assert data is not None
        back = ref.unpack(data)
        assert back is not None and mw.same_fields(back, mw.expected_after_wire(m)), kind
        key = "m%02d" % n
        out[key + "_kind"] = np.array(kind)
        out[key + "_robot"] = np.array(m["robot"])
        for f in ("vid", "vest", "readings", "laser4", "eft", "eest", "einfo", "closures"):
            if f in m:
                out[key + "_" + f] = np.asarray(m[f])
        if "node_id" in m:
            out[key + "_node_id"] = np.array(m["node_id"])
        out[key + "_bytes"] = np.frombuffer(data, dtype=np.uint8)
        n += 1
out["count"] = np.array(n)
path = os.path.join(ROOT, "tests", "golden", "msg_wire.npz")
np.savez_compressed(path, **out)
print("wrote", path, n, "messages", os.path.getsize(path), "bytes")
