"""ctypes access to the flat message interface of oracle/msg_shim.cpp (TEST INFRASTRUCTURE), which
exists in two builds: the reference's msg_factory.cpp compiled verbatim (oracle/_ref/libref_msg.so)
and this repository's include/cgm/msg_factory.hpp (build/libcgm_msg_shim.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_msg.so")
PRODUCT_SO = os.path.join(ROOT, "build", "libcgm_msg_shim.so")
TYPES = {"vertices": 1, "laser": 2, "combo": 4, "edges": 5, "closures": 6, "condensed": 7, "graph": 8}

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def build_product():
    os.makedirs(os.path.dirname(PRODUCT_SO), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "include", "cgm", "ref_names"),
                           os.path.join(ROOT, "oracle", "msg_shim.cpp"), "-o", PRODUCT_SO])
    return PRODUCT_SO


class Wire:
    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.msg_pack.restype = C.c_int
        self.lib.msg_unpack.restype = C.c_int

    def pack(self, msg, bsize=100000):
        """msg: dict with type, robot, and the fields its type carries. Returns bytes or None."""
        vid = np.ascontiguousarray(msg.get("vid", np.zeros(0)), dtype=np.int32)
        vest = np.ascontiguousarray(msg.get("vest", np.zeros((0, 3))), dtype=np.float64)
        readings = np.ascontiguousarray(msg.get("readings", np.zeros(0)), dtype=np.float64)
        laser4 = np.ascontiguousarray(msg.get("laser4", np.zeros(4)), dtype=np.float64)
        eft = np.ascontiguousarray(msg.get("eft", np.zeros((0, 2))), dtype=np.int32)
        eest = np.ascontiguousarray(msg.get("eest", np.zeros((0, 3))), dtype=np.float64)
        einfo = np.ascontiguousarray(msg.get("einfo", np.zeros((0, 6))), dtype=np.float64)
        closures = np.ascontiguousarray(msg.get("closures", np.zeros(0)), dtype=np.int32)
        buf = C.create_string_buffer(max(bsize, 1))
        n = self.lib.msg_pack(
            TYPES[msg["type"]], int(msg["robot"]), len(vid), vid.ctypes.data_as(_ip), vest.ctypes.data_as(_dp),
            int(msg.get("node_id", -1)), len(readings), readings.ctypes.data_as(_dp), laser4.ctypes.data_as(_dp),
            len(eft), eft.ctypes.data_as(_ip), eest.ctypes.data_as(_dp), einfo.ctypes.data_as(_dp),
            len(closures), closures.ctypes.data_as(_ip), buf, bsize)
        return None if n < 0 else buf.raw[:n]

    def unpack(self, data, cap=4096):
        counts = np.array([cap, cap, cap, cap, 0, 0], dtype=np.int32)
        vid = np.zeros(cap, dtype=np.int32)
        vest = np.zeros((cap, 3))
        node_id = C.c_int(-1)
        readings = np.zeros(cap)
        laser4 = np.zeros(4)
        eft = np.zeros((cap, 2), dtype=np.int32)
        eest = np.zeros((cap, 3))
        einfo = np.zeros((cap, 6))
        closures = np.zeros(cap, dtype=np.int32)
        rc = self.lib.msg_unpack(data, len(data), counts.ctypes.data_as(_ip), vid.ctypes.data_as(_ip),
                                 vest.ctypes.data_as(_dp), C.byref(node_id), readings.ctypes.data_as(_dp),
                                 laser4.ctypes.data_as(_dp), eft.ctypes.data_as(_ip), eest.ctypes.data_as(_dp),
                                 einfo.ctypes.data_as(_dp), closures.ctypes.data_as(_ip))
        if rc < 0:
            return None
        t, robot, nv, nr, ne, nc = (int(v) for v in counts)
        name = {v: k for k, v in TYPES.items()}[t]
        return {"type": name, "robot": robot, "vid": vid[:nv].copy(), "vest": vest[:nv].copy(),
                "node_id": node_id.value, "readings": readings[:nr].copy(), "laser4": laser4.copy(),
                "eft": eft[:ne].copy(), "eest": eest[:ne].copy(), "einfo": einfo[:ne].copy(),
                "closures": closures[:nc].copy(), "consumed": rc}


HAS = {"vertices": ("v",), "laser": ("l",), "combo": ("v", "l"), "edges": ("e",), "closures": ("c",),
       "condensed": ("e", "c"), "graph": ("v", "e", "c")}


def random_message(kind, rng, robot=None, big=False):
    """A seeded message of one kind with values whose float32 rounding is visible."""
    m = {"type": kind, "robot": int(rng.integers(0, 8)) if robot is None else robot}
    parts = HAS[kind]
    if "v" in parts:
        n = int(rng.integers(0, 40 if big else 6))
        base = 10000 * m["robot"]
        m["vid"] = (base + rng.integers(0, 5000, size=n)).astype(np.int32)
        m["vest"] = np.column_stack([rng.uniform(-80, 80, n), rng.uniform(-80, 80, n),
                                     rng.uniform(-np.pi, np.pi, n)]).reshape(n, 3)
    if "l" in parts:
        n = int(rng.integers(0, 1081 if big else 12))
        m["node_id"] = int(rng.integers(0, 90000))
        m["readings"] = rng.uniform(0.02, 30.0, n)
        m["laser4"] = np.array([-np.pi / 2, np.pi / 360, 8.0 + rng.uniform(0, 1), 0.01])
    if "e" in parts:
        n = int(rng.integers(0, 30 if big else 5))
        m["eft"] = rng.integers(0, 90000, size=(n, 2)).astype(np.int32)
        m["eest"] = np.column_stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(-np.pi, np.pi, n)]).reshape(n, 3)
        a = rng.normal(size=(n, 3, 3))
        om = np.einsum("nij,nkj->nik", a, a) * 100.0 + np.eye(3) * 10.0
        m["einfo"] = np.stack([om[:, 0, 0], om[:, 0, 1], om[:, 0, 2], om[:, 1, 1], om[:, 1, 2], om[:, 2, 2]],
                              axis=1).reshape(n, 6)
    if "c" in parts:
        n = int(rng.integers(0, 25 if big else 5))
        m["closures"] = rng.integers(0, 90000, size=n).astype(np.int32)
    return m


def expected_after_wire(m):
    """What a receiver must see: ints unchanged, every double rounded through float32."""
    f = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)
    out = {"type": m["type"], "robot": m["robot"]}
    parts = HAS[m["type"]]
    if "v" in parts:
        out["vid"], out["vest"] = np.asarray(m["vid"]), f(m["vest"])
    if "l" in parts:
        out["node_id"], out["readings"], out["laser4"] = m["node_id"], f(m["readings"]), f(m["laser4"])
    if "e" in parts:
        out["eft"], out["eest"], out["einfo"] = np.asarray(m["eft"]), f(m["eest"]), f(m["einfo"])
    if "c" in parts:
        out["closures"] = np.asarray(m["closures"])
    return out


def same_fields(got, want):
    for k, v in want.items():
        g = got[k]
        if isinstance(v, np.ndarray):
            if g.shape != v.shape or not np.array_equal(g, v):
                return False
        elif g != v:
            return False
    return True
