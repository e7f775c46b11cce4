"""Inter-robot wire format (src/mrslam/msg_factory.{h,cpp}): this repository's
message classes (include/cgm/msg_factory.hpp) against golden datagrams written by the reference's
own msg_factory.cpp, and -- where the verbatim reference build exists -- against that build on
fresh seeded messages. Bar: identical bytes; identical fields after the float32 round trip."""
import os
import struct

import numpy as np
import pytest

import msg_wire_util as mw

GOLDEN = os.path.join(mw.ROOT, "tests", "golden", "msg_wire.npz")


@pytest.fixture(scope="module")
def product():
    return mw.Wire(mw.build_product())


def golden_messages():
    z = np.load(GOLDEN, allow_pickle=False)
    for i in range(int(z["count"])):
        key = "m%02d_" % i
        m = {"type": str(z[key + "kind"]), "robot": int(z[key + "robot"])}
        for f in ("vid", "vest", "readings", "laser4", "eft", "eest", "einfo", "closures"):
            if key + f in z.files:
                m[f] = z[key + f]
        if key + "node_id" in z.files:
            m["node_id"] = int(z[key + "node_id"])
        yield m, z[key + "bytes"].tobytes()


def test_golden_bytes(product):
    n = 0
    for m, want in golden_messages():
        got = product.pack(m)
        assert got == want, (m["type"], len(got or b""), len(want))
        n += 1
    assert n == 14


def test_golden_parse(product):
    for m, data in golden_messages():
        back = product.unpack(data)
        assert back is not None and back["consumed"] == len(data)
        assert mw.same_fields(back, mw.expected_after_wire(m)), m["type"]


def test_layout_by_hand(product):
    """The layout spelled out: int32 type, int32 robot, size_t count, then int32 ids and float32s."""
    m = {"type": "condensed", "robot": 3, "eft": [[10007, 10042]], "eest": [[1.25, -2.5, 0.1]],
         "einfo": [[1000.0, 0.5, 0.25, 1000.0, 0.125, 10000.0]], "closures": [20011, 20012]}
    want = struct.pack("<ii", 7, 3) + struct.pack("<Q", 1) + struct.pack("<ii", 10007, 10042) + \
        struct.pack("<3f", 1.25, -2.5, 0.1) + struct.pack("<6f", 1000.0, 0.5, 0.25, 1000.0, 0.125, 10000.0) + \
        struct.pack("<Q", 2) + struct.pack("<ii", 20011, 20012)
    assert product.pack(m) == want
    back = product.unpack(want)
    assert back["eest"][0, 2] == float(np.float32(0.1)) != 0.1   # the float32 rounding is visible


def test_buffer_too_small_and_bad_input(product):
    rng = np.random.default_rng(5)
    m = mw.random_message("graph", rng, big=True)
    full = product.pack(m)
    assert full is not None
    assert product.pack(m, bsize=len(full)) == full
    for cut in (len(full) - 1, len(full) // 2, 8, 7, 0):
        assert product.pack(m, bsize=cut) is None          # toCharArray returns 0
    assert product.unpack(full[:-1]) is None               # truncated datagram
    assert product.unpack(full + b"\0") is None            # trailing bytes
    assert product.unpack(struct.pack("<ii", 3, 0) + full[8:]) is None   # type 3 is not registered
    huge = struct.pack("<ii", 6, 0) + struct.pack("<Q", 1 << 60)       # count no datagram can back
    assert product.unpack(huge) is None


def test_garbled_datagrams_never_crash(product):
    """Datagrams arrive from the network: truncated, extended and bit-flipped versions of valid
    messages, and pure noise, must parse to a message of exactly the datagram's length or be
    refused -- never crash, never read past the buffer (counts are checked against the bytes left)."""
    rng = np.random.default_rng(99)
    valid = [product.pack(mw.random_message(k, rng, big=b)) for k in mw.TYPES for b in (False, True)]
    accepted = refused = 0
    for trial in range(3000):
        d = bytearray(valid[trial % len(valid)])
        kind = trial % 4
        if kind == 0 and len(d) > 1:
            d = d[:int(rng.integers(0, len(d)))]
        elif kind == 1:
            d += bytes(rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8))
        elif kind == 2:
            for _ in range(int(rng.integers(1, 6))):
                d[int(rng.integers(0, len(d)))] ^= 1 << int(rng.integers(0, 8))
        else:
            d = bytearray(rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8))
            if len(d) >= 4 and trial % 8 == 3:
                d[:4] = struct.pack("<i", [1, 2, 4, 5, 6, 7, 8][trial % 7])
        back = product.unpack(bytes(d), cap=1 << 16)
        if back is None:
            refused += 1
        else:
            accepted += 1
            assert back["consumed"] == len(d)
    assert refused > 1000 and accepted > 100, (accepted, refused)


def test_parser_under_sanitizers(tmp_path):
    """The same kind of input through a C++ harness built with AddressSanitizer + UBSan, every
    datagram in a heap buffer of exactly its size."""
    import subprocess
    exe = str(tmp_path / "wire_fuzz")
    build = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined",
                            "-fno-omit-frame-pointer", "-I" + os.path.join(mw.ROOT, "include"),
                            os.path.join(mw.ROOT, "tests", "cpp", "wire_fuzz.cpp"), "-o", exe],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower():
        pytest.skip("sanitizer runtime not available")
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "Sanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-2000:]
    accepted, refused = [int(x) for x in run.stdout.split()[1::2]]
    assert accepted > 10000 and refused > 10000


@pytest.mark.skipif(not os.path.exists(mw.REF_SO), reason="verbatim reference build not present")
def test_against_reference_build(product):
    ref = mw.Wire(mw.REF_SO)
    rng = np.random.default_rng(77)
    for trial in range(60):
        kind = list(mw.TYPES)[trial % len(mw.TYPES)]
        m = mw.random_message(kind, rng, big=trial % 3 == 0)
        a, b = product.pack(m), ref.pack(m)
        assert a == b, (kind, trial)
        # each side parses the other's datagram to the same fields
        pa, pb = product.unpack(b), ref.unpack(a)
        want = mw.expected_after_wire(m)
        assert mw.same_fields(pa, want) and mw.same_fields(pb, want), (kind, trial)
