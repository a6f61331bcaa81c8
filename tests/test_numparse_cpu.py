"""The product's JSON-number parser (host build of pie_numparse.cuh, the code the ingest kernels compile) against
Python's float(), which is correctly rounded: random bit patterns printed with 17 and with shortest digits, typed
decimals, long digit strings, halfway cases, exponents at both ends of the range, subnormals."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def numparse():
    src = os.path.join(HERE, "native", "numparse_host.cpp")
    so = os.path.join(HERE, "native", "libnumparse_host.so")
    csrc = os.path.join(HERE, "..", "sph_pie_b200", "csrc")
    deps = [src, os.path.join(csrc, "pie_numparse.cuh"), os.path.join(csrc, "pow5_128_table.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)

    def parse(texts):
        enc = [t.encode() for t in texts]
        offs = np.zeros(len(enc) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in enc], out=offs[1:])
        blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()
        n = len(enc)
        values, status, used = np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int64)
        lib.numparse_host_batch(C.c_void_p(blob.ctypes.data), C.c_void_p(offs.ctypes.data), C.c_int64(n),
                                C.c_void_p(values.ctypes.data), C.c_void_p(status.ctypes.data), C.c_void_p(used.ctypes.data))
        return values, status, used

    return parse


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def check(parse, texts, allow_undecided=0.0):
    values, status, used = parse(texts)
    undecided = 0
    for t, v, st, u in zip(texts, values.tolist(), status.tolist(), used.tolist()):
        if st == 2:
            undecided += 1
            continue
        assert st == 0 and u == len(t), (t, st, u)
        assert bits(v) == bits(float(t)), (t, v, float(t))
    assert undecided <= allow_undecided * len(texts), (undecided, len(texts))
    return undecided


def test_known_answers(numparse):
    texts = ["0", "-0", "0.0", "-0.0e5", "1", "-1", "12.5", "0.1", "0.3", "1e22", "1e23", "123456789012345678",
             "9007199254740993", "9007199254740992", "9007199254740991", "1.7976931348623157e308", "1.7976931348623159e308",
             "2e308", "4.9e-324", "2.4703282292062327e-324", "2.4703282292062328e-324", "5e-324", "1e-400", "1e400",
             "2.2250738585072011e-308", "2.2250738585072014e-308", "0.000001", "1E5", "1e+5", "1e-5", "123e0",
             "8.41e21", "1e21", "0.5", "3.7", "100", "1704067200000", "0.30000000000000004"]
    assert check(numparse, texts) == 0
    v, st, used = numparse(["-", "01", "1.", ".5", "1e", "1e+", "+1", "", "abc", "-a", "1.e5"])
    assert st.tolist() == [1] * 11
    v, st, used = numparse(["12,3", "5]", "7 ", "1e5}", "0x10"])  # stops at the first byte that is not part of the number
    assert st.tolist() == [0] * 5 and used.tolist() == [2, 1, 1, 3, 1]


def test_random_doubles_round_trip(numparse):
    rng = np.random.default_rng(5)
    xs = rng.integers(0, 2 ** 64, 120000, dtype=np.uint64).view(np.float64)
    xs = xs[np.isfinite(xs)]
    assert check(numparse, [repr(float(x)) for x in xs]) == 0          # shortest digits (JSON.stringify's form)
    assert check(numparse, ["%.17e" % x for x in xs[:60000]]) == 0     # 17 significant digits
    assert check(numparse, ["%.25e" % x for x in xs[:40000]], allow_undecided=0.001) <= 40   # > 19 digits


def test_typed_decimals_and_long_strings(numparse):
    rng = np.random.default_rng(6)
    texts = []
    for _ in range(60000):
        n = int(rng.integers(1, 10 ** int(rng.integers(1, 19))))
        d = int(rng.integers(0, 12))
        s = str(n)
        texts.append(s if d == 0 else (s[:-d] or "0") + "." + s[-d:].rjust(d, "0") if len(s) > d else "0." + s.rjust(d, "0"))
    for _ in range(20000):  # huge / tiny exponents, many digits, trailing zeros
        mant = "".join(rng.choice(list("0123456789"), int(rng.integers(1, 40))))
        mant = mant.lstrip("0") or "0"
        e = int(rng.integers(-400, 400))
        texts.append(f"{mant[0]}.{mant[1:] or '0'}e{e}")
    assert check(numparse, texts, allow_undecided=0.002) <= 160


def test_halfway_and_boundaries(numparse):
    texts = []
    for k in range(1, 2000):  # exact ties between neighbouring doubles around 2^53 and small integers + 0.5 ulp
        texts.append(str(2 ** 53 + 2 * k + 1))
        texts.append(str((2 ** 53 + 2 * k) * 2 ** 10 + 2 ** 10))
    for e in range(-330, 310, 7):
        for m in ("1", "9.999999999999999", "1.0000000000000002", "4.4501477170144023"):
            texts.append(f"{m}e{e}")
    assert check(numparse, texts, allow_undecided=0.01) <= 80


def test_numbers_of_the_shape_the_warp_path_defers_are_always_decided(numparse):
    """The warp-per-document ingest (csrc/pie_json_fast.cuh, number_shape) only CHECKS in its first pass that a number has
    no exponent part and at most 19 digits in all, and leaves the conversion to the second pass, which can no longer
    decline a document: the parser must never answer "undecided" for such a number.  (No digit is dropped at <= 19
    digits, and the Eisel-Lemire product is only undecided for powers of ten outside 10^-27 .. 10^55.)"""
    rng = np.random.default_rng(29)
    texts = []
    for _ in range(60000):
        digits = int(rng.integers(1, 20))
        body = "".join(str(int(d)) for d in rng.integers(0, 10, digits))
        cut = int(rng.integers(0, digits + 1))  # digits before the point
        ip, fp = body[:cut], body[cut:]
        ip = ip.lstrip("0") or "0"
        if ip == "0" and cut > 1:
            continue  # the stripped zeros would no longer count as digits of the text: keep the total honest
        t = ip + ("." + fp if fp else "")
        if rng.random() < 0.3:
            t = "-" + t
        texts.append(t)
    # the worst cases for a truncated product: digit strings around powers of two and the halfway points between doubles
    for e in range(0, 64):
        for d in (-1, 0, 1):
            v = 2 ** e + d
            if 0 < v < 10 ** 19:
                texts.append(str(v))
                s = str(v)
                for cut in (1, len(s) // 2, len(s) - 1):
                    if 0 < cut < len(s):
                        texts.append(s[:cut] + "." + s[cut:])
    texts += ["9999999999999999999", "0.9999999999999999999"[:21], "1844674407370955161.5", "0.000000000000000001",
              "9007199254740993", "9007199254740992.5", "4503599627370496.5", "4503599627370497.5"]
    texts = [t for t in texts if sum(c.isdigit() for c in t) <= 19]
    assert len(texts) > 50000
    assert check(numparse, texts, allow_undecided=0.0) == 0
