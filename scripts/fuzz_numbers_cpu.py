"""Long campaign for the two number routines of the path, ON THE CPU (host builds of the very sources the kernels
compile): Number::toString (csrc/pie_numfmt.cuh: typed-decimal shortcut + Ryu; the delaySec cell of a CSV row,
reference server/webhookDispatcher.js:291 / :342 `String(value)`) against Python's repr-derived restatement
(oracle/pie_oracle.js_number_to_string), and the JSON-number parser of the ingest (csrc/pie_numparse.cuh, Eisel-Lemire;
JSON.parse at server/storage/sqlProvider.js:899) against Python's float(), which is correctly rounded.

    python scripts/fuzz_numbers_cpu.py --seed 1 --minutes 10

tests/test_export_rows_cpu.py and tests/test_numparse_cpu.py hold the bounded versions.  Exits 1 at the first difference."""
import argparse
import ctypes as C
import json
import os
import struct
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import pie_oracle as po  # noqa: E402

NATIVE = os.path.join(ROOT, "tests", "native")
CSRC = os.path.join(ROOT, "sph_pie_b200", "csrc")


def host_lib(name, deps):
    src, so = os.path.join(NATIVE, name + ".cpp"), os.path.join(NATIVE, "lib" + name + ".so")
    deps = [src] + [os.path.join(CSRC, d) for d in deps]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def fmt_batch(lib, xs):
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n = len(xs)
    out, lens = np.zeros(n * 32, dtype=np.uint8), np.zeros(n, dtype=np.int32)
    lib.numfmt_host_batch(C.c_void_p(xs.ctypes.data), C.c_int64(n), C.c_void_p(out.ctypes.data), C.c_void_p(lens.ctypes.data))
    b = out.tobytes()
    return [b[i * 32:i * 32 + lens[i]].decode() for i in range(n)]


def parse_batch(lib, texts):
    enc = [t.encode() for t in texts]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in enc], out=offs[1:])
    blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()
    n = len(enc)
    values, status, used = np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int64)
    lib.numparse_host_batch(C.c_void_p(blob.ctypes.data), C.c_void_p(offs.ctypes.data), C.c_int64(n), C.c_void_p(values.ctypes.data),
                            C.c_void_p(status.ctypes.data), C.c_void_p(used.ctypes.data))
    return values, status, used


def samples(rng, n):
    """Doubles of every kind the cell can hold: raw bit patterns, typed decimals and their neighbours one ulp away,
    powers of two and ten, short binary fractions, integers at every size."""
    d = rng.integers(0, 9, n)
    typed = rng.integers(0, 10 ** rng.integers(1, 17, n)).astype(np.float64) / 10.0 ** d
    return np.concatenate([
        rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64), typed, np.nextafter(typed, np.inf), np.nextafter(typed, -np.inf),
        -typed, 2.0 ** rng.integers(-1074, 1024, n), 10.0 ** rng.integers(-320, 309, n),
        np.nextafter(10.0 ** rng.integers(-320, 309, n), rng.choice([-np.inf, np.inf], n)),
        rng.integers(1, 2 ** 52, n, dtype=np.uint64).view(np.float64),  # subnormals
        rng.integers(1, 2 ** 53, n).astype(np.float64) / 2.0 ** rng.integers(1, 64, n),
        rng.random(n) * rng.choice([1e-9, 1e-3, 1, 1e3, 1e15, 1e21, 1e25], n),
        np.round(rng.random(n) * 10.0 ** rng.integers(0, 7, n), rng.integers(0, 4)),
        (1e21 + rng.integers(-10 ** 6, 10 ** 6, n).astype(np.float64) * 131072.0), (1e-6 + (rng.random(n) - 0.5) * 1e-7)])


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=5.0)
    args = ap.parse_args()
    fmt = host_lib("numfmt_host", ["pie_numfmt.cuh", "ryu_tables.h"])
    par = host_lib("numparse_host", ["pie_numparse.cuh", "pow5_128_table.h"])
    rng = np.random.default_rng(args.seed)
    stats = dict(seed=args.seed, formatted=0, parsed=0, undecided_over_19_digits=0)
    t0 = time.time()
    while time.time() - t0 < args.minutes * 60:
        xs = samples(rng, 20000)
        got = fmt_batch(fmt, xs)
        for x, a in zip(xs.tolist(), got):
            want = po.js_number_to_string(x)
            if a != want:
                print("DIFFERENCE Number::toString", repr(x), a, want, flush=True)
                sys.exit(1)
        stats["formatted"] += len(xs)
        # what JSON.stringify wrote comes back as the same double; so do 17 digits, and more digits than a double holds
        fin = xs[np.isfinite(xs)]
        texts = [t for t in got if t not in ("NaN", "Infinity", "-Infinity")]
        texts += ["%.17e" % x for x in fin[:60000]] + ["%.*e" % (int(k), x) for k, x in zip(rng.integers(18, 40, 30000), fin[-30000:])]
        texts += ["%d.%0*d" % (a, int(w), b) for a, b, w in zip(rng.integers(0, 10 ** 9, 20000), rng.integers(0, 10 ** 9, 20000),
                                                                 rng.integers(9, 11, 20000))]
        values, status, used = parse_batch(par, texts)
        for t, v, st, u in zip(texts, values.tolist(), status.tolist(), used.tolist()):
            digits = sum(ch.isdigit() for ch in t.split("e")[0])
            if st == 2 and digits > 19:
                stats["undecided_over_19_digits"] += 1  # the kernels hand such a number to the exact slow path
                continue
            if st != 0 or u != len(t) or bits(v) != bits(float(t)):
                print("DIFFERENCE number parser", t, st, u, repr(v), repr(float(t)), flush=True)
                sys.exit(1)
        stats["parsed"] += len(texts)
        print(json.dumps(stats), flush=True)
    stats["minutes"] = round((time.time() - t0) / 60, 1)
    print("SUMMARY", json.dumps(stats), flush=True)


if __name__ == "__main__":
    main()
