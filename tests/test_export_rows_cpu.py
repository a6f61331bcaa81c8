"""CPU checks for the export-row path: the product's Number::toString code (host build of
pie_numfmt.cuh, the same source the kernels compile) against two independent oracles, and the C
restatement of buildCsvRow against the JSON-level Python restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_c
import pie_oracle as po
from sph_pie_b200.synth import synth_archive, table_to_shows

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def numfmt_host():
    src = os.path.join(HERE, "native", "numfmt_host.cpp")
    so = os.path.join(HERE, "native", "libnumfmt_host.so")
    csrc = os.path.join(HERE, "..", "sph_pie_b200", "csrc")
    deps = [src, os.path.join(csrc, "pie_numfmt.cuh"), os.path.join(csrc, "ryu_tables.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)

    def fmt(xs):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n = len(xs)
        out = np.zeros(n * 32, dtype=np.uint8)
        lens = np.zeros(n, dtype=np.int32)
        lib.numfmt_host_batch(C.c_void_p(xs.ctypes.data), C.c_int64(n), C.c_void_p(out.ctypes.data), C.c_void_p(lens.ctypes.data))
        b = out.tobytes()
        return [b[i * 32:i * 32 + lens[i]].decode() for i in range(n)]

    return fmt


SPECIAL = [0.0, -0.0, 1.0, -1.5, 100.0, 1e21, 1e20, 1e-6, 1e-7, 0.1, 0.3, 0.1 + 0.2, 5e-324, 1.7976931348623157e308,
           2.0 ** 53, 123.456, 1.5e-10, 1.2345678901234567e19, float("nan"), float("inf"), float("-inf"),
           2.2250738585072014e-308, 2.225073858507201e-308, 9007199254740993.0, 4.35, 123456789012345680000.0, 1e22,
           1e23, 8.41e21, 9.5367431640625e-07, 4.940656e-318, 2.98023223876953125e-8, 5.764607523034235e39,
           1.152921504606847e40, 0.5, 12.5, 0.25, 99.99, 1e15, 1e16, 123456789.123456789]


def number_samples(n, seed):
    rng = np.random.default_rng(seed)
    return np.concatenate([
        np.array(SPECIAL), rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64),
        rng.integers(-10 ** 6, 10 ** 6, n).astype(np.float64), np.round(rng.random(n) * 1000, 2),
        rng.random(n) * rng.choice([1e-9, 1e-3, 1, 1e3, 1e15, 1e25], n), 2.0 ** rng.integers(-1074, 1024, n),
        10.0 ** rng.integers(-320, 309, n), rng.integers(1, 2 ** 52, n, dtype=np.uint64).view(np.float64),
        # short exact binary fractions k / 2^f at every size (the exact-decimal shortcut and its 15-digit limit)
        rng.integers(1, 2 ** 20, n).astype(np.float64) / 2.0 ** rng.integers(1, 14, n),
        rng.integers(1, 2 ** 53, n).astype(np.float64) / 2.0 ** rng.integers(1, 64, n),
        (10.0 ** rng.integers(9, 16, n) + rng.integers(0, 1000, n)) / 2.0 ** rng.integers(0, 12, n),
        # typed decimals: integers of every size over 10^d, d = 0..8 (the divide-and-compare shortcut, its d <= 6
        # and 15-digit limits), and their neighbours one ulp away (which must NOT take the same string)
        rng.integers(0, 10 ** rng.integers(1, 17, n)).astype(np.float64) / 10.0 ** rng.integers(0, 9, n),
        np.nextafter(rng.integers(1, 10 ** 6, n).astype(np.float64) / 10.0 ** rng.integers(0, 7, n), np.inf),
        np.nextafter(rng.integers(1, 10 ** 6, n).astype(np.float64) / 10.0 ** rng.integers(0, 7, n), -np.inf),
        -rng.integers(1, 10 ** 9, n).astype(np.float64) / 10.0 ** rng.integers(0, 7, n)])


def test_number_to_string_three_ways(built, numfmt_host):
    xs = number_samples(20000, 3)
    ryu = numfmt_host(xs)                            # product code (Ryu), host build
    printf = oracle_c.number_to_string_batch(xs)     # C oracle: printf/strtod search
    for x, a, b in zip(xs.tolist(), ryu, printf):
        want = po.js_number_to_string(x)             # Python oracle: repr()
        assert a == want and b == want, (x, a, b, want)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c_csv_rows_match_python_oracle(built, seed):
    table = synth_archive(120, seed=seed)
    shows = table_to_shows(table)
    offsets, data = oracle_c.csv_rows(table)
    blob = bytes(data.numpy())
    o = offsets.tolist()
    e = 0
    for show in shows:
        body = []
        for entry in show["entries"]:
            want = po.build_csv_row(po.build_table_row(show, entry))
            got = blob[o[e]:o[e + 1] - 1].decode("utf-8")
            assert blob[o[e + 1] - 1:o[e + 1]] == b"\n"
            assert got == want, (e, got, want)
            body.append(got)
            e += 1
        assert po.export_show_as_csv(show) == "\n".join([",".join(po.EXPORT_COLUMNS)] + body)
    assert e == table.n_entries and o[-1] == len(blob)


def test_csv_edge_rows(built):
    from sph_pie_b200.columnar import pack_shows

    shows = [{"id": 'a"b', "date": "2024-07-04", "time": "21:00", "label": "x,y", "crew": ["A|B", 'q"', ""],
              "leadPilot": "l\np", "monkeyLead": "", "notes": "r\rn",
              "entries": [
                  {"id": "e1", "status": "Completed", "primaryIssue": "Battery", "subIssue": "s", "otherDetail": "o",
                   "severity": "v", "rootCause": "r", "actions": ["x,y", "z"], "delaySec": 0, "notes": '""'},
                  {"id": "e2", "status": "completed", "primaryIssue": "Battery", "delaySec": 1e21, "actions": []},
                  {"id": "e3", "status": "Abort", "delaySec": None, "notes": ","},
                  {"id": "e4", "delaySec": -0.0}, {"id": "e5", "delaySec": float("nan")}, {"delaySec": 1.5e-7}]},
             {"id": "empty", "entries": []}, None]
    table = pack_shows(shows)
    offsets, data = oracle_c.csv_rows(table)
    blob, o = bytes(data.numpy()), offsets.tolist()
    rows = [blob[o[i]:o[i + 1] - 1].decode() for i in range(table.n_entries)]
    want = [po.build_csv_row(po.build_table_row(shows[0], e)) for e in shows[0]["entries"]]
    assert rows == want
    assert rows[0].startswith('"a""b",2024-07-04,21:00,"x,y","A|B|q""|","l\np",,"r\rn",e1,')
    assert rows[0].endswith(',Completed,,,,,,"x,y|z",,,0,,""""""')
    assert ",1e+21," in rows[1] and rows[1].count("Battery") == 1
    assert rows[3].split(",")[-3] == "0" and rows[4].split(",")[-3] == "NaN" and rows[5].split(",")[-3] == "1.5e-7"
