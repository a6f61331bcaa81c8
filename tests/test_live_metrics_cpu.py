"""CPU checks for computeMetrics (reference public/app.js:5024-5047): the JSON-level Python restatement against
the columnar C restatement, and hand-derived known answers for the ECMAScript rules it leans on (toFixed on the
exact binary value, Math.round ties, Object.entries key order, stable sort)."""
import math

import numpy as np
import pytest

import oracle_c
import pie_oracle as po
from sph_pie_b200 import _lib
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows
from test_export_rows_cpu import number_samples


def metrics_from_planes(table, i32, text):
    issue = table.entry_cols["primary_issue"]
    offs, data = issue.offsets.tolist(), bytes(issue.data.numpy())
    i32 = i32.tolist()
    out = []
    for s in range(table.n_shows):
        rows = [i32[_lib.CM_TOP0 + k][s] for k in range(3)]
        out.append({"successRate": i32[0][s], "countCompleted": i32[1][s], "countNoLaunch": i32[2][s],
                    "countAbort": i32[3][s], "avgDelay": bytes(text[s, : i32[7][s]].numpy()).decode(),
                    "topIssues": [data[offs[e]:offs[e + 1]].decode() for e in rows if e >= 0]})
    return out


def test_to_fixed_known_answers():
    # ECMA-262 21.1.3.3: exact binary value, ties pick the larger n (in magnitude), "-" iff x < 0
    assert po.js_to_fixed2(1.005) == "1.00"      # 1.00499999999999989...
    assert po.js_to_fixed2(1.255) == "1.25" and po.js_to_fixed2(2.675) == "2.67"
    assert po.js_to_fixed2(0.125) == "0.13" and po.js_to_fixed2(0.375) == "0.38"  # exact ties round up
    assert po.js_to_fixed2(-0.125) == "-0.13" and po.js_to_fixed2(-0.001) == "-0.00" and po.js_to_fixed2(-0.0) == "0.00"
    assert po.js_to_fixed2(2.0 ** 60) == "1152921504606846976.00"  # exact integer digits, unlike Number::toString
    assert po.js_to_fixed2(1e21) == "1e+21" and po.js_to_fixed2(999999999999999868928.0) == "999999999999999868928.00"
    assert po.js_to_fixed2(float("inf")) == "Infinity" and po.js_to_fixed2(float("nan")) == "NaN"
    assert po.js_to_fixed2(5e-324) == "0.00" and po.js_to_fixed2(0.995) == "0.99" and po.js_to_fixed2(0.005) == "0.01"


def test_math_round_and_key_order_known_answers():
    assert po.js_math_round(2.5) == 3 and po.js_math_round(-2.5) == -2 and po.js_math_round(0.49999999999999994) == 0
    sh = {"entries": [{"status": "Abort", "primaryIssue": "b"}, {"status": "Abort", "primaryIssue": "10"},
                      {"status": "Abort", "primaryIssue": "2"}, {"status": "Abort", "primaryIssue": "02"},
                      {"status": "x", "primaryIssue": "b"}, {"status": "Completed", "primaryIssue": "02"}]}
    # counts b:2, 10:1, 2:1, 02:1; Object.entries lists the array-index keys 2, 10 first, then b, 02
    assert po.compute_metrics(sh)["topIssues"] == ["b", "2", "10"]
    sh = {"entries": [{"status": "", "primaryIssue": k} for k in ["z", "4294967295", "4294967294", "y"]]}
    assert po.compute_metrics(sh)["topIssues"] == ["4294967294", "z", "4294967295"]  # 2^32-1 is not an array index


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_c_metrics_match_python_oracle(built, seed):
    table = synth_archive(400, seed=seed)
    i32, text = oracle_c.compute_metrics(table)
    got = metrics_from_planes(table, i32, text)
    want = [po.compute_metrics(s) for s in table_to_shows(table)]
    assert got == want


def test_c_metrics_edge_shows(built):
    shows = [
        {"entries": [{"planned": "Yes", "status": "Completed", "delaySec": 1},
                     {"planned": "Yes", "status": "Abort", "delaySec": 2, "primaryIssue": "Battery"},
                     {"planned": "No", "status": "No-launch", "delaySec": 0.005, "primaryIssue": "RF link"},
                     {"planned": "Yes", "status": "Abort", "primaryIssue": "Battery"}]},
        {"entries": []}, None,
        {"entries": [{"planned": "yes", "status": "completed", "delaySec": float("nan")}]},
        {"entries": [{"planned": "Yes", "status": "Completed"}] * 2 + [{"planned": "Yes"}] * 5},  # 2/7*100 = 28.57
        {"entries": [{"planned": "Yes", "status": "Completed"}] + [{"planned": "Yes"}] * 7},       # 12.5 -> 13
        {"entries": [{"status": "Abort", "primaryIssue": k} for k in ["b", "10", "2", "02", "b", "a", "a", "a"]]},
        {"entries": [{"delaySec": -0.001}, {"delaySec": -0.002}]},
        {"entries": [{"delaySec": float("inf")}, {"delaySec": 1}]},
        {"entries": [{"status": "Abort", "primaryIssue": f"issue {i % 40}"} for i in range(150)]},  # > 64 entries
    ]
    table = pack_shows(shows)
    i32, text = oracle_c.compute_metrics(table)
    got = metrics_from_planes(table, i32, text)
    assert got == [po.compute_metrics(s) for s in shows]
    assert got[0] == {"successRate": 33, "countCompleted": 1, "countNoLaunch": 1, "countAbort": 2, "avgDelay": "1.00",
                      "topIssues": ["Battery", "RF link"]}
    assert got[1]["avgDelay"] == "0.00" and got[3]["avgDelay"] == "NaN" and got[3]["countCompleted"] == 0
    assert got[4]["successRate"] == 29 and got[5]["successRate"] == 13
    assert got[6]["topIssues"] == ["a", "b", "2"] and got[7]["avgDelay"] == "-0.00" and got[8]["avgDelay"] == "Infinity"


def test_c_to_fixed_matches_decimal_oracle(built):
    """avgDelay of a single-entry show is x.toFixed(2): ~0.4 M doubles through the C restatement (integer
    arithmetic on the mantissa) against the Python one (decimal.Decimal, exact)."""
    xs = number_samples(8000, 21)
    shows = [{"entries": [{"delaySec": float(x)}]} for x in xs]
    table = pack_shows(shows)
    i32, text = oracle_c.compute_metrics(table)
    lens = i32[_lib.CM_AVG_LEN].tolist()
    text = text.numpy()
    for s, x in enumerate(xs.tolist()):
        assert bytes(text[s, : lens[s]]).decode() == po.js_to_fixed2(x), x
