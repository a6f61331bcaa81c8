"""computeMetrics (reference public/app.js:5024-5047) on the GPU against the oracles: int32 planes and the avgDelay
text bit for bit, through the device and the host entry point, and the mirror API against the Python oracle."""
import pytest
import torch

import oracle_c
import pie_oracle as po
from sph_pie_b200 import _lib, ops
from sph_pie_b200.archive import computeMetrics, computeMetricsMany
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows
from test_export_rows_cpu import number_samples

pytestmark = pytest.mark.gpu


def assert_same(got: ops.LiveMetrics, table):
    i32, text = oracle_c.compute_metrics(table)
    assert torch.equal(got.i32.cpu(), i32), "metric planes"
    assert torch.equal(got.text.cpu(), text), "avgDelay text"


@pytest.mark.parametrize("n_shows,seed", [(1, 0), (7, 1), (310, 2), (5000, 3), (40000, 4)])
def test_metrics_match_c_oracle_both_entry_points(cuda, n_shows, seed):
    host = synth_archive(n_shows, seed=seed)
    assert_same(ops.compute_metrics(host), host)
    assert_same(ops.compute_metrics(host.to(cuda)), host)


def test_mirror_api_and_edge_shows(cuda):
    shows = [
        {"entries": [{"planned": "Yes", "status": "Completed", "delaySec": 1},
                     {"planned": "Yes", "status": "Abort", "delaySec": 2, "primaryIssue": "Battery"},
                     {"planned": "No", "status": "No-launch", "delaySec": 0.005, "primaryIssue": "RF link"},
                     {"planned": "Yes", "status": "Abort", "primaryIssue": "Battery"}]},
        {"entries": []}, None,
        {"entries": [{"planned": "yes", "status": "completed", "delaySec": float("nan")}]},
        {"entries": [{"planned": "Yes", "status": "Completed"}] * 2 + [{"planned": "Yes"}] * 5},
        {"entries": [{"planned": "Yes", "status": "Completed"}] + [{"planned": "Yes"}] * 7},
        {"entries": [{"status": "Abort", "primaryIssue": k} for k in ["b", "10", "2", "02", "b", "a", "a", "a"]]},
        {"entries": [{"status": "", "primaryIssue": k} for k in ["z", "4294967295", "4294967294", "y"]]},
        {"entries": [{"delaySec": -0.001}, {"delaySec": -0.002}]},
        {"entries": [{"delaySec": float("inf")}, {"delaySec": 1}]},
        {"entries": [{"delaySec": 2.0 ** 60}, {"delaySec": 2.0 ** 60}]},
        {"entries": [{"status": "Abort", "primaryIssue": f"issue {i % 40}"} for i in range(150)]},
        {"entries": [{"status": "Abort", "primaryIssue": "ü" * (i % 3 + 1)} for i in range(9)]},
    ]
    want = [po.compute_metrics(s) for s in shows]
    assert computeMetricsMany(shows) == want
    assert computeMetrics(shows[0]) == want[0] == {"successRate": 33, "countCompleted": 1, "countNoLaunch": 1,
                                                   "countAbort": 2, "avgDelay": "1.00", "topIssues": ["Battery", "RF link"]}
    assert computeMetrics(None) == po.compute_metrics(None if False else {})
    table = pack_shows(shows)
    assert_same(ops.compute_metrics(table.to(cuda)), table)


def test_to_fixed_on_device(cuda):
    """avgDelay of a single-entry show is x.toFixed(2): ~0.4 M doubles on the device against the exact
    decimal.Decimal oracle and the C oracle."""
    xs = number_samples(8000, 22)
    shows = [{"entries": [{"delaySec": float(x)}]} for x in xs]
    table = pack_shows(shows)
    got = ops.compute_metrics(table.to(cuda))
    assert_same(got, table)
    lens = got.i32[_lib.CM_AVG_LEN].cpu().tolist()
    text = got.text.cpu().numpy()
    for s, x in enumerate(xs.tolist()):
        assert bytes(text[s, : lens[s]]).decode() == po.js_to_fixed2(x), x


def test_sliced_table_metrics(cuda):
    host = synth_archive(3000, seed=7)
    whole = ops.compute_metrics(host.to(cuda))
    part = host.to(cuda).slice_shows(1000, 2200)
    got = ops.compute_metrics(part)
    e0 = int(host.entry_offsets[1000])
    i32 = whole.i32[:, 1000:2200].clone()
    for k in (_lib.CM_TOP0, _lib.CM_TOP1, _lib.CM_TOP2):  # entry rows are numbered inside the batch
        i32[k] = torch.where(i32[k] >= 0, i32[k] - e0, i32[k])
    assert torch.equal(got.i32, i32) and torch.equal(got.text, whole.text[1000:2200])


def test_long_shows_and_look_alike_issues(cuda):
    """Tiles longer than one shared-memory round (the thread-per-show path), shows that straddle CTAs, and issue
    strings that agree in length and first byte (the 16-bit print passes them to the full comparison)."""
    alike = ["Tracking lost", "Tracking lose", "Tracking losé"[:13], "T" + "x" * 12, "Tracking lost", "T" * 300, "T" * 299 + "U",
             "T" * 300, "", "Tracking lost"]
    shows = [{"entries": [{"status": "Abort", "planned": "Yes", "delaySec": float(i % 7), "primaryIssue": alike[i % len(alike)]}
                          for i in range(n)]} for n in (5000, 3, 0, 2049, 2048, 700, 1, 130)]
    shows += [{"entries": [{"status": "Abort" if i % 3 else "Completed", "primaryIssue": alike[(i * 7 + j) % len(alike)]}
                           for i in range(j % 23)]} for j in range(400)]
    table = pack_shows(shows)
    assert_same(ops.compute_metrics(table.to(cuda)), table)
    assert_same(ops.compute_metrics(table), table)
