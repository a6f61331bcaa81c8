"""The two independent CPU restatements against each other: oracle/pie_oracle.py (JSON documents,
JS value semantics) and oracle/pie_oracle.c (columnar layout).  CPU only."""
import pytest
import torch

import oracle_c
from helpers import assert_daily_match_py_oracle, assert_stats_match_py_oracle
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows


@pytest.mark.parametrize("seed,tz,shuffle,missing", [(0, 0, False, 0.0), (1, -480, True, 0.1), (2, 330, True, 0.3),
                                                     (3, 840, False, 1.0), (4, -720, True, 0.0)])
def test_c_oracle_matches_python_oracle(built, seed, tz, shuffle, missing):
    table = synth_archive(240, seed=seed, shuffle_days=shuffle, missing_created_frac=missing)
    shows = table_to_shows(table)
    st, daily, rc, _ = oracle_c.archive_analytics(table, tz)
    assert rc == 0
    assert_stats_match_py_oracle(shows, st)
    assert_daily_match_py_oracle(shows, daily, tz)


def test_pack_round_trip_feeds_both_oracles(built):
    table = synth_archive(60, seed=9)
    shows = table_to_shows(table)
    repacked = pack_shows(shows)
    a, _, _, _ = oracle_c.archive_analytics(table, 0)
    b, _, _, _ = oracle_c.archive_analytics(repacked, 0)
    assert torch.equal(a.i32, b.i32) and torch.equal(a.f64.view(torch.int64), b.f64.view(torch.int64))


def test_edge_shows(built):
    shows = [
        None,                                                     # falsy show: skipped (:3405)
        {"id": "empty", "createdAt": 1720000000000, "entries": []},
        {"id": "no-entries-key", "createdAt": 1720000000001},
        {"id": "ts-from-date", "date": "2024-07-03", "time": "23:59", "entries": [{"status": "Abort"}]},
        {"id": "ts-from-archived", "archivedAt": 1720000500000, "entries": [{"launched": "yes"}]},
        {"id": "ts-from-entries", "entries": [{"ts": 1720090000000.7, "delaySec": -0.0}, {"ts": 1719990000000}]},
        {"id": "no-ts", "entries": [{"status": "Completed", "delaySec": 1e308}, {"delaySec": 1e308}]},
        {"id": "neg-epoch", "createdAt": -1.5, "entries": [{"primaryIssue": " Other "}]},
        {"id": "bad-month", "date": "2024-13-01", "time": "10:00", "archivedAt": 1720000900000, "entries": []},
    ]
    table = pack_shows(shows)
    for tz in (0, -300, 60):
        st, daily, rc, _ = oracle_c.archive_analytics(table, tz)
        assert rc == 0
        assert_stats_match_py_oracle(shows, st)
        assert_daily_match_py_oracle(shows, daily, tz)


def test_threaded_oracle_is_the_same_function(built):
    table = synth_archive(3000, seed=5)
    a = oracle_c.show_stats(table, nthreads=1)
    b = oracle_c.show_stats(table, nthreads=4)
    assert torch.equal(a.i32, b.i32) and torch.equal(a.f64.view(torch.int64), b.f64.view(torch.int64))


def test_range_error_and_unsupported_date(built):
    import pie_oracle as po
    from sph_pie_b200 import _lib

    ok = {"id": "ok", "createdAt": 1.0, "entries": []}
    _, _, rc, who = oracle_c.archive_analytics(pack_shows([ok, {"createdAt": 9e15}, {"createdAt": -9e15}]), 0)
    assert (rc, who) == (_lib.PIE_ERR_RANGE, 1)
    with pytest.raises(po.JsRangeError):
        po.build_archive_daily_groups([ok, {"createdAt": 9e15}], 0)
    _, _, rc, who = oracle_c.archive_analytics(pack_shows([ok, ok, {"date": "July 4, 2024", "time": "21:00"}]), 0)
    assert (rc, who) == (_lib.PIE_ERR_UNSUPPORTED_DATE, 2)
    with pytest.raises(NotImplementedError):
        po.build_archive_daily_groups([{"date": "July 4, 2024", "time": "21:00"}], 0)
