"""Parity of the CUDA path with the oracle, through the C ABI.  Bit-exact: counts and indices are
integers, and every float64 result is a fixed IEEE-754 operation sequence (left-to-right sums, one
division, one multiplication) that the kernels perform in the same order as the reference."""
import pytest
import torch

import oracle_c
import pie_oracle as po
from helpers import assert_analytics_equal, assert_daily_match_py_oracle, assert_stats_match_py_oracle, bits_equal
from sph_pie_b200 import _lib, ops
from sph_pie_b200.archive import (buildArchiveDailyGroups, computeArchiveShowStats, computeArchiveShowStatsMany,
                                  getOrCreateGroupMetricSummary)
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows

pytestmark = pytest.mark.gpu


def oracle(table, tz=0):
    st, daily, rc, who = oracle_c.archive_analytics(table, tz)
    assert rc == 0
    return st, daily


@pytest.mark.parametrize("n_shows,seed,tz,shuffle,missing", [
    (1, 0, 0, False, 0.0), (2, 1, 0, True, 0.0), (310, 2, -480, False, 0.0), (310, 3, 330, True, 0.2),
    (2049, 4, 0, True, 0.0), (5000, 5, -720, True, 0.5), (40000, 6, 840, False, 0.0), (40000, 7, -60, True, 1.0)])
def test_analytics_matches_c_oracle_both_entry_points(cuda, n_shows, seed, tz, shuffle, missing):
    host = synth_archive(n_shows, seed=seed, shuffle_days=shuffle, missing_created_frac=missing)
    ref = oracle(host, tz)
    assert_analytics_equal(ops.archive_analytics(host, tz), ref, "host entry point")
    assert_analytics_equal(ops.archive_analytics(host.to(cuda), tz), ref, "device entry point")


def test_reference_reachable_maximum_matches_python_oracle(cuda):
    """310 shows x <= 21 entries: the largest archive the reference's own rules allow,
    checked against the JSON-level Python oracle."""
    host = synth_archive(310, seed=42, shuffle_days=True, missing_created_frac=0.1)
    shows = table_to_shows(host)
    for tz in (0, -420):
        st, daily = ops.archive_analytics(host, tz)
        assert_stats_match_py_oracle(shows, st)
        assert_daily_match_py_oracle(shows, daily, tz)


def test_million_entry_batch_properties_and_oracle(cuda):
    """Larger than the oracle-in-seconds sizes of the other tests: ~1M entries, C oracle + invariants."""
    host = synth_archive(100_000, seed=8)
    dev = host.to(cuda)
    st, daily = ops.archive_analytics(dev, 0)
    assert_analytics_equal((st, daily), oracle(host, 0), "1M entries")
    i32 = st.i32.cpu()
    assert int(i32[_lib.SI_TOTAL].sum()) == host.n_entries
    assert bool((i32[_lib.SI_COMPLETED] + i32[_lib.SI_NO_LAUNCH] + i32[_lib.SI_ABORT] <= i32[_lib.SI_TOTAL]).all())
    # every show lands in exactly one group and groups ascend
    order = daily.show_order.cpu().long()
    assert torch.equal(torch.sort(order).values, torch.arange(host.n_shows))
    g = daily.group_day_start.cpu()
    assert bool((g[1:] > g[:-1]).all())
    assert int(daily.summary_count[0].sum()) == host.n_shows  # entriesCount is valid for every show


def test_permutation_of_shows_permutes_stats_and_keeps_groups(cuda):
    host = synth_archive(3000, seed=9)
    shows = table_to_shows(host)
    perm = torch.randperm(len(shows), generator=torch.Generator().manual_seed(1)).tolist()
    a = ops.show_stats(host)
    b = ops.show_stats(pack_shows([shows[p] for p in perm]))
    assert bits_equal(a.i32[:, perm], b.i32) and bits_equal(a.f64[:, perm], b.f64)


def test_ragged_and_empty_inputs(cuda):
    empty = pack_shows([])
    st, daily = ops.archive_analytics(empty, 0)
    assert st.i32.shape[1] == 0 and daily.n_groups == 0
    st, daily = ops.archive_analytics(empty.to(cuda), 0)
    assert daily.n_groups == 0
    only_empty_shows = pack_shows([{"id": str(i), "createdAt": 1.7e12 + i, "entries": []} for i in range(70)])
    assert_analytics_equal(ops.archive_analytics(only_empty_shows, 0), oracle(only_empty_shows, 0))
    big = {"id": "big", "createdAt": 1.7e12, "entries": [
        {"status": ["Completed", "Abort", "no-launch"][i % 3], "launched": "Yes" if i % 2 else "no",
         "primaryIssue": po.PRIMARY_ISSUES[i % 10] if i % 4 else "", "delaySec": (i % 97) * 0.1} for i in range(50_000)]}
    ragged = pack_shows([big, {"id": "tiny", "createdAt": 1.7e12 + 9e7, "entries": [{"delaySec": 1}]}, None])
    for table in (ragged, ragged.to(cuda)):
        assert_analytics_equal(ops.archive_analytics(table, 0), oracle(ragged, 0), "one 50k-entry show")
    assert_stats_match_py_oracle([big], ops.show_stats(pack_shows([big])))


def test_dirty_values(cuda):
    shows = [{"id": "d", "createdAt": 1.7e12, "entries": [
        {"status": "COMPLETED", "launched": "YES", "primaryIssue": " Battery ", "delaySec": -0.0},
        {"status": "Completed ", "launched": "yes ", "primaryIssue": " Other　", "delaySec": 1e308},
        {"status": "No-Launch", "launched": "", "primaryIssue": "Propulsión", "delaySec": 1e308},
        {"status": "abort", "launched": "Yes", "primaryIssue": "   ", "delaySec": float("nan")},
        {"status": "Kompleted", "launched": "no", "primaryIssue": "rf link", "delaySec": float("-inf")},
        {"status": "", "launched": "Yes", "primaryIssue": "Software or show control", "delaySec": None},
        {"status": "Abort", "launched": "Yes", "primaryIssue": " Tracking lost", "delaySec": 0.1},
        {"status": "Abort", "launched": "Yes", "primaryIssue": "Failed to launch ", "delaySec": 0.2}]}]
    got = computeArchiveShowStats(shows[0])
    want = po.compute_archive_show_stats(shows[0])
    assert got == want and list(got["issueCounts"]) == list(want["issueCounts"])
    assert got["avgDelaySec"] == float("inf") and got["completedCount"] == 1 and got["launchedCount"] == 5
    groups = buildArchiveDailyGroups(shows, 0)
    s = getOrCreateGroupMetricSummary(groups[0], "avgDelaySec")
    assert s["count"] == 0 and s["average"] is None  # Infinity is not a valid metric value (:4128-4134)


def test_mirror_api_on_json_documents(cuda):
    host = synth_archive(150, seed=12, shuffle_days=True)
    shows = table_to_shows(host)
    many = computeArchiveShowStatsMany(shows)
    for show, got in zip(shows, many):
        assert got == po.compute_archive_show_stats(show)
    groups = buildArchiveDailyGroups(shows, -300)
    want = po.build_archive_daily_groups(shows, -300)
    assert [g["dateKey"] for g in groups] == [g["dateKey"] for g in want]
    assert [g["timestamp"] for g in groups] == [g["timestamp"] for g in want]
    assert [g["midpoint"] for g in groups] == [g["midpoint"] for g in want]
    for g, w in zip(groups, want):
        assert g["totalShows"] == w["totalShows"]
        assert [i["show"]["id"] for i in g["shows"]] == [i["show"]["id"] for i in w["shows"]]
        assert [i["stats"] for i in g["shows"]] == [i["stats"] for i in w["shows"]]
        for key in po.ALL_METRIC_KEYS:
            a, b = getOrCreateGroupMetricSummary(g, key), po.group_metric_summary(w, key)
            assert (a["average"], a["min"], a["max"], a["count"], a["totalShows"]) == \
                   (b["average"], b["min"], b["max"], b["count"], b["totalShows"]), key
            assert [v["value"] for v in a["showValues"]] == b["values"], key
        assert getOrCreateGroupMetricSummary(g, "launchRate") is g["metrics"]["launchRate"]  # memoised
    assert buildArchiveDailyGroups([], 0) == [] and buildArchiveDailyGroups(None, 0) == []


def test_error_behaviour_matches_reference(cuda):
    ok = {"id": "ok", "createdAt": 1.0, "entries": []}
    with pytest.raises(_lib.JsRangeError):  # toISOString on an invalid Date throws RangeError (:3415)
        buildArchiveDailyGroups([ok, {"createdAt": 8.64e15 + 2}], 0)
    with pytest.raises(_lib.JsRangeError):
        ops.archive_analytics(pack_shows([ok, {"createdAt": -8.64e15 - 2}]).to(cuda), 0)
    with pytest.raises(_lib.UnsupportedDateError):
        buildArchiveDailyGroups([{"date": "July 4, 2024", "time": "21:00"}], 0)
    # a non-finite createdAt is simply skipped (:3409-3411), not an error
    assert buildArchiveDailyGroups([{"createdAt": float("inf")}, None], 0) == []
    with pytest.raises(_lib.PieError):
        ops.archive_analytics(pack_shows([ok]), 99999)


def test_shared_reciprocal_division_is_exact(cuda):
    """The rate columns use one IEEE reciprocal + an exact correction instead of 13 divisions per
    show.  Every (count, total) pair the kernels can feed it (total <= 4096) is compared bit-for-bit
    with IEEE division on the device."""
    import ctypes as C

    bad = C.c_uint64(123)
    _lib.check(_lib.load().pie_selftest_fast_div(4096, C.byref(bad)))
    assert bad.value == 0


def test_large_shows_use_true_division_and_counter_flush(cuda):
    """Shows with > 4096 entries (true-division path) and > 255 entries (packed-counter flush)."""
    rows = lambda n: [{"status": ["Completed", "Abort", "No-launch", "x"][i % 4], "launched": "Yes" if i % 3 else "",
                       "primaryIssue": po.PRIMARY_ISSUES[i % 10] if i % 2 else "", "delaySec": (i % 13) * 0.25}
                      for i in range(n)]
    shows = [{"id": f"n{n}", "createdAt": 1.7e12 + k * 9e7, "entries": rows(n)}
             for k, n in enumerate((254, 255, 256, 511, 4096, 4097, 9001))]
    table = pack_shows(shows)
    for t in (table, table.to(cuda)):
        assert_analytics_equal(ops.archive_analytics(t, 0), oracle(table, 0))
    assert_stats_match_py_oracle(shows[:4], ops.show_stats(pack_shows(shows[:4])))


def test_sliced_table_equals_whole(cuda):
    host = synth_archive(4000, seed=13)
    whole = ops.show_stats(host)
    a, b = ops.show_stats(host.slice_shows(0, 1500)), ops.show_stats(host.slice_shows(1500, 4000))
    assert bits_equal(torch.cat([a.i32, b.i32], 1), whole.i32) and bits_equal(torch.cat([a.f64, b.f64], 1), whole.f64)
    dev = host.to(cuda)
    c = ops.show_stats(dev.slice_shows(1500, 4000))
    assert bits_equal(c.i32, b.i32) and bits_equal(c.f64, b.f64)
