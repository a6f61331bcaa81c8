"""Export rows (buildTableRow -> csvEscape -> buildCsvRow) on the GPU against the oracles.
Bit-exact: byte strings and int64 offsets."""
import numpy as np
import pytest
import torch

import oracle_c
import pie_oracle as po
from sph_pie_b200 import _lib, ops
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows
from sph_pie_b200.webhook import EXPORT_COLUMNS, buildCsvRows, buildCsvRowsMany, buildMessagePayload, exportShowAsCsv
from test_export_rows_cpu import number_samples

pytestmark = pytest.mark.gpu


def assert_same_rows(got: ops.CsvRows, table):
    offsets, data = oracle_c.csv_rows(table)
    assert torch.equal(got.row_offsets.cpu(), offsets), "row_offsets"
    assert torch.equal(got.data.cpu(), data), "csv bytes"


@pytest.mark.parametrize("n_shows,seed", [(1, 0), (7, 1), (310, 2), (5000, 3), (40000, 4)])
def test_csv_rows_match_c_oracle_both_entry_points(cuda, n_shows, seed):
    host = synth_archive(n_shows, seed=seed)
    assert_same_rows(ops.csv_rows(host), host)
    assert_same_rows(ops.csv_rows(host.to(cuda)), host)


def test_number_to_string_on_device(cuda):
    """delaySec cells over ~160k doubles (random bit patterns, integers, decimals, powers, subnormals,
    specials): Number::toString on the device vs the Python oracle and the printf oracle."""
    xs = number_samples(20000, 11)
    n = len(xs)
    table = synth_archive(1, seed=0, max_entries=0)
    shows = [{"id": "n", "entries": [{"delaySec": float(x)} for x in xs]}]
    table = pack_shows(shows)
    rows = ops.csv_rows(table.to(cuda)).rows()
    assert len(rows) == n
    printf = oracle_c.number_to_string_batch(xs)
    for x, row, want_c in zip(xs.tolist(), rows, printf):
        cell = row.split(",")[21]
        assert cell == po.js_number_to_string(x) == want_c, (x, cell, want_c)


def test_edge_rows_and_mirror_api(cuda):
    shows = [{"id": 'a"b', "date": "2024-07-04", "time": "21:00", "label": "x,y", "crew": ["A|B", 'q"', ""],
              "leadPilot": "l\np", "monkeyLead": "", "notes": "r\rn",
              "entries": [
                  {"id": "e1", "status": "Completed", "primaryIssue": "Battery", "subIssue": "s", "otherDetail": "o",
                   "severity": "v", "rootCause": "r", "actions": ["x,y", "z"], "delaySec": 0, "notes": '""'},
                  {"id": "e2", "status": "completed", "primaryIssue": "Battery", "delaySec": 1e21, "actions": []},
                  {"id": "e3", "status": "Abort", "delaySec": None, "notes": ","},
                  {"id": "e4", "delaySec": -0.0}, {"id": "e5", "delaySec": float("nan")}, {"delaySec": 1.5e-7},
                  {"id": "ü-ñ-漢字", "notes": "emoji 🚁, ok"}]},
             {"id": "empty", "entries": []}, None,
             {"id": "long", "entries": [{"id": "L", "notes": "x" * 70000 + '"' + "y" * 5000}, {"id": "after"}]}]
    many = buildCsvRowsMany(shows)
    for show, rows in zip(shows, many):
        entries = (show or {}).get("entries", [])
        assert rows == [po.build_csv_row(po.build_table_row(show, e)) for e in entries]
    assert exportShowAsCsv(shows[0]) == po.export_show_as_csv(shows[0])
    assert exportShowAsCsv(shows[1]) == ",".join(EXPORT_COLUMNS)
    assert buildCsvRows(None) == []
    row = po.build_table_row(shows[0], shows[0]["entries"][0])
    assert buildMessagePayload(row) == po.build_message_payload(row)
    dev = pack_shows(shows).to(cuda)
    assert_same_rows(ops.csv_rows(dev), pack_shows(shows))


def test_reference_fixture_row(cuda):
    """The reference's only fixture (scripts/simulate-webhook.js:42-65)."""
    import json
    import os

    fix = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "webhook_fixture.json")))
    show = dict(fix["show"], entries=[fix["entry"]])
    assert buildCsvRows(show) == [fix["expected_csv_row"]]
    assert exportShowAsCsv(show) == ",".join(fix["export_columns"]) + "\n" + fix["expected_csv_row"]


def test_capacity_and_size_query(cuda):
    host = synth_archive(300, seed=5)
    dev = host.to(cuda)
    offsets, data = oracle_c.csv_rows(host)
    sizing = ops.CsvBuffers(dev.n_entries, 0, cuda)
    ops.csv_rows_dev(dev, sizing, size_only=True)
    assert int(sizing.total.cpu()) == data.numel() and torch.equal(sizing.row_offsets.cpu(), offsets)
    small = ops.CsvBuffers(dev.n_entries, data.numel() // 2, cuda)
    small.data.fill_(0xEE)
    ops.csv_rows_dev(dev, small)
    torch.cuda.synchronize()
    assert int(small.total.cpu()) == data.numel()            # still reports the size needed
    # nothing written past the capacity the caller gave (buffer is exactly capacity bytes long, so an
    # overrun would corrupt the allocator's neighbour; check the prefix that was written is right)
    written = small.data.cpu()
    k = int(torch.nonzero(written != 0xEE).max()) + 1 if bool((written != 0xEE).any()) else 0
    assert k <= data.numel() // 2 and torch.equal(written[:k], data[:k])
    import ctypes as C

    view = host.view()
    total = C.c_uint64(0)
    off = torch.empty(host.n_entries + 1, dtype=torch.int64)
    buf = torch.empty(16, dtype=torch.uint8)
    rc = _lib.load().pie_csv_rows_host(C.byref(view), off.data_ptr(), buf.data_ptr(), 16, C.byref(total))
    assert rc == _lib.PIE_ERR_CAPACITY and total.value == data.numel()


def test_round_trip_through_python_csv_module(cuda):
    """Size-independent property at a larger size: the bytes parse back (RFC 4180 reader) into the
    24 cells buildTableRow defines, for every row."""
    import csv
    import io

    host = synth_archive(20000, seed=6)
    shows = table_to_shows(host)
    got = ops.csv_rows(host.to(cuda))
    text = bytes(got.data.cpu().numpy()).decode("utf-8")
    parsed = list(csv.reader(io.StringIO(text, newline=""), lineterminator="\n"))
    assert len(parsed) == host.n_entries
    i = 0
    for show in shows:
        for entry in show["entries"]:
            row = po.build_table_row(show, entry)
            want = ["" if row[c] == "" else po.js_string(row[c]) for c in EXPORT_COLUMNS]
            assert parsed[i] == want, i
            i += 1


def test_sliced_table_rows(cuda):
    host = synth_archive(3000, seed=7)
    whole = ops.csv_rows(host).rows()
    part = host.slice_shows(1000, 2200)
    e0, e1 = int(host.entry_offsets[1000]), int(host.entry_offsets[2200])
    assert ops.csv_rows(part).rows() == whole[e0:e1]
    assert ops.csv_rows(host.to(cuda).slice_shows(1000, 2200)).rows() == whole[e0:e1]


def test_host_pipeline_many_chunks(cuda):
    """The host entry point streams the batch in chunks (upload / kernels / download overlapped on three
    streams, two staging slots).  Force many small chunks and compare with the oracle."""
    lib = _lib.load()
    host = synth_archive(6000, seed=8).pin()
    old = lib.pie_set_csv_chunk_rows(3000)
    try:
        assert lib.pie_set_csv_chunk_rows(0) == 3000
        for _ in range(2):  # second pass reuses the staging arenas
            assert_same_rows(ops.csv_rows(host), host)
        # a show larger than a chunk, empty shows around it, and a capacity overflow in a late chunk
        lib.pie_set_csv_chunk_rows(64)
        shows = ([{"id": "e", "entries": []}] * 3 + [{"id": "big", "entries": [{"id": str(i), "delaySec": i / 8} for i in range(700)]}]
                 + [{"id": "t", "entries": [{"id": "x"}]}] * 150 + [{"id": "e2", "entries": []}])
        table = pack_shows(shows)
        assert_same_rows(ops.csv_rows(table), table)
        import ctypes as C

        offsets, data = oracle_c.csv_rows(table)
        view, total = table.view(), C.c_uint64(0)
        off = torch.empty(table.n_entries + 1, dtype=torch.int64)
        buf = torch.full((data.numel() - 10,), 0xEE, dtype=torch.uint8)
        rc = lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), buf.data_ptr(), buf.numel(), C.byref(total))
        assert rc == _lib.PIE_ERR_CAPACITY and total.value == data.numel() and torch.equal(off, offsets)
    finally:
        lib.pie_set_csv_chunk_rows(old)


def test_slow_path_forced_matches_oracle(cuda):
    """Tiles that do not fit the shared-memory staging go warp-per-row from / to global memory.  Force
    every tile through that path and compare with the oracle (synthetic + the edge-case shows)."""
    lib = _lib.load()
    old = lib.pie_debug_csv_force_slow_path(1)
    try:
        assert lib.pie_debug_csv_force_slow_path(-1) == 1
        for n_shows, seed in [(1, 0), (310, 2), (3000, 9)]:
            host = synth_archive(n_shows, seed=seed)
            dev = host.to(cuda)
            assert_same_rows(ops.csv_rows(dev), host)
        shows = [{"id": 'a"b', "label": "x,y", "crew": ["A|B", 'q"', ""], "notes": "r\rn",
                  "entries": [{"id": "e1", "status": "Completed", "primaryIssue": "Battery", "actions": ["x,y", "z"],
                               "delaySec": 0, "notes": '""' * 40},
                              {"id": "e2", "delaySec": 1e21, "actions": [], "notes": 'q"' * 33 + "tail"}]}]
        table = pack_shows(shows)
        assert_same_rows(ops.csv_rows(table.to(cuda)), table)
    finally:
        lib.pie_debug_csv_force_slow_path(old)
    assert lib.pie_debug_csv_force_slow_path(-1) == old


def test_fast_path_is_the_one_that_runs(cuda):
    """The synthetic archive fits the staging buffers: no tile may fall back to the slow path (the bench
    measures the fast path), while an archive with one oversized cell sends exactly that tile there."""
    host = synth_archive(5000, seed=3)
    dev = host.to(cuda)
    sizing = ops.CsvBuffers(dev.n_entries, 0, cuda)
    ops.csv_rows_dev(dev, sizing, size_only=True)
    total = int(sizing.total.cpu())
    assert ops.csv_slow_tiles(dev, sizing) == 0
    bufs = ops.CsvBuffers(dev.n_entries, total, cuda)
    ops.csv_rows_dev(dev, bufs)
    assert ops.csv_slow_tiles(dev, bufs) == 0
    shows = [{"id": f"s{i}", "entries": [{"id": f"e{i}-{j}", "notes": "n" * (60000 if (i, j) == (40, 3) else 7)}
                                          for j in range(10)]} for i in range(100)]
    table = pack_shows(shows)
    dev = table.to(cuda)
    sizing = ops.CsvBuffers(dev.n_entries, 0, cuda)
    ops.csv_rows_dev(dev, sizing, size_only=True)
    assert ops.csv_slow_tiles(dev, sizing) == 1
    assert_same_rows(ops.csv_rows(dev), table)


def test_many_tiny_shows_and_empty_shows(cuda):
    """More shows than rows in a tile (empty shows between entries) and single-entry shows with dirty
    show-level cells: the show-level cell table of a tile is exercised at and beyond its capacity."""
    shows = []
    for i in range(700):
        shows.append({"id": f"s,{i}" if i % 3 == 0 else f"s{i}", "label": 'The "Late" show' if i % 5 == 0 else "L",
                      "crew": [f"c{i}", "x|y", 'q"z'][: i % 4], "notes": "a\nb" if i % 7 == 0 else "",
                      "entries": [{"id": f"e{i}", "delaySec": i * 0.25, "actions": ["A", "B,C", "D"][: i % 4]}]
                      if i % 2 == 0 or i > 600 else []})
    shows += [{"id": "gap", "entries": []}] * 300 + [{"id": "last", "entries": [{"id": "z"}] * 5}]
    table = pack_shows(shows)
    assert_same_rows(ops.csv_rows(table.to(cuda)), table)
    assert_same_rows(ops.csv_rows(table), table)


def test_archive_step_host_equals_the_two_calls(cuda):
    """pie_archive_step_host = pie_archive_analytics_host + pie_csv_rows_host on one upload: identical statistics,
    daily groups, summaries and CSV bytes, also when the batch is streamed in many chunks."""
    lib = _lib.load()
    host = synth_archive(6000, seed=13, shuffle_days=True, missing_created_frac=0.05).pin()
    st_ref, daily_ref = ops.archive_analytics(host, -300)
    rows_ref = ops.csv_rows(host)

    def same(a, b):
        return torch.equal(a.view(torch.int64) if a.dtype.is_floating_point else a,
                           b.view(torch.int64) if b.dtype.is_floating_point else b)

    old = lib.pie_set_csv_chunk_rows(0)
    try:
        for chunk_rows in (old, 2500, 64):
            lib.pie_set_csv_chunk_rows(chunk_rows)
            st, daily, rows = ops.archive_step(host, -300)
            assert same(st.i32, st_ref.i32) and same(st.f64.contiguous(), st_ref.f64.contiguous())
            assert daily.n_groups == daily_ref.n_groups
            for name in ("show_day_start", "show_order", "group_day_start", "group_offsets", "summary_f64", "summary_count"):
                assert same(getattr(daily, name).contiguous(), getattr(daily_ref, name).contiguous()), name
            assert torch.equal(rows.row_offsets, rows_ref.row_offsets) and torch.equal(rows.data, rows_ref.data)
    finally:
        lib.pie_set_csv_chunk_rows(old)
    # a show outside the time range raises the same error as the analytics call, whatever the rows did
    from sph_pie_b200.columnar import pack_shows as pack
    bad = pack([{"id": "a", "createdAt": 8.64e15 + 1, "entries": [{"id": "x"}]}])
    with pytest.raises(_lib.JsRangeError):
        ops.archive_step(bad, 0)
    # empty batch
    st, daily, rows = ops.archive_step(pack([]), 0)
    assert daily.n_groups == 0 and rows.data.numel() == 0


@pytest.mark.parametrize("fmt", ["csv", "payload"])
def test_bump_area_swept_across_its_brim(cuda, fmt):
    """The shared-memory bump allocator of the row kernel (csvEscape'd / joined / JSON-escaped cells are written out
    behind the staged bytes) has nothing behind it but the delaySec strings and the next stage.  Shared memory cannot
    be inspected from outside, so the canary is the output itself: free text whose every cell needs escaping, grown
    step by step from 'the bump area is hardly used' to 'no tile fits any more', must give the oracle's bytes at every
    step — an allocation that ran past its area would land in the number strings or the neighbouring stage and show
    up in them — and the sweep must actually cross the brim (both kinds of tiles are seen)."""
    from sph_pie_b200.columnar import StrCol
    from sph_pie_b200.synth import strcol_from_codes

    dev = cuda
    seen_fast = seen_slow = False
    for repeat in (1, 3, 6, 9, 12, 16, 24, 40):
        host = synth_archive(1200, seed=60 + repeat, notes_repeat=1)
        E = host.n_entries
        g = torch.Generator().manual_seed(repeat)
        # every note needs escaping (quotes, commas, line breaks, a backslash, a control character), lengths jitter
        vocab = [('say "x", then\n' * k)[: 11 * k + j] + "\\\x01" for k in (repeat,) for j in range(7)]
        for column in ("notes", "other_detail", "sub_issue", "operator_name"):  # two of them are read by each format
            host.entry_cols[column] = strcol_from_codes(torch.randint(0, len(vocab), (E,), generator=g), vocab)
        table = host.to(dev)
        if fmt == "csv":
            ref_off, ref_data = oracle_c.csv_rows(host)
            fn = ops.csv_rows_dev
        else:
            ref_off, ref_data = oracle_c.payload_rows(host)
            fn = ops.archive_payloads_dev
        sizing = ops.CsvBuffers(E, 0, dev)
        fn(table, sizing, size_only=True)
        total = int(sizing.total.cpu())
        assert total == ref_data.numel()
        bufs = ops.CsvBuffers(E, total, dev)
        fn(table, bufs)
        slow = ops.csv_slow_tiles(table, bufs)
        assert torch.equal(bufs.row_offsets.cpu(), ref_off), (fmt, repeat)
        assert torch.equal(bufs.data[:total].cpu(), ref_data), (fmt, repeat)
        n_tiles_min = (E + 159) // 160
        seen_fast |= slow < n_tiles_min
        seen_slow |= slow > 0
    assert seen_fast and seen_slow
