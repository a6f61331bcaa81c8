"""Archive entry payloads — JSON.stringify(buildArchiveEntryPayload(show, entry)) per entry (reference
server/webhookDispatcher.js:315-330, :527-540) — on the GPU against the oracles.  Bit-exact bytes and offsets."""
import json

import pytest
import torch

import oracle_c
import pie_oracle as po
from sph_pie_b200 import _lib, ops
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows
from sph_pie_b200.webhook import archiveEntryPayloadBodies, archiveEntryPayloadBodiesMany, buildArchiveEntryPayload

pytestmark = pytest.mark.gpu


def assert_same_rows(got: ops.CsvRows, table):
    offsets, data = oracle_c.payload_rows(table)
    assert torch.equal(got.row_offsets.cpu(), offsets), "row_offsets"
    assert torch.equal(got.data.cpu(), data), "payload bytes"


@pytest.mark.parametrize("n_shows,seed", [(1, 0), (7, 1), (310, 2), (5000, 3), (40000, 4)])
def test_payloads_match_c_oracle_both_entry_points(cuda, n_shows, seed):
    host = synth_archive(n_shows, seed=seed)
    assert_same_rows(ops.archive_payloads(host), host)
    assert_same_rows(ops.archive_payloads(host.to(cuda)), host)


EDGE_SHOWS = [
    {"id": "a", "date": "2024-07-04", "time": "21:00", "label": 'quote " backslash \\ slash /', "leadPilot": "l\np",
     "monkeyLead": "tab\there", "entries": [
         {"operator": "\x00\x01\x1f\x7f ctl", "unitId": "bell\x07 vt\x0b ff\x0c bs\x08 cr\r", "planned": " YES ",
          "launched": " yes ", "commandRx": "yes!", "primaryIssue": "漢字 🚁 ü", "subIssue": "  "},
         {"planned": "yes", "launched": "no", "commandRx": "Yes\n", "primaryIssue": "", "subIssue": None},
         {"planned": "y e s", "launched": "﻿YES", "commandRx": "", "operator": '""', "unitId": "\\\\"},
         {"planned": "ＹＥＳ", "launched": "yeſ", "commandRx": "\tyEs\t"}]},
    {"id": "empty", "entries": []}, None,
    {"id": "long", "label": "L" * 300, "entries": [{"operator": "x" * 70000 + '"' + "\n" * 50, "planned": "yes"},
                                                     {"operator": "after"}]},
]


def test_edge_payloads_and_mirror_api(cuda):
    many = archiveEntryPayloadBodiesMany(EDGE_SHOWS)
    for show, rows in zip(EDGE_SHOWS, many):
        entries = (show or {}).get("entries", [])
        assert rows == [po.archive_entry_payload_json(show, e) for e in entries]
        for row, e in zip(rows, entries):  # and it is JSON: parses back to the payload object
            assert json.loads(row) == po.build_archive_entry_payload(show, e)
    first = EDGE_SHOWS[0]
    assert buildArchiveEntryPayload(first, first["entries"][1]) == po.build_archive_entry_payload(first, first["entries"][1])
    assert buildArchiveEntryPayload() == po.build_archive_entry_payload()
    assert archiveEntryPayloadBodies(None) == []
    table = pack_shows(EDGE_SHOWS)
    assert_same_rows(ops.archive_payloads(table.to(cuda)), table)
    assert_same_rows(ops.archive_payloads(table), table)


def test_reference_fixture_payload(cuda):
    """The reference's only fixture (scripts/simulate-webhook.js:42-65) through buildArchiveEntryPayload."""
    import os

    fix = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "webhook_fixture.json")))
    show = dict(fix["show"], entries=[fix["entry"]])
    assert archiveEntryPayloadBodies(show) == [fix["expected_archive_payload_json"]]
    assert buildArchiveEntryPayload(fix["show"], fix["entry"]) == fix["expected_archive_entry_payload"]


def test_slow_path_forced_payloads(cuda):
    lib = _lib.load()
    old = lib.pie_debug_csv_force_slow_path(1)
    try:
        for n_shows, seed in [(1, 0), (310, 2), (3000, 9)]:
            host = synth_archive(n_shows, seed=seed)
            assert_same_rows(ops.archive_payloads(host.to(cuda)), host)
        table = pack_shows(EDGE_SHOWS)
        assert_same_rows(ops.archive_payloads(table.to(cuda)), table)
    finally:
        lib.pie_debug_csv_force_slow_path(old)


def test_fast_path_and_json_lines_round_trip(cuda):
    """No tile of the synthetic archive may fall back to the slow path; and the whole output is JSON Lines:
    every line parses to the payload object of its entry (size-independent property at a larger size)."""
    host = synth_archive(20000, seed=6)
    dev = host.to(cuda)
    sizing = ops.CsvBuffers(dev.n_entries, 0, cuda)
    ops.archive_payloads_dev(dev, sizing, size_only=True)
    total = int(sizing.total.cpu())
    assert ops.csv_slow_tiles(dev, sizing) == 0
    bufs = ops.CsvBuffers(dev.n_entries, total, cuda)
    ops.archive_payloads_dev(dev, bufs)
    assert ops.csv_slow_tiles(dev, bufs) == 0
    lines = bytes(bufs.data[:total].cpu().numpy()).decode("utf-8").split("\n")
    assert lines[-1] == "" and len(lines) == host.n_entries + 1
    i = 0
    for show in table_to_shows(host):
        for entry in show["entries"]:
            assert json.loads(lines[i]) == po.build_archive_entry_payload(show, entry), i
            i += 1


def test_payload_columns_only(cuda):
    """The entry point reads 12 string columns; every other column of the view may be NULL."""
    import ctypes as C

    host = synth_archive(500, seed=12)
    dev = host.to(cuda)
    view = dev.view()
    for name in ("show_id", "show_notes", "entry_id", "status", "other_detail", "severity", "root_cause", "battery_id",
                 "notes"):
        col = getattr(view, name)
        col.offsets = None
        col.data = None
    view.crew.list_offsets = None
    view.actions.list_offsets = None
    view.delay_sec = None
    view.delay_valid = None
    offsets, data = oracle_c.payload_rows(host)
    bufs = ops.CsvBuffers(dev.n_entries, data.numel(), cuda)
    lib = _lib.load()
    _lib.check(lib.pie_archive_payloads_dev(C.byref(view), bufs.row_offsets.data_ptr(), bufs.data.data_ptr(), bufs.capacity,
                                            bufs.total.data_ptr(), bufs.scratch.data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(bufs.row_offsets.cpu(), offsets) and torch.equal(bufs.data.cpu(), data)
