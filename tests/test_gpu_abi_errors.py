"""The host-buffer ABI fails loudly, and cleanly, on malformed input: offsets arrays are validated on the device
before any kernel walks them (csrc/validate.cu), every stream is drained before the call returns, and the next call
works.  pie_release gives the library's device memory back."""
import pytest
import torch

from sph_pie_b200 import _lib, ops
from sph_pie_b200.columnar import StrCol
from sph_pie_b200.synth import synth_archive

pytestmark = pytest.mark.gpu


def _break(col: StrCol, row: int, how: str) -> StrCol:
    off = col.offsets.clone()
    if how == "decrease":
        off[row] = off[row + 1] + 7
    else:  # past the heap
        off[row] = int(off[-1]) + 1000
    return StrCol(off, col.data)


@pytest.mark.parametrize("column,how", [("status", "decrease"), ("notes", "beyond"), ("primary_issue", "decrease")])
def test_malformed_entry_offsets_are_reported(cuda, column, how):
    import oracle_c

    host = synth_archive(3000, seed=5)
    good_off, good_csv = oracle_c.csv_rows(host)
    bad = synth_archive(3000, seed=5)
    bad.entry_cols[column] = _break(bad.entry_cols[column], 1234, how)
    with pytest.raises(_lib.PieError) as e:
        ops.csv_rows(bad)
    assert e.value.code == _lib.PIE_ERR_INVALID_ARG and column in e.value.message
    if column in ("status", "primary_issue"):
        with pytest.raises(_lib.PieError) as e:
            ops.archive_analytics(bad, tz_offset_minutes=0)
        assert e.value.code == _lib.PIE_ERR_INVALID_ARG and column in e.value.message
        with pytest.raises(_lib.PieError) as e:
            ops.compute_metrics(bad)
        assert e.value.code == _lib.PIE_ERR_INVALID_ARG
    # nothing sticks: the same call on the well-formed table gives the oracle's bytes
    rows = ops.csv_rows(host)
    assert torch.equal(rows.row_offsets, good_off) and torch.equal(rows.data, good_csv)


def test_malformed_list_and_show_offsets(cuda):
    bad = synth_archive(500, seed=6)
    lo = bad.actions.list_offsets.clone()
    lo[77] = lo[78] + 3
    bad.actions = type(bad.actions)(lo, bad.actions.items)
    with pytest.raises(_lib.PieError) as e:
        ops.csv_rows(bad)
    assert e.value.code == _lib.PIE_ERR_INVALID_ARG and "actions" in e.value.message
    bad = synth_archive(500, seed=6)
    bad.show_cols["show_date"] = _break(bad.show_cols["show_date"], 10, "decrease")
    with pytest.raises(_lib.PieError) as e:
        ops.archive_step(bad, 0)
    assert e.value.code == _lib.PIE_ERR_INVALID_ARG and "show_date" in e.value.message
    eo = synth_archive(500, seed=6)
    eo.entry_offsets = eo.entry_offsets.clone()
    eo.entry_offsets[100] = eo.entry_offsets[101] + 1
    with pytest.raises(_lib.PieError) as e:
        ops.csv_rows(eo)
    assert e.value.code == _lib.PIE_ERR_INVALID_ARG and "entry_offsets" in e.value.message


def test_multi_chunk_pipeline_reports_a_late_chunk(cuda):
    lib = _lib.load()
    old = lib.pie_set_csv_chunk_rows(4096)
    try:
        bad = synth_archive(4000, seed=7)
        bad.entry_cols["notes"] = _break(bad.entry_cols["notes"], bad.n_entries - 50, "decrease")
        with pytest.raises(_lib.PieError) as e:
            ops.archive_step(bad, -480)
        assert e.value.code == _lib.PIE_ERR_INVALID_ARG and "notes" in e.value.message
        good = synth_archive(4000, seed=7)
        st, daily, rows = ops.archive_step(good, -480)
        assert rows.row_offsets.numel() == good.n_entries + 1
    finally:
        lib.pie_set_csv_chunk_rows(old)


def test_release_and_reuse(cuda):
    import oracle_c

    lib = _lib.load()
    host = synth_archive(2000, seed=8)
    ref_off, ref_csv = oracle_c.csv_rows(host)
    ops.archive_step(host, 0)
    free0, _ = torch.cuda.mem_get_info()
    _lib.check(lib.pie_release())
    free1, _ = torch.cuda.mem_get_info()
    assert free1 >= free0  # the arenas went back to the driver
    _, _, rows = ops.archive_step(host, 0)  # everything is allocated again on demand
    assert torch.equal(rows.row_offsets, ref_off) and torch.equal(rows.data, ref_csv)
    docs = ops.JsonDocs.from_texts(['{"id":"a","entries":[{"status":"Completed"}]}'])
    _lib.check(lib.pie_release())
    table, status = ops.ingest_json(docs)
    assert table.n_entries == 1 and status.tolist() == [0]


def test_plain_c_caller_from_stored_texts(cuda, tmp_path):
    """tests/native/c_consumer.c: a C99 program with nothing but malloc'd (pageable) buffers hands the provider's stored
    texts (one JSON.stringify(show) per line) to pie_archive_step_json_host and writes the CSV rows, the statistics
    planes and the dropped-row flags: byte for byte what the C oracle makes of the same shows."""
    import json
    import subprocess

    import numpy as np
    import oracle_c
    from helpers import build_c_consumer
    from sph_pie_b200.columnar import pack_shows
    from sph_pie_b200.synth import table_to_shows

    host = synth_archive(2500, seed=61, shuffle_days=True)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)  # JSON.stringify writes NaN / Infinity as null
    host.delay_valid[lost] = 0
    host.delay_sec[lost] = 0.0
    shows = table_to_shows(host)
    texts = [json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in shows]
    for k, junk in ((7, "not json"), (300, ""), (1999, '"a string"'), (2400, '{"id":"cut')):  # rows the reference maps to null
        shows[k], texts[k] = None, junk
    assert not any("\n" in t for t in texts)
    path = tmp_path / "docs.jsonl"
    path.write_bytes(("\n".join(texts) + "\n").encode("utf-8"))
    exe = build_c_consumer()
    r = subprocess.run([exe, str(path), "-300", str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr)
    ref = pack_shows(shows)
    ref_offsets, ref_csv = oracle_c.csv_rows(ref)
    ref_stats, ref_daily, rc, _ = oracle_c.archive_analytics(ref, tz_offset_minutes=-300)
    assert rc == 0
    n_docs, n_entries, csv_bytes, n_groups, dropped = (int(x) for x in r.stdout.split())
    assert (n_docs, n_entries, csv_bytes, dropped) == (len(texts), ref.n_entries, int(ref_offsets[-1]), 4)
    assert n_groups == ref_daily.n_groups
    assert (tmp_path / "out.csv").read_bytes() == bytes(ref_csv.numpy())
    status = np.frombuffer((tmp_path / "out.status").read_bytes(), dtype=np.uint8)
    assert status.nonzero()[0].tolist() == [7, 300, 1999, 2400]
    stats = np.frombuffer((tmp_path / "out.stats_i32").read_bytes(), dtype=np.int32).reshape(_lib.PIE_SI_COUNT, n_docs)
    assert np.array_equal(stats, ref_stats.i32.numpy())
