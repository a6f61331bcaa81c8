"""_getTimestamp of the documents' time fields, _mapArchiveRow in full, and the archive / purge decisions on the GPU
(csrc/archive_maintenance.cu through the C ABI) against the Python restatement of the reference
(server/storage/sqlProvider.js:758-816, :863-926, :970-1009).  Bit-exact: every value is an integer-valued or
correctly rounded double, every decision a boolean."""
import json
import math
import random

import numpy as np
import pytest
import torch

import pie_oracle as po
from sph_pie_b200 import _lib, ops, storage

pytestmark = pytest.mark.gpu

DAY = 86400000.0
T0 = 1704067200000.0  # 2024-01-01T00:00:00Z

TIME_VALUES = [None, True, False, 0, -0.0, 1.5, T0, -T0, 8.64e15, 8.64e15 + 1, 1e300, "", "  ", " 12 ", "1e3", ".5", "5.", "007",
               "+7", "-7.25", "0x1F", "0b11", "0o17", "1704067200000", "2024-01-01", "2024-01-01T00:00:00.000Z",
               "2024-03-05T10:20", "2024-03-05T10:20:30", "2024-03-05T10:20:30.123+05:30", "2024-13-01", "2024-01-01T25:00",
               " 42　", "123456789012345678901234567890"]


def _nan_to_none(t):
    return [None if math.isnan(x) else x for x in t.cpu().tolist()]


@pytest.mark.parametrize("tz", [0, -480, 330])
def test_get_timestamps_matches_the_oracle(cuda, tz):
    rng = random.Random(11)
    shows = []
    for i in range(400):
        show = {"id": f"s{i}"}
        for key in ("createdAt", "updatedAt", "archivedAt", "deletedAt"):
            if rng.random() < 0.8:
                show[key] = rng.choice(TIME_VALUES)
        shows.append(show)
    texts = [json.dumps(s, ensure_ascii=False) for s in shows] + ["null", "[]", "{}"]
    docs = ops.JsonDocs.from_texts(texts).to(cuda)
    table, status = ops.ingest_json(docs)
    times = ops.get_timestamps(table, docs, tz)
    parsed = [po.map_archive_row(t) for t in texts]
    for name, key in (("created_at", "createdAt"), ("updated_at", "updatedAt"), ("archived_at", "archivedAt"),
                      ("deleted_at", "deletedAt")):
        got = _nan_to_none(getattr(times, name))
        for i, show in enumerate(parsed):
            v = show[key] if isinstance(show, dict) and key in show else po.UNDEFINED
            want = po.get_timestamp_tz(v, tz)
            assert got[i] == want and (want is None or math.copysign(1, got[i]) == math.copysign(1, want)), (name, i, v, got[i], want)


def test_unsupported_time_texts_fail_loudly(cuda):
    for text in ['{"createdAt":"next tuesday"}', '{"updatedAt":"01/02/2024"}', '{"deletedAt":"2024-02-30"}',
                 '{"createdAt":"Infinity"}', '{"createdAt":"12\\u0033"}']:
        docs = ops.JsonDocs.from_texts(['{"id":"ok"}', text]).to(cuda)
        table, _ = ops.ingest_json(docs)
        with pytest.raises(_lib.UnsupportedDateError) as e:
            ops.get_timestamps(table, docs, 0)
        assert "show 1" in e.value.message
    for text in ['{"createdAt":[5]}', '{"archivedAt":{}}']:
        docs = ops.JsonDocs.from_texts([text]).to(cuda)
        table, _ = ops.ingest_json(docs)
        with pytest.raises(_lib.SchemaError):
            ops.get_timestamps(table, docs, 0)
    # a text timestamp, and no documents to read it from
    docs = ops.JsonDocs.from_texts(['{"createdAt":"5"}']).to(cuda)
    table, _ = ops.ingest_json(docs)
    with pytest.raises(_lib.SchemaError):
        ops.get_timestamps(table, None, 0)
    assert ops.get_timestamps(table, docs, 0).created_at.tolist() == [5.0]


@pytest.mark.parametrize("tz,provider", [(0, "sql"), (-420, "sql"), (0, "postgres")])
def test_map_archive_rows_in_full(cuda, tz, provider):
    """storage.mapArchiveRows against _mapArchiveRow as the oracle restates it, row columns included."""
    rng = random.Random(5)
    cols = [None, "1704067200000", " 1704067200001 ", "", "2024-02-03T04:05:06.007Z", 17.5, "0x10"]
    rows = []
    for i in range(300):
        show = {"id": f"s{i}", "entries": [{"status": "Completed"}] * rng.randrange(0, 3)}
        for key in ("createdAt", "archivedAt", "deletedAt", "updatedAt"):
            if rng.random() < 0.7:
                show[key] = rng.choice([None, True, 0, T0 + i, "", " 12 ", "2024-01-01T00:00:00.000Z", "1e3"])
        row = {"data": json.dumps(show)}
        for c in ("archived_at", "created_at", "deleted_at"):
            if rng.random() < 0.7:
                row[c] = rng.choice(cols)
        rows.append(row)
    rows += [{"data": "broken", "created_at": "5"}, {"data": "[]", "archived_at": "7"}, {"data": None}, "null", '{"id":"plain text row"}']
    table, dropped = storage.mapArchiveRows(rows, cuda, tz, provider)
    want = [po.map_archive_row_all(r if isinstance(r, dict) else {"data": r}, tz, provider) for r in rows]
    assert dropped.cpu().tolist() == [w is None for w in want]
    created, archived, deleted = (_nan_to_none(t) for t in (table.created_at, table.archived_at, table.deleted_at))
    kinds = table.time_kind.cpu().tolist()
    for i, w in enumerate(want):
        if w is None:
            continue
        for name, got, f in (("createdAt", created, _lib.TF_CREATED), ("archivedAt", archived, _lib.TF_ARCHIVED),
                             ("deletedAt", deleted, _lib.TF_DELETED)):
            v = w.get(name, po.UNDEFINED)
            finite = po.js_is_number(v) and math.isfinite(v)  # what Number.isFinite(show.<name>) sees downstream
            assert got[i] == (float(v) if finite else None), (i, name, v, got[i])
            if name == "deletedAt" and name not in w:
                assert kinds[i][f] == _lib.TK_ABSENT  # `delete show.deletedAt`
    # the analytics downstream read the mapped createdAt: a document whose createdAt is null counts as the epoch
    t2, _ = storage.mapArchiveRows(['{"id":"z","createdAt":null,"entries":[]}'], cuda, tz)
    _, daily = ops.archive_analytics(t2, tz)
    assert daily.n_groups == 1 and int(daily.show_day_start[0]) == int(po.local_day_start(0.0, tz))


def test_archive_due_matches_the_oracle(cuda):
    rng = random.Random(3)
    now = T0 + 40 * DAY
    dates = ["2024-02-%02d" % d for d in range(1, 20)] + ["", "  ", " 2024-02-01 ", "2024-02-01　", "__undated__", "x" * 70]
    rows = []
    for i in range(3000):
        show = {"id": f"s{i}"}
        r = rng.random()
        if r < 0.85:
            show["date"] = rng.choice(dates)
        elif r < 0.9:
            show["date"] = None
        created = rng.choice([now - rng.randrange(0, int(1.2 * 43200000)), now - 43200000, now - 43199999, None, "abc-skip"])
        if created != "abc-skip":
            show["createdAt"] = created
        if rng.random() < 0.3:
            show["updatedAt"] = rng.choice([now - 50000000, now, str(int(now - 43200000)), None])
        rows.append(json.dumps(show))
        if rng.random() < 0.02:
            rows.append(rng.choice(["oops", "12", "null", "[]", '"text"']))
    due, order = storage.archiveDailyShowsDecision(rows, now, 0, cuda)
    want_due, want_order = po.archive_daily_shows_decision(rows, now, 0)
    assert due == want_due
    assert order == want_order
    assert any(due) and not all(due)
    # hand-checked: one show without any timestamp makes its date group due (Number(null) === 0)
    rows = ['{"date":"d1","createdAt":%d}' % int(now - 1000), '{"date":"d2","createdAt":%d}' % int(now - 1000), '{"date":"d1"}']
    assert storage.archiveDailyShowsDecision(rows, now, 0, cuda) == ([True, False, True], [0, 2])


@pytest.mark.parametrize("tz", [0, -480, 330, 840])
def test_archive_expired_matches_the_oracle(cuda, tz):
    rng = random.Random(tz + 1)
    created = []
    for y in (1969, 1970, 1999, 2000, 2023, 2024, 2100):
        for m in range(1, 13):
            for d in (1, 28, 29, 30, 31):
                if d <= po.days_in_month(y, m):
                    for h in (0, 7.99, 23.999):
                        created.append(po.days_from_civil(y, m, d) * DAY + h * 3600000)
    created += [0.0, -0.0, -1.0, 0.5, 8.64e15, -8.64e15, 8.64e15 - 1, 8.64e15 + 2, 8.6399e15, 1e300, -1e300, math.nan, math.inf, -math.inf]
    created += [rng.uniform(-4e12, 4e12) for _ in range(2000)]
    t = torch.tensor(created, dtype=torch.float64, device=cuda)
    for now in (T0, T0 + 61 * DAY, 4102444800000.0, -1e12, 8.64e15):
        got = ops.archive_expired(t, now, tz).cpu().tolist()
        want = [1 if po.is_archive_expired(c, now, tz) else 0 for c in created]
        assert got == want, [(c, g, w) for c, g, w in zip(created, got, want) if g != w][:5]
    # the boundary itself: now == expiry is expired, one millisecond earlier is not
    for c in created[:400]:
        expiry = po.add_months(c, 2, tz)
        pair = ops.archive_expired(torch.tensor([c, c], dtype=torch.float64, device=cuda), expiry, tz).cpu().tolist()
        assert pair == [1, 1]
        assert ops.archive_expired(torch.tensor([c], dtype=torch.float64, device=cuda), expiry - 1, tz).cpu().tolist() == [0]


def test_purge_decision_end_to_end(cuda):
    d = lambda y, m, dd: float(po.days_from_civil(y, m, dd) * 86400000)
    rows = [{"data": '{"createdAt":%d}' % int(d(2024, 1, 1)), "created_at": "1"},
            {"data": "{}", "created_at": str(int(d(2024, 1, 1)))},
            {"data": "broken", "created_at": None},
            {"data": "{}"},
            {"data": '{"createdAt":"2024-01-01T00:00:00.000Z"}'},
            {"data": "[]", "created_at": str(int(d(2023, 12, 31)))},
            {"data": '"text"', "created_at": str(int(d(2024, 2, 1)))}]
    for now in (d(2024, 3, 1), d(2024, 3, 1) - 1, d(2024, 3, 2), d(2024, 4, 1)):
        assert storage.purgeExpiredArchivesDecision(rows, now, 0, cuda) == po.purge_expired_archives_decision(rows, now, 0)
    assert storage.purgeExpiredArchivesDecision(rows, d(2024, 3, 1), 0, cuda) == [True, True, True, False, True, False, False]


def test_time_fields_of_the_ingested_table(cuda):
    """ABI 2 columns of the GPU ingest (updated_at, deleted_at, time_kind) against the table packer on JSON.parse's
    values — both entry points."""
    from ingest_helpers import assert_tables_equal, oracle_ingest

    rng = random.Random(9)
    docs = []
    for i in range(500):
        show = {"id": f"s{i}", "entries": []}
        for key in ("createdAt", "updatedAt", "archivedAt", "deletedAt"):
            if rng.random() < 0.8:
                show[key] = rng.choice([None, True, False, 0, -0.0, T0 + i, 1.5, "text", "", [1], {"a": 1}])
        docs.append(json.dumps(show))
    docs += ['{"createdAt":1e999,"deletedAt":-1e999}', "null", "[]", '{"updatedAt":5e-324}']
    ref, ref_status = oracle_ingest(docs)
    d = ops.JsonDocs.from_texts(docs)
    for dd in (d, d.to(cuda)):
        table, status = ops.ingest_json(dd)
        assert status.cpu().tolist() == ref_status.tolist()
        assert_tables_equal(table, ref)
        assert torch.equal(table.time_kind.cpu(), ref.time_kind)
        for name in ("updated_at", "deleted_at"):
            a, b = getattr(table, name).cpu().numpy(), getattr(ref, name).numpy()
            assert np.array_equal(np.isnan(a), np.isnan(b))
            assert np.array_equal(a[~np.isnan(a)].view(np.int64), b[~np.isnan(b)].view(np.int64))
