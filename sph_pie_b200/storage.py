"""Host-side mirror of the provider's read and maintenance paths for stored shows (reference server/storage/sqlProvider.js):
`listArchivedShows` (:230-234) selects `data` from show_archive and maps every row through `_mapArchiveRow`
(:892-926 — JSON.parse, null unless an object) before `.filter(Boolean)`.  Here the rows' texts go to the GPU as they
are and come back as the columnar archive table the analytics and export operators read; no document is ever
materialised as an object.  There is no CPU fallback."""
from __future__ import annotations

import datetime
import math
import re
from typing import Iterable, Tuple, Union

import torch

from . import ops
from ._lib import UnsupportedDateError
from .columnar import ArchiveTable

Row = Union[str, bytes, dict]


_MISSING = object()
_JS_WS = "\t\n\v\f\r \u00a0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000\ufeff"
_DECIMAL = re.compile(r"[+-]?([0-9]+\.?[0-9]*([eE][+-]?[0-9]+)?|\.[0-9]+([eE][+-]?[0-9]+)?)")
_RADIX = re.compile(r"0[xX]([0-9a-fA-F]+)|0[oO]([0-7]+)|0[bB]([01]+)")


_ISO = re.compile(r"(\d{4})-(\d{2})-(\d{2})(?:T(\d{2}):(\d{2})(?::(\d{2})(?:\.(\d{3}))?)?(Z|[+-]\d{2}:\d{2})?)?")


def _days_from_civil(y: int, m: int, d: int) -> int:
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    return era * 146097 + yoe * 365 + yoe // 4 - yoe // 100 + doy - 719468


def _date_parse(text: str, tz_offset_minutes: int) -> float:
    """Date.parse for the format ECMA-262 specifies (a row's column that holds an ISO text); anything else is V8's
    legacy parser and raises."""
    m = _ISO.fullmatch(text)
    if not m:
        raise UnsupportedDateError(-4, f"row timestamp {text!r}: neither numeric nor an ECMA-262 date-time string")
    y, mo, d = int(m.group(1)), int(m.group(2)), int(m.group(3))
    h, mi = (int(m.group(4)), int(m.group(5))) if m.group(4) else (0, 0)
    sec, ms, off = int(m.group(6) or 0), int(m.group(7) or 0), m.group(8)
    if off and off != "Z" and (int(off[1:3]) > 23 or int(off[4:6]) > 59):
        return math.nan
    if not (1 <= mo <= 12 and 1 <= d <= 31 and h <= 24 and mi <= 59 and sec <= 59) or (h == 24 and (mi or sec or ms)):
        return math.nan
    dim = 31 if mo in (1, 3, 5, 7, 8, 10, 12) else 30 if mo != 2 else (29 if y % 4 == 0 and (y % 100 or y % 400 == 0) else 28)
    if d > dim:
        raise UnsupportedDateError(-4, f"row timestamp {text!r}: a day past the end of the month")
    local = ((_days_from_civil(y, mo, d) * 24 + h) * 60 + mi) * 60000 + sec * 1000 + ms
    if m.group(4) is None or off == "Z":
        return float(local)
    if off:
        return float(local - (-1 if off[0] == "-" else 1) * (int(off[1:3]) * 60 + int(off[4:6])) * 60000)
    return float(local - tz_offset_minutes * 60000)


def _row_timestamp(v, tz_offset_minutes: int = 0) -> float:
    """_getTimestamp(value) (sqlProvider.js:970-985) for a row's column: the number, or NaN for null (the JS function's
    `null`).  A column the SELECT did not fetch is undefined -> NaN; SQL NULL is JS null -> Number(null) = 0; a text
    goes through Number(text) (StringToNumber), then Date.parse (the ECMA-262 format; any other text raises instead of
    guessing what V8's legacy parser would make of it)."""
    if v is _MISSING:
        return math.nan
    if v is None:
        return 0.0
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, (int, float)):
        f = float(v)
        return f if math.isfinite(f) else math.nan
    if isinstance(v, datetime.datetime):  # a TIMESTAMPTZ column of the Postgres provider: pg hands back a Date, Number(date) = its time value
        if v.tzinfo is None:
            v = v.replace(tzinfo=datetime.timezone.utc)
        delta = v - datetime.datetime(1970, 1, 1, tzinfo=datetime.timezone.utc)
        return float(delta.days * 86400000 + delta.seconds * 1000 + delta.microseconds // 1000)
    if isinstance(v, str):
        t = v.strip(_JS_WS)
        if t == "":
            return 0.0
        m = _RADIX.fullmatch(t)
        if m:
            return float(int(m.group(1) or m.group(2) or m.group(3), 16 if m.group(1) else 8 if m.group(2) else 2))
        if _DECIMAL.fullmatch(t):
            f = float(t)
            if math.isfinite(f):
                return f
        parsed = _date_parse(v, tz_offset_minutes)
        return parsed if math.isfinite(parsed) else math.nan
    return math.nan


def _text(row: Row):
    if isinstance(row, dict):  # a result row of `SELECT data, ... FROM show_archive` (sqlProvider.js:232)
        row = row.get("data")
    if row is None:
        return b"null"  # JSON.parse(null) is null: the row is dropped
    return row


def mapArchiveRows(rows: Iterable[Row], device="cuda", tz_offset_minutes: int = 0,
                   provider: str = "sql") -> Tuple[ArchiveTable, torch.Tensor]:
    """rows.map(row => this._mapArchiveRow(row)) for a batch (sqlProvider.js:892-926): (table, dropped) where dropped[i]
    is True for the rows the reference maps to null (text that is not JSON, or not an object); their table rows are
    empty shows.  Rows given as dicts may carry the `archived_at` / `created_at` / `deleted_at` columns of the SELECT.
    Every timestamp goes through _getTimestamp as the reference's does — the document's own fields on the GPU
    (pie_get_timestamps_dev: null is 0, numeric text its number, an ISO text Date.parse), the row's columns here:
      archivedAt = _getTimestamp(row.archived_at) ?? _getTimestamp(show.archivedAt)   set when not null
      createdAt  = _getTimestamp(show.createdAt) ?? _getTimestamp(row.created_at)     set when not null
      deletedAt  = _getTimestamp(row.deleted_at) ?? _getTimestamp(show.deletedAt)     set when not null, else deleted
    provider="postgres" restates postgresProvider.js:711-741 instead, which prefers the ROW's created_at.
    The table stays in HBM for the operators that follow; there is no CPU path."""
    from . import _lib

    rows = list(rows)
    want_host = str(device) == "cpu"  # the work is the GPU's either way; "cpu" only says where the table should end up
    docs = ops.JsonDocs.from_texts([_text(r) for r in rows]).to("cuda" if want_host else device)
    table, status = ops.ingest_json(docs)
    dropped = status.bool()
    dev = table.created_at.device
    times = ops.get_timestamps(table, docs, tz_offset_minutes)

    def column(name):
        vals = [_row_timestamp(r.get(name, _MISSING), tz_offset_minutes) if isinstance(r, dict) else math.nan for r in rows]
        return torch.tensor(vals, dtype=torch.float64).to(dev)

    nan = torch.full((len(rows),), math.nan, dtype=torch.float64, device=dev)
    keep = ~dropped.to(dev)

    def first(a, b):  # a ?? b on NaN-for-null
        return torch.where(torch.isfinite(a), a, b)

    archived = torch.where(keep, first(column("archived_at"), times.archived_at), nan)
    if provider == "postgres":
        created = torch.where(keep, first(column("created_at"), times.created_at), nan)
    else:
        created = torch.where(keep, first(times.created_at, column("created_at")), nan)
    deleted = torch.where(keep, first(column("deleted_at"), times.deleted_at), nan)
    kind = table.time_kind.clone()
    number = torch.full_like(kind[:, 0], _lib.TK_NUMBER)
    # a field the reference sets is a number now; one it leaves alone keeps whatever it held (which Number.isFinite
    # rejects: NaN in the value column); deletedAt is deleted when null
    kind[:, _lib.TF_ARCHIVED] = torch.where(torch.isfinite(archived), number, kind[:, _lib.TF_ARCHIVED])
    kind[:, _lib.TF_CREATED] = torch.where(torch.isfinite(created), number, kind[:, _lib.TF_CREATED])
    kind[:, _lib.TF_DELETED] = torch.where(torch.isfinite(deleted), number, torch.zeros_like(number))
    table.archived_at, table.created_at, table.deleted_at, table.time_kind = archived, created, deleted, kind
    if want_host:
        return table.to("cpu"), dropped.cpu()
    return table, dropped


def listArchivedShows(rows: Iterable[Row], device="cuda") -> ArchiveTable:
    """rows.map(_mapArchiveRow).filter(Boolean): the table of the rows that survive, in row order."""
    table, dropped = mapArchiveRows(rows, device)
    if not bool(dropped.any()):
        return table
    keep = (~dropped).nonzero().flatten().cpu().tolist()
    # dropped rows are empty shows: removing them only touches the show-level columns
    return _select_shows(table, keep)


def _select_shows(table: ArchiveTable, keep) -> ArchiveTable:
    from .columnar import StrCol, StrListCol

    dev = table.entry_offsets.device
    idx = torch.tensor(keep, dtype=torch.int64, device=dev)

    def sel_offsets(off: torch.Tensor) -> torch.Tensor:
        # a dropped show is empty, so offsets[i] == offsets[i+1] there: the kept starts plus the final end
        return torch.cat([off[idx], off[table.n_shows:table.n_shows + 1]]).contiguous()

    return ArchiveTable(
        n_shows=len(keep), n_entries=table.n_entries, entry_offsets=sel_offsets(table.entry_offsets),
        show_cols={k: StrCol(sel_offsets(c.offsets), c.data) for k, c in table.show_cols.items()},
        crew=StrListCol(sel_offsets(table.crew.list_offsets), table.crew.items),
        created_at=table.created_at[idx].contiguous(), archived_at=table.archived_at[idx].contiguous(),
        entry_cols=table.entry_cols, actions=table.actions, delay_sec=table.delay_sec, delay_valid=table.delay_valid,
        entry_ts=table.entry_ts,
        updated_at=None if table.updated_at is None else table.updated_at[idx].contiguous(),
        deleted_at=None if table.deleted_at is None else table.deleted_at[idx].contiguous(),
        time_kind=None if table.time_kind is None else table.time_kind[idx].contiguous())


def archiveDailyShowsDecision(rows: Iterable[Row], now_ms: float, tz_offset_minutes: int = 0, device="cuda"):
    """Which rows of `shows` _archiveDailyShows (sqlProvider.js:758-816) archives at `now_ms`, and in which order it
    saves / dispatches them: (due: list[bool], order: list[int]).  JSON.parse, the date grouping, `createdAt ??
    updatedAt` through _getTimestamp and the 12 h rule all run on the GPU; the order is (first row of the date group,
    row) — a Map keeps its keys in insertion order."""
    rows = list(rows)
    docs = ops.JsonDocs.from_texts([_text(r) for r in rows]).to(device)
    table, status = ops.ingest_json(docs)
    times = ops.get_timestamps(table, docs, tz_offset_minutes)
    created = torch.where(torch.isfinite(times.created_at), times.created_at, times.updated_at)
    due, first = ops.archive_due(table, created, now_ms, status)
    due, first = due.cpu().bool(), first.cpu()
    idx = due.nonzero().flatten().tolist()
    order = sorted(idx, key=lambda i: (int(first[i]), i))
    return due.tolist(), order


def purgeExpiredArchivesDecision(rows: Iterable[dict], now_ms: float, tz_offset_minutes: int = 0, device="cuda"):
    """Which rows of `show_archive` _purgeExpiredArchives (sqlProvider.js:863-890) deletes at `now_ms`: list[bool].
    rows: dicts with `data` and (optionally) `created_at`.  A text that does not parse, or parses to something that is
    not an object, has no createdAt of its own (`show?.createdAt` is undefined): the row's column decides."""
    rows = list(rows)
    docs = ops.JsonDocs.from_texts([_text(r) for r in rows]).to(device)
    table, status = ops.ingest_json(docs)
    times = ops.get_timestamps(table, docs, tz_offset_minutes)
    dev = table.created_at.device
    col = torch.tensor([_row_timestamp(r.get("created_at", _MISSING), tz_offset_minutes) if isinstance(r, dict) else math.nan
                        for r in rows], dtype=torch.float64).to(dev)
    doc_created = torch.where(status.bool().to(dev), torch.full_like(col, math.nan), times.created_at)
    created = torch.where(torch.isfinite(doc_created), doc_created, col)
    return ops.archive_expired(created, now_ms, tz_offset_minutes).cpu().bool().tolist()
