"""Host-side mirror of the provider's read path for archived shows (reference server/storage/sqlProvider.js):
`listArchivedShows` (:230-234) selects `data` from show_archive and maps every row through `_mapArchiveRow`
(:892-926 — JSON.parse, null unless an object) before `.filter(Boolean)`.  Here the rows' texts go to the GPU as they
are and come back as the columnar archive table the analytics and export operators read; no document is ever
materialised as an object.  There is no CPU fallback."""
from __future__ import annotations

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


def _row_timestamp(v) -> float:
    """_getTimestamp(value) (sqlProvider.js:970-985) for a row's column: the number, or NaN for null (the JS function's
    `null`).  A column the SELECT did not fetch is undefined -> NaN; SQL NULL is JS null -> Number(null) = 0; a text
    goes through Number(text) (StringToNumber).  A text that is not numeric would go to Date.parse, which this
    package does not restate (V8's legacy parser): it raises instead of guessing."""
    if v is _MISSING:
        return math.nan
    if v is None:
        return 0.0
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, (int, float)):
        f = float(v)
        return f if math.isfinite(f) else math.nan
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
        raise UnsupportedDateError(-4, f"row timestamp {v!r} is not numeric: Date.parse is not provided")
    return math.nan


def _text(row: Row):
    if isinstance(row, dict):  # a result row of `SELECT data, ... FROM show_archive` (sqlProvider.js:232)
        row = row.get("data")
    if row is None:
        return b"null"  # JSON.parse(null) is null: the row is dropped
    return row


def mapArchiveRows(rows: Iterable[Row], device="cuda") -> Tuple[ArchiveTable, torch.Tensor]:
    """rows.map(row => this._mapArchiveRow(row)) for a batch: (table, dropped) where dropped[i] is True for the rows
    the reference maps to null (text that is not JSON, or not an object); their table rows are empty shows.  Rows given
    as dicts may carry the `archived_at` / `created_at` columns of the SELECT (texts of epoch milliseconds, or None).
    `device="cuda"` keeps the table in HBM for the operators that follow; "cpu" returns host tensors through the
    host-buffer entry point."""
    rows = list(rows)
    docs = ops.JsonDocs.from_texts([_text(r) for r in rows])
    if str(device) != "cpu":
        docs = docs.to(device)
    table, status = ops.ingest_json(docs)
    dropped = status.bool()
    # the row's own timestamp columns (sqlProvider.js:905-918): archived_at wins over the document's archivedAt; the
    # document's createdAt wins over created_at.  (The document side is what the table holds: a finite JS number, or
    # absent — _getTimestamp's coercions of a null / string field of the DOCUMENT are not modelled; the provider
    # always stores numbers there, sqlProvider.js:361-409.)
    if any(isinstance(r, dict) and ("archived_at" in r or "created_at" in r) for r in rows):
        arch = [_row_timestamp(r.get("archived_at", _MISSING)) if isinstance(r, dict) else math.nan for r in rows]
        crea = [_row_timestamp(r.get("created_at", _MISSING)) if isinstance(r, dict) else math.nan for r in rows]
        dev = table.created_at.device
        arch = torch.tensor(arch, dtype=torch.float64).to(dev)
        crea = torch.tensor(crea, dtype=torch.float64).to(dev)
        keep = ~dropped.to(dev)
        table.archived_at = torch.where(keep & torch.isfinite(arch), arch, table.archived_at)
        table.created_at = torch.where(keep & ~torch.isfinite(table.created_at), crea, table.created_at)
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
        entry_ts=table.entry_ts)
