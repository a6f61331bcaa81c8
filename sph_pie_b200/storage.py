"""Host-side mirror of the provider's read path for archived shows (reference server/storage/sqlProvider.js):
`listArchivedShows` (:230-234) selects `data` from show_archive and maps every row through `_mapArchiveRow`
(:892-926 — JSON.parse, null unless an object) before `.filter(Boolean)`.  Here the rows' texts go to the GPU as they
are and come back as the columnar archive table the analytics and export operators read; no document is ever
materialised as an object.  There is no CPU fallback."""
from __future__ import annotations

from typing import Iterable, Tuple, Union

import torch

from . import ops
from .columnar import ArchiveTable

Row = Union[str, bytes, dict]


def _text(row: Row):
    if isinstance(row, dict):  # a result row of `SELECT data, ... FROM show_archive` (sqlProvider.js:232)
        row = row.get("data")
    if row is None:
        return b"null"  # JSON.parse(null) is null: the row is dropped
    return row


def mapArchiveRows(rows: Iterable[Row], device="cuda") -> Tuple[ArchiveTable, torch.Tensor]:
    """rows.map(row => this._mapArchiveRow(row)) for a batch: (table, dropped) where dropped[i] is True for the rows
    the reference maps to null (text that is not JSON, or not an object); their table rows are empty shows.
    `device="cuda"` keeps the table in HBM for the operators that follow; "cpu" returns host tensors through the
    host-buffer entry point."""
    docs = ops.JsonDocs.from_texts([_text(r) for r in rows])
    if str(device) != "cpu":
        docs = docs.to(device)
    table, status = ops.ingest_json(docs)
    return table, status.bool()


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
