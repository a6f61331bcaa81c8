"""The columnar "archive table"  and the packer from provider-normalised show
documents (reference server/storage/sqlProvider.js:361-409 `_normalizeShow` / `_normalizeEntry`).

Columns are torch tensors (CPU, pinned CPU, or CUDA); torch is used for memory only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib

# show document key -> column
SHOW_KEY_TO_COL = {"id": "show_id", "date": "show_date", "time": "show_time", "label": "show_label",
                   "leadPilot": "lead_pilot", "monkeyLead": "monkey_lead", "notes": "show_notes"}
ENTRY_KEY_TO_COL = {"id": "entry_id", "unitId": "unit_id", "planned": "planned", "launched": "launched",
                    "status": "status", "primaryIssue": "primary_issue", "subIssue": "sub_issue",
                    "otherDetail": "other_detail", "severity": "severity", "rootCause": "root_cause",
                    "operator": "operator_name", "batteryId": "battery_id", "commandRx": "command_rx",
                    "notes": "notes"}


def _ptr(t: torch.Tensor) -> int:
    """data_ptr(), except that an empty VIEW of a real allocation keeps its address (torch reports 0 for every
    tensor without elements; the C ABI takes NULL for 'column absent')."""
    p = t.data_ptr()
    if p == 0 and t.numel() == 0:
        p = t.untyped_storage().data_ptr() + t.storage_offset() * t.element_size() if t.untyped_storage().nbytes() else 0
    return p


@dataclass
class StrCol:
    offsets: torch.Tensor  # int32 [n + 1]
    data: torch.Tensor     # uint8 [bytes]

    def to(self, device, non_blocking=False):
        return StrCol(self.offsets.to(device, non_blocking=non_blocking), self.data.to(device, non_blocking=non_blocking))

    def pin(self):
        return StrCol(self.offsets.pin_memory(), self.data.pin_memory())

    def nbytes(self) -> int:
        return self.offsets.numel() * 4 + self.data.numel()

    def get(self, i: int) -> str:
        o = self.offsets
        return bytes(self.data[int(o[i]):int(o[i + 1])].cpu().numpy()).decode("utf-8")

    def c(self) -> _lib.StrColC:
        return _lib.StrColC(_ptr(self.offsets), _ptr(self.data))


@dataclass
class StrListCol:
    list_offsets: torch.Tensor  # int32 [n + 1]
    items: StrCol

    def to(self, device, non_blocking=False):
        return StrListCol(self.list_offsets.to(device, non_blocking=non_blocking), self.items.to(device, non_blocking))

    def pin(self):
        return StrListCol(self.list_offsets.pin_memory(), self.items.pin())

    def nbytes(self) -> int:
        return self.list_offsets.numel() * 4 + self.items.nbytes()

    def c(self) -> _lib.StrListColC:
        return _lib.StrListColC(_ptr(self.list_offsets), self.items.c())


def strcol_from_strings(values: List[str]) -> StrCol:
    enc = []
    for v in values:
        try:
            enc.append(v.encode("utf-8"))
        except UnicodeEncodeError as e:  # lone surrogate: a JS string that is not well-formed UTF-16
            raise TypeError(f"string {v!r} is not encodable as UTF-8 (lone surrogate)") from e
    lens = np.fromiter((len(b) for b in enc), dtype=np.int64, count=len(enc))
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    if offs[-1] >= 2 ** 31:
        raise ValueError("string column exceeds 2 GiB: split the batch")
    data = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if enc else np.zeros(0, dtype=np.uint8)
    return StrCol(torch.from_numpy(offs.astype(np.int32)), torch.from_numpy(data))


def _norm_str(v, what: str) -> str:
    """null / undefined / '' are one value for every function on the path (`x || ''`)."""
    if v is None:
        return ""
    if isinstance(v, str):
        return v
    raise TypeError(
        f"{what} is {type(v).__name__}, not a string: the archive table holds provider-normalised shows "
        "(sqlProvider.js:361-409)"
    )


def _norm_time(v) -> float:
    """Number.isFinite(v) ? v : NaN (no coercion: only JS numbers qualify)."""
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        return math.nan
    f = float(v)
    return f if math.isfinite(f) else math.nan


_ABSENT = object()


def _time_kind(v, what: str) -> int:
    """What a show's time field holds (_lib.TK_*).  A string is recorded as a string, but not its text: the table
    points into the stored documents for that, which only the JSON ingest can do (pie_get_timestamps_dev reports
    PIE_ERR_SCHEMA when asked to coerce a packed table's text timestamp)."""
    if v is _ABSENT:
        return _lib.TK_ABSENT
    if v is None:
        return _lib.TK_NULL
    if v is True:
        return _lib.TK_TRUE
    if v is False:
        return _lib.TK_FALSE
    if isinstance(v, (int, float)):
        return _lib.TK_NUMBER if math.isfinite(float(v)) else _lib.TK_NONFINITE
    return _lib.TK_STRING if isinstance(v, str) else _lib.TK_OTHER


@dataclass
class ArchiveTable:
    n_shows: int
    n_entries: int
    entry_offsets: torch.Tensor
    show_cols: Dict[str, StrCol]
    crew: StrListCol
    created_at: torch.Tensor
    archived_at: torch.Tensor
    entry_cols: Dict[str, StrCol]
    actions: StrListCol
    delay_sec: torch.Tensor
    delay_valid: torch.Tensor
    entry_ts: torch.Tensor
    # ABI 2 (optional): show.updatedAt / show.deletedAt like created_at, and what each of the four time fields holds
    # when it is not a finite number: uint8 [n_shows, 4] of _lib.TK_* (created, updated, archived, deleted)
    updated_at: Optional[torch.Tensor] = None
    deleted_at: Optional[torch.Tensor] = None
    time_kind: Optional[torch.Tensor] = None
    _keep: list = field(default_factory=list, repr=False)

    @property
    def device(self) -> torch.device:
        return self.entry_offsets.device

    @property
    def is_cuda(self) -> bool:
        return self.entry_offsets.is_cuda

    def _map(self, f_tensor, f_col):
        return ArchiveTable(
            self.n_shows, self.n_entries, f_tensor(self.entry_offsets),
            {k: f_col(c) for k, c in self.show_cols.items()}, f_col(self.crew),
            f_tensor(self.created_at), f_tensor(self.archived_at),
            {k: f_col(c) for k, c in self.entry_cols.items()}, f_col(self.actions),
            f_tensor(self.delay_sec), f_tensor(self.delay_valid), f_tensor(self.entry_ts),
            *(None if t is None else f_tensor(t) for t in (self.updated_at, self.deleted_at, self.time_kind)))

    def to(self, device, non_blocking=False) -> "ArchiveTable":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking), lambda c: c.to(device, non_blocking))

    def pin(self) -> "ArchiveTable":
        return self._map(lambda t: t.pin_memory(), lambda c: c.pin())

    def nbytes(self) -> int:
        n = self.entry_offsets.numel() * 4 + (self.created_at.numel() + self.archived_at.numel()) * 8
        n += self.delay_sec.numel() * 8 + self.delay_valid.numel() + self.entry_ts.numel() * 8
        n += sum(t.numel() * t.element_size() for t in (self.updated_at, self.deleted_at, self.time_kind) if t is not None)
        n += sum(c.nbytes() for c in self.show_cols.values()) + sum(c.nbytes() for c in self.entry_cols.values())
        return n + self.crew.nbytes() + self.actions.nbytes()

    def view(self) -> _lib.ArchiveViewC:
        """pie_archive_view over this table's buffers (valid while the table is alive)."""
        v = _lib.ArchiveViewC()
        v.n_shows = self.n_shows
        v.n_entries = self.n_entries
        v.entry_offsets = _ptr(self.entry_offsets)
        for name in _lib.SHOW_STR_COLS:
            setattr(v, name, self.show_cols[name].c())
        v.crew = self.crew.c()
        v.created_at = _ptr(self.created_at)
        v.archived_at = _ptr(self.archived_at)
        for name in _lib.ENTRY_STR_COLS:
            setattr(v, name, self.entry_cols[name].c())
        v.actions = self.actions.c()
        v.delay_sec = _ptr(self.delay_sec)
        v.delay_valid = _ptr(self.delay_valid)
        v.entry_ts = _ptr(self.entry_ts)
        v.updated_at = _ptr(self.updated_at) if self.updated_at is not None else None
        v.deleted_at = _ptr(self.deleted_at) if self.deleted_at is not None else None
        v.time_kind = _ptr(self.time_kind) if self.time_kind is not None else None
        return v

    def slice_shows(self, s0: int, s1: int) -> "ArchiveTable":
        """Rows [s0, s1) of the show columns and their entries, sharing the byte heaps (string
        offsets keep their absolute values, which the ABI allows).  Used to shard by show range."""
        eo = self.entry_offsets
        e0, e1 = int(eo[s0]), int(eo[s1])

        def sl_col(c: StrCol, a, b):
            return StrCol(c.offsets[a:b + 1], c.data)

        def sl_list(c: StrListCol, a, b):
            return StrListCol(c.list_offsets[a:b + 1], c.items)

        return ArchiveTable(
            s1 - s0, e1 - e0, (eo[s0:s1 + 1] - e0).contiguous(),
            {k: sl_col(c, s0, s1) for k, c in self.show_cols.items()}, sl_list(self.crew, s0, s1),
            self.created_at[s0:s1], self.archived_at[s0:s1],
            {k: sl_col(c, e0, e1) for k, c in self.entry_cols.items()}, sl_list(self.actions, e0, e1),
            self.delay_sec[e0:e1], self.delay_valid[e0:e1], self.entry_ts[e0:e1],
            *(None if t is None else t[s0:s1] for t in (self.updated_at, self.deleted_at, self.time_kind)))


def _list_col(lists: List[List[str]]) -> StrListCol:
    lo = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists)), out=lo[1:])
    flat = [s for x in lists for s in x]
    return StrListCol(torch.from_numpy(lo.astype(np.int32)), strcol_from_strings(flat))


def pack_shows(shows: List[Optional[dict]]) -> ArchiveTable:
    """Pack provider-normalised show documents into an archive table.

    A falsy list element (null) becomes an empty show whose timestamp columns are NaN, so it is
    skipped by the daily grouping exactly like `if(!show) return` (public/app.js:3405-3407).
    """
    show_vals = {c: [] for c in SHOW_KEY_TO_COL.values()}
    entry_vals = {c: [] for c in ENTRY_KEY_TO_COL.values()}
    crew, actions = [], []
    created, archived, delay, delay_valid, ets = [], [], [], [], []
    updated, deleted, kinds = [], [], []
    entry_offsets = [0]
    for si, show in enumerate(shows):
        show = show if isinstance(show, dict) else {}
        for k, c in SHOW_KEY_TO_COL.items():
            show_vals[c].append(_norm_str(show.get(k), f"shows[{si}].{k}"))
        cr = show.get("crew")
        crew.append([_norm_str(x, f"shows[{si}].crew[]") for x in cr] if isinstance(cr, list) else [])
        created.append(_norm_time(show.get("createdAt")))
        archived.append(_norm_time(show.get("archivedAt")))
        updated.append(_norm_time(show.get("updatedAt")))
        deleted.append(_norm_time(show.get("deletedAt")))
        kinds.append([_time_kind(show.get(k, _ABSENT), f"shows[{si}].{k}")
                      for k in ("createdAt", "updatedAt", "archivedAt", "deletedAt")])
        entries = show.get("entries")
        entries = entries if isinstance(entries, list) else []
        for ei, e in enumerate(entries):
            e = e if isinstance(e, dict) else {}
            for k, c in ENTRY_KEY_TO_COL.items():
                entry_vals[c].append(_norm_str(e.get(k), f"shows[{si}].entries[{ei}].{k}"))
            ac = e.get("actions")
            actions.append([_norm_str(x, f"shows[{si}].entries[{ei}].actions[]") for x in ac]
                           if isinstance(ac, list) else [])
            d = e.get("delaySec")
            if d is None:
                delay.append(0.0)
                delay_valid.append(0)
            elif isinstance(d, bool) or not isinstance(d, (int, float)):
                raise TypeError(f"shows[{si}].entries[{ei}].delaySec is {type(d).__name__}; number or null expected")
            else:
                delay.append(float(d))
                delay_valid.append(1)
            ets.append(_norm_time(e.get("ts")))
        entry_offsets.append(entry_offsets[-1] + len(entries))
    return ArchiveTable(
        n_shows=len(shows), n_entries=entry_offsets[-1],
        entry_offsets=torch.tensor(entry_offsets, dtype=torch.int32),
        show_cols={c: strcol_from_strings(v) for c, v in show_vals.items()},
        crew=_list_col(crew),
        created_at=torch.tensor(created, dtype=torch.float64),
        archived_at=torch.tensor(archived, dtype=torch.float64),
        entry_cols={c: strcol_from_strings(v) for c, v in entry_vals.items()},
        actions=_list_col(actions),
        delay_sec=torch.tensor(delay, dtype=torch.float64),
        delay_valid=torch.tensor(delay_valid, dtype=torch.uint8),
        entry_ts=torch.tensor(ets, dtype=torch.float64),
        updated_at=torch.tensor(updated, dtype=torch.float64),
        deleted_at=torch.tensor(deleted, dtype=torch.float64),
        time_kind=torch.tensor(kinds, dtype=torch.uint8).reshape(len(shows), 4),
    )


def _concat_strcols(cols: List[StrCol], rows: List[int]) -> StrCol:
    """Rows [0, rows[i]) of every column, one after the other; offsets rebased onto one compact heap."""
    offs, datas, base = [], [], 0
    for c, n in zip(cols, rows):
        o = c.offsets[:n + 1].to(torch.int64)
        first, last = int(o[0]), int(o[n])
        offs.append(o[:n] - first + base)
        datas.append(c.data[first:last])
        base += last - first
    if base >= 2 ** 31:
        raise ValueError("string column exceeds 2 GiB: split the batch")
    dev = cols[0].offsets.device
    offs.append(torch.tensor([base], dtype=torch.int64, device=dev))
    return StrCol(torch.cat(offs).to(torch.int32), torch.cat(datas) if datas else torch.zeros(0, dtype=torch.uint8, device=dev))


def _concat_lists(cols: List[StrListCol], rows: List[int]) -> StrListCol:
    lo, items, item_rows, base = [], [], [], 0
    for c, n in zip(cols, rows):
        l = c.list_offsets[:n + 1].to(torch.int64)
        first, last = int(l[0]), int(l[n])
        lo.append(l[:n] - first + base)
        items.append(StrCol(c.items.offsets[first:last + 1], c.items.data))
        item_rows.append(last - first)
        base += last - first
    dev = cols[0].list_offsets.device
    lo.append(torch.tensor([base], dtype=torch.int64, device=dev))
    return StrListCol(torch.cat(lo).to(torch.int32), _concat_strcols(items, item_rows))


def concat_tables(parts: List[ArchiveTable]) -> ArchiveTable:
    """The shows of several tables (slices included), one after the other, as one compact table on the parts' device:
    what a rank whose day range spans more than one stored segment assembles (sharding.py)."""
    parts = [p for p in parts if p.n_shows > 0] or parts[:1]
    S = [p.n_shows for p in parts]
    E = [p.n_entries for p in parts]
    dev = parts[0].entry_offsets.device
    eo, base = [], 0
    for p in parts:
        o = p.entry_offsets[:p.n_shows + 1].to(torch.int64)
        eo.append(o[:p.n_shows] - int(o[0]) + base)
        base += p.n_entries
    eo.append(torch.tensor([base], dtype=torch.int64, device=dev))

    def opt(name):
        vals = [getattr(p, name) for p in parts]
        return None if any(v is None for v in vals) else torch.cat([v[:n] for v, n in zip(vals, S)])

    return ArchiveTable(
        n_shows=sum(S), n_entries=sum(E), entry_offsets=torch.cat(eo).to(torch.int32),
        show_cols={k: _concat_strcols([p.show_cols[k] for p in parts], S) for k in parts[0].show_cols},
        crew=_concat_lists([p.crew for p in parts], S),
        created_at=torch.cat([p.created_at[:n] for p, n in zip(parts, S)]),
        archived_at=torch.cat([p.archived_at[:n] for p, n in zip(parts, S)]),
        entry_cols={k: _concat_strcols([p.entry_cols[k] for p in parts], E) for k in parts[0].entry_cols},
        actions=_concat_lists([p.actions for p in parts], E),
        delay_sec=torch.cat([p.delay_sec[:n] for p, n in zip(parts, E)]),
        delay_valid=torch.cat([p.delay_valid[:n] for p, n in zip(parts, E)]),
        entry_ts=torch.cat([p.entry_ts[:n] for p, n in zip(parts, E)]),
        updated_at=opt("updated_at"), deleted_at=opt("deleted_at"), time_kind=opt("time_kind"))
