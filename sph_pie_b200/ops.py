"""Operators over an ArchiveTable: thin Python over the C ABI (include/sph_pie_b200.h).

A CUDA-resident table goes through the `*_dev` entry points on torch's current stream (outputs are
CUDA tensors); a host table goes through the `*_host` entry points, which copy host->device, run
the kernels and copy the results back (outputs are CPU tensors).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .columnar import ArchiveTable


@dataclass
class ShowStats:
    """Plane-major per-show statistics (computeArchiveShowStats, reference public/app.js:3939-3952)."""
    i32: torch.Tensor  # int32 [PIE_SI_COUNT, n_shows]
    f64: torch.Tensor  # float64 [PIE_SF_COUNT, n_shows]


@dataclass
class DailySummary:
    """Daily groups and metric summaries (reference public/app.js:3401-3502)."""
    n_groups: int
    show_day_start: torch.Tensor   # int64 [n_shows]
    show_order: torch.Tensor       # int32 [n_shows]
    group_day_start: torch.Tensor  # int64 [n_groups]
    group_offsets: torch.Tensor    # int32 [n_groups + 1]
    summary_f64: torch.Tensor      # float64 [3, 19, n_groups]   (average, min, max)
    summary_count: torch.Tensor    # int32 [19, n_groups]


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class DailyBuffers:
    """Pre-allocated outputs + scratch for the device entry points (reused across calls)."""

    def __init__(self, n_shows: int, n_entries: int, device):
        lib = _lib.load()
        S = max(n_shows, 1)
        self.S = S
        self.stats_i32 = torch.empty((_lib.PIE_SI_COUNT, S), dtype=torch.int32, device=device)
        self.stats_f64 = torch.empty((_lib.PIE_SF_COUNT, S), dtype=torch.float64, device=device)
        self.show_day_start = torch.empty(S, dtype=torch.int64, device=device)
        self.show_order = torch.empty(S, dtype=torch.int32, device=device)
        self.group_day_start = torch.empty(S, dtype=torch.int64, device=device)
        self.group_offsets = torch.empty(S + 1, dtype=torch.int32, device=device)
        self.summary_f64 = torch.empty((_lib.PIE_DF_COUNT, _lib.PIE_N_METRICS, S), dtype=torch.float64, device=device)
        self.summary_count = torch.empty((_lib.PIE_N_METRICS, S), dtype=torch.int32, device=device)
        self.n_groups = torch.zeros(1, dtype=torch.int64, device=device)
        self.status = torch.zeros(2, dtype=torch.int32, device=device)
        self.daily_scratch = torch.empty(int(lib.pie_daily_scratch_bytes(n_shows)), dtype=torch.uint8, device=device)

    def daily_out(self) -> _lib.DailyOutC:
        return _lib.DailyOutC(self.S, self.show_day_start.data_ptr(), self.show_order.data_ptr(),
                              self.group_day_start.data_ptr(), self.group_offsets.data_ptr(),
                              self.summary_f64.data_ptr(), self.summary_count.data_ptr(),
                              self.n_groups.data_ptr(), self.status.data_ptr())


def show_stats_dev(table: ArchiveTable, bufs: DailyBuffers) -> None:
    """Enqueue the show-statistics kernels on torch's current stream (no sync)."""
    _lib.ensure_init()
    view = table.view()
    _lib.check(_lib.load().pie_show_stats_dev(C.byref(view), bufs.stats_i32.data_ptr(), bufs.stats_f64.data_ptr(),
                                              bufs.S, _stream_ptr()))


def daily_summary_dev(table: ArchiveTable, bufs: DailyBuffers, tz_offset_minutes: int = 0) -> None:
    """Enqueue the daily-group kernels (reads bufs.stats_*) on torch's current stream (no sync)."""
    _lib.ensure_init()
    view = table.view()
    out = bufs.daily_out()
    _lib.check(_lib.load().pie_daily_summary_dev(C.byref(view), bufs.stats_i32.data_ptr(), bufs.stats_f64.data_ptr(),
                                                 bufs.S, tz_offset_minutes, C.byref(out),
                                                 bufs.daily_scratch.data_ptr(), _stream_ptr()))


def _raise_daily_status(code: int, show: int) -> None:
    if code == _lib.PIE_ERR_RANGE:
        raise _lib.JsRangeError(code, f"RangeError: Invalid time value (show {show})")
    if code == _lib.PIE_ERR_UNSUPPORTED_DATE:
        raise _lib.UnsupportedDateError(code, f"show {show}: date/time is not an ECMA-262 date-time string")
    if code != 0:
        raise _lib.PieError(code, f"daily summary failed at show {show}")


def show_stats(table: ArchiveTable) -> ShowStats:
    """computeArchiveShowStats for every show of the table."""
    _lib.ensure_init()
    S = table.n_shows
    if table.is_cuda:
        bufs = DailyBuffers(S, table.n_entries, table.device)
        show_stats_dev(table, bufs)
        return ShowStats(bufs.stats_i32[:, :S], bufs.stats_f64[:, :S])
    Sc = max(S, 1)
    i32 = torch.empty((_lib.PIE_SI_COUNT, Sc), dtype=torch.int32)
    f64 = torch.empty((_lib.PIE_SF_COUNT, Sc), dtype=torch.float64)
    view = table.view()
    _lib.check(_lib.load().pie_show_stats_host(C.byref(view), i32.data_ptr(), f64.data_ptr(), Sc))
    return ShowStats(i32[:, :S], f64[:, :S])


class HostOutputs:
    """Reusable (optionally pinned) host result buffers for archive_analytics on a host table."""

    def __init__(self, n_shows: int, pinned: bool = False):
        S = max(n_shows, 1)
        self.S = S
        kw = dict(pin_memory=pinned)
        self.stats_i32 = torch.empty((_lib.PIE_SI_COUNT, S), dtype=torch.int32, **kw)
        self.stats_f64 = torch.empty((_lib.PIE_SF_COUNT, S), dtype=torch.float64, **kw)
        self.show_day_start = torch.empty(S, dtype=torch.int64, **kw)
        self.show_order = torch.empty(S, dtype=torch.int32, **kw)
        self.group_day_start = torch.empty(S, dtype=torch.int64, **kw)
        self.group_offsets = torch.empty(S + 1, dtype=torch.int32, **kw)
        self.summary_f64 = torch.empty((_lib.PIE_DF_COUNT, _lib.PIE_N_METRICS, S), dtype=torch.float64, **kw)
        self.summary_count = torch.empty((_lib.PIE_N_METRICS, S), dtype=torch.int32, **kw)
        self.n_groups = torch.zeros(1, dtype=torch.int64, **kw)
        self.status = torch.zeros(2, dtype=torch.int32, **kw)

    def daily_out(self) -> _lib.DailyOutC:
        return _lib.DailyOutC(self.S, self.show_day_start.data_ptr(), self.show_order.data_ptr(),
                              self.group_day_start.data_ptr(), self.group_offsets.data_ptr(),
                              self.summary_f64.data_ptr(), self.summary_count.data_ptr(),
                              self.n_groups.data_ptr(), self.status.data_ptr())


def archive_analytics(table: ArchiveTable, tz_offset_minutes: int = 0, bufs=None):
    """Show statistics + daily groups + metric summaries: what buildArchiveDailyGroups followed by
    getOrCreateGroupMetricSummary for every metric computes.  Returns (ShowStats, DailySummary)."""
    _lib.ensure_init()
    S = table.n_shows
    if table.is_cuda:
        b = bufs if bufs is not None else DailyBuffers(S, table.n_entries, table.device)
        show_stats_dev(table, b)
        daily_summary_dev(table, b, tz_offset_minutes)
        code, show = (int(x) for x in b.status.cpu())  # syncs the stream
        _raise_daily_status(code, show)
        G = int(b.n_groups.cpu())
        return (ShowStats(b.stats_i32[:, :S], b.stats_f64[:, :S]),
                DailySummary(G, b.show_day_start[:S], b.show_order[:S], b.group_day_start[:G],
                             b.group_offsets[:G + 1], b.summary_f64[:, :, :G], b.summary_count[:, :G]))
    h = bufs if bufs is not None else HostOutputs(S)
    view = table.view()
    out = h.daily_out()
    _lib.check(_lib.load().pie_archive_analytics_host(C.byref(view), tz_offset_minutes, h.stats_i32.data_ptr(),
                                                      h.stats_f64.data_ptr(), h.S, C.byref(out)))
    G = int(h.n_groups[0])
    return (ShowStats(h.stats_i32[:, :S], h.stats_f64[:, :S]),
            DailySummary(G, h.show_day_start[:S], h.show_order[:S], h.group_day_start[:G], h.group_offsets[:G + 1],
                         h.summary_f64[:, :, :G], h.summary_count[:, :G]))


@dataclass
class CsvRows:
    """Export rows as one string column (buildCsvRow per entry, reference webhookDispatcher.js:340-342):
    row i = data[row_offsets[i] : row_offsets[i+1] - 1]; every row is followed by '\\n'."""
    row_offsets: torch.Tensor  # int64 [n_entries + 1]
    data: torch.Tensor         # uint8 [total_bytes]

    def row(self, i: int) -> str:
        o = self.row_offsets
        return bytes(self.data[int(o[i]):int(o[i + 1]) - 1].cpu().numpy()).decode("utf-8")

    def rows(self):
        o = self.row_offsets.cpu().tolist()
        b = bytes(self.data.cpu().numpy())
        return [b[o[i]:o[i + 1] - 1].decode("utf-8") for i in range(len(o) - 1)]


class CsvBuffers:
    """Pre-allocated device outputs + scratch for pie_csv_rows_dev."""

    def __init__(self, n_entries: int, capacity_bytes: int, device):
        lib = _lib.load()
        self.row_offsets = torch.empty(n_entries + 1, dtype=torch.int64, device=device)
        self.data = torch.empty(max(capacity_bytes, 1), dtype=torch.uint8, device=device)
        self.capacity = capacity_bytes
        self.total = torch.zeros(1, dtype=torch.int64, device=device)
        self.scratch = torch.empty(int(lib.pie_csv_rows_scratch_bytes(n_entries)), dtype=torch.uint8, device=device)


def _rows_dev(entry: str, table: ArchiveTable, bufs: CsvBuffers, size_only: bool) -> None:
    _lib.ensure_init()
    view = table.view()
    fn = getattr(_lib.load(), entry)
    _lib.check(fn(C.byref(view), bufs.row_offsets.data_ptr(), None if size_only else bufs.data.data_ptr(),
                  bufs.capacity, bufs.total.data_ptr(), bufs.scratch.data_ptr(), _stream_ptr()))


def csv_rows_dev(table: ArchiveTable, bufs: CsvBuffers, size_only: bool = False) -> None:
    """Enqueue the export-row kernels on torch's current stream (no sync)."""
    _rows_dev("pie_csv_rows_dev", table, bufs, size_only)


def archive_payloads_dev(table: ArchiveTable, bufs: CsvBuffers, size_only: bool = False) -> None:
    """Enqueue the archive-entry-payload kernels (JSON Lines) on torch's current stream (no sync)."""
    _rows_dev("pie_archive_payloads_dev", table, bufs, size_only)


def csv_slow_tiles(table: ArchiveTable, bufs: CsvBuffers) -> int:
    """Tiles (32..160 rows) of the last csv_rows_dev launch on `bufs` that did not fit the kernel's
    shared-memory staging and took the warp-per-row path (test / tuning hook)."""
    n = C.c_uint32(0)
    _lib.check(_lib.load().pie_debug_csv_slow_tiles(bufs.scratch.data_ptr(), table.n_entries, C.byref(n), _stream_ptr()))
    return int(n.value)


def _rows(table: ArchiveTable, dev_fn, host_entry: str) -> CsvRows:
    _lib.ensure_init()
    E = table.n_entries
    if table.is_cuda:
        sizing = CsvBuffers(E, 0, table.device)
        dev_fn(table, sizing, size_only=True)
        total = int(sizing.total.cpu())
        bufs = CsvBuffers(E, total, table.device)
        dev_fn(table, bufs)
        return CsvRows(bufs.row_offsets, bufs.data[:total])
    view = table.view()
    offsets = torch.empty(E + 1, dtype=torch.int64)
    total = C.c_uint64(0)
    fn = getattr(_lib.load(), host_entry)
    _lib.check(fn(C.byref(view), offsets.data_ptr(), None, 0, C.byref(total)))
    data = torch.empty(max(int(total.value), 1), dtype=torch.uint8)
    _lib.check(fn(C.byref(view), offsets.data_ptr(), data.data_ptr(), int(total.value), C.byref(total)))
    return CsvRows(offsets, data[:int(total.value)])


def csv_rows(table: ArchiveTable) -> CsvRows:
    """buildCsvRow(buildTableRow(show, entry)) for every entry of every show."""
    return _rows(table, csv_rows_dev, "pie_csv_rows_host")


def archive_payloads(table: ArchiveTable) -> CsvRows:
    """JSON.stringify(buildArchiveEntryPayload(show, entry)) for every entry of every show (one row each,
    reference webhookDispatcher.js:315-330 and :527-540)."""
    return _rows(table, archive_payloads_dev, "pie_archive_payloads_host")


@dataclass
class LiveMetrics:
    """computeMetrics(show) for every show (reference public/app.js:5024-5047): int32 planes
    [PIE_CM_COUNT][S] (success rate, status counts, topIssues as entry rows, length of avgDelay) and the
    avgDelay text, 32 bytes per show."""
    i32: torch.Tensor   # [PIE_CM_COUNT, S]
    text: torch.Tensor  # uint8 [S, PIE_CM_TEXT]

    def avg_delay(self, s: int) -> str:
        n = int(self.i32[_lib.CM_AVG_LEN][s])
        return bytes(self.text[s, :n].cpu().numpy()).decode("ascii")


def compute_metrics_dev(table: ArchiveTable, i32: torch.Tensor, text: torch.Tensor) -> None:
    """Enqueue the live-metrics kernel on torch's current stream (no sync)."""
    _lib.ensure_init()
    view = table.view()
    _lib.check(_lib.load().pie_compute_metrics_dev(C.byref(view), i32.data_ptr(), text.data_ptr(), i32.shape[1],
                                                   _stream_ptr()))


def compute_metrics(table: ArchiveTable) -> LiveMetrics:
    _lib.ensure_init()
    S = table.n_shows
    Sc = max(S, 1)
    dev = table.device if table.is_cuda else "cpu"
    i32 = torch.empty((_lib.PIE_CM_COUNT, Sc), dtype=torch.int32, device=dev)
    text = torch.empty((Sc, _lib.PIE_CM_TEXT), dtype=torch.uint8, device=dev)
    if table.is_cuda:
        compute_metrics_dev(table, i32, text)
    else:
        view = table.view()
        _lib.check(_lib.load().pie_compute_metrics_host(C.byref(view), i32.data_ptr(), text.data_ptr(), Sc))
    return LiveMetrics(i32[:, :S], text[:S])


def archive_step(table: ArchiveTable, tz_offset_minutes: int = 0, out: "HostOutputs" = None,
                 row_offsets: torch.Tensor = None, data: torch.Tensor = None):
    """Host table in, everything the archive workspace shows out, through ONE pipelined upload: show statistics,
    daily groups + metric summaries and the CSV rows (pie_archive_step_host).  With `row_offsets` / `data`
    preallocated (pinned, data.numel() = the CSV size) one call does the whole job; otherwise the CSV size is
    queried first.  Returns (ShowStats, DailySummary, CsvRows)."""
    _lib.ensure_init()
    assert not table.is_cuda
    lib = _lib.load()
    S, E = table.n_shows, table.n_entries
    h = out if out is not None else HostOutputs(S)
    view = table.view()
    total = C.c_uint64(0)
    if row_offsets is None:
        row_offsets = torch.empty(E + 1, dtype=torch.int64)
    if data is None:
        _lib.check(lib.pie_csv_rows_host(C.byref(view), row_offsets.data_ptr(), None, 0, C.byref(total)))
        data = torch.empty(max(int(total.value), 1), dtype=torch.uint8)
    dout = h.daily_out()
    _lib.check(lib.pie_archive_step_host(C.byref(view), tz_offset_minutes, h.stats_i32.data_ptr(), h.stats_f64.data_ptr(),
                                         h.S, C.byref(dout), row_offsets.data_ptr(), data.data_ptr(), data.numel(),
                                         C.byref(total)))
    G = int(h.n_groups[0])
    return (ShowStats(h.stats_i32[:, :S], h.stats_f64[:, :S]),
            DailySummary(G, h.show_day_start[:S], h.show_order[:S], h.group_day_start[:G], h.group_offsets[:G + 1],
                         h.summary_f64[:, :, :G], h.summary_count[:, :G]),
            CsvRows(row_offsets, data[:int(total.value)]))


# ---- JSON ingest: stored show documents -> archive table (pie_ingest_*) ---------------------------------------
@dataclass
class JsonDocs:
    """The `data` texts of a batch of rows (show_archive.data, sqlProvider.js:696): document s is
    data[offsets[s] : offsets[s+1]] (UTF-8 JSON).  Tensors on the CPU (pinned or not) or on the GPU."""
    offsets: torch.Tensor  # int64 [n + 1]
    data: torch.Tensor     # uint8

    @property
    def n_docs(self) -> int:
        return self.offsets.numel() - 1

    @property
    def is_cuda(self) -> bool:
        return self.data.is_cuda

    def to(self, device, non_blocking=False) -> "JsonDocs":
        return JsonDocs(self.offsets.to(device, non_blocking=non_blocking), self.data.to(device, non_blocking=non_blocking))

    def pin(self) -> "JsonDocs":
        return JsonDocs(self.offsets.pin_memory(), self.data.pin_memory())

    def c(self) -> _lib.JsonDocsC:
        return _lib.JsonDocsC(self.n_docs, self.offsets.data_ptr(), self.data.data_ptr())

    @staticmethod
    def from_texts(texts) -> "JsonDocs":
        import numpy as np

        enc = [t.encode("utf-8") if isinstance(t, str) else bytes(t) for t in texts]
        offs = np.zeros(len(enc) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in enc], out=offs[1:])
        data = np.zeros(int(offs[-1]) + 8, dtype=np.uint8)  # +8: never an empty tensor
        data[:offs[-1]] = np.frombuffer(b"".join(enc), dtype=np.uint8)
        return JsonDocs(torch.from_numpy(offs), torch.from_numpy(data))


def alloc_ingest_table(n_docs: int, totals, device) -> ArchiveTable:
    """An archive table with exactly the room pie_ingest_measure_* counted (totals[PIE_INGEST_TOTALS])."""
    from .columnar import StrCol, StrListCol

    E, crew_items, action_items = (int(totals[_lib.PIE_IT_ENTRIES]), int(totals[_lib.PIE_IT_CREW_ITEMS]),
                                   int(totals[_lib.PIE_IT_ACTION_ITEMS]))

    def col(rows, heap):
        return StrCol(torch.empty(rows + 1, dtype=torch.int32, device=device),
                      torch.empty(int(totals[heap]) + 8, dtype=torch.uint8, device=device))

    # one spare element: a column of no rows still has an address (the ABI rejects NULL columns)
    f64 = lambda n: torch.empty(n + 1, dtype=torch.float64, device=device)[:n]
    i32 = lambda n: torch.empty(n, dtype=torch.int32, device=device)
    return ArchiveTable(
        n_shows=n_docs, n_entries=E, entry_offsets=i32(n_docs + 1),
        show_cols={name: col(n_docs, h) for h, name in enumerate(_lib.SHOW_STR_COLS)},
        crew=StrListCol(i32(n_docs + 1), col(crew_items, 7)),
        created_at=f64(n_docs), archived_at=f64(n_docs),
        entry_cols={name: col(E, 8 + h) for h, name in enumerate(_lib.ENTRY_STR_COLS)},
        actions=StrListCol(i32(E + 1), col(action_items, 22)),
        delay_sec=f64(E), delay_valid=torch.empty(E + 1, dtype=torch.uint8, device=device)[:E], entry_ts=f64(E),
        updated_at=f64(n_docs), deleted_at=f64(n_docs),
        time_kind=torch.zeros((n_docs + 1, _lib.PIE_TF_COUNT), dtype=torch.uint8, device=device)[:n_docs])


def _raise_ingest_status(code: int, doc: int) -> None:
    if code == 0:
        return
    what = {_lib.PIE_ERR_SCHEMA: "is not a provider-normalised show (a text field that is not a string, delaySec that "
                                 "is not a number, or a lone surrogate)",
            _lib.PIE_ERR_UNSUPPORTED_JSON: "is JSON the ingest path does not decide (duplicate known key, nesting "
                                           "deeper than 64, not UTF-8, or a number on a rounding boundary)",
            _lib.PIE_ERR_CAPACITY: "a string heap or row count reaches 2 GiB: split the batch"}.get(code, "failed")
    cls = {_lib.PIE_ERR_SCHEMA: _lib.SchemaError, _lib.PIE_ERR_UNSUPPORTED_JSON: _lib.UnsupportedJsonError}.get(
        code, _lib.PieError)
    err = cls(code, (f"document {doc} " if doc >= 0 else "") + what)
    err.doc = doc
    raise err


class IngestBuffers:
    """Scratch and small outputs of the two ingest calls for up to n_docs documents (reused across calls)."""

    def __init__(self, n_docs: int, device):
        lib = _lib.load()
        self.scratch = torch.empty(int(lib.pie_ingest_scratch_bytes(n_docs)), dtype=torch.uint8, device=device)
        self.doc_status = torch.empty(max(n_docs, 1), dtype=torch.uint8, device=device)
        self.totals = torch.zeros(_lib.PIE_INGEST_TOTALS, dtype=torch.int64, device=device)
        self.status = torch.zeros(2, dtype=torch.int32, device=device)
        self.fill_scratch = None  # the entry rows of the second walk: sized when the number of entries is known


def ingest_measure_dev(docs: JsonDocs, bufs: IngestBuffers) -> None:
    """First walk + scans on torch's current stream (no sync): bufs.totals / bufs.status / bufs.doc_status."""
    _lib.ensure_init()
    d = docs.c()
    _lib.check(_lib.load().pie_ingest_measure_dev(C.byref(d), bufs.scratch.data_ptr(), bufs.doc_status.data_ptr(),
                                                  bufs.totals.data_ptr(), bufs.status.data_ptr(), _stream_ptr()))


def ingest_fill_dev(docs: JsonDocs, bufs: IngestBuffers, table: ArchiveTable) -> None:
    """Second walk on torch's current stream (no sync) into a table allocated from bufs.totals."""
    d = docs.c()
    view = table.view()
    lib = _lib.load()
    need = int(lib.pie_ingest_fill_scratch_bytes(table.n_entries))
    if bufs.fill_scratch is None or bufs.fill_scratch.numel() < need:
        bufs.fill_scratch = torch.empty(need, dtype=torch.uint8, device=docs.data.device)
    _lib.check(lib.pie_ingest_fill_dev(C.byref(d), bufs.scratch.data_ptr(), bufs.doc_status.data_ptr(), C.byref(view),
                                       bufs.fill_scratch.data_ptr(), _stream_ptr()))


def set_ingest_warp_path(on: int) -> int:
    """Debug knob: the warp-per-document path of the ingest on (1) / off (0), < 0 only queries; returns the previous
    value.  Results are identical either way (tests/test_gpu_ingest.py); off = the thread-per-document walk alone."""
    return int(_lib.load().pie_debug_ingest_warp_path(int(on)))


def ingest_declined(bufs: IngestBuffers, n_docs: int) -> int:
    """How many documents of the last ingest_measure_dev on `bufs` the warp path declined (synchronises)."""
    n = C.c_uint32(0)
    _lib.check(_lib.load().pie_debug_ingest_declined(bufs.scratch.data_ptr(), n_docs, C.byref(n), _stream_ptr()))
    return int(n.value)


def ingest_json(docs: JsonDocs, bufs: IngestBuffers = None):
    """Stored show documents -> (ArchiveTable, doc_status uint8[n]): what
    `rows.map(row => this._mapArchiveRow(row))` (sqlProvider.js:230-234, :892-926) parses, laid out as the table
    every other operator reads.  doc_status[s] = 1 marks a row the reference drops (`.filter(Boolean)`); its
    table row is the empty show.  CUDA-resident docs give a CUDA-resident table (pie_ingest_measure_dev /
    pie_ingest_fill_dev); host docs go through pie_ingest_host and give a host table."""
    _lib.ensure_init()
    lib = _lib.load()
    n = docs.n_docs
    if docs.is_cuda:
        bufs = bufs or IngestBuffers(n, docs.data.device)
        ingest_measure_dev(docs, bufs)
        totals = bufs.totals.cpu().tolist()  # synchronises: the table is sized from it
        code, doc = bufs.status.cpu().tolist()
        _raise_ingest_status(code, doc)
        table = alloc_ingest_table(n, totals, docs.data.device)
        ingest_fill_dev(docs, bufs, table)
        return table, bufs.doc_status[:n]
    d = docs.c()
    view = _lib.ArchiveViewC()
    doc_status = torch.empty(max(n, 1), dtype=torch.uint8)
    totals = torch.zeros(_lib.PIE_INGEST_TOTALS, dtype=torch.int64)
    bad = C.c_int64(-1)
    rc = lib.pie_ingest_host(C.byref(d), C.byref(view), doc_status.data_ptr(), totals.data_ptr(), C.byref(bad))
    if rc in (_lib.PIE_ERR_SCHEMA, _lib.PIE_ERR_UNSUPPORTED_JSON, _lib.PIE_ERR_CAPACITY):
        _raise_ingest_status(rc, int(bad.value))
    _lib.check(rc)
    table = _table_from_library_view(view, n, totals.tolist())
    return table, doc_status[:n]


def _table_from_library_view(view: "_lib.ArchiveViewC", n_docs: int, totals) -> ArchiveTable:
    """Copies the library-owned pinned result of pie_ingest_host into tensors the caller owns."""
    import numpy as np
    from .columnar import StrCol, StrListCol

    def arr(ptr, n, dtype):
        if n == 0 or not ptr:
            return torch.zeros(0, dtype=dtype)
        size = n * torch.empty(0, dtype=dtype).element_size()
        buf = (C.c_uint8 * size).from_address(ptr)
        return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8).copy()).view(dtype)

    E, crew_items, action_items = (int(totals[_lib.PIE_IT_ENTRIES]), int(totals[_lib.PIE_IT_CREW_ITEMS]),
                                   int(totals[_lib.PIE_IT_ACTION_ITEMS]))

    def col(c, rows, heap):
        return StrCol(arr(c.offsets, rows + 1, torch.int32), arr(c.data, int(totals[heap]), torch.uint8))

    return ArchiveTable(
        n_shows=n_docs, n_entries=E, entry_offsets=arr(view.entry_offsets, n_docs + 1, torch.int32),
        show_cols={name: col(getattr(view, name), n_docs, h) for h, name in enumerate(_lib.SHOW_STR_COLS)},
        crew=StrListCol(arr(view.crew.list_offsets, n_docs + 1, torch.int32), col(view.crew.items, crew_items, 7)),
        created_at=arr(view.created_at, n_docs, torch.float64), archived_at=arr(view.archived_at, n_docs, torch.float64),
        entry_cols={name: col(getattr(view, name), E, 8 + h) for h, name in enumerate(_lib.ENTRY_STR_COLS)},
        actions=StrListCol(arr(view.actions.list_offsets, E + 1, torch.int32), col(view.actions.items, action_items, 22)),
        delay_sec=arr(view.delay_sec, E, torch.float64), delay_valid=arr(view.delay_valid, E, torch.uint8),
        entry_ts=arr(view.entry_ts, E, torch.float64),
        updated_at=arr(view.updated_at, n_docs, torch.float64), deleted_at=arr(view.deleted_at, n_docs, torch.float64),
        time_kind=arr(view.time_kind, n_docs * _lib.PIE_TF_COUNT, torch.uint8).reshape(n_docs, _lib.PIE_TF_COUNT))


def archive_step_from_json(docs: JsonDocs, tz_offset_minutes: int = 0, out: "HostOutputs" = None,
                           row_offsets: torch.Tensor = None, data: torch.Tensor = None, device="cuda"):
    """The archive workspace computed from the provider's stored texts: host documents in (pinned for speed), GPU
    ingest, show statistics + daily summaries + CSV rows on the device-resident table, host results out.  The table
    itself never leaves the GPU and no show object is ever built — the composition a binding of the reference would
    make of pie_ingest_*_dev, pie_show_stats_dev, pie_daily_summary_dev and pie_csv_rows_dev.
    Returns (ShowStats, DailySummary, CsvRows, dropped bool[n_docs]); `out` / `row_offsets` / `data` are reused when
    given (data.numel() must be at least the CSV size)."""
    _lib.ensure_init()
    assert not docs.is_cuda
    dev = torch.device(device)
    n = docs.n_docs
    d = docs.to(dev, non_blocking=True)
    ib = IngestBuffers(n, dev)
    ingest_measure_dev(d, ib)
    totals = ib.totals.cpu().tolist()
    code, doc = ib.status.cpu().tolist()
    _raise_ingest_status(code, doc)
    table = alloc_ingest_table(n, totals, dev)
    ingest_fill_dev(d, ib, table)
    S, E = n, table.n_entries
    db = DailyBuffers(S, E, dev)
    show_stats_dev(table, db)
    daily_summary_dev(table, db, tz_offset_minutes)
    sizing = CsvBuffers(E, 0, dev)
    csv_rows_dev(table, sizing, size_only=True)
    csv_total = int(sizing.total.cpu())  # synchronises
    cb = CsvBuffers(E, csv_total, dev)
    cb.scratch = sizing.scratch
    csv_rows_dev(table, cb)
    h = out if out is not None else HostOutputs(S, pinned=True)
    if row_offsets is None:
        row_offsets = torch.empty(E + 1, dtype=torch.int64, pin_memory=True)
    if data is None:
        data = torch.empty(max(csv_total, 1), dtype=torch.uint8, pin_memory=True)
    assert data.numel() >= csv_total and row_offsets.numel() >= E + 1
    Sc = db.S
    for name in ("stats_i32", "stats_f64", "summary_f64", "summary_count"):
        getattr(h, name)[..., :Sc].copy_(getattr(db, name), non_blocking=True)
    for name in ("show_day_start", "show_order", "group_day_start"):
        getattr(h, name)[:Sc].copy_(getattr(db, name), non_blocking=True)
    h.group_offsets[:Sc + 1].copy_(db.group_offsets, non_blocking=True)
    h.n_groups.copy_(db.n_groups, non_blocking=True)
    h.status.copy_(db.status, non_blocking=True)
    row_offsets[:E + 1].copy_(cb.row_offsets, non_blocking=True)
    data[:csv_total].copy_(cb.data[:csv_total], non_blocking=True)
    dropped = ib.doc_status[:n].bool().cpu()
    torch.cuda.synchronize(dev)
    _raise_daily_status(int(h.status[0]), int(h.status[1]))
    G = int(h.n_groups[0])
    stats = ShowStats(h.stats_i32[:, :S], h.stats_f64[:, :S])
    daily = DailySummary(G, h.show_day_start[:S], h.show_order[:S], h.group_day_start[:G], h.group_offsets[:G + 1],
                         h.summary_f64[:, :, :G], h.summary_count[:, :G])
    return stats, daily, CsvRows(row_offsets[:E + 1], data[:csv_total]), dropped


_PIPE_STREAMS = {}


def archive_step_from_json_pipelined(docs: JsonDocs, tz_offset_minutes: int = 0, out: "HostOutputs" = None,
                                     row_offsets: torch.Tensor = None, data: torch.Tensor = None, device="cuda",
                                     chunk_docs: int = 131072):
    """archive_step_from_json with the batch cut into chunks of documents that move through three streams: while
    chunk c is ingested and its statistics / CSV rows are computed, chunk c+1 uploads and the CSV of chunk c-1
    downloads (PCIe is full duplex).  Shows are independent, so a chunk is a complete little batch; only the daily
    grouping needs all of them, and it reads a few show-level columns that are kept and joined at the end.
    `row_offsets` (int64 [>= n_entries + 1]) and `data` (uint8 [>= CSV bytes]) must be given (pinned): their sizes
    come from a first unpipelined call.  Same results as archive_step_from_json.  Chunks must stay large: a lane
    needs ~3.5 ms for a 4 KB document, so a walk of fewer than ~10^5 documents leaves the GPU waiting on latency."""
    from .columnar import StrCol, StrListCol

    _lib.ensure_init()
    lib = _lib.load()
    assert not docs.is_cuda and row_offsets is not None and data is not None
    dev = torch.device(device)
    n = docs.n_docs
    h = out if out is not None else HostOutputs(n, pinned=True)
    # the same three streams on every call: torch's allocator caches blocks per stream
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _PIPE_STREAMS:
        _PIPE_STREAMS[key] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    s_up, s_run, s_down = _PIPE_STREAMS[key]
    torch.cuda.current_stream(dev).synchronize()
    host_off = docs.offsets
    bounds = list(range(0, n, chunk_docs)) + [n]
    chunks = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
    max_bytes = max((int(host_off[b]) - int(host_off[a]) for a, b in chunks), default=0)
    with torch.cuda.stream(s_up):
        offs_dev = host_off.to(dev, non_blocking=True)
        slots = [torch.empty(max_bytes + 64, dtype=torch.uint8, device=dev) for _ in range(2)]
    up_done = [torch.cuda.Event() for _ in chunks]
    slot_free = [torch.cuda.Event(), torch.cuda.Event()]
    run_done = [torch.cuda.Event() for _ in chunks]
    db = DailyBuffers(n, 0, dev)  # the statistics planes of the whole batch; chunks write their column ranges
    totals_host = torch.zeros(_lib.PIE_INGEST_TOTALS + 2, dtype=torch.int64, pin_memory=True)
    csv_total_host = torch.zeros(1, dtype=torch.int64, pin_memory=True)
    ibufs = [IngestBuffers(min(chunk_docs, max(n, 1)), dev) for _ in range(2)]
    keep, doc_status = [], torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    ev = torch.cuda.Event()

    def upload(c):
        a, b = chunks[c]
        first, last = int(host_off[a]), int(host_off[b])
        with torch.cuda.stream(s_up):
            if c >= 2:
                s_up.wait_event(slot_free[c % 2])  # the chunk that used this slot is through its second walk
            skew = first & 7
            slots[c % 2][skew:skew + last - first].copy_(docs.data[first:last], non_blocking=True)
            up_done[c].record(s_up)
        return _lib.JsonDocsC(b - a, offs_dev.data_ptr() + 8 * a, slots[c % 2].data_ptr() + skew - first)

    e_base, csv_base = 0, 0
    pending = upload(0) if chunks else None
    for c, (a, b) in enumerate(chunks):
        d = pending
        if c + 1 < len(chunks):
            pending = upload(c + 1)
        ib = ibufs[c % 2]
        with torch.cuda.stream(s_run):
            s_run.wait_event(up_done[c])
            _lib.check(lib.pie_ingest_measure_dev(C.byref(d), ib.scratch.data_ptr(), doc_status.data_ptr() + a,
                                                  ib.totals.data_ptr(), ib.status.data_ptr(), s_run.cuda_stream))
            totals_host[:_lib.PIE_INGEST_TOTALS].copy_(ib.totals, non_blocking=True)
            totals_host[_lib.PIE_INGEST_TOTALS:].copy_(ib.status, non_blocking=True)
            ev.record(s_run)
        ev.synchronize()
        totals = totals_host[:_lib.PIE_INGEST_TOTALS].tolist()
        code, doc = int(totals_host[-2]), int(totals_host[-1])
        if code:
            torch.cuda.synchronize(dev)
            _raise_ingest_status(code, doc + a if doc >= 0 else doc)
        with torch.cuda.stream(s_run):
            table = alloc_ingest_table(b - a, totals, dev)
            view = table.view()
            need = int(lib.pie_ingest_fill_scratch_bytes(table.n_entries))
            if ib.fill_scratch is None or ib.fill_scratch.numel() < need:
                ib.fill_scratch = torch.empty(need, dtype=torch.uint8, device=dev)
            _lib.check(lib.pie_ingest_fill_dev(C.byref(d), ib.scratch.data_ptr(), doc_status.data_ptr() + a, C.byref(view),
                                               ib.fill_scratch.data_ptr(), s_run.cuda_stream))
            slot_free[c % 2].record(s_run)
            E_c = table.n_entries
            _lib.check(lib.pie_show_stats_dev(C.byref(view), db.stats_i32.data_ptr() + 4 * a, db.stats_f64.data_ptr() + 8 * a,
                                              db.S, s_run.cuda_stream))
            cb = CsvBuffers(E_c, 0, dev)
            _lib.check(lib.pie_csv_rows_dev(C.byref(view), cb.row_offsets.data_ptr(), None, 0, cb.total.data_ptr(),
                                            cb.scratch.data_ptr(), s_run.cuda_stream))
            csv_total_host.copy_(cb.total, non_blocking=True)
            ev.record(s_run)
        ev.synchronize()
        csv_c = int(csv_total_host[0])
        assert csv_base + csv_c <= data.numel() and e_base + E_c + 1 <= row_offsets.numel(), "row_offsets / data too small"
        with torch.cuda.stream(s_run):
            csv_dev = torch.empty(max(csv_c, 1), dtype=torch.uint8, device=dev)
            _lib.check(lib.pie_csv_rows_dev(C.byref(view), cb.row_offsets.data_ptr(), csv_dev.data_ptr(), csv_c,
                                            cb.total.data_ptr(), cb.scratch.data_ptr(), s_run.cuda_stream))
            if csv_base:
                cb.row_offsets += csv_base
            run_done[c].record(s_run)
        with torch.cuda.stream(s_down):
            s_down.wait_event(run_done[c])
            data[csv_base:csv_base + csv_c].copy_(csv_dev[:csv_c], non_blocking=True)
            row_offsets[e_base:e_base + E_c + 1].copy_(cb.row_offsets, non_blocking=True)
        keep.append((table, (cb, csv_dev, totals), None, e_base))
        e_base += E_c
        csv_base += csv_c
    # the daily groups read show-level columns of every chunk: join them (offsets rebased) and run the summary once
    with torch.cuda.stream(s_run):
        def join_col(name):
            heap = _lib.SHOW_STR_COLS.index(name)
            cols = [StrCol(t.show_cols[name].offsets, t.show_cols[name].data[:int(k[2][heap])]) for t, k, _, _ in keep]
            sizes = [int(c.data.numel()) for c in cols]
            bases = [0]
            for z in sizes[:-1]:
                bases.append(bases[-1] + z)
            offs = [c.offsets[:t.n_shows] + bs for c, bs, (t, _, _, _) in zip(cols, bases, keep)]
            last = cols[-1].offsets[keep[-1][0].n_shows:keep[-1][0].n_shows + 1] + bases[-1]
            return StrCol(torch.cat(offs + [last]).to(torch.int32), torch.cat([c.data for c in cols]))

        if keep:
            eo = torch.cat([t.entry_offsets[:t.n_shows] + eb for t, _, _, eb in keep]
                           + [torch.tensor([e_base], dtype=torch.int32, device=dev)]).to(torch.int32)
            empty = StrCol(torch.zeros(n + 1, dtype=torch.int32, device=dev), torch.zeros(8, dtype=torch.uint8, device=dev))
            skinny = ArchiveTable(
                n_shows=n, n_entries=e_base, entry_offsets=eo,
                show_cols={k: (join_col(k) if k in ("show_date", "show_time") else empty) for k in _lib.SHOW_STR_COLS},
                crew=StrListCol(torch.zeros(n + 1, dtype=torch.int32, device=dev), empty),
                created_at=torch.cat([t.created_at for t, _, _, _ in keep]),
                archived_at=torch.cat([t.archived_at for t, _, _, _ in keep]),
                entry_cols={k: empty for k in _lib.ENTRY_STR_COLS}, actions=StrListCol(eo, empty),
                delay_sec=torch.zeros(1, dtype=torch.float64, device=dev), delay_valid=torch.zeros(1, dtype=torch.uint8, device=dev),
                entry_ts=torch.cat([t.entry_ts for t, _, _, _ in keep] + [torch.zeros(1, dtype=torch.float64, device=dev)]))
            daily_summary_dev(skinny, db, tz_offset_minutes)
        Sc = db.S
        for name in ("stats_i32", "stats_f64", "summary_f64", "summary_count"):
            getattr(h, name)[..., :Sc].copy_(getattr(db, name), non_blocking=True)
        for name in ("show_day_start", "show_order", "group_day_start"):
            getattr(h, name)[:Sc].copy_(getattr(db, name), non_blocking=True)
        h.group_offsets[:Sc + 1].copy_(db.group_offsets, non_blocking=True)
        h.n_groups.copy_(db.n_groups, non_blocking=True)
        h.status.copy_(db.status, non_blocking=True)
    dropped = doc_status[:n].bool().cpu()
    torch.cuda.synchronize(dev)
    if not keep:
        h.n_groups.zero_()
        h.status.zero_()
        row_offsets[0] = 0
    _raise_daily_status(int(h.status[0]), int(h.status[1]))
    G = int(h.n_groups[0])
    stats = ShowStats(h.stats_i32[:, :n], h.stats_f64[:, :n])
    daily = DailySummary(G, h.show_day_start[:n], h.show_order[:n], h.group_day_start[:G], h.group_offsets[:G + 1],
                         h.summary_f64[:, :, :G], h.summary_count[:, :G])
    return stats, daily, CsvRows(row_offsets[:e_base + 1], data[:csv_base]), dropped


def archive_step_json_host(docs: JsonDocs, tz_offset_minutes: int = 0, out: "HostOutputs" = None,
                           row_offsets: torch.Tensor = None, data: torch.Tensor = None):
    """pie_archive_step_json_host: the same as archive_step_from_json, as ONE call of the C ABI with host buffers in
    and out (what a binding of the reference would call).  Without `row_offsets` / `data` the sizes are queried with
    a first call."""
    _lib.ensure_init()
    assert not docs.is_cuda
    lib = _lib.load()
    n = docs.n_docs
    h = out if out is not None else HostOutputs(n)
    d = docs.c()
    dout = h.daily_out()
    status = torch.empty(max(n, 1), dtype=torch.uint8)
    n_entries, total, bad = C.c_int64(0), C.c_uint64(0), C.c_int64(-1)

    def call(off, dat):
        rc = lib.pie_archive_step_json_host(
            C.byref(d), tz_offset_minutes, status.data_ptr(), h.stats_i32.data_ptr(), h.stats_f64.data_ptr(), h.S,
            C.byref(dout), off.data_ptr() if off is not None else None, off.numel() if off is not None else 0,
            dat.data_ptr() if dat is not None else None, dat.numel() if dat is not None else 0,
            C.byref(n_entries), C.byref(total), C.byref(bad))
        if rc in (_lib.PIE_ERR_SCHEMA, _lib.PIE_ERR_UNSUPPORTED_JSON):
            _raise_ingest_status(rc, int(bad.value))
        _lib.check(rc)

    if row_offsets is None or data is None:
        call(None, None)
        row_offsets = torch.empty(n_entries.value + 1, dtype=torch.int64)
        data = torch.empty(max(total.value, 1), dtype=torch.uint8)
    call(row_offsets, data)
    E, G = n_entries.value, int(h.n_groups[0])
    stats = ShowStats(h.stats_i32[:, :n], h.stats_f64[:, :n])
    daily = DailySummary(G, h.show_day_start[:n], h.show_order[:n], h.group_day_start[:G], h.group_offsets[:G + 1],
                         h.summary_f64[:, :, :G], h.summary_count[:, :G])
    return stats, daily, CsvRows(row_offsets[:E + 1], data[:total.value]), status[:n].bool()


# ---- _getTimestamp of the documents' time fields; archive maintenance decisions (pie_get_timestamps_dev, pie_archive_*) ----
@dataclass
class DocTimes:
    """_getTimestamp(show.createdAt / updatedAt / archivedAt / deletedAt) (sqlProvider.js:970-985) per show: float64
    [n_shows] each, NaN where the reference's function returns null."""
    created_at: torch.Tensor
    updated_at: torch.Tensor
    archived_at: torch.Tensor
    deleted_at: torch.Tensor


def _raise_times_status(code: int, show: int) -> None:
    if code == 0:
        return
    if code == _lib.PIE_ERR_UNSUPPORTED_DATE:
        raise _lib.UnsupportedDateError(code, f"show {show}: a text timestamp that is neither numeric nor an ECMA-262 date-time string")
    if code == _lib.PIE_ERR_SCHEMA:
        raise _lib.SchemaError(code, f"show {show}: a time field holds an array / object, or a text and the documents were not given")
    raise _lib.PieError(code, f"timestamps failed at show {show}")


def get_timestamps(table: ArchiveTable, docs: "JsonDocs" = None, tz_offset_minutes: int = 0) -> DocTimes:
    """_getTimestamp for the four time fields of every show of a CUDA-resident table (`docs`: the CUDA-resident
    documents it was ingested from, needed when a field holds a string)."""
    _lib.ensure_init()
    assert table.is_cuda, "pie_get_timestamps_dev works on a device-resident table"
    dev, S = table.device, table.n_shows
    out = [torch.empty(max(S, 1), dtype=torch.float64, device=dev) for _ in range(4)]
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    view = table.view()
    d = docs.c() if docs is not None else None
    times = _lib.DocTimesC(*(t.data_ptr() for t in out))
    _lib.check(_lib.load().pie_get_timestamps_dev(C.byref(view), C.byref(d) if d is not None else None, tz_offset_minutes,
                                                  C.byref(times), status.data_ptr(), _stream_ptr()))
    code, show = status.cpu().tolist()
    _raise_times_status(code, show)
    return DocTimes(*(t[:S] for t in out))


def archive_due(table: ArchiveTable, created: torch.Tensor, now_ms: float, doc_status: torch.Tensor = None):
    """The decision of _archiveDailyShows (sqlProvider.js:758-816) for the rows of a CUDA-resident table: (due uint8[S],
    group_first int32[S]).  created[s] = _getTimestamp(createdAt) ?? _getTimestamp(updatedAt) (NaN = null)."""
    _lib.ensure_init()
    assert table.is_cuda
    lib, dev, S = _lib.load(), table.device, table.n_shows
    due = torch.zeros(max(S, 1), dtype=torch.uint8, device=dev)
    first = torch.full((max(S, 1),), -1, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.pie_archive_due_scratch_bytes(S)), dtype=torch.uint8, device=dev)
    view = table.view()
    created = created.to(dev, torch.float64).contiguous()
    _lib.check(lib.pie_archive_due_dev(C.byref(view), doc_status.data_ptr() if doc_status is not None else None,
                                       created.data_ptr() if S else None, float(now_ms), due.data_ptr(), first.data_ptr(),
                                       scratch.data_ptr(), _stream_ptr()))
    return due[:S], first[:S]


def archive_expired(created: torch.Tensor, now_ms: float, tz_offset_minutes: int = 0) -> torch.Tensor:
    """_purgeExpiredArchives' decision (sqlProvider.js:863-890, _addMonths :999-1009): expired uint8[n] for
    created[n] = _getTimestamp(show?.createdAt) ?? _getTimestamp(row.created_at) (NaN = null), on the GPU."""
    _lib.ensure_init()
    assert created.is_cuda
    created = created.to(torch.float64).contiguous()
    n = created.numel()
    out = torch.zeros(max(n, 1), dtype=torch.uint8, device=created.device)
    _lib.check(_lib.load().pie_archive_expired_dev(created.data_ptr() if n else None, n, float(now_ms), tz_offset_minutes,
                                                   out.data_ptr() if n else None, _stream_ptr()))
    return out[:n]


# ---- the schemaVersion 2 show payload (pie_show_payloads_dev) ---------------------------------------------------------
@dataclass
class ShowPayloads:
    """One JSON document per show: document s = data[doc_offsets[s] : doc_offsets[s+1]]."""
    doc_offsets: torch.Tensor  # int64 [n_shows + 1]
    data: torch.Tensor         # uint8

    def documents(self):
        o = self.doc_offsets.cpu().tolist()
        b = bytes(self.data.cpu().numpy())
        return [b[o[i]:o[i + 1]].decode("utf-8") for i in range(len(o) - 1)]


def show_payloads(table: ArchiveTable, head: bytes, tail: bytes) -> ShowPayloads:
    """JSON.stringify of dispatchShowEvent's schemaVersion 2 payload (webhookDispatcher.js:545-584) for every show of a
    CUDA-resident table; `head` / `tail` are the serialised texts around the per-show part (webhook.payload_frame)."""
    _lib.ensure_init()
    assert table.is_cuda
    lib, dev, S = _lib.load(), table.device, table.n_shows
    h = torch.tensor(list(head) or [0], dtype=torch.uint8, device=dev)
    t = torch.tensor(list(tail) or [0], dtype=torch.uint8, device=dev)
    offs = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.pie_show_payloads_scratch_bytes(S)) + 256, dtype=torch.uint8, device=dev)
    sptr = (scratch.data_ptr() + 255) & ~255
    view = table.view()

    def call(out, cap):
        _lib.check(lib.pie_show_payloads_dev(C.byref(view), h.data_ptr(), len(head), t.data_ptr(), len(tail), offs.data_ptr(),
                                             out, cap, total.data_ptr(), status.data_ptr(), sptr, _stream_ptr()))
        code, show = status.cpu().tolist()
        if code == _lib.PIE_ERR_SCHEMA:
            raise _lib.SchemaError(code, f"show {show}: a time field holds a text, an array or an object (the table does not hold the value)")
        if code:
            raise _lib.PieError(code, f"show payloads failed at show {show}")

    call(None, 0)
    n = int(total.cpu())
    data = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    call(data.data_ptr(), n)
    return ShowPayloads(offs, data[:n])
