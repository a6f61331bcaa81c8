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
