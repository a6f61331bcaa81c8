"""ctypes loader for oracle/libpie_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by the product.
It reuses the product's ctypes struct definitions so both sides see one layout."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

from sph_pie_b200 import _lib
from sph_pie_b200.columnar import ArchiveTable
from sph_pie_b200.ops import DailySummary, HostOutputs, ShowStats

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpie_oracle.so")
_so = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pie_oracle.c")
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpie_oracle.so"], stdout=subprocess.DEVNULL)
    return SO_PATH


def load():
    global _so
    if _so is None:
        build()  # (re)builds when pie_oracle.c is newer
        so = C.CDLL(SO_PATH)
        so.oracle_show_stats.restype = C.c_int
        so.oracle_show_stats.argtypes = [C.POINTER(_lib.ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        so.oracle_daily_summary.restype = C.c_int
        so.oracle_daily_summary.argtypes = [C.POINTER(_lib.ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_int32, C.POINTER(_lib.DailyOutC)]
        so.oracle_max_threads.restype = C.c_int
        _so = so
    return _so


def max_threads() -> int:
    return int(load().oracle_max_threads())


def show_stats(table: ArchiveTable, nthreads: int = 1, out: HostOutputs = None) -> ShowStats:
    assert not table.is_cuda
    S = table.n_shows
    h = out if out is not None else HostOutputs(S)
    view = table.view()
    rc = load().oracle_show_stats(C.byref(view), h.stats_i32.data_ptr(), h.stats_f64.data_ptr(), h.S, nthreads)
    assert rc == 0
    return ShowStats(h.stats_i32[:, :S], h.stats_f64[:, :S])


def archive_analytics(table: ArchiveTable, tz_offset_minutes: int = 0, nthreads: int = 1, out: HostOutputs = None):
    """Returns (ShowStats, DailySummary, status_code, status_show)."""
    assert not table.is_cuda
    S = table.n_shows
    h = out if out is not None else HostOutputs(S)
    st = show_stats(table, nthreads, h)
    view = table.view()
    out = h.daily_out()
    rc = load().oracle_daily_summary(C.byref(view), h.stats_i32.data_ptr(), h.stats_f64.data_ptr(), h.S,
                                     tz_offset_minutes, C.byref(out))
    if rc != 0:
        return st, None, rc, int(h.status[1])
    G = int(h.n_groups[0])
    return st, DailySummary(G, h.show_day_start[:S], h.show_order[:S], h.group_day_start[:G], h.group_offsets[:G + 1],
                            h.summary_f64[:, :, :G], h.summary_count[:, :G]), 0, -1


def _rows(entry: str, table: ArchiveTable, nthreads: int, offsets, data):
    assert not table.is_cuda
    fn = getattr(load(), entry)
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(_lib.ArchiveViewC), C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
    view = table.view()
    total = C.c_uint64(0)
    if offsets is None:
        offsets = torch.empty(table.n_entries + 1, dtype=torch.int64)
    if data is None:
        fn(C.byref(view), offsets.data_ptr(), None, 0, C.byref(total), nthreads)
        data = torch.empty(max(int(total.value), 1), dtype=torch.uint8)
    fn(C.byref(view), offsets.data_ptr(), data.data_ptr(), data.numel(), C.byref(total), nthreads)
    return offsets, data[: int(total.value)]


def csv_rows(table: ArchiveTable, nthreads: int = 1, offsets=None, data=None):
    """(row_offsets int64[E+1], data uint8[total]) from the C restatement of buildCsvRow.
    With preallocated `offsets`/`data` (timed baseline) one call does the whole job."""
    return _rows("oracle_csv_rows_mt", table, nthreads, offsets, data)


def payload_rows(table: ArchiveTable, nthreads: int = 1, offsets=None, data=None):
    """Same for JSON.stringify(buildArchiveEntryPayload(show, entry)) + '\\n' per entry."""
    return _rows("oracle_payload_rows_mt", table, nthreads, offsets, data)


def compute_metrics(table: ArchiveTable):
    """(int32[PIE_CM_COUNT][S], uint8[S][32]) from the C restatement of computeMetrics."""
    assert not table.is_cuda
    so = load()
    so.oracle_compute_metrics.restype = C.c_int
    so.oracle_compute_metrics.argtypes = [C.POINTER(_lib.ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64]
    S = max(table.n_shows, 1)
    i32 = torch.zeros((_lib.PIE_CM_COUNT, S), dtype=torch.int32)
    text = torch.zeros((S, _lib.PIE_CM_TEXT), dtype=torch.uint8)
    view = table.view()
    so.oracle_compute_metrics(C.byref(view), i32.data_ptr(), text.data_ptr(), S)
    return i32[:, : table.n_shows], text[: table.n_shows]


def number_to_string_batch(xs):
    import numpy as np

    so = load()
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    n = len(xs)
    out = np.zeros(n * 32, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    so.oracle_number_to_string_batch.restype = None
    so.oracle_number_to_string_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    so.oracle_number_to_string_batch(xs.ctypes.data, n, out.ctypes.data, lens.ctypes.data)
    b = out.tobytes()
    return [b[i * 32:i * 32 + lens[i]].decode() for i in range(n)]


def ingest(texts, nthreads: int = 1):
    """The C restatement of JSON.parse + projection (oracle_ingest_*): (table | None, doc_status, (pie_status, doc)).
    `texts`: list of str / bytes, or an ops.JsonDocs on the host."""
    import numpy as np

    from sph_pie_b200 import ops

    docs = texts if isinstance(texts, ops.JsonDocs) else ops.JsonDocs.from_texts(texts)
    so = load()
    n = docs.n_docs
    rows = np.zeros((max(n, 1), _lib.PIE_INGEST_TOTALS), dtype=np.uint32)
    doc_status = np.zeros(max(n, 1), dtype=np.uint8)
    totals = np.zeros(_lib.PIE_INGEST_TOTALS, dtype=np.int64)
    status = np.zeros(2, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    so.oracle_ingest_measure(C.c_void_p(docs.data.data_ptr()), C.c_void_p(docs.offsets.data_ptr()), C.c_int64(n), p(rows),
                             p(doc_status), p(totals), p(status), C.c_int(nthreads))
    if status[0] != 0:
        return None, doc_status[:n], (int(status[0]), int(status[1]))
    table = ops.alloc_ingest_table(n, totals.tolist(), "cpu")
    view = table.view()
    so.oracle_ingest_fill(C.c_void_p(docs.data.data_ptr()), C.c_void_p(docs.offsets.data_ptr()), C.c_int64(n), p(rows),
                          p(doc_status), p(totals), C.byref(view), C.c_int(nthreads))
    return table, doc_status[:n], (0, -1)
