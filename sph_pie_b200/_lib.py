"""ctypes binding of libsphpie_b200.so (include/sph_pie_b200.h).

The library is the product: if it is missing or no sm_100 GPU is present, every compute call
raises.  There is no CPU fallback in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIE_LIB_PATH") or os.path.join(_HERE, "libsphpie_b200.so")  # env: tuning sweeps only

PIE_N_ISSUES = 10
PIE_N_METRICS = 19
PIE_SI_COUNT = 18
PIE_SF_COUNT = 15
PIE_DF_COUNT = 3
PIE_CM_COUNT = 8
PIE_CM_TEXT = 32
PIE_DAY_NONE = -(2 ** 63)

PIE_OK = 0
PIE_ERR_CUDA = -1
PIE_ERR_INVALID_ARG = -2
PIE_ERR_RANGE = -3
PIE_ERR_UNSUPPORTED_DATE = -4
PIE_ERR_CAPACITY = -5
PIE_ERR_NO_DEVICE = -6
PIE_ERR_SCHEMA = -7
PIE_ERR_UNSUPPORTED_JSON = -8

# JSON ingest: totals[0..22] = bytes of the string heaps in table order, then the row counts
PIE_INGEST_HEAPS = 23
PIE_IT_ENTRIES, PIE_IT_CREW_ITEMS, PIE_IT_ACTION_ITEMS = 23, 24, 25
PIE_INGEST_TOTALS = 26

# time fields of a show and what they can hold (pie_archive_view.time_kind)
TF_CREATED, TF_UPDATED, TF_ARCHIVED, TF_DELETED = range(4)
PIE_TF_COUNT = 4
(TK_ABSENT, TK_NUMBER, TK_NULL, TK_TRUE, TK_FALSE, TK_STRING, TK_OTHER, TK_NONFINITE) = range(8)

# plane indices (include/sph_pie_b200.h)
SI_TOTAL, SI_COMPLETED, SI_NO_LAUNCH, SI_ABORT, SI_LAUNCHED, SI_DELAY_COUNT = range(6)
SI_ISSUE_COUNT0 = 6
SI_ISSUE_ORDER_LO = 16
SI_ISSUE_ORDER_HI = 17
SF_AVG_DELAY, SF_MAX_DELAY, SF_COMPLETION_RATE, SF_LAUNCH_RATE, SF_ABORT_RATE = range(5)
SF_ISSUE_RATE0 = 5
DF_AVERAGE, DF_MIN, DF_MAX = range(3)
(CM_SUCCESS_RATE, CM_COMPLETED, CM_NO_LAUNCH, CM_ABORT, CM_TOP0, CM_TOP1, CM_TOP2, CM_AVG_LEN) = range(8)


class StrColC(C.Structure):
    _fields_ = [("offsets", C.c_void_p), ("data", C.c_void_p)]


class StrListColC(C.Structure):
    _fields_ = [("list_offsets", C.c_void_p), ("items", StrColC)]


SHOW_STR_COLS = ["show_id", "show_date", "show_time", "show_label", "lead_pilot", "monkey_lead", "show_notes"]
ENTRY_STR_COLS = ["entry_id", "unit_id", "planned", "launched", "status", "primary_issue", "sub_issue",
                  "other_detail", "severity", "root_cause", "operator_name", "battery_id", "command_rx", "notes"]


class ArchiveViewC(C.Structure):
    _fields_ = (
        [("n_shows", C.c_int64), ("n_entries", C.c_int64), ("entry_offsets", C.c_void_p)]
        + [(n, StrColC) for n in SHOW_STR_COLS]
        + [("crew", StrListColC), ("created_at", C.c_void_p), ("archived_at", C.c_void_p)]
        + [(n, StrColC) for n in ENTRY_STR_COLS]
        + [("actions", StrListColC), ("delay_sec", C.c_void_p), ("delay_valid", C.c_void_p),
           ("entry_ts", C.c_void_p)]
        # ABI 2: the document's other two time fields and what the four hold when it is not a finite number
        + [("updated_at", C.c_void_p), ("deleted_at", C.c_void_p), ("time_kind", C.c_void_p)]
    )


class DocTimesC(C.Structure):
    _fields_ = [("created_at", C.c_void_p), ("updated_at", C.c_void_p), ("archived_at", C.c_void_p),
                ("deleted_at", C.c_void_p)]


class JsonDocsC(C.Structure):
    _fields_ = [("n_docs", C.c_int64), ("offsets", C.c_void_p), ("data", C.c_void_p)]


class DailyOutC(C.Structure):
    _fields_ = [
        ("stride", C.c_int64),
        ("show_day_start", C.c_void_p),
        ("show_order", C.c_void_p),
        ("group_day_start", C.c_void_p),
        ("group_offsets", C.c_void_p),
        ("summary_f64", C.c_void_p),
        ("summary_count", C.c_void_p),
        ("n_groups", C.c_void_p),
        ("status", C.c_void_p),
    ]


class PieError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[pie {code}] {message}")
        self.code = code
        self.message = message


class JsRangeError(PieError, ValueError):
    """RangeError: Invalid time value — what the reference throws from toISOString
    (public/app.js:3415) for a show whose timestamp is outside the ECMAScript time range."""


class UnsupportedDateError(PieError, NotImplementedError):
    """show.date/time is outside the ECMA-262 date-time grammar (V8 legacy parsing not provided)."""


class SchemaError(PieError, TypeError):
    """A stored document is not a provider-normalised show (what pack_shows raises TypeError for)."""


class UnsupportedJsonError(PieError, ValueError):
    """Valid JSON the ingest kernels decline to decide (duplicate known key, depth > 64, not UTF-8, undecided number)."""


_lib = None

# every function include/sph_pie_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pie_abi_version": (C.c_int, []),
    "pie_last_error": (C.c_char_p, []),
    "pie_init": (C.c_int, [C.c_int]),
    "pie_device_sm_count": (C.c_int, []),
    "pie_host_alloc": (C.c_void_p, [C.c_uint64]),
    "pie_host_free": (None, [C.c_void_p]),
    "pie_last_transfer_bytes": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pie_kernel_launch_count": (C.c_uint64, []),
    "pie_release": (C.c_int, []),
    "pie_show_stats_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pie_show_stats_host": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64]),
    "pie_daily_scratch_bytes": (C.c_uint64, [C.c_int64]),
    "pie_daily_summary_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.POINTER(DailyOutC), C.c_void_p, C.c_void_p]),
    "pie_selftest_fast_div": (C.c_int, [C.c_int32, C.POINTER(C.c_uint64)]),
    "pie_csv_rows_scratch_bytes": (C.c_uint64, [C.c_int64]),
    "pie_csv_rows_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "pie_set_csv_chunk_rows": (C.c_int64, [C.c_int64]),
    "pie_set_json_chunk_docs": (C.c_int64, [C.c_int64]),
    "pie_archive_step_host": (C.c_int, [C.POINTER(ArchiveViewC), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.POINTER(DailyOutC), C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "pie_compute_metrics_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pie_compute_metrics_host": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_int64]),
    "pie_archive_payloads_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "pie_archive_payloads_host": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_uint64,
                                            C.POINTER(C.c_uint64)]),
    "pie_debug_csv_force_slow_path": (C.c_int, [C.c_int]),
    "pie_debug_csv_slow_tiles": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint32), C.c_void_p]),
    "pie_debug_ingest_warp_path": (C.c_int, [C.c_int]),
    "pie_debug_ingest_declined": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint32), C.c_void_p]),
    "pie_csv_rows_host": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_uint64,
                                    C.POINTER(C.c_uint64)]),
    "pie_ingest_scratch_bytes": (C.c_uint64, [C.c_int64]),
    "pie_ingest_measure_dev": (C.c_int, [C.POINTER(JsonDocsC), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "pie_ingest_fill_scratch_bytes": (C.c_uint64, [C.c_int64]),
    "pie_ingest_fill_dev": (C.c_int, [C.POINTER(JsonDocsC), C.c_void_p, C.c_void_p, C.POINTER(ArchiveViewC),
                                      C.c_void_p, C.c_void_p]),
    "pie_ingest_host": (C.c_int, [C.POINTER(JsonDocsC), C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_int64)]),
    "pie_ingest_host_release": (None, []),
    "pie_archive_step_json_host": (C.c_int, [C.POINTER(JsonDocsC), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.POINTER(DailyOutC), C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64,
                                             C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]),
    "pie_show_payloads_scratch_bytes": (C.c_uint64, [C.c_int64]),
    "pie_show_payloads_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pie_get_timestamps_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.POINTER(JsonDocsC), C.c_int32, C.POINTER(DocTimesC),
                                         C.c_void_p, C.c_void_p]),
    "pie_archive_due_scratch_bytes": (C.c_uint64, [C.c_int64]),
    "pie_archive_due_dev": (C.c_int, [C.POINTER(ArchiveViewC), C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "pie_archive_expired_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "pie_archive_analytics_host": (C.c_int, [C.POINTER(ArchiveViewC), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.POINTER(DailyOutC)]),
}


def load():
    """Load the shared library (no GPU needed to load; compute calls need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "sph_pie_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().pie_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == PIE_OK:
        return
    msg = last_error()
    if rc == PIE_ERR_RANGE:
        raise JsRangeError(rc, msg)
    if rc == PIE_ERR_UNSUPPORTED_DATE:
        raise UnsupportedDateError(rc, msg)
    if rc == PIE_ERR_SCHEMA:
        raise SchemaError(rc, msg)
    if rc == PIE_ERR_UNSUPPORTED_JSON:
        raise UnsupportedJsonError(rc, msg)
    raise PieError(rc, msg)


_initialised = False


def init(device: int = -1) -> None:
    """pie_init: verifies an sm_100 device; raises PieError(PIE_ERR_NO_DEVICE) otherwise."""
    global _initialised
    check(load().pie_init(device))
    _initialised = True


def ensure_init() -> None:
    if not _initialised:
        init(-1)


def last_transfer_bytes():
    h, d = C.c_uint64(0), C.c_uint64(0)
    load().pie_last_transfer_bytes(C.byref(h), C.byref(d))
    return int(h.value), int(d.value)
