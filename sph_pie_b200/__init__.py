"""sph_pie_b200 — B200 (sm_100a) implementation of the sph-pie archive-analytics / export-row path.

Read DESIGN.md §0 first: the reference has no data-parallel hot path at the sizes it can reach;
this package is the SURVEY §7 fallback scope, built to the parity + measurement bar, and labelled
as such.  The compute lives in libsphpie_b200.so (CUDA, C ABI in include/sph_pie_b200.h); this
package is the host-side mirror of the reference's functions.  There is no CPU fallback.
"""
from . import _lib
from ._lib import JsRangeError, PieError, SchemaError, UnsupportedDateError, UnsupportedJsonError
from .archive import (ALL_METRIC_KEYS, ARCHIVE_METRIC_KEYS, PRIMARY_ISSUES, buildArchiveDailyGroups,
                      computeArchiveShowStats, computeArchiveShowStatsMany, computeMetrics, computeMetricsMany,
                      getOrCreateGroupMetricSummary)
from .columnar import ArchiveTable, StrCol, StrListCol, pack_shows
from .storage import listArchivedShows, mapArchiveRows
from .webhook import (EXPORT_COLUMNS, archiveEntryPayloadBodies, archiveEntryPayloadBodiesMany,
                      buildArchiveEntryPayload, buildCsvRows, buildCsvRowsMany, buildMessagePayload, exportShowAsCsv)

__all__ = [
    "ALL_METRIC_KEYS", "ARCHIVE_METRIC_KEYS", "PRIMARY_ISSUES", "ArchiveTable", "StrCol", "StrListCol",
    "JsRangeError", "PieError", "UnsupportedDateError", "buildArchiveDailyGroups", "computeArchiveShowStats",
    "computeArchiveShowStatsMany", "getOrCreateGroupMetricSummary", "pack_shows", "computeMetrics",
    "computeMetricsMany", "EXPORT_COLUMNS", "archiveEntryPayloadBodies", "archiveEntryPayloadBodiesMany",
    "buildArchiveEntryPayload", "buildCsvRows", "buildCsvRowsMany", "buildMessagePayload", "exportShowAsCsv",
    "SchemaError", "UnsupportedJsonError", "listArchivedShows", "mapArchiveRows",
]
