"""Host-side mirror of the reference's export-row functions (server/webhookDispatcher.js), same
names and argument meaning.  The per-entry CSV strings are produced on the GPU through the C ABI;
the pure re-shaping functions (object <-> column order) are plain host code, as in the reference.

    EXPORT_COLUMNS                       server/webhookDispatcher.js:15-19
    buildCsvRows(show)                   tableRows.map(buildCsvRow) of dispatchShowEvent (:556, :571)
    exportShowAsCsv(show)                public/app.js:5558-5570 (returns the CSV text)
    buildMessagePayload(rowObject)       server/webhookDispatcher.js:307-313
    archiveEntryPayloadBodies(show)      JSON.stringify(buildArchiveEntryPayload(show, entry)) for every entry: the
                                         bodies dispatchShowEvent('show.archived') posts one by one (:520-540)
    buildArchiveEntryPayload(show, entry)  :315-330, as an object (parsed back from the GPU's JSON text)
    showEventPayloadBodies(shows, event, ...)  JSON.stringify of the schemaVersion 2 payload dispatchShowEvent builds for
                                         every other event (:545-584), one request body per show
    buildShowSummary(show)               :472-488, as an object (parsed back from the GPU's JSON text)
"""
from __future__ import annotations

from typing import List, Optional

from .columnar import pack_shows
from .ops import archive_payloads, csv_rows

EXPORT_COLUMNS = [
    "showId", "showDate", "showTime", "showLabel", "crew", "leadPilot", "monkeyLead", "showNotes",
    "entryId", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail",
    "severity", "rootCause", "actions", "operator", "batteryId", "delaySec", "commandRx", "notes",
]


def buildCsvRowsMany(shows: List[Optional[dict]]) -> List[List[str]]:
    """csv.rows of every show in one launch: buildCsvRow(buildTableRow(show, entry)) per entry."""
    table = pack_shows(shows)
    rows = csv_rows(table).rows()
    eo = table.entry_offsets.tolist()
    return [rows[eo[s]:eo[s + 1]] for s in range(len(shows))]


def buildCsvRows(show: Optional[dict]) -> List[str]:
    return buildCsvRowsMany([show])[0]


def exportShowAsCsv(show: dict) -> str:
    """Header line + one line per entry, joined by '\\n' (public/app.js:5567).  The header cells
    are EXPORT_COLUMNS through csvEscape, which leaves these identifiers unchanged."""
    return "\n".join([",".join(EXPORT_COLUMNS)] + buildCsvRows(show))


def buildMessagePayload(rowObject: Optional[dict] = None) -> dict:
    row = rowObject if isinstance(rowObject, dict) else {}
    return {c: ("" if row.get(c) is None else row[c]) for c in EXPORT_COLUMNS}


def archiveEntryPayloadBodiesMany(shows: List[Optional[dict]]) -> List[List[str]]:
    """Request bodies of every entry of every show in one launch."""
    table = pack_shows(shows)
    rows = archive_payloads(table).rows()
    eo = table.entry_offsets.tolist()
    return [rows[eo[s]:eo[s + 1]] for s in range(len(shows))]


def archiveEntryPayloadBodies(show: Optional[dict]) -> List[str]:
    return archiveEntryPayloadBodiesMany([show])[0]


def buildArchiveEntryPayload(show: Optional[dict] = None, entry: Optional[dict] = None) -> dict:
    """The payload object of ONE entry (reference :315-330).  Computed on the GPU like the bulk call and
    parsed back, so the mirror has no second implementation of toYesNoBoolean / the `|| ''` defaults."""
    import json

    s = dict(show) if isinstance(show, dict) else {}
    s["entries"] = [entry if isinstance(entry, dict) else {}]
    return json.loads(archiveEntryPayloadBodies(s)[0])


def _json_quote(text: str) -> str:
    """QuoteJSONString (ECMA-262 25.5.2.3) of a caller-supplied string (event name, URL ...)."""
    import json

    return json.dumps(text, ensure_ascii=False)


def payload_frame(event: str, dispatchedAt: str, targetUrl: str, targetMethod: str, meta: Optional[dict] = None):
    """The texts around the per-show part of the schemaVersion 2 payload (webhookDispatcher.js:557-565, :580-583):
    (head, tail) as UTF-8 bytes.  `meta` is normalizeMeta's result: a non-empty plain object is appended, anything
    else leaves the key out; its values go through Python's json (strings, numbers, booleans, null, nesting)."""
    import json

    head = ('{"event":' + _json_quote(event) + ',"schemaVersion":2,"dispatchedAt":' + _json_quote(dispatchedAt) +
            ',"target":{"url":' + _json_quote(targetUrl) + ',"method":' + _json_quote(targetMethod) + '},')
    tail = "}"
    if isinstance(meta, dict) and meta:
        tail = ',"meta":' + json.dumps(meta, ensure_ascii=False, separators=(",", ":")) + "}"
    return head.encode("utf-8"), tail.encode("utf-8")


def showEventPayloadBodies(shows: List[Optional[dict]], event: str, dispatchedAt: str, targetUrl: str, targetMethod: str,
                           meta: Optional[dict] = None, device="cuda") -> List[str]:
    """The request body dispatchShowEvent(event, show, meta) posts for every show (any event but 'show.archived'), in one
    launch.  Shows are provider-normalised documents (their entries carry the 17 stored keys)."""
    from .ops import show_payloads

    table = pack_shows(shows).to(device)
    head, tail = payload_frame(event, dispatchedAt, targetUrl, targetMethod, meta)
    return show_payloads(table, head, tail).documents()


def buildShowSummary(show: Optional[dict] = None, device="cuda") -> dict:
    """buildShowSummary(show) (reference :472-488), computed on the GPU like the bulk call and parsed back."""
    import json

    body = showEventPayloadBodies([show if isinstance(show, dict) else {}], "", "", "", "", None, device)[0]
    return json.loads(body)["show"]
