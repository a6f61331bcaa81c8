"""Host-side mirror of the reference's archive-analytics functions (public/app.js), same names and
argument meaning, computed on the GPU through the C ABI.

    computeArchiveShowStats(show)                     public/app.js:3898-3953
    buildArchiveDailyGroups(shows)                    public/app.js:3401-3443
    getOrCreateGroupMetricSummary(group, metricKey)   public/app.js:3445-3502 (numeric part)
    computeMetrics(show)                              public/app.js:5024-5047 (live show header)

Shows are provider-normalised documents (dicts as `json.loads` returns them).  JS `null` is None.
Display-only members of the reference objects (`displayDate`, `label`, `shortLabel`, `formatted`:
locale formatting for the Chart.js tooltip) are UI and are not produced.
"""
from __future__ import annotations

from typing import List, Optional

from . import _lib
from .columnar import pack_shows
from .ops import archive_analytics, compute_metrics, show_stats

PRIMARY_ISSUES = [  # public/app.js:1-13
    "Tracking lost", "Failed to launch", "Command delay", "RF link", "Battery", "Motor or prop",
    "Sensor or IMU", "Software or show control", "Operator input", "Other",
]
ARCHIVE_METRIC_KEYS = [  # public/app.js:21-86, insertion order
    "entriesCount", "completedCount", "noLaunchCount", "abortCount", "avgDelaySec", "maxDelaySec",
    "completionRate", "launchRate", "abortRate",
]
ISSUE_METRIC_PREFIX = "issue:"  # public/app.js:98
ALL_METRIC_KEYS = ARCHIVE_METRIC_KEYS + [ISSUE_METRIC_PREFIX + i for i in PRIMARY_ISSUES]


def _stats_dict(i32, f64, s: int) -> dict:
    """Plane values of show s -> the object computeArchiveShowStats returns (:3939-3952)."""
    total = int(i32[_lib.SI_TOTAL][s])
    delay_n = int(i32[_lib.SI_DELAY_COUNT][s])
    # property insertion order of issueCounts = order of first occurrence: packed nibbles (k+1), 0 ends
    code = (int(i32[_lib.SI_ISSUE_ORDER_LO][s]) & 0xFFFFFFFF) | ((int(i32[_lib.SI_ISSUE_ORDER_HI][s]) & 0xFF) << 32)
    order = []
    while code & 0xF:
        order.append((code & 0xF) - 1)
        code >>= 4

    def rate(plane):
        return float(f64[plane][s]) if total else None

    return {
        "totalEntries": total,
        "completedCount": int(i32[_lib.SI_COMPLETED][s]),
        "noLaunchCount": int(i32[_lib.SI_NO_LAUNCH][s]),
        "abortCount": int(i32[_lib.SI_ABORT][s]),
        "launchedCount": int(i32[_lib.SI_LAUNCHED][s]),
        "avgDelaySec": float(f64[_lib.SF_AVG_DELAY][s]) if delay_n else None,
        "maxDelaySec": float(f64[_lib.SF_MAX_DELAY][s]) if delay_n else None,
        "completionRate": rate(_lib.SF_COMPLETION_RATE),
        "launchRate": rate(_lib.SF_LAUNCH_RATE),
        "abortRate": rate(_lib.SF_ABORT_RATE),
        "issueCounts": {PRIMARY_ISSUES[k]: int(i32[_lib.SI_ISSUE_COUNT0 + k][s]) for k in order},
        "issueRates": {PRIMARY_ISSUES[k]: rate(_lib.SF_ISSUE_RATE0 + k) for k in range(_lib.PIE_N_ISSUES)},
    }


def computeArchiveShowStats(show: Optional[dict]) -> dict:
    table = pack_shows([show])
    st = show_stats(table)
    return _stats_dict(st.i32.numpy(), st.f64.numpy(), 0)


def computeArchiveShowStatsMany(shows: List[Optional[dict]]) -> List[dict]:
    """computeArchiveShowStats mapped over a list of shows in one launch."""
    table = pack_shows(shows)
    st = show_stats(table)
    i32, f64 = st.i32.numpy(), st.f64.numpy()
    return [_stats_dict(i32, f64, s) for s in range(len(shows))]


def _civil_from_days(z: int):
    z += 719468
    era = (z if z >= 0 else z - 146096) // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + (3 if mp < 10 else -9)
    return y + (m <= 2), m, d


def _iso_date_key(ms: int) -> str:
    """new Date(ms).toISOString().slice(0, 10)  (public/app.js:3415)."""
    y, m, d = _civil_from_days(ms // 86400000)
    if 0 <= y <= 9999:
        return f"{y:04d}-{m:02d}-{d:02d}"
    return f"{'+' if y > 0 else '-'}{abs(y):06d}-{m:02d}-{d:02d}"[:10]


def buildArchiveDailyGroups(shows, tz_offset_minutes: int = 0) -> List[dict]:
    """Groups of shows per local calendar day, ascending; `tz_offset_minutes` is the fixed offset
    east of UTC that stands in for the browser's local zone (setHours(0,0,0,0), :3413)."""
    lst = shows if isinstance(shows, list) else []
    if not lst:
        return []
    table = pack_shows(lst)
    st, daily = archive_analytics(table, tz_offset_minutes)
    i32, f64 = st.i32.numpy(), st.f64.numpy()
    order = daily.show_order.numpy()
    goff = daily.group_offsets.numpy()
    gday = daily.group_day_start.numpy()
    sf, sc = daily.summary_f64.numpy(), daily.summary_count.numpy()
    groups = []
    for g in range(daily.n_groups):
        start = int(gday[g])
        members = [int(order[i]) for i in range(int(goff[g]), int(goff[g + 1]))]
        summaries = {}
        for m, key in enumerate(ALL_METRIC_KEYS):
            n = int(sc[m][g])
            summaries[key] = {
                "average": float(sf[_lib.DF_AVERAGE][m][g]) if n else None,
                "min": float(sf[_lib.DF_MIN][m][g]) if n else None,
                "max": float(sf[_lib.DF_MAX][m][g]) if n else None,
                "count": n,
            }
        groups.append({
            "dateKey": _iso_date_key(start),
            "timestamp": start,
            "midpoint": start + 12 * 60 * 60 * 1000,
            "shows": [{"show": lst[s], "stats": _stats_dict(i32, f64, s)} for s in members],
            "metrics": {},
            "totalShows": len(members),
            "_summaries": summaries,
        })
    return groups


def _metric_value(stats: dict, key: str):
    if key.startswith(ISSUE_METRIC_PREFIX):
        return stats["issueRates"].get(key[len(ISSUE_METRIC_PREFIX):])
    field = {"entriesCount": "totalEntries"}.get(key, key)
    return stats[field]


def getOrCreateGroupMetricSummary(group: Optional[dict], metricKey: str) -> Optional[dict]:
    """average / min / max / count / totalShows / showValues / valueMap of one metric over one
    daily group; memoised in group['metrics'] like the reference."""
    if not group:
        return None
    group.setdefault("metrics", {})
    if metricKey in group["metrics"]:
        return group["metrics"][metricKey]
    if metricKey not in group["_summaries"]:
        raise KeyError(f"unknown archive metric {metricKey!r}")
    base = group["_summaries"][metricKey]
    show_values = []
    for item in group["shows"]:
        v = _metric_value(item["stats"], metricKey)
        numeric = float(v) if (v is not None and v == v and abs(v) != float("inf")) else None
        show_values.append({"showId": item["show"].get("id"), "value": numeric})
    summary = dict(base)
    summary["totalShows"] = len(group["shows"])
    summary["showValues"] = show_values
    summary["valueMap"] = {e["showId"]: e for e in show_values if e["showId"]}
    group["metrics"][metricKey] = summary
    return summary


def computeMetricsMany(shows: List[Optional[dict]]) -> List[dict]:
    """computeMetrics for every show in one launch."""
    table = pack_shows(shows)
    m = compute_metrics(table)
    i32 = m.i32.cpu().tolist()
    text = m.text.cpu().numpy()
    issue = table.entry_cols["primary_issue"]
    offs, data = issue.offsets.tolist(), bytes(issue.data.cpu().numpy())
    out = []
    for s in range(table.n_shows):
        rows = [i32[_lib.CM_TOP0 + k][s] for k in range(3)]
        out.append({
            "successRate": i32[_lib.CM_SUCCESS_RATE][s],
            "countCompleted": i32[_lib.CM_COMPLETED][s],
            "countNoLaunch": i32[_lib.CM_NO_LAUNCH][s],
            "countAbort": i32[_lib.CM_ABORT][s],
            "avgDelay": bytes(text[s, : i32[_lib.CM_AVG_LEN][s]]).decode("ascii"),
            "topIssues": [data[offs[e]:offs[e + 1]].decode("utf-8") for e in rows if e >= 0],
        })
    return out


def computeMetrics(show: Optional[dict]) -> dict:
    return computeMetricsMany([show])[0]
