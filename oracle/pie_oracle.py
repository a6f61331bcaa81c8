"""CPU restatement of the reference's archive-analytics and export-row functions.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sph_pie_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
use it, and only as the checker.

PARITY STATUS: **parity unpinned** for values.  The reference is JavaScript and no
JS engine exists in this image (node, d8, quickjs, ... all probed absent), so the
reference cannot be executed to produce golden vectors.  The reference's own test
material for this path is one fixture (``scripts/simulate-webhook.js:42-65``) that
is checked only against the *same module's* builders (``:75-95``), i.e. it pins
column order and shape, not values.  ``tests/golden/`` holds that fixture plus the
row this restatement derives from it by reading the source; it is a
hand-derivation, not reference output.

Every function cites the reference lines it restates.  Inputs are the Python
images of ``JSON.parse`` output: dict / list / str / int / float / bool / None,
plus ``UNDEFINED`` for a missing property where JS distinguishes it from null.
"""
from __future__ import annotations

import math
from decimal import Decimal

# public/app.js:1-13  (ISSUE_MAP keys, in insertion order)
PRIMARY_ISSUES = [
    "Tracking lost",
    "Failed to launch",
    "Command delay",
    "RF link",
    "Battery",
    "Motor or prop",
    "Sensor or IMU",
    "Software or show control",
    "Operator input",
    "Other",
]

# server/webhookDispatcher.js:15-19 == public/app.js:16-20
EXPORT_COLUMNS = [
    "showId", "showDate", "showTime", "showLabel", "crew", "leadPilot", "monkeyLead", "showNotes",
    "entryId", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail",
    "severity", "rootCause", "actions", "operator", "batteryId", "delaySec", "commandRx", "notes",
]

# public/app.js:21-86 (ARCHIVE_METRIC_DEFS keys, insertion order) + issue metrics (:98, :3955-3994)
ARCHIVE_METRIC_KEYS = [
    "entriesCount", "completedCount", "noLaunchCount", "abortCount", "avgDelaySec",
    "maxDelaySec", "completionRate", "launchRate", "abortRate",
]
ISSUE_METRIC_PREFIX = "issue:"
ALL_METRIC_KEYS = ARCHIVE_METRIC_KEYS + [ISSUE_METRIC_PREFIX + i for i in PRIMARY_ISSUES]

_STAT_FIELD_FOR_METRIC = {
    "entriesCount": "totalEntries",
    "completedCount": "completedCount",
    "noLaunchCount": "noLaunchCount",
    "abortCount": "abortCount",
    "avgDelaySec": "avgDelaySec",
    "maxDelaySec": "maxDelaySec",
    "completionRate": "completionRate",
    "launchRate": "launchRate",
    "abortRate": "abortRate",
}


class _Undefined:
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
        return cls._inst

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEFINED = _Undefined()

# ECMAScript WhiteSpace + LineTerminator code points (String.prototype.trim)
JS_WHITESPACE = (
    "\t\n\v\f\r \u00a0\u1680"
    "\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a"
    "\u2028\u2029\u202f\u205f\u3000\ufeff"
)


# --------------------------------------------------------------------------- JS value semantics
def js_get(obj, key):
    """obj?.key for a JSON object: missing property -> undefined."""
    if isinstance(obj, dict):
        return obj.get(key, UNDEFINED)
    return UNDEFINED


def js_is_number(v) -> bool:
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def js_truthy(v) -> bool:
    if v is None or v is UNDEFINED:
        return False
    if isinstance(v, bool):
        return v
    if js_is_number(v):
        return not (v == 0 or (isinstance(v, float) and math.isnan(v)))
    if isinstance(v, str):
        return len(v) > 0
    return True  # objects / arrays


def js_or(v, default):
    """``v || default``"""
    return v if js_truthy(v) else default


def number_is_finite(v) -> bool:
    """Number.isFinite: no coercion; only numbers qualify."""
    return js_is_number(v) and math.isfinite(v)


def js_number_to_string(x) -> str:
    """Number::toString(x) radix 10 (ECMA-262 6.1.6.1.20): shortest round-trip digits,
    fixed notation for 1e-7 < |x| < 1e21, exponent notation otherwise."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if x == 0:
        return "0"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    sign = "-" if x < 0 else ""
    d = Decimal(repr(abs(x)))  # repr() is the shortest round-trip decimal (David Gay)
    _, digits, exp = d.as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:  # repr may keep a trailing ".0"
        digits.pop()
        exp += 1
    k = len(digits)
    n = k + exp  # value = 0.d1..dk * 10^n
    s = "".join(str(c) for c in digits)
    if k <= n <= 21:
        body = s + "0" * (n - k)
    elif 0 < n <= 21:
        body = s[:n] + "." + s[n:]
    elif -6 < n <= 0:
        body = "0." + "0" * (-n) + s
    else:
        e = n - 1
        es = ("+" if e >= 0 else "-") + str(abs(e))
        body = (s if k == 1 else s[0] + "." + s[1:]) + "e" + es
    return sign + body


def js_string(v) -> str:
    """String(v) for JSON-shaped values."""
    if isinstance(v, str):
        return v
    if v is None:
        return "null"
    if v is UNDEFINED:
        return "undefined"
    if isinstance(v, bool):
        return "true" if v else "false"
    if js_is_number(v):
        return js_number_to_string(v)
    if isinstance(v, list):
        return js_array_join(v, ",")
    return "[object Object]"


def js_array_join(arr, sep: str) -> str:
    """Array.prototype.join: null/undefined elements become ''."""
    return sep.join("" if (e is None or e is UNDEFINED) else js_string(e) for e in arr)


def js_trim(s: str) -> str:
    return s.strip(JS_WHITESPACE)


def js_to_lower(s: str) -> str:
    # For the ASCII targets compared on this path ('completed', 'no-launch', 'abort', 'yes', 'no')
    # only code points whose lowercase is an ASCII letter matter; Python and ECMAScript agree there
    # (both follow Unicode SpecialCasing; U+212A KELVIN SIGN -> 'k' in both).
    return s.lower()


# --------------------------------------------------------------------------- archive analytics
def compute_archive_show_stats(show):
    """public/app.js:3898-3953 computeArchiveShowStats(show)."""
    entries = js_get(show, "entries")
    entries = entries if isinstance(entries, list) else []
    completed = no_launch = abort = launched = 0
    delay_values = []
    issue_counts = {}
    for entry in entries:
        status = js_to_lower(js_string(js_or(js_get(entry, "status"), "")))  # :3907
        if status == "completed":
            completed += 1
        elif status == "no-launch":
            no_launch += 1
        elif status == "abort":
            abort += 1
        if js_to_lower(js_string(js_or(js_get(entry, "launched"), ""))) == "yes":  # :3915
            launched += 1
        d = js_get(entry, "delaySec")
        if number_is_finite(d):  # :3918
            delay_values.append(float(d))
        pi = js_get(entry, "primaryIssue")
        issue = js_trim(pi) if isinstance(pi, str) else ""  # :3921
        if issue:
            normalized = issue if issue in PRIMARY_ISSUES else "Other"  # :3923
            issue_counts[normalized] = issue_counts.get(normalized, 0) + 1
    total = len(entries)
    delay_sum = 0.0
    for v in delay_values:  # :3928 left-to-right reduce, initial 0
        delay_sum = delay_sum + v
    avg_delay = delay_sum / len(delay_values) if delay_values else None
    max_delay = js_math_max(delay_values) if delay_values else None
    completion_rate = (completed / total) * 100 if total else None
    launch_rate = (launched / total) * 100 if total else None
    abort_rate = (abort / total) * 100 if total else None
    issue_rates = {}
    for issue in PRIMARY_ISSUES:  # :3935-3938
        count = issue_counts.get(issue, 0)
        issue_rates[issue] = (count / total) * 100 if total else None
    return {
        "totalEntries": total,
        "completedCount": completed,
        "noLaunchCount": no_launch,
        "abortCount": abort,
        "launchedCount": launched,
        "avgDelaySec": avg_delay,
        "maxDelaySec": max_delay,
        "completionRate": completion_rate,
        "launchRate": launch_rate,
        "abortRate": abort_rate,
        "issueCounts": issue_counts,
        "issueRates": issue_rates,
    }


def js_math_max(values):
    """Math.max(...values): +0 > -0; values here are finite."""
    best = values[0]
    for v in values[1:]:
        if v > best or (v == best and v == 0 and math.copysign(1.0, best) < 0 and math.copysign(1.0, v) > 0):
            best = v
    return best


def js_math_min(values):
    """Math.min(...values): -0 < +0."""
    best = values[0]
    for v in values[1:]:
        if v < best or (v == best and v == 0 and math.copysign(1.0, v) < 0 and math.copysign(1.0, best) > 0):
            best = v
    return best


def parse_show_date_time(date_str, time_str, tz_offset_minutes=0):
    """public/app.js:4118-4126 parseShowDateTime.  ``Date.parse`` of ``${date}T${time}``: only the
    ECMA-262 date-time string format is specified (21.4.1.32); a date-time form without an offset is
    local time.  Anything else falls to V8's legacy parser, which is implementation-defined and is
    NOT restated: such strings raise NotImplementedError here."""
    if not isinstance(date_str, str) or not date_str:
        return None
    time = time_str if (isinstance(time_str, str) and time_str) else "00:00"
    iso = f"{date_str}T{time}"
    ms = _parse_iso_local(iso, tz_offset_minutes)
    return ms


def _parse_iso_local(iso: str, tz_offset_minutes: int):
    import re

    m = re.fullmatch(r"(\d{4})-(\d{2})-(\d{2})T(\d{2}):(\d{2})(?::(\d{2})(?:\.(\d{1,3}))?)?", iso)
    if not m:
        raise NotImplementedError(f"non-ISO date string {iso!r}: V8 legacy Date.parse is not restated")
    y, mo, d, h, mi = (int(m.group(i)) for i in range(1, 6))
    s = int(m.group(6)) if m.group(6) else 0
    ms = int((m.group(7) or "0").ljust(3, "0"))
    if not (1 <= mo <= 12 and 1 <= d <= days_in_month(y, mo) and h <= 24 and mi <= 59 and s <= 59):
        return None
    if h == 24 and (mi or s or ms):
        return None
    days = days_from_civil(y, mo, d)
    local = ((days * 24 + h) * 60 + mi) * 60000 + s * 1000 + ms
    return float(local - tz_offset_minutes * 60000)


def days_in_month(y, m):
    if m == 2:
        return 29 if (y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)) else 28
    return 30 if m in (4, 6, 9, 11) else 31


def days_from_civil(y, m, d):
    """Days since 1970-01-01 (proleptic Gregorian)."""
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


def civil_from_days(z):
    z += 719468
    era = (z if z >= 0 else z - 146096) // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + (3 if mp < 10 else -9)
    return (y + (m <= 2), m, d)


def get_show_timestamp(show, tz_offset_minutes=0):
    """public/app.js:4092-4116 getShowTimestamp(show)."""
    if not js_truthy(show):
        return None
    created = js_get(show, "createdAt")
    if number_is_finite(created):
        return float(created)
    parsed = parse_show_date_time(js_get(show, "date"), js_get(show, "time"), tz_offset_minutes)
    if parsed is not None:
        return parsed
    archived = js_get(show, "archivedAt")
    if number_is_finite(archived):
        return float(archived)
    entries = js_get(show, "entries")
    if isinstance(entries, list) and entries:
        ts = sorted(float(js_get(e, "ts")) for e in entries if number_is_finite(js_get(e, "ts")))
        if ts:
            return ts[0]
    return None


class JsRangeError(ValueError):
    """RangeError: Invalid time value (Date.prototype.toISOString on an invalid Date)."""


MAX_TIME_MS = 8.64e15


def local_day_start(timestamp: float, tz_offset_minutes=0):
    """``d = new Date(ts); d.setHours(0,0,0,0); d.getTime()`` (public/app.js:3412-3414) in a zone
    that is a fixed ``tz_offset_minutes`` east of UTC.  Returns None for an invalid Date."""
    if not math.isfinite(timestamp) or abs(timestamp) > MAX_TIME_MS:
        return None  # TimeClip -> NaN
    t = int(timestamp)  # TimeClip: ToIntegerOrInfinity truncates toward zero
    off = tz_offset_minutes * 60000
    local = t + off
    start = (local // 86400000) * 86400000 - off
    if abs(start) > MAX_TIME_MS:
        return None
    return start


def iso_date_key(ms: int) -> str:
    """``new Date(ms).toISOString().slice(0, 10)`` (public/app.js:3415)."""
    y, m, d = civil_from_days(ms // 86400000)
    if 0 <= y <= 9999:
        full = f"{y:04d}-{m:02d}-{d:02d}"
    else:
        full = f"{'+' if y > 0 else '-'}{abs(y):06d}-{m:02d}-{d:02d}"
    return full[:10]


def metric_value(metric_key, stats):
    """metricDef.getValue(stats, show): public/app.js:21-86 and :3978-3988."""
    if metric_key in _STAT_FIELD_FOR_METRIC:
        return stats[_STAT_FIELD_FOR_METRIC[metric_key]]
    if metric_key.startswith(ISSUE_METRIC_PREFIX):
        issue = metric_key[len(ISSUE_METRIC_PREFIX):]
        rates = stats.get("issueRates")
        if rates is not None and issue in rates:
            v = rates[issue]
            return v if number_is_finite(v) else (0 if v == 0 and v is not None else None)
        return None
    raise KeyError(metric_key)


def is_valid_metric_value(v) -> bool:
    """public/app.js:4128-4134 (for number-or-null inputs)."""
    if v is None or v is UNDEFINED:
        return False
    return math.isfinite(float(v))


def build_archive_daily_groups(shows, tz_offset_minutes=0):
    """public/app.js:3401-3443 buildArchiveDailyGroups(shows), without the display label."""
    groups = {}
    lst = shows if isinstance(shows, list) else []
    for show in lst:
        if not js_truthy(show):
            continue
        ts = get_show_timestamp(show, tz_offset_minutes)
        if ts is None or not math.isfinite(ts):
            continue
        start = local_day_start(ts, tz_offset_minutes)
        if start is None:
            raise JsRangeError("Invalid time value")  # :3415 toISOString throws
        key = iso_date_key(start)
        g = groups.get(key)
        if g is None:
            g = {"dateKey": key, "timestamp": start, "midpoint": start + 12 * 60 * 60 * 1000,
                 "shows": [], "metrics": {}, "totalShows": 0}
            groups[key] = g
        g["shows"].append({"show": show, "stats": compute_archive_show_stats(show)})
    out = sorted(groups.values(), key=lambda g: g["timestamp"])  # stable, :3435
    for g in out:
        g["totalShows"] = len(g["shows"])
    return out


def group_metric_summary(group, metric_key):
    """public/app.js:3445-3502 getOrCreateGroupMetricSummary, numeric part
    (average / min / max / count / totalShows and the per-show numeric values)."""
    values = []
    numeric = []
    for item in group["shows"]:
        v = metric_value(metric_key, item["stats"])
        n = float(v) if is_valid_metric_value(v) else None
        values.append(n)
        if n is not None:
            numeric.append(n)
    total = 0.0
    for v in numeric:  # :3481 left-to-right reduce, initial 0
        total = total + v
    return {
        "average": total / len(numeric) if numeric else None,
        "min": js_math_min(numeric) if numeric else None,
        "max": js_math_max(numeric) if numeric else None,
        "count": len(numeric),
        "totalShows": len(group["shows"]),
        "values": values,
    }


def js_is_array_index(key: str) -> bool:
    """A property key that is an array index (ECMA-262 6.1.7): the canonical decimal string of an integer in
    0 .. 2^32 - 2.  OrdinaryOwnPropertyKeys lists such keys first, in ascending numeric order, before the string
    keys in insertion order — which is the order Object.entries walks."""
    return key.isascii() and key.isdigit() and (key == "0" or key[0] != "0") and len(key) <= 10 and int(key) <= 2 ** 32 - 2


def js_object_entries(d: dict):
    """Object.entries of an object whose properties were created in the dict's insertion order."""
    idx = sorted((k for k in d if js_is_array_index(k)), key=int)
    return [(k, d[k]) for k in idx] + [(k, v) for k, v in d.items() if not js_is_array_index(k)]


def compute_metrics(show):
    """public/app.js:5024-5047 computeMetrics(show) (live show header)."""
    entries = js_or(js_get(show, "entries"), [])
    planned_yes = sum(1 for e in entries if js_get(e, "planned") == "Yes")
    completed = sum(1 for e in entries if js_get(e, "status") == "Completed")
    no_launch = sum(1 for e in entries if js_get(e, "status") == "No-launch")
    abort = sum(1 for e in entries if js_get(e, "status") == "Abort")
    delays = [float(js_get(e, "delaySec")) for e in entries if js_is_number(js_get(e, "delaySec"))]
    if delays:
        total = 0.0
        for v in delays:
            total = total + v
        avg = js_to_fixed2(total / len(delays))
    else:
        avg = "0.00"
    issues = {}
    for e in entries:
        pi = js_get(e, "primaryIssue")
        if js_get(e, "status") != "Completed" and js_truthy(pi):
            k = js_string(pi)
            issues[k] = issues.get(k, 0) + 1
    top = [k for k, _ in sorted(js_object_entries(issues), key=lambda kv: -kv[1])[:3]]  # Array.prototype.sort is stable
    success = js_math_round((completed / planned_yes) * 100) if planned_yes else 0
    return {"successRate": success, "countCompleted": completed, "countNoLaunch": no_launch,
            "countAbort": abort, "avgDelay": avg, "topIssues": top}


def js_math_round(x: float):
    """Math.round: the integer closest to x, ties toward +inf (floor(x + 0.5) would round the double just
    below 0.5 up, because the sum is not exact)."""
    if not math.isfinite(x):
        return x
    r = math.floor(x)
    return r + 1 if x - r >= 0.5 else r


def js_to_fixed2(x: float) -> str:
    """Number.prototype.toFixed(2): exact decimal expansion of the double, ties away from zero
    on the exact value ("let n be an integer for which n/10^f - x is as close to zero as possible;
    if there are two such n, pick the larger n")."""
    if math.isnan(x):
        return "NaN"
    if abs(x) >= 1e21:
        return js_number_to_string(x)
    d = Decimal(abs(x)) * 100
    n = int(d.to_integral_value(rounding="ROUND_HALF_UP"))
    s = f"{n // 100}.{n % 100:02d}"
    return ("-" + s) if x < 0 else s  # step 6: "if x < 0, set s to '-' and x to -x": (-0.001).toFixed(2) is "-0.00", (-0).toFixed(2) "0.00"


# --------------------------------------------------------------------------- export rows
def build_table_row(show=UNDEFINED, entry=UNDEFINED):
    """server/webhookDispatcher.js:276-305 buildTableRow == public/app.js:5582-5612 buildWebhookRow."""
    show = {} if show is UNDEFINED else show
    entry = {} if entry is UNDEFINED else entry
    crew = js_get(show, "crew")
    crew = crew if isinstance(crew, list) else []
    actions = js_get(entry, "actions")
    actions = actions if isinstance(actions, list) else []
    completed = js_get(entry, "status") == "Completed"  # strict ===, case-sensitive

    def g(o, k):
        return js_or(js_get(o, k), "")

    d = js_get(entry, "delaySec")
    return {
        "showId": g(show, "id"),
        "showDate": g(show, "date"),
        "showTime": g(show, "time"),
        "showLabel": g(show, "label"),
        "crew": js_array_join(crew, "|"),
        "leadPilot": g(show, "leadPilot"),
        "monkeyLead": g(show, "monkeyLead"),
        "showNotes": g(show, "notes"),
        "entryId": g(entry, "id"),
        "unitId": g(entry, "unitId"),
        "planned": g(entry, "planned"),
        "launched": g(entry, "launched"),
        "status": g(entry, "status"),
        "primaryIssue": "" if completed else g(entry, "primaryIssue"),
        "subIssue": "" if completed else g(entry, "subIssue"),
        "otherDetail": "" if completed else g(entry, "otherDetail"),
        "severity": "" if completed else g(entry, "severity"),
        "rootCause": "" if completed else g(entry, "rootCause"),
        "actions": js_array_join(actions, "|"),
        "operator": g(entry, "operator"),
        "batteryId": g(entry, "batteryId"),
        "delaySec": "" if (d is None or d is UNDEFINED) else d,
        "commandRx": g(entry, "commandRx"),
        "notes": g(entry, "notes"),
    }


def build_message_payload(row=UNDEFINED):
    """server/webhookDispatcher.js:307-313 buildMessagePayload."""
    row = {} if row is UNDEFINED else row
    out = {}
    for c in EXPORT_COLUMNS:
        v = js_get(row, c)
        out[c] = "" if (v is UNDEFINED or v is None) else v
    return out


def csv_escape(value) -> str:
    """server/webhookDispatcher.js:332-338 csvEscape == public/app.js:6025-6034."""
    s = "" if (value is None or value is UNDEFINED) else js_string(value)
    if '"' in s or "," in s or "\n" in s or "\r" in s:
        return '"' + s.replace('"', '""') + '"'
    return s


def build_csv_row(row) -> str:
    """server/webhookDispatcher.js:340-342 buildCsvRow."""
    cells = []
    for c in EXPORT_COLUMNS:
        v = js_get(row, c)
        cells.append(csv_escape("" if (v is None or v is UNDEFINED) else v))  # ?? ''
    return ",".join(cells)


def export_show_as_csv(show) -> str:
    """public/app.js:5558-5570 exportShowAsCsv: header line + one line per entry, '\\n'-joined."""
    entries = js_or(js_get(show, "entries"), [])
    lines = [",".join(csv_escape(c) for c in EXPORT_COLUMNS)]
    for e in entries:
        lines.append(build_csv_row(build_table_row(show, e)))
    return "\n".join(lines)


def to_yes_no_boolean(value) -> bool:
    """server/webhookDispatcher.js:60-77 toYesNoBoolean."""
    if isinstance(value, str):
        n = js_to_lower(js_trim(value))
        if n == "yes":
            return True
        if n == "no":
            return False
    if isinstance(value, bool):
        return value
    if js_is_number(value):
        return (value != 0) if math.isfinite(value) else False
    return False


def build_archive_entry_payload(show=UNDEFINED, entry=UNDEFINED):
    """server/webhookDispatcher.js:315-330 buildArchiveEntryPayload."""
    show = {} if show is UNDEFINED else show
    entry = {} if entry is UNDEFINED else entry

    def g(o, k):
        return js_or(js_get(o, k), "")

    return {
        "showDate": g(show, "date"),
        "showTime": g(show, "time"),
        "showNumber": g(show, "label"),
        "leadPilot": g(show, "leadPilot"),
        "monkeyLead": g(show, "monkeyLead"),
        "operator": g(entry, "operator"),
        "monkeyId": g(entry, "unitId"),
        "planned": to_yes_no_boolean(js_get(entry, "planned")),
        "launched": to_yes_no_boolean(js_get(entry, "launched")),
        "commandReceived": to_yes_no_boolean(js_get(entry, "commandRx")),
        "primaryIssue": g(entry, "primaryIssue"),
        "subIssue": g(entry, "subIssue"),
    }


def json_quote(s: str) -> str:
    """QuoteJSONString (ECMA-262 25.5.2.3) for a string of well-formed code points: \" \\ and the
    five short escapes, other code units below U+0020 as \\u00xx (lower-case hex), the rest verbatim.
    (Lone surrogates, which JSON.stringify writes as \\udxxx, cannot occur in UTF-8 columns.)"""
    short = {8: "\\b", 9: "\\t", 10: "\\n", 12: "\\f", 13: "\\r", 0x22: '\\"', 0x5C: "\\\\"}
    out = ['"']
    for ch in s:
        cp = ord(ch)
        if cp in short:
            out.append(short[cp])
        elif cp < 0x20:
            out.append("\\u%04x" % cp)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def archive_entry_payload_json(show=UNDEFINED, entry=UNDEFINED) -> str:
    """JSON.stringify(buildArchiveEntryPayload(show, entry)) — the request body sendWebhookPayload hands to
    axios for every entry of an archived show (server/webhookDispatcher.js:527-540).  SerializeJSONObject
    walks the object's own string keys in insertion order, i.e. the order of the literal at :316-329; the
    values are strings and booleans only."""
    payload = build_archive_entry_payload(show, entry)
    parts = []
    for key, value in payload.items():
        if isinstance(value, bool):
            text = "true" if value else "false"
        else:
            text = json_quote(js_string(value))
        parts.append(json_quote(key) + ":" + text)
    return "{" + ",".join(parts) + "}"


# ---------------------------------------------------------------------------------------------------------
# stored documents -> shows: JSON.parse(row.data) as _mapArchiveRow does it (server/storage/sqlProvider.js:892-926)
# ---------------------------------------------------------------------------------------------------------
SHOW_DOC_KEYS = ("id", "date", "time", "label", "leadPilot", "monkeyLead", "notes", "crew", "createdAt", "archivedAt",
                 "entries", "updatedAt", "deletedAt")
ENTRY_DOC_KEYS = ("id", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail",
                  "severity", "rootCause", "operator", "batteryId", "commandRx", "notes", "actions", "delaySec", "ts")
MAX_JSON_DEPTH = 64  # deeper nesting is reported, not parsed (the kernel's kind stack)


class UnsupportedJson(ValueError):
    """Valid JSON the ingest path declines to decide: PIE_ERR_UNSUPPORTED_JSON."""


class JsObject(dict):
    """JSON.parse's object: a dict that remembers which keys the text held more than once."""
    dup_keys: frozenset = frozenset()


def _pairs_hook(pairs):
    obj = JsObject()
    dups = set()
    for k, v in pairs:
        if k in obj:
            dups.add(k)
        obj[k] = v  # the last one wins, as in JSON.parse
    obj.dup_keys = frozenset(dups)
    return obj


def _reject_constant(name):
    raise ValueError(f"{name} is not JSON")  # Python's json accepts NaN / Infinity / -Infinity; ECMA-404 does not


def _nesting_depth(text: str) -> int:
    depth = deepest = 0
    in_string = escaped = False
    for ch in text:
        if in_string:
            if escaped:
                escaped = False
            elif ch == "\\":
                escaped = True
            elif ch == '"':
                in_string = False
        elif ch == '"':
            in_string = True
        elif ch in "[{":
            depth += 1
            deepest = max(deepest, depth)
        elif ch in "]}":
            depth -= 1
    return deepest


def js_json_parse(text: str):
    """JSON.parse(text) on the Python image of JS values: every number is a Number (binary64, correctly rounded,
    overflow -> Infinity), objects are JsObject.  Raises ValueError where JSON.parse throws SyntaxError.  Python's
    json module is the parser — an implementation of the same grammar that shares nothing with the kernel."""
    import json

    return json.loads(text, parse_int=float, parse_float=float, parse_constant=_reject_constant,
                      object_pairs_hook=_pairs_hook)


def map_archive_row(data):
    """_mapArchiveRow(row) (sqlProvider.js:892-926) as far as row.data goes: the parsed show, or None when the
    text does not parse or the value is not an object (`!show || typeof show !== 'object'`; arrays ARE objects).
    `data` is the column's UTF-8 bytes or a str.  Raises UnsupportedJson for what the ingest path reports instead of
    deciding: bytes that are not UTF-8, nesting deeper than 64, a known key twice in the show or in an entry."""
    bad_utf8 = False
    if isinstance(data, (bytes, bytearray, memoryview)):
        try:
            text = bytes(data).decode("utf-8")
        except UnicodeDecodeError:
            bad_utf8 = True
            text = bytes(data).decode("utf-8", errors="replace")
    else:
        text = data
    try:
        show = js_json_parse(text)
    except RecursionError:
        raise UnsupportedJson("nesting deeper than 64")
    except ValueError:
        return None
    if bad_utf8:
        raise UnsupportedJson("the text is not UTF-8")
    if _nesting_depth(text) > MAX_JSON_DEPTH:
        raise UnsupportedJson("nesting deeper than 64")
    if show is None or not isinstance(show, (dict, list)):
        return None
    if isinstance(show, JsObject):
        if show.dup_keys & set(SHOW_DOC_KEYS):
            raise UnsupportedJson("a show key twice")
        entries = show.get("entries")
        if isinstance(entries, list):
            for e in entries:
                if isinstance(e, JsObject) and e.dup_keys & set(ENTRY_DOC_KEYS):
                    raise UnsupportedJson("an entry key twice")
    return show


def js_json_stringify(value) -> str:
    """JSON.stringify(value) for the values a stored show is made of (ECMA-262 25.5.2): objects in insertion order,
    no whitespace, Number::toString for finite numbers, null for NaN / Infinity, QuoteJSONString for strings."""
    if value is None:
        return "null"
    if value is True:
        return "true"
    if value is False:
        return "false"
    if isinstance(value, (int, float)):
        f = float(value)
        return js_number_to_string(f) if math.isfinite(f) else "null"
    if isinstance(value, str):
        return json_quote(value)
    if isinstance(value, (list, tuple)):
        return "[" + ",".join(js_json_stringify(v) for v in value) + "]"
    if isinstance(value, dict):
        return "{" + ",".join(json_quote(k) + ":" + js_json_stringify(v) for k, v in value.items()
                              if v is not UNDEFINED) + "}"
    raise TypeError(type(value).__name__)


# ---------------------------------------------------------------------------------------------------------
# _mapArchiveRow in full: the row's timestamp columns over the document's fields (sqlProvider.js:905-918)
# ---------------------------------------------------------------------------------------------------------
_JS_WS = "\t\n\v\f\r                  　﻿"


def js_string_to_number(s: str) -> float:
    """StringToNumber (ECMA-262 7.1.4.1.1): white space trimmed, '' -> 0, decimal literals (optional sign, `.5`, `5.`,
    exponents), Infinity, 0x / 0o / 0b integers; anything else NaN."""
    import re

    t = s.strip(_JS_WS)
    if t == "":
        return 0.0
    m = re.fullmatch(r"0[xX]([0-9a-fA-F]+)|0[oO]([0-7]+)|0[bB]([01]+)", t)
    if m:
        base = 16 if m.group(1) else 8 if m.group(2) else 2
        return float(int(m.group(1) or m.group(2) or m.group(3), base))
    if re.fullmatch(r"[+-]?Infinity", t):
        return -math.inf if t[0] == "-" else math.inf
    if re.fullmatch(r"[+-]?([0-9]+\.?[0-9]*([eE][+-]?[0-9]+)?|\.[0-9]+([eE][+-]?[0-9]+)?)", t):
        return float(t)  # Python's float(): the same correctly rounded value
    return math.nan


def js_to_number(v) -> float:
    """ToNumber for the values a row or a parsed document can hold."""
    if v is UNDEFINED:
        return math.nan
    if v is None:
        return 0.0
    if v is True:
        return 1.0
    if v is False:
        return 0.0
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, str):
        return js_string_to_number(v)
    if isinstance(v, list):  # ToPrimitive: join(',') — [] -> '' -> 0, [5] -> '5' -> 5
        return js_string_to_number(js_array_join(v, ","))
    return math.nan  # objects: '[object Object]'


def get_timestamp(v):
    """_getTimestamp(value) (sqlProvider.js:970-985): a finite number as it is; else Number(value) when finite (so null
    is 0, '12' is 12, true is 1); else Date.parse of a string.  Date.parse is NOT restated here (V8's legacy parser):
    a string that is not numeric raises NotImplementedError instead of guessing."""
    if js_is_number(v) and math.isfinite(v):
        return float(v)
    n = js_to_number(v)
    if math.isfinite(n):
        return n
    if isinstance(v, str):
        raise NotImplementedError(f"Date.parse({v!r})")
    return None


def map_archive_row_full(row):
    """_mapArchiveRow(row) (sqlProvider.js:892-926) for a row {data, archived_at?, created_at?, deleted_at?}: a missing
    key is undefined, None is SQL NULL."""
    if not row:
        return None
    show = map_archive_row(row.get("data") if row.get("data") is not None else "null")
    if show is None:
        return None
    if isinstance(show, list):  # an array: properties are set on it, it has no fields of its own
        show = JsObject()
    get = lambda o, k: o[k] if k in o else UNDEFINED  # noqa: E731
    archived = get_timestamp(get(row, "archived_at"))
    if archived is None:
        archived = get_timestamp(get(show, "archivedAt"))
    stored_created = get_timestamp(get(row, "created_at"))
    created = get_timestamp(get(show, "createdAt"))
    if created is None:
        created = stored_created
    if archived is not None:
        show["archivedAt"] = archived
    if created is not None:
        show["createdAt"] = created
    if not isinstance(show.get("entries"), list):
        show["entries"] = []
    if not isinstance(show.get("crew"), list):
        show["crew"] = []
    return show


# ---------------------------------------------------------------------------------------------------------
# Date.parse for the one format ECMA-262 specifies (21.4.1.32), as far as _getTimestamp (sqlProvider.js:978-982)
# can meet it in a stored document: YYYY-MM-DD (UTC) and YYYY-MM-DDTHH:mm[:ss[.sss]][Z|+HH:mm|-HH:mm] (local time
# when no offset is given).  Everything else is V8's legacy parser and is not restated: NotImplementedError.
# ---------------------------------------------------------------------------------------------------------
def js_date_parse(s: str, tz_offset_minutes: int = 0) -> float:
    import re

    m = re.fullmatch(r"(\d{4})-(\d{2})-(\d{2})(?:T(\d{2}):(\d{2})(?::(\d{2})(?:\.(\d{3}))?)?(Z|[+-]\d{2}:\d{2})?)?", s)
    if not m:
        raise NotImplementedError(f"Date.parse({s!r}): outside the ECMA-262 date-time format")
    y, mo, d = int(m.group(1)), int(m.group(2)), int(m.group(3))
    date_only = m.group(4) is None
    h, mi = (int(m.group(4)), int(m.group(5))) if not date_only else (0, 0)
    sec = int(m.group(6)) if m.group(6) else 0
    ms = int(m.group(7)) if m.group(7) else 0
    off = m.group(8)
    if off and off != "Z":
        oh, om = int(off[1:3]), int(off[4:6])
        if oh > 23 or om > 59:
            return math.nan
    if not (1 <= mo <= 12 and 1 <= d <= 31 and h <= 24 and mi <= 59 and sec <= 59):
        return math.nan
    if h == 24 and (mi or sec or ms):
        return math.nan
    if d > days_in_month(y, mo):
        raise NotImplementedError(f"Date.parse({s!r}): a day past the end of the month (engines disagree)")
    local = ((days_from_civil(y, mo, d) * 24 + h) * 60 + mi) * 60000 + sec * 1000 + ms
    if date_only or off == "Z":
        return float(local)
    if off:
        sign = -1 if off[0] == "-" else 1
        return float(local - sign * (int(off[1:3]) * 60 + int(off[4:6])) * 60000)
    return float(local - tz_offset_minutes * 60000)


def get_timestamp_tz(v, tz_offset_minutes: int = 0):
    """_getTimestamp(value) (sqlProvider.js:970-985) with its Date.parse leg for the ECMA-262 format."""
    if js_is_number(v) and math.isfinite(v):
        return float(v)
    n = js_to_number(v)
    if math.isfinite(n):
        return n
    if isinstance(v, str):
        parsed = js_date_parse(v, tz_offset_minutes)
        if math.isfinite(parsed):
            return parsed
    return None


def _nullish(a, b):
    return b if a is None else a


def map_archive_row_all(row, tz_offset_minutes: int = 0, provider: str = "sql"):
    """_mapArchiveRow(row) (sqlProvider.js:892-926) in full: the row's archived_at / created_at / deleted_at columns
    against the document's own fields, every one through _getTimestamp (null -> 0, numeric text -> its number, an
    ISO date-time text -> Date.parse); deletedAt set or deleted; entries / crew made arrays."""
    if not row:
        return None
    show = map_archive_row(row.get("data") if row.get("data") is not None else "null")
    if show is None:
        return None
    if isinstance(show, list):
        show = JsObject()
    get = lambda o, k: o[k] if k in o else UNDEFINED  # noqa: E731
    ts = lambda v: get_timestamp_tz(v, tz_offset_minutes)  # noqa: E731
    archived = _nullish(ts(get(row, "archived_at")), ts(get(show, "archivedAt")))
    if provider == "postgres":  # postgresProvider.js:721 asks the row first
        created = _nullish(ts(get(row, "created_at")), ts(get(show, "createdAt")))
    else:                       # sqlProvider.js:906-907
        created = _nullish(ts(get(show, "createdAt")), ts(get(row, "created_at")))
    deleted = _nullish(ts(get(row, "deleted_at")), ts(get(show, "deletedAt")))
    if archived is not None:
        show["archivedAt"] = archived
    if created is not None:
        show["createdAt"] = created
    if deleted is not None:
        show["deletedAt"] = deleted
    else:
        show.pop("deletedAt", None)
    if not isinstance(show.get("entries"), list):
        show["entries"] = []
    if not isinstance(show.get("crew"), list):
        show["crew"] = []
    return show


# ---------------------------------------------------------------------------------------------------------
# archive maintenance decisions (server/storage/sqlProvider.js:758-816, :863-890, :991-1009)
# ---------------------------------------------------------------------------------------------------------
AUTO_ARCHIVE_WINDOW_MS = 12 * 60 * 60 * 1000  # sqlProvider.js:9
ARCHIVE_RETENTION_MONTHS = 2                  # sqlProvider.js:10
UNDATED_KEY = "__undated__"                   # sqlProvider.js:774


def archive_daily_shows_decision(rows_data, now_ms: float, tz_offset_minutes: int = 0):
    """The decision of _archiveDailyShows (sqlProvider.js:758-816) for the rows of `shows`: which rows are archived
    now, and in which order they are saved / dispatched.  Rows whose text does not parse to an object are skipped
    (:767-772).  Groups are keyed by show.date.trim() (or '__undated__'), in order of first appearance (a Map); a
    group is due when now - earliest >= 12 h, earliest = min over the group of _getTimestamp(item.createdAt) where
    item.createdAt = _getTimestamp(show.createdAt) ?? _getTimestamp(show.updatedAt) — and _getTimestamp(null) is 0
    (Number(null)), so a show without any usable timestamp counts as created at the epoch (:783-792).
    Returns (due: list[bool] per row, order: row indices in the order the reference archives them)."""
    ts = lambda v: get_timestamp_tz(v, tz_offset_minutes)  # noqa: E731
    groups = {}
    for i, data in enumerate(rows_data):
        show = map_archive_row(data if data is not None else "null")  # JSON.parse + `typeof show !== 'object'`
        if show is None:
            continue
        date = show.get("date") if isinstance(show, dict) else UNDEFINED
        key = js_trim(date) if isinstance(date, str) and js_trim(date) else UNDATED_KEY
        get = lambda k: show[k] if isinstance(show, dict) and k in show else UNDEFINED  # noqa: E731
        created = _nullish(ts(get("createdAt")), ts(get("updatedAt")))
        groups.setdefault(key, []).append((i, created))
    due = [False] * len(rows_data)
    order = []
    for items in groups.values():
        earliest = None
        for _, created in items:
            value = ts(created)  # created is a number or None; _getTimestamp(null) === 0
            if value is None:
                continue
            if earliest is None or value < earliest:
                earliest = value
        if earliest is None:
            continue
        if now_ms - earliest >= AUTO_ARCHIVE_WINDOW_MS:
            for i, _ in items:
                due[i] = True
                order.append(i)
    return due, order


def js_time_clip(t: float) -> float:
    """TimeClip (ECMA-262 21.4.1.31): NaN beyond +-8.64e15, else truncated toward zero."""
    if not math.isfinite(t) or abs(t) > 8.64e15:
        return math.nan
    return float(math.trunc(t)) + 0.0


def add_months(timestamp, months: int, tz_offset_minutes: int = 0):
    """_addMonths(timestamp, months) (sqlProvider.js:999-1009): new Date(timestamp), setMonth(getMonth() + months) in
    LOCAL time (a fixed-offset zone here), getTime().  setMonth keeps the day of the month and lets MakeDay carry an
    overflow into the following month (31 Dec + 2 months = 3 Mar, or 2 Mar in a leap year)."""
    if not (js_is_number(timestamp) and math.isfinite(timestamp)):
        return timestamp
    t = js_time_clip(float(timestamp))
    if math.isnan(t):
        return timestamp
    off = tz_offset_minutes * 60000
    local = int(t) + off
    day, in_day = local // 86400000, local % 86400000
    y, m, d = civil_from_days(day)
    m0 = (m - 1) + months
    ym, mn = y + m0 // 12, m0 % 12
    new_day = days_from_civil(ym, mn + 1, 1) + d - 1
    v = float(new_day * 86400000 + in_day - off)
    return js_time_clip(v)


def is_archive_expired(created_at, now_ms: float, tz_offset_minutes: int = 0) -> bool:
    """_isArchiveExpired (sqlProvider.js:991-997): now >= _addMonths(createdAt, 2); false for a non-finite createdAt
    (and for an expiry that TimeClip turned into NaN)."""
    if not (js_is_number(created_at) and math.isfinite(created_at)):
        return False
    expiry = add_months(created_at, ARCHIVE_RETENTION_MONTHS, tz_offset_minutes)
    return bool(now_ms >= expiry)  # NaN compares false


def purge_expired_archives_decision(rows, now_ms: float, tz_offset_minutes: int = 0):
    """_purgeExpiredArchives (sqlProvider.js:863-890) for rows {data, created_at}: expired[i].  A text that does not
    parse leaves show null; `show?.createdAt` of anything that is not an object is undefined, so the row's own
    created_at column decides."""
    ts = lambda v: get_timestamp_tz(v, tz_offset_minutes)  # noqa: E731
    out = []
    for row in rows:
        data = row.get("data")
        try:
            show = js_json_parse(data if isinstance(data, str) else bytes(data).decode("utf-8"))
        except (ValueError, RecursionError, TypeError, AttributeError):
            show = None
        doc_created = show["createdAt"] if isinstance(show, dict) and "createdAt" in show else UNDEFINED
        created = _nullish(ts(doc_created), ts(row["created_at"] if "created_at" in row else UNDEFINED))
        out.append(False if created is None else is_archive_expired(created, now_ms, tz_offset_minutes))
    return out


# ---------------------------------------------------------------------------------------------------------
# the schemaVersion 2 show payload (server/webhookDispatcher.js:460-496, :545-585)
# ---------------------------------------------------------------------------------------------------------
def normalize_entry_list(show):
    """normalizeEntryList (webhookDispatcher.js:460-470): {...entry, actions: Array.isArray(entry?.actions) ?
    entry.actions : []} for every element when show.entries is an array, else []."""
    if not js_truthy(show):
        return []
    entries = js_get(show, "entries")
    if not isinstance(entries, list):
        return []
    out = []
    for entry in entries:
        e = dict(entry) if isinstance(entry, dict) else {}  # spreading a non-object copies no own properties
        actions = js_get(entry, "actions")
        e["actions"] = actions if isinstance(actions, list) else []  # an existing key keeps its place
        out.append(e)
    return out


def build_show_summary(show=UNDEFINED):
    """buildShowSummary (webhookDispatcher.js:472-488): `|| ''` for the texts, `?? null` for the four timestamps."""
    show = {} if show is UNDEFINED else show
    crew = js_get(show, "crew")

    def g(k):
        return js_or(js_get(show, k), "")

    def n(k):
        v = js_get(show, k)
        return None if (v is UNDEFINED or v is None) else v

    return {"id": g("id"), "label": g("label"), "date": g("date"), "time": g("time"),
            "crew": crew if isinstance(crew, list) else [], "leadPilot": g("leadPilot"), "monkeyLead": g("monkeyLead"),
            "notes": g("notes"), "createdAt": n("createdAt"), "updatedAt": n("updatedAt"), "archivedAt": n("archivedAt"),
            "deletedAt": n("deletedAt")}


def normalize_meta(meta):
    """normalizeMeta (webhookDispatcher.js:490-496): a shallow copy of a non-empty plain object, else null."""
    if not isinstance(meta, dict) or not meta:
        return None
    return dict(meta)


def dispatch_show_payload(event, show, dispatched_at: str, target_url, target_method, meta=UNDEFINED):
    """The payload object dispatchShowEvent builds for every event but 'show.archived'
    (webhookDispatcher.js:545-584): schemaVersion 2, the table / csv / message views of every entry, the show summary
    twice and the entries as they are.  `dispatched_at` stands for new Date().toISOString(), the target for
    activeConfig.url / .method."""
    show = show if isinstance(show, dict) else {}
    crew = js_get(show, "crew")
    normalized = dict(show)
    normalized["crew"] = crew if isinstance(crew, list) else []
    normalized["entries"] = normalize_entry_list(show)
    summary = build_show_summary(normalized)
    table_rows = [build_table_row(normalized, e) for e in normalized["entries"]]

    def cell(row, c):
        v = js_get(row, c)
        return "" if (v is UNDEFINED or v is None) else v  # ?? ''

    payload = {
        "event": event, "schemaVersion": 2, "dispatchedAt": dispatched_at,
        "target": {"url": target_url, "method": target_method},
        "table": {"columns": list(EXPORT_COLUMNS), "rows": [[cell(r, c) for c in EXPORT_COLUMNS] for r in table_rows]},
        "csv": {"header": list(EXPORT_COLUMNS), "rows": [build_csv_row(r) for r in table_rows]},
        "message": {"show": summary, "entries": table_rows},
        "show": summary,
        "entries": normalized["entries"],
    }
    m = normalize_meta(meta)
    if m:
        payload["meta"] = m
    return payload


def show_payload_json(event, show, dispatched_at: str, target_url, target_method, meta=UNDEFINED) -> str:
    """JSON.stringify of that payload — the request body axios sends (sendWebhookPayload, :362-373)."""
    return js_json_stringify(dispatch_show_payload(event, show, dispatched_at, target_url, target_method, meta))
