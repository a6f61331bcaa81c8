"""The oracle against the reference's only fixture (scripts/simulate-webhook.js:42-65) and against
hand-derived known answers.  The reference's own check (:75-95) is restated in
test_reference_self_check; it pins column order and shape only, hence "parity unpinned" for values."""
import json
import math
import os

import pie_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, "golden", "webhook_fixture.json")))

# written by hand from server/webhookDispatcher.js:276-342, not produced by any code in this repo
HAND_ROW = ["simulation-show", "2024-07-04", "21:00", "Independence Demo", "Alex|Nazar", "Alex", "Nazar",
            "Verification run", "entry-001", "Drone-01", "Yes", "Yes", "Completed", "", "", "", "", "",
            "Logged only", "Alex", "B-12", 0, "Yes", "Green across the board"]
HAND_CSV = ("simulation-show,2024-07-04,21:00,Independence Demo,Alex|Nazar,Alex,Nazar,Verification run,"
            "entry-001,Drone-01,Yes,Yes,Completed,,,,,,Logged only,Alex,B-12,0,Yes,Green across the board")


def test_export_columns_order():
    assert FIX["export_columns"] == po.EXPORT_COLUMNS
    assert len(po.EXPORT_COLUMNS) == 24 and po.EXPORT_COLUMNS[0] == "showId" and po.EXPORT_COLUMNS[-1] == "notes"


def test_reference_self_check():
    """scripts/simulate-webhook.js:75-95: table.row == EXPORT_COLUMNS.map(c => rowMap[c] ?? '') and
    message == buildMessagePayload(rowMap), after a JSON round trip."""
    row_map = po.build_table_row(FIX["show"], FIX["entry"])
    row = [row_map[c] for c in po.EXPORT_COLUMNS]
    assert json.loads(json.dumps(row)) == row
    msg = po.build_message_payload(row_map)
    assert list(msg) == po.EXPORT_COLUMNS
    assert json.loads(json.dumps(msg)) == msg


def test_fixture_hand_derived_values():
    row_map = po.build_table_row(FIX["show"], FIX["entry"])
    assert [row_map[c] for c in po.EXPORT_COLUMNS] == HAND_ROW == FIX["expected_table_row"]
    assert po.build_csv_row(row_map) == HAND_CSV == FIX["expected_csv_row"]
    assert po.build_archive_entry_payload(FIX["show"], FIX["entry"]) == {
        "showDate": "2024-07-04", "showTime": "21:00", "showNumber": "Independence Demo", "leadPilot": "Alex",
        "monkeyLead": "Nazar", "operator": "Alex", "monkeyId": "Drone-01", "planned": True, "launched": True,
        "commandReceived": True, "primaryIssue": "", "subIssue": ""}


def test_fixture_show_stats():
    st = po.compute_archive_show_stats({**FIX["show"], "entries": [FIX["entry"]]})
    assert st["totalEntries"] == 1 and st["completedCount"] == 1 and st["launchedCount"] == 1
    assert st["avgDelaySec"] == 0 and st["maxDelaySec"] == 0
    assert st["completionRate"] == 100 and st["launchRate"] == 100 and st["abortRate"] == 0
    assert st["issueCounts"] == {} and all(v == 0 for v in st["issueRates"].values())
    assert st == FIX["expected_show_stats"]


def test_csv_escape_known_answers():
    # server/webhookDispatcher.js:332-338
    assert po.csv_escape(None) == "" and po.csv_escape(po.UNDEFINED) == ""
    assert po.csv_escape("plain") == "plain"
    assert po.csv_escape("a,b") == '"a,b"'
    assert po.csv_escape('say "hi"') == '"say ""hi"""'
    assert po.csv_escape("l1\nl2") == '"l1\nl2"' and po.csv_escape("a\rb") == '"a\rb"'
    assert po.csv_escape(0) == "0" and po.csv_escape(1.5) == "1.5" and po.csv_escape(float("nan")) == "NaN"
    assert po.csv_escape(True) == "true"


def test_completed_blanks_issue_fields():
    e = {"status": "Completed", "primaryIssue": "Battery", "subIssue": "swelling", "otherDetail": "x",
         "severity": "Major visible", "rootCause": "Hardware"}
    r = po.build_table_row({}, e)
    assert [r[k] for k in ("primaryIssue", "subIssue", "otherDetail", "severity", "rootCause")] == [""] * 5
    e["status"] = "completed"  # strict === 'Completed' (:293): lower case does NOT blank
    r = po.build_table_row({}, e)
    assert r["primaryIssue"] == "Battery" and r["severity"] == "Major visible"


def test_delay_sec_cell():
    assert po.build_table_row({}, {"delaySec": None})["delaySec"] == ""
    assert po.build_table_row({}, {})["delaySec"] == ""
    assert po.build_table_row({}, {"delaySec": 0})["delaySec"] == 0
    assert po.build_csv_row(po.build_table_row({}, {"delaySec": 12.5})).split(",")[21] == "12.5"


def test_number_to_string_known_answers():
    # ECMA-262 Number::toString examples
    cases = {0: "0", -0.0: "0", 1: "1", -1.5: "-1.5", 100: "100", 1e21: "1e+21", 1e20: "100000000000000000000",
             123456789012345680000: "123456789012345680000", 1e-6: "0.000001", 1e-7: "1e-7", 0.1: "0.1",
             0.1 + 0.2: "0.30000000000000004", 5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e+308",
             2 ** 53: "9007199254740992", 123.456: "123.456", 1.5e-10: "1.5e-10", 12345678901234567890: "12345678901234567000"}
    for x, s in cases.items():
        assert po.js_number_to_string(x) == s, (x, s)
    assert po.js_number_to_string(float("inf")) == "Infinity" and po.js_number_to_string(float("-inf")) == "-Infinity"


def test_status_matching_is_case_insensitive_and_untrimmed():
    mk = lambda s: {"entries": [{"status": s}]}
    assert po.compute_archive_show_stats(mk("COMPLETED"))["completedCount"] == 1
    assert po.compute_archive_show_stats(mk("Completed "))["completedCount"] == 0  # no trim on status
    assert po.compute_archive_show_stats(mk("No-Launch"))["noLaunchCount"] == 1
    assert po.compute_archive_show_stats(mk("abort"))["abortCount"] == 1


def test_issue_normalisation():
    mk = lambda *iss: {"entries": [{"primaryIssue": i} for i in iss]}
    st = po.compute_archive_show_stats(mk(" Battery ", "Weather", "battery", "Other", "", "  ", "RF link", " RF link　"))
    assert st["issueCounts"] == {"Battery": 1, "Other": 3, "RF link": 2}
    assert list(st["issueCounts"]) == ["Battery", "Other", "RF link"]  # insertion order
    assert st["issueRates"]["Other"] == (3 / 8) * 100 and st["issueRates"]["Tracking lost"] == 0


def test_delay_rules():
    sh = {"entries": [{"delaySec": 3}, {"delaySec": None}, {"delaySec": float("nan")}, {"delaySec": "5"},
                      {"delaySec": 0.1}, {"delaySec": 0.2}, {"delaySec": True}]}
    st = po.compute_archive_show_stats(sh)
    assert st["avgDelaySec"] == ((0.0 + 3) + 0.1 + 0.2) / 3 and st["maxDelaySec"] == 3
    assert po.compute_archive_show_stats({"entries": []})["avgDelaySec"] is None
    assert po.compute_archive_show_stats({})["completionRate"] is None


def test_daily_groups_local_midnight_and_date_key_quirk():
    day = 86400000
    t0 = 1720000000000  # 2024-07-03T09:46:40Z
    shows = [{"id": "a", "createdAt": t0, "entries": []}, {"id": "b", "createdAt": t0 + 1000, "entries": []},
             {"id": "c", "createdAt": t0 - day, "entries": []}]
    g = po.build_archive_daily_groups(shows, 0)
    assert [x["dateKey"] for x in g] == ["2024-07-02", "2024-07-03"]
    assert [[i["show"]["id"] for i in x["shows"]] for x in g] == [["c"], ["a", "b"]]
    assert g[1]["timestamp"] == (t0 // day) * day and g[1]["midpoint"] == g[1]["timestamp"] + day // 2
    # east of UTC the ISO (UTC) date of local midnight is the previous day (:3415 quirk)
    g = po.build_archive_daily_groups(shows[:1], 120)
    assert g[0]["timestamp"] == (t0 + 7200000) // day * day - 7200000 and g[0]["dateKey"] == "2024-07-02"


def test_timestamp_chain():
    assert po.get_show_timestamp({"createdAt": 5.0}) == 5.0
    assert po.get_show_timestamp({"createdAt": "5", "date": "2024-07-04", "time": "21:00"}, 0) == 1720126800000.0
    assert po.get_show_timestamp({"date": "2024-07-04", "time": ""}, -60) == 1720051200000.0 + 3600000
    assert po.get_show_timestamp({"date": "2024-13-04", "archivedAt": 7}) == 7.0
    assert po.get_show_timestamp({"entries": [{"ts": 9}, {"ts": 4}, {"ts": None}]}) == 4.0
    assert po.get_show_timestamp({"entries": []}) is None and po.get_show_timestamp(None) is None


def test_out_of_range_timestamp_throws():
    import pytest

    with pytest.raises(po.JsRangeError):
        po.build_archive_daily_groups([{"createdAt": 8.64e15 + 1}], 0)
    assert po.build_archive_daily_groups([{"createdAt": float("inf")}], 0) == []


def test_summary_min_max_signed_zero_and_order():
    grp = {"shows": [{"show": {}, "stats": po.compute_archive_show_stats({"entries": [{"delaySec": d}]})}
                     for d in (0.0, -0.0, 0.1, 0.2, 0.3)]}
    s = po.group_metric_summary(grp, "avgDelaySec")
    assert s["count"] == 5 and s["average"] == ((((0.0 + 0.0) + -0.0) + 0.1 + 0.2) + 0.3) / 5
    assert math.copysign(1, s["min"]) == 1 and s["max"] == 0.3  # (0 + -0)/1 is +0: an average is never -0
    s = po.group_metric_summary(grp, "maxDelaySec")  # Math.max(-0) is -0, and Math.min(0, -0) is -0
    assert math.copysign(1, s["min"]) == -1 and s["min"] == 0 and s["max"] == 0.3


def test_compute_metrics_known_answers():
    sh = {"entries": [{"planned": "Yes", "status": "Completed", "delaySec": 1},
                      {"planned": "Yes", "status": "Abort", "delaySec": 2, "primaryIssue": "Battery"},
                      {"planned": "No", "status": "No-launch", "delaySec": 0.005, "primaryIssue": "RF link"},
                      {"planned": "Yes", "status": "Abort", "primaryIssue": "Battery"}]}
    m = po.compute_metrics(sh)
    assert m == {"successRate": 33, "countCompleted": 1, "countNoLaunch": 1, "countAbort": 2, "avgDelay": "1.00",
                 "topIssues": ["Battery", "RF link"]}
    assert po.compute_metrics({})["avgDelay"] == "0.00" and po.compute_metrics({})["successRate"] == 0
    assert po.js_to_fixed2(1.005) == "1.00" and po.js_to_fixed2(1.255) == "1.25" and po.js_to_fixed2(2.675) == "2.67"
