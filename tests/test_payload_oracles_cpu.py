"""CPU checks for the archive-entry-payload path: the JSON-level Python restatement, the columnar C
restatement and Python's own json encoder (an independent implementation of the same string grammar) must agree;
known answers for QuoteJSONString and toYesNoBoolean are written out by hand from the spec / the reference."""
import json
import os

import pytest

import oracle_c
import pie_oracle as po
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "webhook_fixture.json")))


def test_fixture_payload_json_known_answer():
    assert po.archive_entry_payload_json(FIX["show"], FIX["entry"]) == FIX["expected_archive_payload_json"]
    assert json.loads(FIX["expected_archive_payload_json"]) == FIX["expected_archive_entry_payload"]


def test_quote_json_string_known_answers():
    # ECMA-262 25.5.2.3 QuoteJSONString, Table 73: the escapes JSON.stringify uses
    assert po.json_quote('a"b') == '"a\\"b"'
    assert po.json_quote("back\\slash") == '"back\\\\slash"'
    assert po.json_quote("\b\t\n\f\r") == '"\\b\\t\\n\\f\\r"'
    assert po.json_quote("\x00\x01\x0b\x1f") == '"\\u0000\\u0001\\u000b\\u001f"'
    assert po.json_quote("\x7f / \u2028 ü 漢 🚁") == '"\x7f / \u2028 ü 漢 🚁"'  # verbatim: not escaped by JSON.stringify
    assert po.json_quote("") == '""'


def test_to_yes_no_boolean_known_answers():
    # server/webhookDispatcher.js:60-77
    yes = ["yes", "Yes", "YES", " yes ", "\tyEs\n", "\ufeffyes", "\u00a0yes\u3000"]
    no = ["no", "", "y", "yes!", "y e s", "ＹＥＳ", "yeſ", "true", "1", None]
    for v in yes:
        assert po.to_yes_no_boolean(v) is True, repr(v)
    for v in no:
        assert po.to_yes_no_boolean(v) is False, repr(v)
    assert po.to_yes_no_boolean(True) is True and po.to_yes_no_boolean(2) is True
    assert po.to_yes_no_boolean(0) is False and po.to_yes_no_boolean(float("nan")) is False


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c_payload_rows_match_python_oracle_and_json_module(built, seed):
    table = synth_archive(150, seed=seed)
    shows = table_to_shows(table)
    offsets, data = oracle_c.payload_rows(table)
    blob, o = bytes(data.numpy()), offsets.tolist()
    e = 0
    for show in shows:
        for entry in show["entries"]:
            want = po.archive_entry_payload_json(show, entry)
            assert blob[o[e]:o[e + 1]].decode("utf-8") == want + "\n", e
            obj = po.build_archive_entry_payload(show, entry)
            assert json.dumps(obj, ensure_ascii=False, separators=(",", ":")) == want
            e += 1
    assert e == table.n_entries and o[-1] == len(blob)
    # threaded C variant (the CPU baseline) gives the same bytes
    off2, data2 = oracle_c.payload_rows(table, nthreads=4)
    assert off2.tolist() == o and bytes(data2.numpy()) == blob


def test_c_payload_edge_rows(built):
    shows = [{"id": "a", "date": "d\"q", "time": "t\\b", "label": "l\nf", "leadPilot": "\x01", "monkeyLead": "\x1f\x7f",
              "entries": [{"operator": "ü", "unitId": "\t", "planned": " YES ", "launched": "no", "commandRx": "yes!",
                           "primaryIssue": "\r", "subIssue": "\x0c\x08"},
                          {}]},
             {"id": "empty", "entries": []}, None]
    table = pack_shows(shows)
    offsets, data = oracle_c.payload_rows(table)
    blob, o = bytes(data.numpy()), offsets.tolist()
    rows = [blob[o[i]:o[i + 1] - 1].decode("utf-8") for i in range(table.n_entries)]
    assert rows == [po.archive_entry_payload_json(shows[0], e) for e in shows[0]["entries"]]
    assert rows[0] == ('{"showDate":"d\\"q","showTime":"t\\\\b","showNumber":"l\\nf","leadPilot":"\\u0001",'
                       '"monkeyLead":"\\u001f\x7f","operator":"ü","monkeyId":"\\t","planned":true,"launched":false,'
                       '"commandReceived":false,"primaryIssue":"\\r","subIssue":"\\f\\b"}')
    assert rows[1] == ('{"showDate":"d\\"q","showTime":"t\\\\b","showNumber":"l\\nf","leadPilot":"\\u0001",'
                       '"monkeyLead":"\\u001f\x7f","operator":"","monkeyId":"","planned":false,"launched":false,'
                       '"commandReceived":false,"primaryIssue":"","subIssue":""}')


def test_show_payload_oracle_on_the_fixture_and_its_rules():
    """The schemaVersion 2 payload (reference server/webhookDispatcher.js:460-496, :545-584): the hand-written body of
    the reference's fixture, and the rules that differ from the row builders (`?? null` against `|| ''`)."""
    import os

    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "webhook_fixture.json")))
    show = {**fx["show"], "entries": [fx["entry"]]}
    body = po.show_payload_json("show.updated", show, "2024-07-05T04:00:00.000Z", "http://127.0.0.1:4101/hook", "POST")
    assert body == fx["expected_show_payload_json"]
    parsed = json.loads(body)
    assert list(parsed) == ["event", "schemaVersion", "dispatchedAt", "target", "table", "csv", "message", "show", "entries"]
    assert parsed["table"]["rows"][0] == fx["expected_table_row"] and parsed["csv"]["rows"][0] == fx["expected_csv_row"]
    assert parsed["message"]["entries"][0] == fx["expected_message"]
    # buildShowSummary: texts `|| ''`, timestamps `?? null` (0 and false survive, undefined and null become null)
    s = po.build_show_summary({"id": 0, "label": None, "crew": "x", "createdAt": 0, "updatedAt": False, "archivedAt": None})
    assert s == {"id": "", "label": "", "date": "", "time": "", "crew": [], "leadPilot": "", "monkeyLead": "", "notes": "",
                 "createdAt": 0, "updatedAt": False, "archivedAt": None, "deletedAt": None}
    # normalizeEntryList: a missing / non-array actions becomes [], an existing key keeps its place, a new one goes last
    assert po.normalize_entry_list({"entries": [{"a": 1, "actions": "x", "b": 2}, {"a": 1}, None, 5]}) == [
        {"a": 1, "actions": [], "b": 2}, {"a": 1, "actions": []}, {"actions": []}, {"actions": []}]
    assert list(po.normalize_entry_list({"entries": [{"a": 1, "actions": "x", "b": 2}]})[0]) == ["a", "actions", "b"]
    assert po.normalize_entry_list({"entries": {"0": 1}}) == [] and po.normalize_entry_list(None) == []
    # meta: only a non-empty plain object is attached; a number that is not finite is null in JSON
    assert "meta" not in po.dispatch_show_payload("e", show, "t", "u", "m", {})
    assert "meta" not in po.dispatch_show_payload("e", show, "t", "u", "m", [1])
    assert po.dispatch_show_payload("e", show, "t", "u", "m", {"k": 1})["meta"] == {"k": 1}
    nan_show = {"entries": [{"delaySec": float("nan"), "status": "Abort"}, {"delaySec": float("inf")}]}
    p = json.loads(po.show_payload_json("e", nan_show, "t", "u", "m"))
    assert p["table"]["rows"][0][21] is None and p["entries"][1]["delaySec"] is None
    assert p["csv"]["rows"][0].split(",")[21] == "NaN" and p["csv"]["rows"][1].split(",")[21] == "Infinity"
