"""Writes tests/golden/webhook_fixture.json.

The `show` and `entry` objects are the reference's only test fixture
(/root/reference/scripts/simulate-webhook.js:42-65).  The reference checks them only against its
own builders (:75-95), so it pins column ORDER and shape; it holds no expected values.  The
`expected_*` members below are therefore a hand derivation from reading
server/webhookDispatcher.js:276-342, computed with oracle/pie_oracle.py and reviewed by eye; they
are NOT output of the reference (no JS engine exists in the build image).
Run from the repo root:  python tests/golden/make_fixture.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pie_oracle as po  # noqa: E402

show = {
    "id": "simulation-show", "date": "2024-07-04", "time": "21:00", "label": "Independence Demo",
    "crew": ["Alex", "Nazar"], "leadPilot": "Alex", "monkeyLead": "Nazar", "notes": "Verification run",
}
entry = {
    "id": "entry-001", "unitId": "Drone-01", "planned": "Yes", "launched": "Yes", "status": "Completed",
    "actions": ["Logged only"], "operator": "Alex", "batteryId": "B-12", "delaySec": 0, "commandRx": "Yes",
    "notes": "Green across the board",
}
row = po.build_table_row(show, entry)
# written out by hand from the object literal at server/webhookDispatcher.js:316-329 (not computed):
PAYLOAD_JSON = ('{"showDate":"2024-07-04","showTime":"21:00","showNumber":"Independence Demo","leadPilot":"Alex",'
                '"monkeyLead":"Nazar","operator":"Alex","monkeyId":"Drone-01","planned":true,"launched":true,'
                '"commandReceived":true,"primaryIssue":"","subIssue":""}')
assert po.archive_entry_payload_json(show, entry) == PAYLOAD_JSON
# the text the provider would store for this show (JSON.stringify, sqlProvider.js:682), written out by hand:
STORED_TEXT = ('{"id":"simulation-show","date":"2024-07-04","time":"21:00","label":"Independence Demo",'
               '"crew":["Alex","Nazar"],"leadPilot":"Alex","monkeyLead":"Nazar","notes":"Verification run",'
               '"entries":[{"id":"entry-001","unitId":"Drone-01","planned":"Yes","launched":"Yes","status":"Completed",'
               '"actions":["Logged only"],"operator":"Alex","batteryId":"B-12","delaySec":0,"commandRx":"Yes",'
               '"notes":"Green across the board"}]}')
assert po.js_json_stringify({**show, "entries": [entry]}) == STORED_TEXT
assert po.map_archive_row(STORED_TEXT) == {**show, "entries": [entry]}
# the schemaVersion 2 body dispatchShowEvent('show.updated', {...show, entries: [entry]}) would post
# (server/webhookDispatcher.js:545-584), written out by hand from the object literals there:
COLS = ('["showId","showDate","showTime","showLabel","crew","leadPilot","monkeyLead","showNotes","entryId","unitId","planned",'
        '"launched","status","primaryIssue","subIssue","otherDetail","severity","rootCause","actions","operator","batteryId",'
        '"delaySec","commandRx","notes"]')
SUMMARY = ('{"id":"simulation-show","label":"Independence Demo","date":"2024-07-04","time":"21:00","crew":["Alex","Nazar"],'
           '"leadPilot":"Alex","monkeyLead":"Nazar","notes":"Verification run","createdAt":null,"updatedAt":null,'
           '"archivedAt":null,"deletedAt":null}')
SHOW_PAYLOAD_JSON = (
    '{"event":"show.updated","schemaVersion":2,"dispatchedAt":"2024-07-05T04:00:00.000Z",'
    '"target":{"url":"http://127.0.0.1:4101/hook","method":"POST"},'
    '"table":{"columns":' + COLS + ',"rows":[["simulation-show","2024-07-04","21:00","Independence Demo","Alex|Nazar","Alex",'
    '"Nazar","Verification run","entry-001","Drone-01","Yes","Yes","Completed","","","","","","Logged only","Alex","B-12",0,"Yes",'
    '"Green across the board"]]},'
    '"csv":{"header":' + COLS + ',"rows":["simulation-show,2024-07-04,21:00,Independence Demo,Alex|Nazar,Alex,Nazar,'
    'Verification run,entry-001,Drone-01,Yes,Yes,Completed,,,,,,Logged only,Alex,B-12,0,Yes,Green across the board"]},'
    '"message":{"show":' + SUMMARY + ',"entries":[{"showId":"simulation-show","showDate":"2024-07-04","showTime":"21:00",'
    '"showLabel":"Independence Demo","crew":"Alex|Nazar","leadPilot":"Alex","monkeyLead":"Nazar","showNotes":"Verification run",'
    '"entryId":"entry-001","unitId":"Drone-01","planned":"Yes","launched":"Yes","status":"Completed","primaryIssue":"",'
    '"subIssue":"","otherDetail":"","severity":"","rootCause":"","actions":"Logged only","operator":"Alex","batteryId":"B-12",'
    '"delaySec":0,"commandRx":"Yes","notes":"Green across the board"}]},'
    '"show":' + SUMMARY + ','
    '"entries":[{"id":"entry-001","unitId":"Drone-01","planned":"Yes","launched":"Yes","status":"Completed",'
    '"actions":["Logged only"],"operator":"Alex","batteryId":"B-12","delaySec":0,"commandRx":"Yes",'
    '"notes":"Green across the board"}]}')
assert po.show_payload_json("show.updated", {**show, "entries": [entry]}, "2024-07-05T04:00:00.000Z",
                            "http://127.0.0.1:4101/hook", "POST") == SHOW_PAYLOAD_JSON
assert json.loads(SHOW_PAYLOAD_JSON)["schemaVersion"] == 2
doc = {
    "source": "scripts/simulate-webhook.js:42-65 (show, entry); expected_* hand-derived, see make_fixture.py",
    "export_columns": po.EXPORT_COLUMNS,
    "show": show,
    "entry": entry,
    "expected_table_row": [row[c] for c in po.EXPORT_COLUMNS],
    "expected_message": po.build_message_payload(row),
    "expected_csv_row": po.build_csv_row(row),
    "expected_archive_entry_payload": po.build_archive_entry_payload(show, entry),
    "expected_archive_payload_json": PAYLOAD_JSON,
    "stored_text": STORED_TEXT,
    "expected_show_payload_json": SHOW_PAYLOAD_JSON,
    "expected_show_stats": po.compute_archive_show_stats({**show, "entries": [entry]}),
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "webhook_fixture.json"), "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps(doc, indent=1))
