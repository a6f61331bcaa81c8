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


# ---- the DEVICE code of the show payload, run on the CPU (tests/native/payload_host.cpp) --------------------------------
ENTRY_KEYS = ["id", "ts", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail", "severity",
              "rootCause", "actions", "operator", "batteryId", "delaySec", "commandRx", "notes"]  # sqlProvider.js:386-408


def normalised(show):
    """The shape _normalizeShow / _normalizeEntry store (key order included)."""
    out = dict(show)
    out["entries"] = [{k: e.get(k, [] if k == "actions" else "" if k not in ("ts", "delaySec") else None) for k in ENTRY_KEYS}
                      for e in show.get("entries", [])]
    return out


_payload_host = None


def payload_host_bodies(shows, event, at, url, method, meta=None, use_stage=1):
    """(bodies, shows that went through the shared-memory stage, status) of pie_show_payload.cuh run on the CPU: its 32
    lanes as fibers, driven like show_payload.cu (measure, exclusive sum, write)."""
    import ctypes as C
    import subprocess

    import numpy as np
    from sph_pie_b200 import webhook
    from sph_pie_b200.columnar import pack_shows

    global _payload_host
    if _payload_host is None:
        here = os.path.dirname(os.path.abspath(__file__))
        src, so = os.path.join(here, "native", "payload_host.cpp"), os.path.join(here, "native", "libpayload_host.so")
        csrc = os.path.join(here, "..", "sph_pie_b200", "csrc")
        deps = [src, os.path.join(here, "native", "cuda_shim", "cuda_runtime.h"), os.path.join(here, "..", "include", "sph_pie_b200.h")]
        deps += [os.path.join(csrc, f) for f in ("pie_show_payload.cuh", "pie_device.cuh", "pie_numfmt.cuh", "ryu_tables.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(here, "native", "cuda_shim"),
                                   "-o", so, src])
        _payload_host = C.CDLL(so)
    lib = _payload_host
    table = pack_shows(shows)
    head, tail = webhook.payload_frame(event, at, url, method, meta)
    view = table.view()
    n = table.n_shows
    offs = np.zeros(n + 1, dtype=np.int64)
    staged, status = C.c_int64(0), np.zeros(2, dtype=np.int32)
    h, t = np.frombuffer(head or b"\0", dtype=np.uint8).copy(), np.frombuffer(tail or b"\0", dtype=np.uint8).copy()
    p = lambda a: C.c_void_p(a.ctypes.data)
    args = (C.byref(view), p(h), C.c_int(len(head)), p(t), C.c_int(len(tail)), p(offs))
    rc = lib.payload_host(*args, None, C.c_uint64(0), C.c_int(use_stage), C.byref(staged), p(status))
    assert rc == 0, rc
    if status[0]:
        return None, staged.value, status.tolist()
    total, guard = int(offs[-1]), 64
    out = np.full(total + 2 * guard, 0xEE, dtype=np.uint8)
    rc = lib.payload_host(*args, C.c_void_p(out.ctypes.data + guard), C.c_uint64(total), C.c_int(use_stage), C.byref(staged), p(status))
    assert rc == 0, rc
    assert (out[:guard] == 0xEE).all() and (out[guard + total:] == 0xEE).all(), "bytes written outside the documents"
    blob, o = bytes(out[guard:guard + total]), offs.tolist()
    return [blob[o[i]:o[i + 1]].decode("utf-8") for i in range(n)], staged.value, status.tolist()


def _check_payload_host(shows, event="show.updated", at="2024-07-05T04:00:00.000Z", url="https://hooks.example/pie?x=1&y=\"2\"",
                        method="POST", meta=None, staged_at_least=None):
    want = [po.show_payload_json(event, normalised(s) if isinstance(s, dict) else {}, at, url, method,
                                 meta if meta is not None else po.UNDEFINED) for s in shows]
    for use_stage in (1, 0):
        bodies, staged, status = payload_host_bodies(shows, event, at, url, method, meta, use_stage)
        assert status == [0, -1]
        assert staged == 0 if not use_stage else staged >= (len(shows) if staged_at_least is None else staged_at_least)
        for i, (body, w) in enumerate(zip(bodies, want)):
            assert body == w, (use_stage, i)
    return want


def test_show_payload_device_code_on_the_cpu_synthetic_archive():
    import random

    shows = [normalised(s) for s in table_to_shows(synth_archive(80, seed=3, missing_created_frac=0.1))]
    rng = random.Random(3)
    for s in shows:
        s["updatedAt"] = rng.choice([None, 1704067200000.5, 0, True, False, 1e21, -0.0])
        if rng.random() < 0.3:
            s["deletedAt"] = rng.choice([None, 1704067200001.0, float("inf")])
    want = _check_payload_host(shows, meta={"automation": {"source": "daily-archive", "totalShows": 80, "showIndex": 0, "showId": None}})
    for w in want[:20]:
        json.loads(w)


def test_show_payload_device_code_on_the_cpu_hostile_strings_and_edge_shapes():
    import random

    rng = random.Random(7)
    alphabet = ['"', ",", "\n", "\r", "\\", "\t", "\b", "\f", "\x00", "\x1f", "|", "é", "漢", "🚁", " ", "a", "B", "7", "'", "/", "{", "]"]
    t = lambda: "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
    shows = []
    for i in range(60):
        show = {"id": t(), "date": t(), "time": t(), "label": t(), "crew": [t() for _ in range(rng.randrange(0, 4))],
                "leadPilot": t(), "monkeyLead": t(), "notes": t(), "createdAt": rng.choice([None, 1.5, 1e-7, 123456789012345680000.0]),
                "entries": []}
        for _ in range(rng.randrange(0, 5)):
            show["entries"].append({"id": t(), "ts": rng.choice([None, 0.0, 1704067200123.0]), "unitId": t(), "planned": t(),
                                    "launched": t(), "status": rng.choice(["Completed", "Abort", "completed", t()]),
                                    "primaryIssue": t(), "subIssue": t(), "otherDetail": t(), "severity": t(), "rootCause": t(),
                                    "actions": [t() for _ in range(rng.randrange(0, 3))], "operator": t(), "batteryId": t(),
                                    "delaySec": rng.choice([None, 0.0, -0.0, 12.5, 1e21, 1e-7, float("nan"), float("-inf"), 1 / 3]),
                                    "commandRx": t(), "notes": t()})
        shows.append(show)
    shows += [{"entries": []}, None, {"id": "only a show", "crew": []}]
    _check_payload_host(shows, "ev\"ent\n", "t", "u\\", "GET")
    # cells at the brims of the one-round paths (29 / 31 / 32 bytes): 27 .. 34 plain bytes, and the same with one byte that needs an escape
    brim = []
    for n in range(27, 35):
        for extra in ("", '"', ",", "\\", "\n", "é"):
            text = ("x" * n + extra)[:n] if not extra else "x" * (n - 1) + extra
            brim.append({"id": text, "label": text, "crew": [text], "entries": [{"id": text, "notes": text, "actions": [text], "status": "Completed",
                                                                                  "rootCause": text, "delaySec": 0.5},
                                                                                 {"unitId": text, "actions": [text, text], "delaySec": None}]})
    _check_payload_host(brim)


def test_show_payload_device_code_on_the_cpu_shows_that_do_not_fit_the_stage():
    """More entries than number slots (33), more bytes than the stage holds (6 KB), both: the same documents either way."""
    entry = lambda i, text="": {"id": "e%d" % i, "ts": 1704067200000.0 + i, "unitId": "u", "status": "Abort", "notes": text,
                                "actions": ["a", "b"] if i % 3 == 0 else [], "delaySec": i / 7 if i % 2 else None}
    shows = [{"id": "33 entries", "entries": [entry(i) for i in range(33)]},
             {"id": "32 entries", "entries": [entry(i) for i in range(32)]},
             {"id": "long notes", "notes": "n" * 7000, "entries": [entry(0, "m" * 100)]},
             {"id": "many and long", "entries": [entry(i, "z" * 90) for i in range(70)]},
             {"id": "just fits?", "entries": [entry(i, "y" * 150) for i in range(14)]},
             {"id": "small", "entries": [entry(1)]}]
    want = _check_payload_host(shows, staged_at_least=3)
    bodies, staged, _ = payload_host_bodies(shows, "show.updated", "2024-07-05T04:00:00.000Z", "https://hooks.example/pie?x=1&y=\"2\"", "POST")
    assert 3 <= staged < len(shows)  # the long ones were emitted from the caller's view
    assert bodies == want


def test_show_payload_device_code_on_the_cpu_schema_error():
    bodies, _, status = payload_host_bodies([{"id": "ok"}, {"id": "x", "createdAt": "2024-01-01"}], "e", "t", "u", "m")
    assert bodies is None and status == [_lib_codes().PIE_ERR_SCHEMA, 1]


def _lib_codes():
    from sph_pie_b200 import _lib

    return _lib
