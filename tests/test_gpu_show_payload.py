"""The schemaVersion 2 show payload on the GPU (csrc/show_payload.cu through pie_show_payloads_dev) against the Python
restatement of dispatchShowEvent (reference server/webhookDispatcher.js:545-584): byte for byte on provider-normalised
shows, and through Python's json module as a third reading of the same text."""
import json
import os
import random

import pytest
import torch

import pie_oracle as po
from sph_pie_b200 import _lib, ops, webhook
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows

pytestmark = pytest.mark.gpu

ENTRY_KEYS = ["id", "ts", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail", "severity",
              "rootCause", "actions", "operator", "batteryId", "delaySec", "commandRx", "notes"]  # sqlProvider.js:386-408


def normalised(show):
    """The shape _normalizeShow / _normalizeEntry store (key order included)."""
    out = dict(show)
    out["entries"] = [{k: e.get(k, [] if k == "actions" else "" if k not in ("ts", "delaySec") else None) for k in ENTRY_KEYS}
                      for e in show.get("entries", [])]
    return out


def check(shows, event="show.updated", at="2024-07-05T04:00:00.000Z", url="https://hooks.example/pie?x=1&y=\"2\"", method="POST",
          meta=None, device="cuda:0"):
    bodies = webhook.showEventPayloadBodies(shows, event, at, url, method, meta, device)
    assert len(bodies) == len(shows)
    for show, body in zip(shows, bodies):
        want = po.show_payload_json(event, show, at, url, method, meta if meta is not None else po.UNDEFINED)
        assert body == want
        assert json.loads(body) == json.loads(want)
    return bodies


def test_fixture_body(cuda):
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "webhook_fixture.json")))
    show = {**fx["show"], "entries": [fx["entry"]]}
    body = webhook.showEventPayloadBodies([show], "show.updated", "2024-07-05T04:00:00.000Z", "http://127.0.0.1:4101/hook", "POST")[0]
    got, want = json.loads(body), json.loads(fx["expected_show_payload_json"])
    # everything the reference derives from the row builders and the summary is the hand-written body's ...
    for key in ("event", "schemaVersion", "dispatchedAt", "target", "table", "csv", "message", "show"):
        assert got[key] == want[key], key
    assert list(got) == list(want)
    # ... and `entries` are the stored entries in the provider's normalised shape (the fixture's entry is not stored
    # text: it lacks ts and the issue keys, which the table cannot tell from '')
    assert got["entries"] == [normalised(show)["entries"][0]]
    assert body == po.show_payload_json("show.updated", normalised(show), "2024-07-05T04:00:00.000Z", "http://127.0.0.1:4101/hook", "POST")


@pytest.mark.parametrize("n_shows,seed", [(1, 1), (300, 2), (2500, 3)])
def test_synthetic_archive_byte_for_byte(cuda, n_shows, seed):
    shows = [normalised(s) for s in table_to_shows(synth_archive(n_shows, seed=seed, missing_created_frac=0.1))]
    rng = random.Random(seed)
    for s in shows:  # the other two timestamps, and what `?? null` lets through
        s["updatedAt"] = rng.choice([None, 1704067200000.5, 0, True, False, 1e21, -0.0])
        if rng.random() < 0.3:
            s["deletedAt"] = rng.choice([None, 1704067200001.0, float("inf")])
    check(shows, meta={"automation": {"source": "daily-archive", "totalShows": n_shows, "showIndex": 0, "showId": None}})
    check(shows[:5])


def test_hostile_strings_and_edge_shapes(cuda):
    rng = random.Random(7)
    alphabet = ['"', ",", "\n", "\r", "\\", "\t", "\b", "\f", "\x00", "\x1f", "|", "é", "漢", "🚁", " ", "a", "B", "7", "'", "/", "{", "]"]
    t = lambda: "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
    shows = []
    for i in range(200):
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
    want_shows = [s if isinstance(s, dict) else {} for s in shows]
    bodies = webhook.showEventPayloadBodies(shows, "ev\"ent\n", "t", "u\\", "GET", None)
    for show, body in zip(want_shows, bodies):
        assert body == po.show_payload_json("ev\"ent\n", normalised(show), "t", "u\\", "GET")
        json.loads(body)
    assert webhook.showEventPayloadBodies([], "e", "t", "u", "m") == []


def test_summary_mirror_and_schema_error(cuda):
    show = {"id": "s", "label": "L", "crew": ["a", "b"], "createdAt": 5, "updatedAt": False}
    assert webhook.buildShowSummary(show) == po.build_show_summary(show)
    # a text in a time field cannot be reproduced from the table (it holds the kind, not the text): fails loudly
    table = pack_shows([{"id": "x", "createdAt": "2024-01-01"}]).to(cuda)
    head, tail = webhook.payload_frame("e", "t", "u", "m")
    with pytest.raises(_lib.SchemaError) as e:
        ops.show_payloads(table, head, tail)
    assert "show 0" in e.value.message


def test_shows_that_do_not_fit_the_stage(cuda):
    """The kernels stage a show in 6 KB of shared memory and keep the numbers of up to 32 entries (pie_show_payload.cuh): more
    entries than number slots, more bytes than the stage holds, both — emitted from global memory through the caller's
    view, the same documents.  (Also run on the CPU: tests/test_payload_oracles_cpu.py.)"""
    entry = lambda i, text="": {"id": "e%d" % i, "ts": 1704067200000.0 + i, "unitId": "u", "status": "Abort", "notes": text,
                                "actions": ["a", "b"] if i % 3 == 0 else [], "delaySec": i / 7 if i % 2 else None}
    shows = [{"id": "33 entries", "entries": [entry(i) for i in range(33)]},
             {"id": "32 entries", "entries": [entry(i) for i in range(32)]},
             {"id": "long notes", "notes": "n" * 7000, "entries": [entry(0, "m" * 100)]},
             {"id": "many and long", "entries": [entry(i, "z" * 90) for i in range(70)]},
             {"id": "just fits?", "entries": [entry(i, "y" * 150) for i in range(14)]},
             {"id": "small", "entries": [entry(1)]}]
    check([normalised(s) for s in shows])
