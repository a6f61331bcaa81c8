"""The ingest kernels' document walk (sph_pie_b200/csrc/pie_json_walk.cuh), built for the HOST by
tests/native/ingest_host.cpp, against the oracle: pie_oracle.map_archive_row (JSON.parse as _mapArchiveRow applies it,
reference server/storage/sqlProvider.js:892-926; Python's json module is the parser) + the table packer.  Bit-exact
tables.  The same cases run on the GPU in tests/test_gpu_ingest.py; this file needs no GPU."""
import json
import os
import random

import numpy as np
import pytest
import torch

import oracle_c
import pie_oracle as po
from ingest_helpers import (ROUTE_FAST, ROUTE_FAST_BIG, ROUTE_LONG, ROUTE_RECORDS, ROUTE_SLOW, assert_tables_equal, fast_host_ingest,
                            host_ingest, oracle_ingest, stored_doc)
from sph_pie_b200 import _lib
from sph_pie_b200.synth import synth_archive, table_to_shows

HOSTILE = ['"', "\\", "/", "\b", "\f", "\n", "\r", "\t", "\x00", "\x1f", "\x7f", "é", "ß", "漢", "字", "🚁", "𝄞", " ",
           "﻿", "a", "Z", " ", ",", ":", "{", "}", "[", "]", "\\u0041", "\\n", "null", "true", "1e5"]


def hostile_text(rng, n):
    return "".join(rng.choice(HOSTILE) for _ in range(n))


def hostile_show(rng, n_entries):
    t = lambda: hostile_text(rng, rng.randrange(0, 12))
    show = {"id": t(), "date": "2024-03-0%d" % rng.randrange(1, 9), "time": t(), "label": t(), "showNumber": rng.choice([None, 3.0, 1e21]),
            "calendarEventId": t(), "eventName": t(), "crew": [t() for _ in range(rng.randrange(0, 4))], "leadPilot": t(),
            "monkeyLead": t(), "notes": t(), "disciplineId": "drones", "entries": [],
            "createdAt": rng.choice([1704067200000.0, 1.5, -0.0, 1e-7, 123456789012345680000.0, None]),
            "updatedAt": 1704067200001.0, "archivedAt": rng.choice([1704067200000.5, None, 4.9e-324])}
    for _ in range(n_entries):
        show["entries"].append({
            "id": t(), "ts": rng.choice([1704067200123.0, None, 0.0]), "unitId": t(), "planned": rng.choice(["Yes", "No", ""]),
            "launched": t(), "status": rng.choice(["Completed", "Abort", t()]), "primaryIssue": t(), "subIssue": t(),
            "otherDetail": t(), "severity": t(), "rootCause": t(), "actions": [t() for _ in range(rng.randrange(0, 3))],
            "operator": t(), "batteryId": t(),
            "delaySec": rng.choice([None, 0.0, 12.5, -3.0, 1e300, 5e-324, 0.1, 123456.789, 2.0 ** 53, 1 / 3]),
            "commandRx": t(), "notes": t()})
    return show


def check(docs, what=""):
    """Four implementations on the same texts: the Python oracle (json module + packer), the C oracle (recursive
    descent + strtod, oracle/pie_oracle.c), the kernels' thread-per-document walk built for the host, and the kernels'
    warp-per-document path built for the host (its 32 lanes as fibers; with the records its first pass leaves for the
    second, and without them, when the second pass parses again)."""
    ref_table, ref_status = oracle_ingest(docs)
    table, status, err = host_ingest(docs)
    assert err == (0, -1), (what, err)
    assert np.array_equal(status, ref_status), f"{what} doc_status"
    assert_tables_equal(table, ref_table, what)
    for pool in (288, 0):
        ftable, fstatus, ferr, _ = fast_host_ingest(docs, pool_units_per_doc=pool)
        assert ferr == (0, -1), (what, ferr)
        assert np.array_equal(fstatus, ref_status), f"{what} doc_status (warp path, pool {pool})"
        assert_tables_equal(ftable, ref_table, f"{what} warp path, pool {pool}")
    for threads in (1, 3):
        ctable, cstatus, cerr = oracle_c.ingest(docs, nthreads=threads)
        assert cerr == (0, -1), (what, cerr)
        assert np.array_equal(cstatus, ref_status), f"{what} doc_status (C oracle)"
        assert_tables_equal(ctable, ref_table, what + " C oracle")
    return table


@pytest.mark.parametrize("style", ["stringify", "ascii", "pretty", "shuffled"])
def test_synthetic_archive_round_trips(style):
    rng = random.Random(3)
    host = synth_archive(310, seed=11, missing_created_frac=0.1)
    shows = table_to_shows(host)
    docs = [stored_doc(s, rng, style) for s in shows]
    table = check(docs, style)
    # and it is the table the documents came from, except what JSON cannot carry: a NaN / Infinity delaySec is
    # written as null (JSON.stringify), i.e. comes back absent
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    if style == "stringify":
        host.delay_sec[host.delay_sec == 0] = 0.0  # JSON.stringify(-0) is "0"
    assert_tables_equal(table, host, style + " vs the table the documents came from")


@pytest.mark.parametrize("style", ["stringify", "ascii", "pretty", "shuffled"])
def test_hostile_strings_and_numbers(style):
    rng = random.Random(5)
    shows = [hostile_show(rng, rng.randrange(0, 6)) for _ in range(200)]
    check([stored_doc(s, rng, style) for s in shows], style)


def test_projection_rules():
    docs = [
        '{}', '[]', '[1,{"id":"x"}]', 'null', 'true', '12', '"text"', '', '   ', '{"id":"a"} ', ' \t\r\n{"id":"b"}\n',
        '{"id":null,"date":null,"crew":null,"entries":null,"createdAt":null}',
        '{"crew":{"0":"a"},"entries":{"length":1},"createdAt":"1704067200000","archivedAt":true}',
        '{"crew":["a",null,"b"],"entries":[null,1,"x",[{"id":"no"}],{},true,{"id":"yes","actions":[null,"go",""]}]}',
        '{"entries":[{"delaySec":null},{"delaySec":0},{"delaySec":-0},{"delaySec":1e400},{"delaySec":-1e400},{"ts":1e400}]}',
        '{"unknown":{"id":"inner","entries":[{"id":"deep"}]},"id":"outer","more":[[[],[{}]],{"a":{"b":[1,2,{"c":null}]}}]}',
        '{"entries":[{"actions":"not a list","ts":"12"},{"actions":{"a":1},"ts":false},{"actions":[]}]}',
        '{"id":"\\ud83d\\ude81 \\u00e9\\u6f22 \\"q\\" \\\\ \\/ \\b\\f\\n\\r\\t \\u0000"}',
        '{"i\\u0064":"escaped key","\\u0065ntries":[{"\\u0069d":"e1"}]}',
        '{"ID":"case matters","Id":"x","id ":"y"," id":"z","entries ":[{}]}',
        '{"createdAt":1704067200000,"archivedAt":1.7040672e12,"entries":[{"ts":1704067200000.5,"delaySec":12.5}]}',
        '{"createdAt":0.1e1,"archivedAt":-0,"entries":[{"delaySec":1E2},{"delaySec":1e+2},{"delaySec":100e-2}]}',
        '{"a":1,}', '{"a" 1}', '{"a":1 "b":2}', '{,}', '[,]', '[1,]', '{"a"}', '{"a":}', '{a:1}', "{'a':1}", '{"a":01}', '{"a":1.}',
        '{"a":.5}', '{"a":+1}', '{"a":1e}', '{"a":0x10}', '{"a":tru}', '{"a":nul}', '{"a":True}', '{"a":NaN}', '{"a":Infinity}',
        '{"a":-Infinity}', '{"a":-}', '{"a":"\\x41"}', '{"a":"\\u12"}', '{"a":"\\u12G4"}', '{"a":"tab\there"}', '{"a":"nl\nhere"}',
        '{"a":"unterminated}', '{"a":"x"', '{"a":"x"}}', '{"a":"x"}]', '{"a":"x"} {}', '{"a":"x"}x', '}{', ']', '{]', '[}', '{"a":[}',
        '{"a":{]}', '[1 2]', '{"a":1:2}', '{"a"::1}', '{"a":1,,"b":2}', '﻿{"id":"bom"}', '{"a":"\\"}', '{"a":"\\',
        '{"id":"a"}\x00', '\x00{"id":"a"}', '{"id":"a",\x0b"date":"b"}', '{"id":"a",\xa0"date":"b"}',
        '{"id":"del \x7f ok"}', '{"entries":[{"id":"1"},{"id":"2"}],"id":"after entries","crew":["z"]}',
        '1e5', '-', '-0', '0.0', '"a" "b"', 'nulll', 'null null',
    ]
    check(docs)


def test_every_prefix_and_single_byte_damage_of_a_document():
    rng = random.Random(9)
    show = hostile_show(rng, 2)
    show["label"] = 'q"\\\né\U0001F681'
    doc = stored_doc(show, rng, "ascii")  # pure ASCII: any cut / substitution leaves valid UTF-8
    docs = [doc[:k] for k in range(len(doc) + 1)]
    for _ in range(1500):
        k = rng.randrange(len(doc))
        docs.append(doc[:k] + rng.choice('"\\{}[]:,0-9.eE+tfn ux\n\t\x01a') + doc[k + 1:])
    for _ in range(500):
        k = rng.randrange(len(doc))
        docs.append(doc[:k] + doc[k + 1:])
    # a damaged document may also be a schema / duplicate-key case: those are compared one by one below
    keep, special = [], []
    for d in docs:
        try:
            oracle_ingest([d])
            keep.append(d)
        except (TypeError, po.UnsupportedJson):
            special.append(d)
    check(keep)
    for d in special:
        _, _, err = host_ingest([d])
        assert err[0] in (_lib.PIE_ERR_SCHEMA, _lib.PIE_ERR_UNSUPPORTED_JSON) and err[1] == 0, (d, err)
        assert oracle_c.ingest([d])[2][1] == 0, d


SCHEMA_DOCS = [
    '{"id":5}', '{"label":true}', '{"notes":{}}', '{"date":[]}', '{"crew":[1]}', '{"crew":[{}]}', '{"crew":["a",false]}',
    '{"entries":[{"status":1}]}', '{"entries":[{"delaySec":"12"}]}', '{"entries":[{"delaySec":true}]}', '{"entries":[{"delaySec":[]}]}',
    '{"entries":[{"delaySec":{}}]}', '{"entries":[{"actions":[1]}]}', '{"entries":[{"actions":[[]]}]}', '{"entries":[{"notes":false}]}',
    '{"id":"\\ud800"}', '{"id":"\\udc00x"}', '{"id":"\\ud800\\u0041"}', '{"entries":[{"id":"\\ud83dx\\ude81"}]}', '{"crew":["\\udfff"]}',
]
UNSUPPORTED_DOCS = [
    '{"id":"a","id":"b"}', '{"entries":[],"entries":[]}', '{"entries":[{"ts":1,"ts":2}]}', '{"entries":[{"id":"a","\\u0069d":"b"}]}',
    "[" * 65 + "]" * 65, '{"x":' + "[" * 64 + "]" * 64 + "}",
    b'{"id":"\xff"}', b'{"id":"\xc0\xaf"}', b'{"id":"\xe2\x82"}', b'{"id":"\xed\xa0\x80"}', b'{"x":"\xf4\x90\x80\x80"}', b'{"\x80":1}',
]


def test_warp_path_on_the_cpu_routes_and_damage():
    """The device code of the warp-per-document path, run on the CPU: which documents it takes (by route), and that
    whatever it takes of damaged documents of the provider's shape is what the oracle makes of them."""
    from sph_pie_b200.synth import synth_archive, table_to_shows
    import torch

    def canonical(n, seed, max_entries=21):
        host = synth_archive(n, seed=seed, missing_created_frac=0.1, max_entries=max_entries)
        lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
        host.delay_valid[lost] = 0
        return table_to_shows(host)

    rng = random.Random(12)
    shows = canonical(60, 31, max_entries=70)
    docs = [stored_doc(s, rng, "stringify") for s in shows]
    sizes = [len(d.encode()) for d in docs]
    _, _, err, routes = fast_host_ingest(docs)
    assert err == (0, -1)
    for d, n, sh, r in zip(docs, sizes, shows, routes.tolist()[:-1]):  # (the last document is always parsed twice)
        if n <= 8000:
            assert r == ROUTE_RECORDS, (n, r)
        elif 9500 <= n <= 15500 and len(sh["entries"]) <= 63:
            assert r in (ROUTE_RECORDS, ROUTE_FAST_BIG), (n, r)
        elif n > 16500:
            assert r == ROUTE_LONG, (n, r)
    assert routes[-1] in (ROUTE_FAST, ROUTE_FAST_BIG, ROUTE_LONG, ROUTE_SLOW)
    check(docs, "canonical, long documents among them")
    pretty = [stored_doc(s, rng, "pretty") for s in shows[:8]]
    assert set(fast_host_ingest(pretty)[3].tolist()) <= {ROUTE_SLOW, ROUTE_LONG}
    # with a pool too small for all of them, the later documents are parsed again by the second pass
    _, _, _, routes = fast_host_ingest(docs[:20], pool_units_per_doc=60)
    assert ROUTE_RECORDS in routes.tolist() and (ROUTE_FAST in routes.tolist() or ROUTE_FAST_BIG in routes.tolist())
    # damage: every prefix of one document, single-character edits of several
    small = [d for d, n in zip(docs, sizes) if n < 2500]
    damaged = []
    if small:
        doc = small[0]
        damaged += [doc[:k] for k in range(len(doc) + 1)]
    alphabet = '"\\{}[]:,0-9.eE+tfn ux\n\t\x01a\u00e9'
    for doc in [d for d, n in zip(docs, sizes) if n < 9000][:12]:
        for _ in range(120):
            k = rng.randrange(len(doc))
            damaged.append(doc[:k] + rng.choice(alphabet) + doc[k + 1:])
        for _ in range(20):
            k = rng.randrange(len(doc))
            damaged.append(doc[:k] + doc[k + 1:])
            damaged.append(doc[:k] + rng.choice(alphabet) + doc[k:])
    keep = []
    for d in damaged:
        try:
            oracle_ingest([d])
            keep.append(d)
        except (TypeError, po.UnsupportedJson):
            _, _, err, _ = fast_host_ingest([d])
            assert err[0] in (_lib.PIE_ERR_SCHEMA, _lib.PIE_ERR_UNSUPPORTED_JSON), (d[:80], err)
    assert len(keep) > 1500
    ref_table, ref_status = oracle_ingest(keep)
    table, status, err, routes = fast_host_ingest(keep)
    assert err == (0, -1)
    assert np.array_equal(status, ref_status)
    assert_tables_equal(table, ref_table, "damaged canonical documents, warp path on the CPU")
    assert (routes == ROUTE_RECORDS).sum() > 100  # damage inside a value leaves the shape alone: the warp path decides those


def structural_variant_texts(seed=77, archive_seed=41):
    """Documents of the provider's shape with one structural change each (see the test below; scripts/fuzz_ingest_cpu.py
    runs it with other seeds)."""
    import copy

    rng = random.Random(seed)
    host = synth_archive(30, seed=archive_seed, missing_created_frac=0.1, max_entries=6)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    shows = [s for s in table_to_shows(host) if s["entries"]]
    other_values = [0, -1.5, 1e21, 12345678901234567890123, None, True, False, [], {}, ["a"], [1], [None], {"a": 1}, {"a": {"b": []}}, "",
                    "text", "\u00e9\"\\", [[]], [{}], ["a", ["b"]], 1e-7, "1700000000000", "2024-01-01T00:00:00Z"]

    def variants(show):
        out = []
        for _ in range(40):
            sh = copy.deepcopy(show)
            target = sh if rng.random() < 0.35 else rng.choice(sh["entries"])
            keys = list(target)
            what = rng.randrange(8)
            if what == 0:
                del target[rng.choice(keys)]
            elif what == 1:
                target[rng.choice(keys)] = copy.deepcopy(rng.choice(other_values))
            elif what == 2:
                target[rng.choice(["extra", "x", "showNumber", "Notes", "iD", "delaysec", "actions2"])] = copy.deepcopy(rng.choice(other_values))
            elif what == 3:
                rng.shuffle(keys)
                for k in keys:
                    target[k] = target.pop(k)
            elif what == 4:
                k = rng.choice(keys)
                target[k + rng.choice(["", "_", "é"])[:1] or "k"] = target.pop(k)
            elif what == 5:
                sh["entries"] = rng.choice([[], None, {}, "x", 3, [{}], [None], [1, {}], sh["entries"] + [{}], [[]], sh["entries"] * 2])
            elif what == 6:
                sh["crew"] = rng.choice([None, [], ["A"], ["A", None], [1], "x", {}, [["A"]], ["A", "B", "C"] * 3])
            else:
                e = rng.choice(sh["entries"])
                e["actions"] = rng.choice([None, [], ["a", None], [1], "x", {}, [["a"]], ["a"] * 7, [{"a": 1}]])
            out.append(sh)
        return out

    texts = []
    for show in shows[:12]:
        for v in variants(show):
            texts.append(po.js_json_stringify(v))
            # a member twice: splice a copy of the text of one member behind itself
            if rng.random() < 0.15:
                t = texts[-1]
                k = t.find('"status":"')
                if k > 0:
                    end = t.find(",", k)
                    texts.append(t[:end + 1] + t[k:end + 1] + t[end + 1:])
    return texts


def test_structural_variants_of_the_providers_documents():
    """Documents of the provider's shape with ONE structural change each — a key missing, twice, renamed, reordered; a value
    of another type (number, null, boolean, list, object) where text / a number / a list belongs; nesting under a foreign
    key; lists with other things than strings; empty objects and lists — through all four implementations.  This is where
    a recogniser that decides only one shape could accept what it must not: whatever the warp path takes must be what
    the oracle makes of it, and what the oracle refuses must fail the same way on both of the kernels' paths."""
    texts = structural_variant_texts()
    assert len(texts) > 450
    keep, refused = [], 0
    for t in texts:
        try:
            oracle_ingest([t])
            keep.append(t)
        except (TypeError, po.UnsupportedJson):
            refused += 1
            walk_err = host_ingest([t])[2]
            warp_err = fast_host_ingest([t])[2]
            assert walk_err[0] in (_lib.PIE_ERR_SCHEMA, _lib.PIE_ERR_UNSUPPORTED_JSON), (t[:120], walk_err)
            assert warp_err == walk_err, (t[:120], warp_err, walk_err)
    assert refused > 30 and len(keep) > 250
    check(keep, "structural variants")
    routes = fast_host_ingest(keep)[3]
    assert (routes == ROUTE_RECORDS).sum() > 20 and (routes == ROUTE_SLOW).sum() > 50  # both decide their share


def test_warp_path_at_the_brims_of_its_lists():
    """Documents that fill a list of the warp path exactly, and by one more: quotes (1472 / 3584), members (480 / 1152),
    numbers (64 / 128), lists with elements (32 / 64), entries (63), bytes (16 KB).  One over the small configuration's
    brim is the roomy one's document, one over that the walk's — and the table is the oracle's every time."""
    entry_keys = ["id", "ts", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail", "severity",
                  "rootCause", "actions", "operator", "batteryId", "delaySec", "commandRx", "notes"]

    def entry(i, actions=(), text=""):
        e = {k: text for k in entry_keys}
        e.update(id="e%d" % i if text is not None else None, ts=1700000000000 + i if text is not None else i, delaySec=None,
                 actions=list(actions))
        return e

    docs, expect = [], []

    def add(show, routes):
        docs.append(po.js_json_stringify(show))
        expect.append(routes)

    both = (ROUTE_RECORDS, ROUTE_FAST, ROUTE_FAST_BIG)
    # numbers: unknown keys with numeric values at the show's level
    for n, routes in ((64, (ROUTE_RECORDS, ROUTE_FAST)), (65, both), (128, both), (129, (ROUTE_SLOW,))):
        add({("n%d" % i): i for i in range(n)}, routes)
    # strings: 2 quotes for a key, 2 for a string value
    for strings, routes in ((736, (ROUTE_RECORDS, ROUTE_FAST)), (737, both), (1792, both), (1793, (ROUTE_SLOW,))):
        show = {("k%d" % i): "v" for i in range(strings // 2)}
        if strings % 2:
            show["odd"] = 1
        add(show, routes)
    # members
    for members, routes in ((480, (ROUTE_RECORDS, ROUTE_FAST)), (481, both), (1152, both), (1153, (ROUTE_SLOW,))):
        add({("m%d" % i): None for i in range(members)}, routes)
    # entries, and entries whose actions hold an element
    for n, routes in ((20, (ROUTE_RECORDS, ROUTE_FAST)), (40, both), (57, both), (58, (ROUTE_SLOW,))):  # 62 quotes an entry
        add({"id": "s", "entries": [entry(i) for i in range(n)]}, routes)
    for n, routes in ((63, both), (64, (ROUTE_SLOW, ROUTE_LONG))):  # null texts: the entry index (or the 16 KB) is what runs out
        add({"id": "s", "entries": [entry(i, text=None) for i in range(n)]}, routes)
    # lists with elements: an entry is 17 members, so the members (480 / 1152) run out before the list of lists (32 / 64)
    # can: 28 entries are the small configuration's, 29 the roomy one's, and both still leave records
    for n, routes in ((28, (ROUTE_RECORDS,)), (29, (ROUTE_RECORDS,)), (40, (ROUTE_RECORDS,))):
        add({"id": "s", "entries": [entry(i, ["go"], text=None) for i in range(n)]}, routes)
    # bytes: one long text value up to the 16 KB a warp takes (less the 31 bytes its alignment may cost)
    for n, routes in ((8000, both), (16000, both), (16300, (ROUTE_RECORDS, ROUTE_FAST, ROUTE_FAST_BIG, ROUTE_LONG)), (16500, (ROUTE_LONG,))):
        add({"id": "s", "notes": "x" * (n - 22)}, routes)
    docs.append('{"id":"the last document is parsed twice"}')
    expect.append((ROUTE_FAST,))
    got = fast_host_ingest(docs, pool_units_per_doc=4000)[3].tolist()  # (a pool that does not run out: that is another test)
    for d, r, want in zip(docs, got, expect):
        assert r in want, (len(d), d[:60], r, want)
    check(docs, "brims")
    # and each of them alone, i.e. as the last document of its batch (no records: the second pass parses it again)
    for d in docs[:-1]:
        ref_table, ref_status = oracle_ingest([d])
        table, status, err, routes = fast_host_ingest([d])
        assert err == (0, -1) and np.array_equal(status, ref_status)
        assert_tables_equal(table, ref_table, d[:40])


def test_warp_path_reads_only_the_words_that_hold_the_documents():
    """The warp path reads the text 32 aligned bytes at a time (the walk: 8).  Run on the CPU with a forbidden page right
    behind the aligned 32-byte word that holds the last byte of the text, and right before the one that holds the first
    byte, at several positions of that byte in its word: a read outside kills the (child) process.  A negative control
    proves that the forbidden pages are."""
    import subprocess
    import sys

    body = """
import sys, random, ctypes
sys.path[:0] = [%r, %r, %r]
import torch
from ingest_helpers import fast_host_ingest, oracle_ingest, assert_tables_equal, stored_doc, _guarded_text
from sph_pie_b200.synth import synth_archive, table_to_shows
import numpy as np
if sys.argv[1] == "control":
    m, text = _guarded_text(np.zeros(100, dtype=np.uint8), 100, "end", 0)
    ctypes.string_at(text.ctypes.data + 100, 1)  # the first byte of the forbidden page
    sys.exit(0)
rng = random.Random(3)
host = synth_archive(14, seed=5, missing_created_frac=0.1)
lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
host.delay_valid[lost] = 0
docs = [stored_doc(s, rng, "stringify") for s in table_to_shows(host)] + ['{"id":"x"}', ' {"id": "pretty"} ', '{"id":"\\u00e9"}']
ref, ref_status = oracle_ingest(docs)
for guard in ("end", "start"):
    for shift in (0, 1, 7, 8, 9, 15, 16, 24, 31):
        for pool in (288, 0):
            table, status, err, routes = fast_host_ingest(docs, pool_units_per_doc=pool, guard=guard, shift=shift)
            assert err == (0, -1) and np.array_equal(status, ref_status)
            assert_tables_equal(table, ref, guard)
print("ok")
""" % (os.path.dirname(os.path.abspath(__file__)), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"),
       os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    control = subprocess.run([sys.executable, "-c", body, "control"], capture_output=True, text=True)
    assert control.returncode < 0, "the forbidden page can be read: the guard does not guard"
    run = subprocess.run([sys.executable, "-c", body, "run"], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok"), (run.returncode, run.stderr[-2000:])


def test_schema_and_unsupported_documents_fail_loudly():
    good = '{"id":"fine","entries":[{"id":"e"}]}'
    for d in SCHEMA_DOCS:
        with pytest.raises(TypeError):
            oracle_ingest([good, d])
        assert fast_host_ingest([good, d, good, '{"id":7}'])[2] == host_ingest([good, d, good, '{"id":7}'])[2], d
        _, _, err = host_ingest([good, d, good, '{"id":7}'])
        assert err == (_lib.PIE_ERR_SCHEMA, 1), (d, err)
        assert oracle_c.ingest([good, d, good, '{"id":7}'])[2] == (_lib.PIE_ERR_SCHEMA, 1), d
    for d in UNSUPPORTED_DOCS:
        with pytest.raises(po.UnsupportedJson):
            oracle_ingest([good, good, d])
        assert fast_host_ingest([good, good, d, '{"id":7}'])[2] == host_ingest([good, good, d, '{"id":7}'])[2], d
        _, _, err = host_ingest([good, good, d, '{"id":7}'])
        assert err == (_lib.PIE_ERR_UNSUPPORTED_JSON, 2), (d, err)
        assert oracle_c.ingest([good, good, d, '{"id":7}'])[2] == (_lib.PIE_ERR_UNSUPPORTED_JSON, 2), d
    # not errors: the same things where the table does not look, nesting of exactly 64, a dropped row that also has them
    fine = ['{"x":{"id":5,"id":6},"y":[{"delaySec":"12"}]}', "[" * 64 + "]" * 64, '{"x":' + "[" * 63 + "]" * 63 + "}",
            '{"id":5', '{"id":"a","id":"b"', '{"showNumber":5,"updatedAt":"x","entries":[{"extra":{"status":1}}]}']
    check(fine)


def test_numbers_are_correctly_rounded():
    rng = np.random.default_rng(17)
    xs = rng.integers(0, 2 ** 64, 4000, dtype=np.uint64).view(np.float64)
    xs = xs[np.isfinite(xs)]
    texts = [repr(float(x)) for x in xs] + ["%.17e" % x for x in xs[:1500]] + [po.js_number_to_string(float(x)) for x in xs[:1500]]
    texts += ["0", "-0", "1e22", "1e23", "9007199254740993", "123456789012345678901234567890", "0.000001", "1E5", "5e-324",
              "2.4703282292062327e-324", "1.7976931348623157e308", "1.7976931348623159e308", "4.4501477170144023e-308"]
    docs = ['{"createdAt":%s,"entries":[{"delaySec":%s,"ts":%s},{"delaySec":%s}]}' % (a, a, a, b)
            for a, b in zip(texts, reversed(texts))]
    check(docs)


def test_empty_batch_and_ragged_documents():
    check([])
    check([""])
    check(["{}"] * 3)
    big = {"id": "big", "notes": "x" * 70000, "entries": [{"notes": "y" * 5000, "actions": ["a"] * 300}] * 40}
    docs = [json.dumps(big), "{}", json.dumps({"entries": [{}] * 1000}), "[]", json.dumps(big)[:-1], '{"id":"z"}']
    check(docs)


def test_reference_fixture_as_stored_text():
    """The reference's one fixture (scripts/simulate-webhook.js:42-65) as the provider would store it."""
    import os

    from sph_pie_b200.columnar import pack_shows

    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "webhook_fixture.json")))
    table = check([fx["stored_text"]], "fixture")
    assert_tables_equal(table, pack_shows([{**fx["show"], "entries": [fx["entry"]]}]), "fixture vs the objects")
    assert table.entry_cols["unit_id"].get(0) == "Drone-01" and table.crew.items.get(1) == "Nazar"
    assert float(table.delay_sec[0]) == 0.0 and int(table.delay_valid[0]) == 1


def random_json_value(rng, depth=0):
    """Any JSON value, biased towards the shapes the projection looks at (strings / numbers / null / arrays /
    objects with known and unknown keys in any position), with random whitespace."""
    ws = lambda: rng.choice(["", "", "", " ", "\n", "\t ", "\r\n  "])
    kind = rng.random()
    if depth > 5:
        kind = min(kind, 0.55)
    if kind < 0.25:
        return json.dumps(hostile_text(rng, rng.randrange(0, 8)), ensure_ascii=rng.random() < 0.5)
    if kind < 0.40:
        return rng.choice(["0", "-0", "12", "1.5", "-3e2", "1E+2", "0.000001", "1704067200000", "9007199254740993", "1e400", "5e-324"])
    if kind < 0.55:
        return rng.choice(["null", "true", "false"])
    if kind < 0.75:
        items = [random_json_value(rng, depth + 1) for _ in range(rng.randrange(0, 4))]
        return "[" + ws() + ("," + ws()).join(items) + ws() + "]"
    keys = list(po.SHOW_DOC_KEYS) + list(po.ENTRY_DOC_KEYS) + ["x", "showNumber", "", "é", "entries ", "ID"]
    members = []
    for k in rng.sample(keys, rng.randrange(0, 7)):  # sample: no key twice in one object
        members.append(json.dumps(k) + ws() + ":" + ws() + random_json_value(rng, depth + 1))
    return "{" + ws() + ("," + ws()).join(members) + ws() + "}"


def random_document(rng):
    """A show-shaped document whose every value is random JSON: most are schema errors, many are fine."""
    return random_json_value(rng, 0) if rng.random() < 0.3 else (
        "{" + ",".join(json.dumps(k) + ":" + (
            "[" + ",".join(random_json_value(rng, 2) for _ in range(rng.randrange(0, 4))) + "]" if k == "entries" and rng.random() < 0.7
            else random_json_value(rng, 1)) for k in rng.sample(list(po.SHOW_DOC_KEYS) + ["x", "y"], rng.randrange(0, 8))) + "}")


def classify(docs):
    """Split random documents by what the oracle says of them: fine / schema error / unsupported."""
    fine, schema, unsupported = [], [], []
    for d in docs:
        try:
            oracle_ingest([d])
            fine.append(d)
        except TypeError:
            schema.append(d)
        except po.UnsupportedJson:
            unsupported.append(d)
    return fine, schema, unsupported


def test_random_json_three_ways():
    rng = random.Random(2024)
    docs = [random_document(rng) for _ in range(6000)]
    fine, schema, unsupported = classify(docs)
    assert len(fine) > 1000 and len(schema) > 1000, (len(fine), len(schema), len(unsupported))
    check(fine, "random JSON")
    for d in schema[:1500]:
        assert host_ingest([d])[2] == (_lib.PIE_ERR_SCHEMA, 0), d
        assert oracle_c.ingest([d])[2] == (_lib.PIE_ERR_SCHEMA, 0), d
    for d in unsupported:  # (a document may also hold a schema error: which of the two is met first is not specified)
        assert host_ingest([d])[2] in ((_lib.PIE_ERR_UNSUPPORTED_JSON, 0), (_lib.PIE_ERR_SCHEMA, 0)), d
        assert oracle_c.ingest([d])[2] in ((_lib.PIE_ERR_UNSUPPORTED_JSON, 0), (_lib.PIE_ERR_SCHEMA, 0)), d


def test_row_timestamps_host_logic():
    """storage._row_timestamp (the product's host side of _getTimestamp, sqlProvider.js:970-985) against the oracle."""
    import math

    from sph_pie_b200 import storage

    values = [storage._MISSING, None, True, False, 5, 5.5, -0.0, "", "  ", " 12 ", "1e3", ".5", "5.", "0x1F", "0b11", "0o17", "+7",
              "-7.25", "1704067200000", "1e+21", float("inf"), float("nan"), "\ufeff42\u3000", "00012"]
    for v in values:
        got = storage._row_timestamp(v)
        want = po.get_timestamp(po.UNDEFINED if v is storage._MISSING else v)
        assert (want is None and math.isnan(got)) or got == want, (v, got, want)
    for v in ["abc", "Infinity", "-Infinity", "１２", "-0x1", "1_000", "12px", "0x", "1e", "2024-02-30", "01/02/2024",
              "2024-01-01 00:00:00", "Mon, 01 Jan 2024 00:00:00 GMT"]:
        with pytest.raises(NotImplementedError):
            storage._row_timestamp(v)
        with pytest.raises(NotImplementedError):
            po.get_timestamp_tz(v)
    # the Date.parse leg, for the one format ECMA-262 specifies (a date-only form is UTC, a date-time without offset
    # is local time): the product's host side against the oracle and against hand-derived values
    day = 1704067200000.0  # 2024-01-01T00:00:00Z
    for text, tz, want in [("2024-01-01", 0, day), ("2024-01-01", -480, day), ("2024-01-01T00:00:00.000Z", 330, day),
                           ("2024-01-01T00:00", 0, day), ("2024-01-01T00:00", -480, day + 480 * 60000),
                           ("2024-01-01T05:30:15.250+05:30", 0, day + 15250), ("2024-01-01T24:00", 0, day + 86400000),
                           ("2024-13-01", 0, None), ("2024-01-01T25:00", 0, None), ("2024-01-01T24:00:01", 0, None),
                           ("2024-01-01T00:00+24:00", 0, None)]:
        got = storage._row_timestamp(text, tz)
        assert po.get_timestamp_tz(text, tz) == want, text
        assert (want is None and math.isnan(got)) or got == want, (text, got, want)


def test_time_fields_and_their_kinds():
    """ABI 2: updatedAt / deletedAt and what each of the four time fields holds when it is not a finite number — the
    walker (built for the host) against the table packer on JSON.parse's values; a text is recorded by its place."""
    docs = [
        '{"id":"a","createdAt":1704067200000,"updatedAt":1704067200001.5,"archivedAt":null,"deletedAt":true}',
        '{"createdAt":null,"updatedAt":false,"deletedAt":null,"entries":[]}',
        '{"id":"c"}',
        '{"createdAt":1e999,"updatedAt":-1e999,"archivedAt":0,"deletedAt":-0}',
        '[1,2]', 'null', '{"updatedAt":5e-324,"deletedAt":123456789012345680000}',
    ]
    got, status, err = host_ingest(docs)
    assert err == (0, -1)
    ref, ref_status = oracle_ingest(docs)
    assert np.array_equal(status, ref_status)
    assert_tables_equal(got, ref)
    assert torch.equal(got.time_kind, ref.time_kind)
    for name in ("updated_at", "deleted_at"):
        a, b = getattr(got, name).numpy(), getattr(ref, name).numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.int64), b[~np.isnan(b)].view(np.int64))
    K = _lib
    assert got.time_kind.tolist()[0] == [K.TK_NUMBER, K.TK_NUMBER, K.TK_NULL, K.TK_TRUE]
    assert got.time_kind.tolist()[3] == [K.TK_NONFINITE, K.TK_NONFINITE, K.TK_NUMBER, K.TK_NUMBER]
    assert got.time_kind.tolist()[4] == [0, 0, 0, 0] and got.time_kind.tolist()[5] == [0, 0, 0, 0]

    # strings, arrays and objects: the kind, and for a string the byte offset of its text in the NaN's payload
    docs = ['{"createdAt":"1704067200000","updatedAt":"2024-01-01T00:00:00.000Z","archivedAt":[5],"deletedAt":{"a":1}}',
            '{"id":"x","deletedAt":" 12 ","createdAt":""}']
    got, status, err = host_ingest(docs)
    assert err == (0, -1) and status.tolist() == [0, 0]
    assert got.time_kind.tolist() == [[K.TK_STRING, K.TK_STRING, K.TK_OTHER, K.TK_OTHER], [K.TK_STRING, K.TK_ABSENT, K.TK_ABSENT, K.TK_STRING]]
    text = "".join(docs).encode()
    def text_at(x):
        bits = int(np.float64(x).view(np.int64)) & 0x7FFFFFFFFFFFF
        return text[bits:text.index(b'"', bits)].decode()
    assert np.isnan(got.created_at[0].item()) and text_at(got.created_at[0].item()) == "1704067200000"
    assert text_at(got.updated_at[0].item()) == "2024-01-01T00:00:00.000Z"
    assert text_at(got.deleted_at[1].item()) == " 12 " and text_at(got.created_at[1].item()) == ""
    # the packer records the same kinds (it has no text to point into: its NaNs carry no payload)
    ref, _ = oracle_ingest(docs)
    assert_tables_equal(got, ref)
    assert torch.equal(got.time_kind, ref.time_kind)
    # a time key twice is as undecided as any other known key twice
    _, _, err = host_ingest(['{"updatedAt":1,"updatedAt":2}'])
    assert err[0] == _lib.PIE_ERR_UNSUPPORTED_JSON


def test_archive_maintenance_oracle_rules():
    """The Python restatement of _archiveDailyShows' decision, _addMonths and _purgeExpiredArchives on hand-derived
    cases (reference server/storage/sqlProvider.js:758-816, :863-890, :991-1009)."""
    H = 3600 * 1000
    now = 1704067200000.0 + 12 * H  # 2024-01-01T12:00Z
    rows = ['{"date":"2024-01-01","createdAt":1704067200000}',            # exactly 12 h old: due (>=)
            '{"date":" 2024-01-01 ","createdAt":1704067200001}',          # same group after trim: due with it
            '{"date":"2024-01-02","createdAt":1704067200001}',            # 12 h minus 1 ms: not due
            '{"date":"2024-01-03","updatedAt":"1704067200000"}',          # createdAt absent -> updatedAt, numeric text
            '{"date":"2024-01-04"}',                                       # no timestamp at all: null -> 0 -> due
            '{"date":"2024-01-02","createdAt":null}',                      # Number(null) = 0: drags its group along
            'not json', '7', '{"createdAt":1}', '{"date":"  ","createdAt":9e15}', '[]']
    due, order = po.archive_daily_shows_decision(rows, now)
    assert due == [True, True, True, True, True, True, False, False, True, True, True]
    # groups in order of first appearance, rows in row order; '__undated__' holds rows 8, 9 and the array
    assert order == [0, 1, 2, 5, 3, 4, 8, 9, 10]
    # setMonth keeps the day of the month and carries an overflow: 31 Dec + 2 months = 2 Mar in a leap year
    d = lambda y, m, dd: float(po.days_from_civil(y, m, dd) * 86400000)
    assert po.add_months(d(2023, 12, 31), 2) == d(2024, 3, 2)
    assert po.add_months(d(2022, 12, 31), 2) == d(2023, 3, 3)
    assert po.add_months(d(2024, 1, 15) + 5.75, 2) == d(2024, 3, 15) + 5  # new Date() truncates the fraction
    assert po.add_months(d(2024, 11, 30), 2) == d(2025, 1, 30)
    assert po.add_months(-1.0, 2) == d(1970, 3, 4) - 1  # 1969-12-31T23:59:59.999 -> 31 Feb 1970 = 3 Mar, same time of day
    # local time: 2024-01-01T02:00Z is still 31 Dec, 18:00, at UTC-8 -> 31 Feb -> 2 Mar 18:00 local = 3 Mar 02:00Z
    assert po.add_months(d(2024, 1, 1) + 2 * H, 2, -480) == d(2024, 3, 3) + 2 * H
    assert po.add_months(8.64e15 + 1, 2) == 8.64e15 + 1 and np.isnan(po.add_months(8.64e15, 2))
    assert po.is_archive_expired(d(2024, 1, 1), d(2024, 3, 1)) and not po.is_archive_expired(d(2024, 1, 1), d(2024, 3, 1) - 1)
    rows = [{"data": '{"createdAt":%d}' % int(d(2024, 1, 1)), "created_at": "1"},   # the document's own field wins
            {"data": "{}", "created_at": str(int(d(2024, 1, 1)))},                  # else the row's column (a text)
            {"data": "broken", "created_at": None},                                  # Number(null) = 0: long expired
            {"data": "{}"},                                                          # neither: skipped
            {"data": '{"createdAt":"2024-01-01T00:00:00.000Z"}'}]                    # Date.parse
    assert po.purge_expired_archives_decision(rows, d(2024, 3, 1)) == [True, True, True, False, True]
    assert po.purge_expired_archives_decision(rows, d(2024, 3, 1) - 1) == [False, False, True, False, False]


def test_add_months_against_the_calendar_of_pythons_datetime():
    """_addMonths (sqlProvider.js:999-1009) restated with days_from_civil arithmetic, against an independent calendar:
    Python's datetime.  setMonth(getMonth() + 2) in a fixed-offset zone is "the first of the month two months on, plus
    (day of the month - 1) days, at the same local time of day" (MakeDay carries a day past the month's end over)."""
    import datetime as dt

    rng = random.Random(23)
    epoch = dt.datetime(1970, 1, 1)
    for _ in range(20000):
        tz = rng.choice([0, -480, 330, 765, -720, 60, -210])
        y, m = rng.randrange(1800, 2400), rng.randrange(1, 13)
        if rng.random() < 0.5:  # the last days of a month, where the overflow happens
            first_next = dt.datetime(y + m // 12, m % 12 + 1, 1)
            local = first_next - dt.timedelta(days=rng.randrange(1, 5), milliseconds=rng.randrange(0, 86400000))
        else:
            local = dt.datetime(y, m, 1) + dt.timedelta(milliseconds=rng.randrange(0, 31 * 86400000))
        m0 = local.month - 1 + 2
        first = dt.datetime(local.year + m0 // 12, m0 % 12 + 1, 1)
        want_local = first + dt.timedelta(days=local.day - 1, hours=local.hour, minutes=local.minute, seconds=local.second,
                                          microseconds=local.microsecond)
        to_ms = lambda t: (t - epoch) // dt.timedelta(milliseconds=1) - tz * 60000  # noqa: E731
        t, want = float(to_ms(local)), float(to_ms(want_local))
        assert po.add_months(t, 2, tz) == want, (local, tz, po.add_months(t, 2, tz), want)
        assert po.is_archive_expired(t, want, tz) and not po.is_archive_expired(t, want - 1, tz)


def test_date_parse_and_string_to_number_against_pythons_own():
    """The Date.parse leg of _getTimestamp (sqlProvider.js:978-982) restated for the ECMA-262 date-time format, against an
    independent implementation — Python's datetime — on random well-formed texts: date-only forms are UTC, a date-time without
    an offset is local time of the fixed-offset zone, Z / +-HH:mm are honoured.  And StringToNumber (Number(value), :974)
    against float() / int() on the spellings a stored document can hold."""
    import datetime as dt

    rng = random.Random(31)
    epoch = dt.datetime(1970, 1, 1, tzinfo=dt.timezone.utc)
    for _ in range(20000):
        tz = rng.choice([0, -480, 330, 765, -720, 60])
        y, mo = rng.randrange(1, 9999), rng.randrange(1, 13)
        d = rng.randrange(1, po.days_in_month(y, mo) + 1)
        h, mi, sec, ms = rng.randrange(24), rng.randrange(60), rng.randrange(60), rng.randrange(1000)
        form = rng.randrange(4)
        date = "%04d-%02d-%02d" % (y, mo, d)
        if form == 0:
            text, local, off = date, dt.datetime(y, mo, d), 0
        else:
            clock = "%02d:%02d" % (h, mi) + (":%02d" % sec if form >= 2 else "") + (".%03d" % ms if form == 3 else "")
            local = dt.datetime(y, mo, d, h, mi, sec if form >= 2 else 0, (ms if form == 3 else 0) * 1000)
            zone = rng.choice(["", "Z", "+05:30", "-08:00", "+00:00", "-11:45", "+14:00"])
            text = date + "T" + clock + zone
            off = tz if zone == "" else 0 if zone == "Z" else (1 if zone[0] == "+" else -1) * (int(zone[1:3]) * 60 + int(zone[4:6]))
        want = (local.replace(tzinfo=dt.timezone.utc) - epoch) // dt.timedelta(milliseconds=1) - off * 60000
        assert po.js_date_parse(text, tz) == float(want), (text, tz)
        assert po.get_timestamp_tz(text, tz) == float(want)
    assert np.isnan(po.js_date_parse("2024-13-01")) and np.isnan(po.js_date_parse("2024-01-01T25:00")) and np.isnan(po.js_date_parse("2024-01-01T10:60"))
    assert po.js_date_parse("2024-01-01T24:00") == po.js_date_parse("2024-01-02") and np.isnan(po.js_date_parse("2024-01-01T24:00:01"))
    for bad in ("2024-1-1", "01/02/2024", "2024-01-01 10:00", "2024-01-01T10", "Jan 1 2024", "2024-02-30"):
        with pytest.raises(NotImplementedError):
            po.js_date_parse(bad)
    # StringToNumber
    for _ in range(20000):
        k = rng.randrange(6)
        if k == 0:
            body = str(rng.randrange(10 ** rng.randrange(1, 25)))
        elif k == 1:
            body = "%s.%s" % (rng.randrange(10 ** 6), rng.randrange(10 ** rng.randrange(1, 12)))
        elif k == 2:
            body = repr(rng.uniform(-1, 1) * 10.0 ** rng.randrange(-300, 300)).lstrip("-")
        elif k == 3:
            body = rng.choice([".5", "5.", "1e3", "1E-3", "0.1e+2", "00012", "1.e2", "Infinity"])
        elif k == 4:
            n = rng.randrange(1 << rng.randrange(1, 70))
            base, prefix = rng.choice([(16, "0x"), (16, "0X"), (8, "0o"), (2, "0b"), (2, "0B")])
            body = prefix + {16: "%x", 8: "%o", 2: "{:b}"}[base].replace("{:b}", "%s") % (n if base != 2 else bin(n)[2:])
            want = float(n)
            pad = rng.choice(["", " ", "\t\n", " ", "﻿", "　 "])
            assert po.js_string_to_number(pad + body + pad) == want, body
            continue
        else:
            body = rng.choice(["", " ", "abc", "1 2", "1,5", "0x", "0xg", "1e", "e5", "+-1", "--1", "1_000", "١٢", "infinity", "NaN", "+0x10", "-0b1"])
            got = po.js_string_to_number(body)
            assert (got == 0.0) if body.strip() == "" else np.isnan(got), body
            continue
        sign = rng.choice(["", "+", "-"])
        pad = rng.choice(["", " ", "\r\n", " ", " "])
        got = po.js_string_to_number(pad + sign + body + pad)
        want = float(sign + body.replace("Infinity", "inf"))
        assert got == want and np.copysign(1, got) == np.copysign(1, want), (sign, body, got, want)
