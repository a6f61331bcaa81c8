"""JSON ingest on the GPU — the stored show documents (reference server/storage/sqlProvider.js:682/:696 write them,
:230-234 / :892-926 read them back with JSON.parse) parsed into the columnar table by pie_ingest_* — against the
oracle (pie_oracle.map_archive_row + the table packer).  Bit-exact tables through both the device-resident and the
host-buffer entry points; the cases are those of tests/test_ingest_cpu.py plus sizes only the GPU handles."""
import json
import random

import numpy as np
import pytest
import torch

import oracle_c
import pie_oracle as po
import test_ingest_cpu as cases
from ingest_helpers import assert_tables_equal, oracle_ingest, stored_doc
from sph_pie_b200 import _lib, ops
from sph_pie_b200.synth import synth_archive, table_to_shows

pytestmark = pytest.mark.gpu


def gpu_check(cuda, docs, what="", host_too=True):
    ref_table, ref_status = oracle_ingest(docs)
    jd = ops.JsonDocs.from_texts(docs)
    table, status = ops.ingest_json(jd.to(cuda))
    torch.cuda.synchronize()
    assert table.is_cuda
    assert np.array_equal(status.cpu().numpy(), ref_status), f"{what} doc_status (dev)"
    assert_tables_equal(table, ref_table, what + " dev")
    if host_too:
        htable, hstatus = ops.ingest_json(jd)
        assert not htable.is_cuda
        assert np.array_equal(hstatus.numpy(), ref_status), f"{what} doc_status (host)"
        assert_tables_equal(htable, ref_table, what + " host")
    return table


@pytest.fixture(params=["warp", "walk"])
def both_paths(request):
    """Runs a test with the warp-per-document path on and off (off = the thread-per-document walk takes everything)."""
    old = ops.set_ingest_warp_path(1 if request.param == "warp" else 0)
    yield request.param
    ops.set_ingest_warp_path(old)


def canonical_shows(n, seed, max_entries=21):
    """Shows the way the provider normalises them (sqlProvider.js:361-409): every entry holds all of its 17 keys."""
    host = synth_archive(n, seed=seed, missing_created_frac=0.1, max_entries=max_entries)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    return table_to_shows(host)


def declined_of(cuda, docs):
    jd = ops.JsonDocs.from_texts(docs).to(cuda)
    bufs = ops.IngestBuffers(jd.n_docs, cuda)
    ops.ingest_measure_dev(jd, bufs)
    return ops.ingest_declined(bufs, jd.n_docs)


def test_warp_path_takes_the_providers_documents_and_declines_the_rest(cuda):
    old = ops.set_ingest_warp_path(1)
    try:
        rng = random.Random(2)
        shows = canonical_shows(300, 5)
        docs = [stored_doc(s, rng, "stringify") for s in shows]
        small = [d for d in docs if len(d.encode()) <= 8000]
        assert len(small) > 250
        assert declined_of(cuda, small) == 0  # every document of the provider's own shape is decided by a warp
        gpu_check(cuda, docs, "canonical", host_too=False)
        pretty = [stored_doc(s, rng, "pretty") for s in shows[:50]]
        assert declined_of(cuda, pretty) == 50  # whitespace between tokens: the walk's business
        assert declined_of(cuda, ["", "{", "[]", "null", '{"id":1}', '{"entries":[{}]}', '{"entries":[{"id":"a"}]}']) == 7
        # shapes the warp path takes although they are not what the provider writes
        assert declined_of(cuda, ["{}", '{"id":"a"}', '{"id":null,"x":1,"y":"z","entries":[],"crew":[]}', '{"crew":["a","b"]}']) == 0
    finally:
        ops.set_ingest_warp_path(old)


def test_long_documents_take_the_roomy_lists(cuda):
    """Documents of 8.5 - 16 KB (shows of 25 - 45 entries) are declined by the lists nearly every document fits and taken
    by the roomy configuration of the same kernels, not by the thread-per-document walk; longer ones (and shows of more
    than 63 entries) are the walk's.  Same table whichever takes them."""
    old = ops.set_ingest_warp_path(1)
    try:
        rng = random.Random(8)
        shows = canonical_shows(160, 23, max_entries=80)
        docs = [stored_doc(s, rng, "stringify") for s in shows]
        sizes = [len(d.encode()) for d in docs]
        mid = [d for d, n, sh in zip(docs, sizes, shows) if 9000 <= n <= 15500 and len(sh["entries"]) <= 63]
        big = [d for d, n in zip(docs, sizes) if n > 17000]
        assert len(mid) > 15 and len(big) > 15
        assert declined_of(cuda, mid) == 0
        assert declined_of(cuda, big) == len(big)
        gpu_check(cuda, docs, "long documents", host_too=False)
        gpu_check(cuda, mid[:3], "long documents only (the last one is parsed twice, by the roomy lists)", host_too=False)
    finally:
        ops.set_ingest_warp_path(old)


def test_structural_variants_of_the_providers_documents(cuda, both_paths):
    """The CPU test's documents (a key missing / twice / renamed / reordered, a value of another type, foreign nesting,
    lists of other things than strings) on the GPU: the same table as the oracle's, and what the oracle refuses fails
    the call with the same status whichever path takes the document."""
    texts = cases.structural_variant_texts()
    keep = []
    for t in texts:
        try:
            oracle_ingest([t])
            keep.append(t)
        except (TypeError, po.UnsupportedJson):
            pass
    gpu_check(cuda, keep, "structural variants", host_too=False)
    refused = [t for t in texts if t not in set(keep)][:40]
    good = keep[0]
    for t in refused:
        with pytest.raises((_lib.SchemaError, _lib.UnsupportedJsonError)) as ei:
            ops.ingest_json(ops.JsonDocs.from_texts([good, t, good]).to(cuda))
        assert ei.value.doc == 1


def test_record_pool_under_pressure(cuda):
    """Pass 1 leaves records for pass 2 in a pool sized for average documents (288 x 8 bytes each): a batch of nothing
    but the largest documents overflows it, the documents that find it full are parsed again by pass 2 — same table."""
    old = ops.set_ingest_warp_path(1)
    try:
        rng = random.Random(6)
        shows = [s for s in canonical_shows(600, 17) if len(s["entries"]) >= 18]
        assert len(shows) > 40
        for sh in shows:
            sh["crew"] = ["Ann", "Bob", "Cy"]
            for e in sh["entries"]:
                e["actions"] = ["Swap battery", "Reboot", 'Call "lead"']
        docs = [stored_doc(s, rng, "stringify") for s in shows]
        docs = [d for d in docs if len(d.encode()) <= 8000]
        assert declined_of(cuda, docs) == 0
        gpu_check(cuda, docs, "pool under pressure", host_too=False)
        gpu_check(cuda, docs[:1], "one document (the last document is always parsed twice)", host_too=False)
    finally:
        ops.set_ingest_warp_path(old)


def test_damaged_canonical_documents(cuda, both_paths):
    """Every prefix and thousands of single-byte damages of documents of the provider's shape: whatever the warp path
    accepts must be what the oracle makes of it; the rest goes to the walk.  Strings with escapes, non-ASCII text,
    crew and actions lists are part of the documents."""
    rng = random.Random(41)
    shows = canonical_shows(40, 9)
    notes = ['line\none "quoted" \\ back', "caf\u00e9 \u00fc \u20ac", "tab\there", "", "\u4e2d\u6587 \U0001f600 \x7f"]
    for k, sh in enumerate(shows):
        sh["notes"] = notes[k % 5]
        sh["crew"] = ["Ann", 'B "ob"', "Zo\u00eb"][:k % 4]
        for j, e in enumerate(sh["entries"]):
            if j % 3 == 0:
                e["actions"] = ["Swap battery", "Re\u00efnit", 'say "hi"\n'][:1 + j % 3]
    docs = []
    for sh in shows:
        doc = stored_doc(sh, rng, "stringify")
        if len(doc) < 2500 and len(docs) < 8000:
            docs += [doc[:k] for k in range(len(doc) + 1)]
    alphabet = '"\\{}[]:,0-9.eE+tfn ux\n\t\x01a\u00e9'
    for sh in shows:
        doc = stored_doc(sh, rng, "stringify")
        docs.append(doc)
        for _ in range(150):
            k = rng.randrange(len(doc))
            docs.append(doc[:k] + rng.choice(alphabet) + doc[k + 1:])
        for _ in range(30):  # a character less, a character more
            k = rng.randrange(len(doc))
            docs.append(doc[:k] + doc[k + 1:])
            docs.append(doc[:k] + rng.choice(alphabet) + doc[k:])
    keep = []
    for d in docs:
        try:
            oracle_ingest([d])
            keep.append(d)
        except (TypeError, po.UnsupportedJson):
            pass
    assert len(keep) > 5000
    gpu_check(cuda, keep, "damaged canonical", host_too=False)


def _outcome(jd):
    try:
        table, status = ops.ingest_json(jd)
        torch.cuda.synchronize()
        return ("ok", table, status.cpu())
    except (_lib.UnsupportedJsonError, _lib.SchemaError) as e:
        return ("error", type(e).__name__, e.doc)


def test_raw_byte_damage_and_alignment(cuda):
    """Damage below the character level (bytes that break UTF-8): both paths must do the same with every batch — the
    same table, or the same error with the same document.  Then documents at every alignment of the text."""
    rng = random.Random(43)
    shows = canonical_shows(12, 13)
    for sh in shows:
        sh["label"] = "caf\u00e9 \u4e2d \U0001f600"
    nasty = [0x80, 0xBF, 0xC0, 0xC2, 0xE0, 0xED, 0xF0, 0xF4, 0xF5, 0xFF, 0x00, 0x1F, 0x22, 0x5C, 0xA0, 0x9F, 0x8F, 0x90]
    good = stored_doc(shows[0], rng, "stringify").encode()
    old = ops.set_ingest_warp_path(1)
    try:
        n_err = n_ok = 0
        for sh in shows:
            raw = stored_doc(sh, rng, "stringify").encode("utf-8")
            hi = [i for i, c in enumerate(raw) if c >= 0x80]
            batch = []
            for _ in range(60):
                k = rng.choice(hi) + rng.randrange(-2, 3) if hi and rng.random() < 0.7 else rng.randrange(len(raw))
                k = min(max(k, 0), len(raw) - 1)
                batch.append(raw[:k] + bytes([rng.choice(nasty)]) + raw[k + 1:])
            for b in batch:
                jd = ops.JsonDocs.from_texts([good, b, good]).to(cuda)
                ops.set_ingest_warp_path(1)
                got = _outcome(jd)
                ops.set_ingest_warp_path(0)
                ref = _outcome(jd)
                assert got[0] == ref[0], b
                if got[0] == "error":
                    assert got[1:] == ref[1:], b
                    n_err += 1
                else:
                    assert torch.equal(got[2], ref[2]), b
                    assert_tables_equal(got[1], ref[1], repr(b))
                    n_ok += 1
        assert n_err > 50 and n_ok > 50
        ops.set_ingest_warp_path(1)
        # every alignment: the same documents behind 0..40 bytes of another document
        base = [stored_doc(sh, rng, "stringify") for sh in shows[:4]]
        for pad in range(0, 41):
            gpu_check(cuda, ["x" * pad] + base, f"pad {pad}", host_too=False)
    finally:
        ops.set_ingest_warp_path(old)


@pytest.mark.parametrize("style", ["stringify", "ascii", "pretty", "shuffled"])
def test_synthetic_archive_round_trips(cuda, style):
    rng = random.Random(3)
    host = synth_archive(310, seed=11, missing_created_frac=0.1)
    docs = [stored_doc(s, rng, style) for s in table_to_shows(host)]
    table = gpu_check(cuda, docs, style)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    if style == "stringify":
        host.delay_sec[host.delay_sec == 0] = 0.0  # JSON.stringify(-0) is "0"
    assert_tables_equal(table, host, style + " vs the source table")


@pytest.mark.parametrize("style", ["stringify", "ascii", "pretty", "shuffled"])
def test_hostile_strings_and_numbers(cuda, style):
    rng = random.Random(5)
    shows = [cases.hostile_show(rng, rng.randrange(0, 6)) for _ in range(400)]
    gpu_check(cuda, [stored_doc(s, rng, style) for s in shows], style)


def test_projection_rules_and_damaged_documents(cuda):
    import inspect

    # the document list of the CPU test, verbatim
    src = inspect.getsource(cases.test_projection_rules)
    env = {}
    exec(src.replace("def test_projection_rules():", "def docs_only():").replace("    check(docs)", "    return docs"), env)
    gpu_check(cuda, env["docs_only"](), "projection rules")
    rng = random.Random(9)
    show = cases.hostile_show(rng, 2)
    doc = stored_doc(show, rng, "ascii")
    docs = [doc[:k] for k in range(len(doc) + 1)]
    for _ in range(3000):
        k = rng.randrange(len(doc))
        docs.append(doc[:k] + rng.choice('"\\{}[]:,0-9.eE+tfn ux\n\t\x01a') + doc[k + 1:])
    keep = []
    for d in docs:
        try:
            oracle_ingest([d])
            keep.append(d)
        except (TypeError, po.UnsupportedJson):
            pass
    gpu_check(cuda, keep, "damaged")


def test_schema_and_unsupported_documents_fail_loudly(cuda):
    good = '{"id":"fine","entries":[{"id":"e"}]}'
    for d in cases.SCHEMA_DOCS:
        for docs in (ops.JsonDocs.from_texts([good, d, good, '{"id":7}']),):
            for jd in (docs, docs.to(cuda)):
                with pytest.raises(_lib.SchemaError) as ei:
                    ops.ingest_json(jd)
                assert ei.value.code == _lib.PIE_ERR_SCHEMA and ei.value.doc == 1, d
                assert isinstance(ei.value, TypeError)
    for d in cases.UNSUPPORTED_DOCS:
        docs = ops.JsonDocs.from_texts([good, good, d, '{"id":7}'])
        for jd in (docs, docs.to(cuda)):
            with pytest.raises(_lib.UnsupportedJsonError) as ei:
                ops.ingest_json(jd)
            assert ei.value.doc == 2, d


def test_numbers_and_ragged_documents(cuda):
    rng = np.random.default_rng(17)
    xs = rng.integers(0, 2 ** 64, 20000, dtype=np.uint64).view(np.float64)
    xs = xs[np.isfinite(xs)]
    texts = [repr(float(x)) for x in xs] + ["%.17e" % x for x in xs[:5000]]
    docs = ['{"createdAt":%s,"entries":[{"delaySec":%s,"ts":%s},{"delaySec":%s}]}' % (a, a, a, b)
            for a, b in zip(texts, reversed(texts))]
    gpu_check(cuda, docs, "numbers")
    gpu_check(cuda, [], "empty batch")
    gpu_check(cuda, [""], "one empty text")
    big = {"id": "big", "notes": "x" * 70000, "entries": [{"notes": "y" * 5000, "actions": ["a"] * 300}] * 40}
    docs = [json.dumps(big), "{}", json.dumps({"entries": [{}] * 1000}), "[]", json.dumps(big)[:-1], '{"id":"z"}'] * 3
    gpu_check(cuda, docs, "ragged")


def test_larger_archive_and_the_operators_downstream(cuda):
    """40k shows: the ingested table is the source table, and the operators give the same results on it."""
    host = synth_archive(40000, seed=21)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    host.delay_sec[lost] = 0.0
    docs = [json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in table_to_shows(host)]
    jd = ops.JsonDocs.from_texts(docs)
    table, status = ops.ingest_json(jd.to(cuda))
    assert int(status.sum()) == 0
    assert_tables_equal(table, host, "40k shows")
    htable, _ = ops.ingest_json(jd)
    assert_tables_equal(htable, host, "40k shows, host entry point")
    ref_stats, ref_daily, rc, _ = oracle_c.archive_analytics(host, tz_offset_minutes=60)
    assert rc == 0
    stats, daily = ops.archive_analytics(table, tz_offset_minutes=60)
    assert torch.equal(stats.i32.cpu(), ref_stats.i32) and torch.equal(stats.f64.cpu().view(torch.int64), ref_stats.f64.view(torch.int64))
    rows = ops.csv_rows(table)
    ref_offsets, ref_csv = oracle_c.csv_rows(host)
    assert torch.equal(rows.row_offsets.cpu(), ref_offsets) and torch.equal(rows.data.cpu(), ref_csv)


def test_replicated_batch_at_scale(cuda):
    """Size-independent property at a batch the oracle cannot walk: k copies of a batch ingest to k copies of its
    table (2^18 documents, ~1.2 GB of text)."""
    host = synth_archive(4096, seed=33)
    docs = [json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in table_to_shows(host)]
    docs[100] = docs[100][:-1]  # one dropped row per copy
    one = ops.JsonDocs.from_texts(docs)
    n, nbytes = one.n_docs, int(one.offsets[-1])
    k = 64
    text = one.data[:nbytes].to(cuda).repeat(k)
    text = torch.cat([text, torch.zeros(8, dtype=torch.uint8, device=cuda)])
    lens = (one.offsets[1:] - one.offsets[:-1]).to(cuda).repeat(k)
    offsets = torch.zeros(n * k + 1, dtype=torch.int64, device=cuda)
    torch.cumsum(lens, 0, out=offsets[1:])
    base, base_status = ops.ingest_json(one.to(cuda))
    table, status = ops.ingest_json(ops.JsonDocs(offsets, text))
    torch.cuda.synchronize()
    assert table.n_shows == n * k and table.n_entries == base.n_entries * k
    assert torch.equal(status.view(k, n), base_status.view(1, n).expand(k, n))
    E = base.n_entries
    eo = table.entry_offsets[:-1].view(k, n) - (torch.arange(k, device=cuda, dtype=torch.int32) * E).view(k, 1)
    assert torch.equal(eo, base.entry_offsets[:-1].view(1, n).expand(k, n))
    for name, col in base.entry_cols.items():
        big = table.entry_cols[name]
        nb = int(col.offsets[E])
        assert int(big.offsets[E * k]) == nb * k, name
        assert torch.equal(big.data[:nb * k].view(k, nb), col.data[:nb].view(1, nb).expand(k, nb)), name
        lens_big = (big.offsets[1:E * k + 1] - big.offsets[:E * k]).view(k, E)
        assert torch.equal(lens_big, (col.offsets[1:E + 1] - col.offsets[:E]).view(1, E).expand(k, E)), name
    assert torch.equal(table.delay_sec.view(torch.int64).view(k, E), base.delay_sec.view(torch.int64).view(1, E).expand(k, E))
    assert torch.equal(table.delay_valid.view(k, E), base.delay_valid.view(1, E).expand(k, E))


def test_mirror_api_list_archived_shows(cuda):
    """listArchivedShows(rows): JSON.parse per row, rows that map to null filtered out (sqlProvider.js:230-234)."""
    from sph_pie_b200 import listArchivedShows, mapArchiveRows
    from sph_pie_b200.columnar import pack_shows

    shows = table_to_shows(synth_archive(50, seed=2))
    rows = [{"data": po.js_json_stringify(s)} for s in shows]
    rows[3] = {"data": "{not json"}
    rows[10] = {"data": "null"}
    rows[11] = "42"
    rows[20] = {"data": None}
    rows.append({"data": "[]"})  # an array is an object: kept, as an empty show
    table, dropped = mapArchiveRows(rows)
    assert dropped.nonzero().flatten().tolist() == [3, 10, 11, 20]
    kept = [po.map_archive_row((r["data"] if r["data"] is not None else "null") if isinstance(r, dict) else r) for r in rows]
    kept = [k for k in kept if k is not None]
    for device in ("cuda", "cpu"):
        got = listArchivedShows(rows, device=device)
        assert got.n_shows == len(rows) - 4
        assert_tables_equal(got, pack_shows(kept), "listArchivedShows " + device)
        st = ops.show_stats(got)
        ref = oracle_c.show_stats(pack_shows(kept))
        assert torch.equal(st.i32.cpu(), ref.i32)


def test_archive_step_from_stored_texts(cuda):
    """Host texts -> GPU ingest -> statistics, daily summaries and CSV rows -> host: the same results as the
    operators on the source table (which the C oracle checks)."""
    host = synth_archive(3000, seed=41, shuffle_days=True)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    host.delay_sec[lost] = 0.0
    docs = ops.JsonDocs.from_texts([json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in table_to_shows(host)])
    stats, daily, rows, dropped = ops.archive_step_from_json(docs.pin(), tz_offset_minutes=-300)
    assert not bool(dropped.any())
    ref_stats, ref_daily, rc, _ = oracle_c.archive_analytics(host, tz_offset_minutes=-300)
    assert rc == 0
    from helpers import assert_analytics_equal

    assert_analytics_equal((stats, daily), (ref_stats, ref_daily), "from JSON")
    ref_offsets, ref_csv = oracle_c.csv_rows(host)
    assert torch.equal(rows.row_offsets, ref_offsets) and torch.equal(rows.data, ref_csv)


def test_fill_walk_stays_inside_the_table(cuda):
    """The second walk writes through positions it computes itself: every byte of slack behind every column it is
    handed must come back untouched (compute-sanitizer is not available on the GPU pool)."""
    rng = random.Random(77)
    shows = [cases.hostile_show(rng, rng.randrange(0, 9)) for _ in range(700)]
    docs = [stored_doc(s, rng, rng.choice(["stringify", "ascii", "pretty", "shuffled"])) for s in shows]
    docs += ["", "{", '{"id":"cut', "[]", json.dumps({"notes": "x" * 70000, "entries": [{}] * 700})]
    jd = ops.JsonDocs.from_texts(docs).to(cuda)
    bufs = ops.IngestBuffers(jd.n_docs, cuda)
    ops.ingest_measure_dev(jd, bufs)
    totals = bufs.totals.cpu().tolist()
    assert bufs.status.cpu().tolist()[0] == 0
    slack = 64
    padded = [t + slack for t in totals]  # rows and bytes: every column gets 64 spare elements
    table = ops.alloc_ingest_table(jd.n_docs + 0, padded, cuda)
    tensors = [table.entry_offsets, table.created_at, table.archived_at, table.delay_sec, table.delay_valid, table.entry_ts,
               table.crew.list_offsets, table.crew.items.offsets, table.crew.items.data, table.actions.list_offsets,
               table.actions.items.offsets, table.actions.items.data]
    for c in list(table.show_cols.values()) + list(table.entry_cols.values()):
        tensors += [c.offsets, c.data]
    for t in tensors:
        t.view(torch.uint8).fill_(0xA5)
    table.n_entries = totals[_lib.PIE_IT_ENTRIES]  # the columns are longer than the table
    ops.ingest_fill_dev(jd, bufs, table)
    torch.cuda.synchronize()
    S, E = jd.n_docs, totals[_lib.PIE_IT_ENTRIES]
    ci, ai = totals[_lib.PIE_IT_CREW_ITEMS], totals[_lib.PIE_IT_ACTION_ITEMS]

    def untouched(t, used, what):
        tail = t[used:].contiguous().view(torch.uint8)  # (the show-level columns are exactly n_docs (+ 1) long: no tail)
        assert bool((tail == 0xA5).all()), what

    # the table was allocated for E + 64 entries etc.: rows used are known from the true totals
    untouched(table.entry_offsets, S + 1, "entry_offsets")
    untouched(table.created_at, S, "created_at")
    untouched(table.delay_sec, E, "delay_sec")
    untouched(table.delay_valid, E, "delay_valid")
    untouched(table.entry_ts, E, "entry_ts")
    untouched(table.crew.list_offsets, S + 1, "crew.list_offsets")
    untouched(table.crew.items.offsets, ci + 1, "crew.items.offsets")
    untouched(table.crew.items.data, totals[7], "crew.items.data")
    untouched(table.actions.list_offsets, E + 1, "actions.list_offsets")
    untouched(table.actions.items.offsets, ai + 1, "actions.items.offsets")
    untouched(table.actions.items.data, totals[22], "actions.items.data")
    for h, name in enumerate(_lib.SHOW_STR_COLS):
        untouched(table.show_cols[name].offsets, S + 1, name)
        untouched(table.show_cols[name].data, totals[h], name)
    for h, name in enumerate(_lib.ENTRY_STR_COLS):
        untouched(table.entry_cols[name].offsets, E + 1, name)
        untouched(table.entry_cols[name].data, totals[8 + h], name)
    # and what it did write is the oracle's table
    ref, _ = oracle_ingest(docs)
    table.n_entries = E
    table.delay_sec, table.delay_valid, table.entry_ts = table.delay_sec[:E], table.delay_valid[:E], table.entry_ts[:E]
    assert_tables_equal(table, ref, "padded table")


@pytest.mark.parametrize("chunk_docs", [700, 1 << 16])
def test_pipelined_step_from_stored_texts(cuda, chunk_docs):
    """The chunked, three-stream composition gives what the single-batch one gives (dropped rows, shuffled days and
    an empty show in the middle of a chunk included)."""
    host = synth_archive(3000, seed=43, shuffle_days=True, missing_created_frac=0.05)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    texts = [json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in table_to_shows(host)]
    texts[5] = "{broken"
    texts[699] = "null"
    texts[700] = '{"id":"no entries","date":"2024-02-03","time":"10:00"}'
    texts[2999] = texts[2999][:-2]
    docs = ops.JsonDocs.from_texts(texts).pin()
    s1, d1, r1, x1 = ops.archive_step_from_json(docs, tz_offset_minutes=120)
    off = torch.empty(r1.row_offsets.numel(), dtype=torch.int64, pin_memory=True)
    csv = torch.empty(r1.data.numel(), dtype=torch.uint8, pin_memory=True)
    s2, d2, r2, x2 = ops.archive_step_from_json_pipelined(docs, 120, None, off, csv, chunk_docs=chunk_docs)
    if chunk_docs == 700:  # the one-call C entry point gives the same, with and without the size query
        for args in ((), (None, torch.empty(off.numel(), dtype=torch.int64), torch.empty(csv.numel(), dtype=torch.uint8))):
            s3, d3, r3, x3 = ops.archive_step_json_host(docs, 120, *args)
            assert torch.equal(x1, x3)
            from helpers import assert_analytics_equal as same

            same((s3, d3), (s1, d1), "pie_archive_step_json_host")
            assert torch.equal(r1.row_offsets, r3.row_offsets) and torch.equal(r1.data, r3.data)
        with pytest.raises(_lib.PieError) as ei:  # too small a CSV buffer: PIE_ERR_CAPACITY, sizes reported
            ops.archive_step_json_host(docs, 120, None, torch.empty(off.numel(), dtype=torch.int64), torch.empty(10, dtype=torch.uint8))
        assert ei.value.code == _lib.PIE_ERR_CAPACITY
        empty = ops.archive_step_json_host(ops.JsonDocs.from_texts([]), 0)
        assert empty[1].n_groups == 0 and empty[2].row_offsets.tolist() == [0]
        # ... and so does its pipelined form (chunks of documents over three streams, the chunks' show-level columns
        # joined for the daily summary): chunk borders inside a day, a dropped row first / last in a chunk
        from helpers import assert_analytics_equal as same

        lib = _lib.load()
        for chunk in (700, 256, 2999):
            old = lib.pie_set_json_chunk_docs(chunk)
            try:
                for _ in range(2):  # the second call reuses every buffer of the first
                    s4, d4, r4, x4 = ops.archive_step_json_host(docs, 120, None, torch.empty(off.numel(), dtype=torch.int64),
                                                                torch.empty(csv.numel(), dtype=torch.uint8))
                    assert torch.equal(x1, x4)
                    same((s4, d4), (s1, d1), f"pie_archive_step_json_host in chunks of {chunk}")
                    assert torch.equal(r1.row_offsets, r4.row_offsets) and torch.equal(r1.data, r4.data)
                with pytest.raises(_lib.PieError) as ei:
                    ops.archive_step_json_host(docs, 120, None, torch.empty(off.numel(), dtype=torch.int64), torch.empty(100000, dtype=torch.uint8))
                assert ei.value.code == _lib.PIE_ERR_CAPACITY
                bad = ops.JsonDocs.from_texts(texts[:1500] + ['{"entries":[{"delaySec":"text"}]}'] + texts[1500:])
                with pytest.raises(_lib.SchemaError) as es:
                    ops.archive_step_json_host(bad, 120, None, torch.empty(off.numel() + 8, dtype=torch.int64), torch.empty(csv.numel() + 64, dtype=torch.uint8))
                assert es.value.doc == 1500
            finally:
                lib.pie_set_json_chunk_docs(old)
    from helpers import assert_analytics_equal

    assert torch.equal(x1, x2) and x1.nonzero().flatten().tolist() == [5, 699, 2999]
    assert_analytics_equal((s2, d2), (s1, d1), "pipelined")
    assert torch.equal(r1.row_offsets, r2.row_offsets) and torch.equal(r1.data, r2.data)
    # and against the oracle on what the documents hold
    ref_table, _ = oracle_ingest(texts)
    ref_stats, ref_daily, rc, _ = oracle_c.archive_analytics(ref_table, tz_offset_minutes=120)
    assert rc == 0
    assert_analytics_equal((s2, d2), (ref_stats, ref_daily), "pipelined vs oracle")
    ref_offsets, ref_csv = oracle_c.csv_rows(ref_table)
    assert torch.equal(r2.row_offsets, ref_offsets) and torch.equal(r2.data, ref_csv)


def test_reference_fixture_and_a_two_megabyte_document(cuda):
    """The reference's fixture as stored text; and one document at the reference's body limit (2 MB, index.js:69) —
    a single lane walks it (~1 s): slow, but right."""
    import os

    from sph_pie_b200.columnar import pack_shows

    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "webhook_fixture.json")))
    table = gpu_check(cuda, [fx["stored_text"]], "fixture")
    assert_tables_equal(table, pack_shows([{**fx["show"], "entries": [fx["entry"]]}]), "fixture vs the objects")
    rng = random.Random(1)
    big = cases.hostile_show(rng, 0)
    big["entries"] = [cases.hostile_show(rng, 1)["entries"][0] for _ in range(4500)]
    text = stored_doc(big, rng, "stringify")
    assert 1.5e6 < len(text.encode()) < 2.1e6
    gpu_check(cuda, ["{}", text, fx["stored_text"]], "2 MB document", host_too=False)


def test_random_json_on_the_gpu(cuda):
    """The random documents of tests/test_ingest_cpu.py (any JSON value anywhere, random whitespace) on the GPU."""
    rng = random.Random(2025)
    docs = [cases.random_document(rng) for _ in range(8000)]
    fine, schema, unsupported = cases.classify(docs)
    assert len(fine) > 1500 and len(schema) > 1500
    gpu_check(cuda, fine, "random JSON", host_too=False)
    mixed = fine[:50] + [schema[0]] + fine[50:100] + [schema[1]]
    with pytest.raises(_lib.SchemaError) as ei:
        ops.ingest_json(ops.JsonDocs.from_texts(mixed).to(cuda))
    assert ei.value.doc == 50
    for d in schema[:300]:
        with pytest.raises(_lib.SchemaError):
            ops.ingest_json(ops.JsonDocs.from_texts([d]).to(cuda))


def test_map_archive_rows_with_the_rows_timestamp_columns(cuda):
    """_mapArchiveRow in full (sqlProvider.js:892-926): archived_at of the row over the document's archivedAt, the
    document's createdAt over created_at, NULL columns (Number(null) = 0), a column the SELECT left out."""
    from sph_pie_b200 import mapArchiveRows
    from sph_pie_b200.columnar import pack_shows

    shows = table_to_shows(synth_archive(40, seed=9, missing_created_frac=0.5))
    rows = []
    for i, s in enumerate(shows):
        # a timestamp the document does not have is ABSENT here (a null would count as 0 — Number(null) — which
        # tests/test_gpu_maintenance.py covers with the full restatement of _mapArchiveRow)
        s = {k: v for k, v in s.items() if not (k in ("createdAt", "archivedAt") and v is None)}
        row = {"data": po.js_json_stringify(s)}
        if i % 4 != 3:
            row["archived_at"] = str(1704067200000 + i) if i % 5 else " 1.7040672e12 "
        if i % 3 == 0:
            row["created_at"] = None
        elif i % 3 == 1:
            row["created_at"] = po.js_number_to_string(1700000000000.0 + i)
        rows.append(row)
    rows[7]["data"] = "{broken"
    rows[8]["data"] = "[]"
    want = [po.map_archive_row_full(r) for r in rows]
    for device in ("cuda", "cpu"):
        table, dropped = mapArchiveRows(rows, device=device)
        assert dropped.nonzero().flatten().tolist() == [7]
        ref = pack_shows(want)
        assert_tables_equal(table, ref, "mapArchiveRows with row columns " + device)
    # a column that holds an ISO text goes through Date.parse (a date-only form is UTC); any other text is V8's
    # legacy parser, which is not restated: it raises
    t, _ = mapArchiveRows([{"data": "{}", "archived_at": "2024-01-01"}])
    assert t.archived_at.tolist() == [1704067200000.0]
    with pytest.raises(NotImplementedError):
        mapArchiveRows([{"data": "{}", "archived_at": "next tuesday"}])


def test_documents_that_do_not_start_at_offset_zero(cuda):
    """A slice of a larger text buffer: offsets[0] != 0, every alignment of the first byte (both entry points and the
    one-call step)."""
    rng = random.Random(12)
    shows = [cases.hostile_show(rng, rng.randrange(0, 5)) for _ in range(40)]
    for i, s in enumerate(shows):  # timestamps the daily grouping accepts (an out-of-range one is a RangeError there)
        s["createdAt"], s["archivedAt"] = 1704067200000.0 + 3600000.0 * i, None
    texts = [stored_doc(s, rng, "stringify") for s in shows]
    ref, ref_status = oracle_ingest(texts)
    base = ops.JsonDocs.from_texts(texts)
    nbytes = int(base.offsets[-1])
    for lead in (1, 5, 8, 13, 64):
        data = torch.cat([torch.full((lead,), ord("}"), dtype=torch.uint8), base.data[:nbytes], torch.full((9,), ord("{"), dtype=torch.uint8)])
        docs = ops.JsonDocs(base.offsets + lead, data)
        for d in (docs, docs.to(cuda)):
            table, status = ops.ingest_json(d)
            assert np.array_equal(status.cpu().numpy(), ref_status)
            assert_tables_equal(table, ref, f"lead {lead}")
        st, dl, rows, dropped = ops.archive_step_json_host(docs, 0)
        ref_rows = ops.csv_rows(ref)
        assert torch.equal(rows.row_offsets, ref_rows.row_offsets) and torch.equal(rows.data, ref_rows.data)
        # a window inside the buffer: the middle documents only
        sub = ops.JsonDocs(docs.offsets[10:31].clone(), docs.data)
        table, status = ops.ingest_json(sub.to(cuda))
        assert_tables_equal(table, oracle_ingest(texts[10:30])[0], f"lead {lead} middle")
