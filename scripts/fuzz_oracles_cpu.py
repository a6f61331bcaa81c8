"""Long cross-check of the two CPU oracles on hostile tables (no GPU): oracle/pie_oracle.c — the columnar C restatement the
GPU tests and bench.py's `parity_checked` compare the kernels with — against oracle/pie_oracle.py, the JSON-level Python
restatement of the same reference functions (public/app.js:3898-3953, :3401-3502, :5024-5047; server/webhookDispatcher.js:
276-342, :315-330).  Tables come from tests/test_gpu_fuzz.py's generator (quotes, commas, CR/LF, control bytes, multi-byte
UTF-8, ragged shows, odd status / yes-no / issue spellings, every kind of number) with random timestamps and dates added.

    python scripts/fuzz_oracles_cpu.py --seed 1 --minutes 10
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_c  # noqa: E402
import pie_oracle as po  # noqa: E402
from helpers import assert_daily_match_py_oracle, assert_stats_match_py_oracle  # noqa: E402
from sph_pie_b200.columnar import pack_shows  # noqa: E402
from test_gpu_fuzz import rand_shows  # noqa: E402
from test_live_metrics_cpu import metrics_from_planes  # noqa: E402

DAY = 86400000


def add_times(rng, shows):
    """getShowTimestamp (public/app.js:4092-4116) asks createdAt, then date + time, then archivedAt, then the entries' ts."""
    for sh in shows:
        if not sh:
            continue
        base = 1704067200000 + rng.randrange(0, 40) * DAY + rng.randrange(0, DAY)
        k = rng.random()
        if k < 0.6:
            sh["createdAt"] = float(base)  # the date cell stays hostile text: it is never asked
        elif k < 0.7:
            sh["archivedAt"] = float(base) + 0.5
            sh["date"] = ""
        elif k < 0.85:
            sh["date"] = "2024-%02d-%02d" % (rng.randrange(1, 13), rng.randrange(1, 29))
            sh["time"] = rng.choice(["", "21:30", "07:05:09"])
        else:
            sh["date"], sh["time"] = "", ""  # the entries' ts, or no usable timestamp: the show is skipped
        for e in sh["entries"]:
            if rng.random() < 0.3:
                e["ts"] = float(base + rng.randrange(0, 3600000))
    if rng.random() < 0.08:  # a date text outside the ECMA-262 format: both oracles must refuse the batch
        victims = [sh for sh in shows if sh and "createdAt" not in sh]
        if victims:
            rng.choice(victims)["date"] = rng.choice(["July 4, 2024", "04/07/2024", "2024-7-4", "yesterday"])
    return shows


def one_round(rng, stats):
    shows = add_times(rng, rand_shows(rng, rng.choice([40, 120]), rng.choice([2, 8, 21, 60]), rng.choice([0.0, 0.002])))
    tz = rng.choice([0, -480, 330, 765, -210])
    table = pack_shows(shows)
    docs = [s if s else {"entries": []} for s in shows]
    threads = rng.choice([1, 3])
    st, daily, rc, bad = oracle_c.archive_analytics(table, tz_offset_minutes=tz, nthreads=threads)
    try:
        groups = po.build_archive_daily_groups(shows, tz)
        py_rc = 0
    except (po.JsRangeError, NotImplementedError) as e:  # RangeError of toISOString / a date text V8's legacy parser would take
        py_rc = type(e).__name__
        groups = None
    if rc != 0 or py_rc != 0:
        assert rc != 0 and py_rc != 0, ("error status", rc, py_rc)
        stats["date_errors"] += 1
        st = oracle_c.show_stats(table, nthreads=threads)
        assert_stats_match_py_oracle(docs, st)
    else:
        assert_stats_match_py_oracle(docs, st)
        assert_daily_match_py_oracle(shows, daily, tz)
        stats["groups"] += len(groups)
    off, data = oracle_c.csv_rows(table, nthreads=threads)
    poff, pdata = oracle_c.payload_rows(table, nthreads=threads)
    blob, o = bytes(data.numpy()), off.tolist()
    pblob, pofs = bytes(pdata.numpy()), poff.tolist()
    e = 0
    for sh in shows:
        if not sh:
            continue
        for entry in sh["entries"]:
            want = po.build_csv_row(po.build_table_row(sh, entry))
            assert blob[o[e]:o[e + 1]].decode("utf-8") == want + "\n", ("csv row", e)
            want = po.archive_entry_payload_json(sh, entry)
            assert pblob[pofs[e]:pofs[e + 1]].decode("utf-8") == want + "\n", ("payload row", e)
            e += 1
    assert e == table.n_entries
    i32, text = oracle_c.compute_metrics(table)
    got = metrics_from_planes(table, i32, text)
    want = [po.compute_metrics(s if s else {"entries": []}) for s in shows]
    assert got == want, "computeMetrics"
    stats["shows"] += len(shows)
    stats["entries"] += e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=5.0)
    args = ap.parse_args()
    oracle_c.build()
    rng = random.Random(args.seed)
    stats = dict(seed=args.seed, rounds=0, shows=0, entries=0, groups=0, date_errors=0)
    t0 = time.time()
    while time.time() - t0 < args.minutes * 60:
        state = rng.getstate()
        try:
            one_round(rng, stats)
        except AssertionError as e:
            print("DISAGREEMENT in round", stats["rounds"], "of seed", args.seed, ":", str(e)[:600], flush=True)
            sys.exit(1)
        stats["rounds"] += 1
        print(json.dumps(stats), flush=True)
    stats["minutes"] = round((time.time() - t0) / 60, 1)
    print("SUMMARY", json.dumps(stats), flush=True)


if __name__ == "__main__":
    main()
