"""Long differential fuzz of the JSON ingest ON THE CPU (no GPU needed): the device code of the warp-per-document path
(tests/native/fast_host.cpp: pie_json_fast.cuh with its 32 lanes as fibers) and the thread-per-document walk
(tests/native/ingest_host.cpp) against the Python oracle (pie_oracle.map_archive_row, reference
server/storage/sqlProvider.js:892-926) on documents of the provider's shape and damaged copies of them: single and
multiple byte edits, deletions, insertions, splices of two documents, hostile strings in every text cell.

    python scripts/fuzz_ingest_cpu.py --seed 1 --minutes 10 [--mode variants]

Prints one line per round and a JSON summary; exits 1 on the first disagreement (the document is written to
gpurun_out/fuzz_ingest_fail_<seed>.json).  tests/test_ingest_cpu.py holds the bounded version of the same comparisons."""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import pie_oracle as po  # noqa: E402
from ingest_helpers import (ROUTE_RECORDS, ROUTE_SLOW, assert_tables_equal, fast_host_ingest, host_ingest, oracle_ingest,  # noqa: E402
                            stored_doc)
from sph_pie_b200 import _lib  # noqa: E402
from sph_pie_b200.synth import synth_archive, table_to_shows  # noqa: E402
from test_ingest_cpu import hostile_text, structural_variant_texts  # noqa: E402

ALPHABET = '"\\{}[]:,0-9.eE+tfn ux\n\t\x01aé-lrs'


def canonical_docs(rng, n):
    host = synth_archive(n, seed=rng.randrange(1 << 30), missing_created_frac=0.1, max_entries=rng.choice([3, 8, 21, 40]))
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    shows = table_to_shows(host)
    for sh in shows:  # hostile text in some cells: escapes of every kind inside the provider's shape
        if rng.random() < 0.4:
            for target in [sh] + sh["entries"]:
                for k, v in target.items():
                    if isinstance(v, str) and k not in ("date",) and rng.random() < 0.15:
                        target[k] = hostile_text(rng, rng.randrange(0, 14))
                    elif isinstance(v, list) and v and isinstance(v[0], str) and rng.random() < 0.15:
                        target[k] = [hostile_text(rng, rng.randrange(0, 6)) for _ in range(rng.randrange(0, 5))]
    return [stored_doc(s, rng, "stringify") for s in shows]


def damage(rng, doc, other):
    what = rng.randrange(7)
    k = rng.randrange(len(doc))
    if what == 0:
        return doc[:k] + rng.choice(ALPHABET) + doc[k + 1:]
    if what == 1:
        return doc[:k] + doc[k + 1:]
    if what == 2:
        return doc[:k] + rng.choice(ALPHABET) + doc[k:]
    if what == 3:  # several edits
        d = list(doc)
        for _ in range(rng.randrange(2, 5)):
            d[rng.randrange(len(d))] = rng.choice(ALPHABET)
        return "".join(d)
    if what == 4:  # a splice of two documents at token-ish boundaries
        j = rng.randrange(len(other))
        return doc[:k] + other[j:]
    if what == 5:  # a run removed
        return doc[:k] + doc[k + rng.randrange(1, 40):]
    j = rng.randrange(len(doc))  # a run repeated
    a, b = min(j, k), max(j, k)
    return doc[:b] + doc[a:b][:60] + doc[b:]


def one_round(rng, stats, mode):
    if mode == "variants":  # one STRUCTURAL change each: a key missing / twice / renamed / reordered, values of other types
        cases = structural_variant_texts(seed=rng.randrange(1 << 30), archive_seed=rng.randrange(1 << 30))
    else:
        docs = canonical_docs(rng, rng.choice([12, 24]))
        small = [d for d in docs if len(d) < 6000] or docs
        cases = list(docs)
        for _ in range(600):
            cases.append(damage(rng, rng.choice(small), rng.choice(small)))
    keep = []
    for d in cases:
        try:
            oracle_ingest([d])
            keep.append(d)
        except (TypeError, po.UnsupportedJson):
            stats["refused"] += 1
            walk_err = host_ingest([d])[2]
            warp_err = fast_host_ingest([d])[2]
            if walk_err[0] not in (_lib.PIE_ERR_SCHEMA, _lib.PIE_ERR_UNSUPPORTED_JSON) or warp_err != walk_err:
                return d, "refused by the oracle: walk %r, warp path %r" % (walk_err, warp_err)
    ref_table, ref_status = oracle_ingest(keep)
    for pool in (288, 0):
        table, status, err, routes = fast_host_ingest(keep, pool_units_per_doc=pool)
        try:
            assert err == (0, -1), err
            assert np.array_equal(status, ref_status), "doc_status"
            assert_tables_equal(table, ref_table, "warp path pool %d" % pool)
        except AssertionError as e:
            # find the document
            for d in keep:
                rt, rs = oracle_ingest([d])
                t, s, er, _ = fast_host_ingest([d] * 2, pool_units_per_doc=pool)
                rt2, rs2 = oracle_ingest([d] * 2)
                try:
                    assert er == (0, -1) and np.array_equal(s, rs2)
                    assert_tables_equal(t, rt2)
                except AssertionError as e2:
                    return d, "warp path (pool %d): %s" % (pool, e2)
            return keep[0], "warp path (pool %d), only in the batch: %s" % (pool, e)
        if pool:
            stats["records"] += int((routes == ROUTE_RECORDS).sum())
            stats["declined"] += int((routes == ROUTE_SLOW).sum())
    table, status, err = host_ingest(keep)
    try:
        assert err == (0, -1), err
        assert np.array_equal(status, ref_status), "doc_status"
        assert_tables_equal(table, ref_table, "walk")
    except AssertionError as e:
        return keep[0], "walk: %s" % e
    stats["documents"] += len(keep)
    stats["dropped_rows"] += int(ref_status.sum())
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=5.0)
    ap.add_argument("--mode", choices=["damage", "variants"], default="damage")
    args = ap.parse_args()
    rng = random.Random(args.seed)
    stats = dict(seed=args.seed, mode=args.mode, rounds=0, documents=0, refused=0, dropped_rows=0, records=0, declined=0)
    t0 = time.time()
    while time.time() - t0 < args.minutes * 60:
        bad = one_round(rng, stats, args.mode)
        if bad:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            path = os.path.join(ROOT, "gpurun_out", "fuzz_ingest_fail_%d.json" % args.seed)
            json.dump({"doc": bad[0], "why": bad[1]}, open(path, "w"))
            print("DISAGREEMENT:", bad[1], "->", path, flush=True)
            sys.exit(1)
        stats["rounds"] += 1
        print(json.dumps(stats), flush=True)
    stats["minutes"] = round((time.time() - t0) / 60, 1)
    print("SUMMARY", json.dumps(stats), flush=True)


if __name__ == "__main__":
    main()
