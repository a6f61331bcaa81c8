"""Long campaign for the show-payload device code ON THE CPU (sph_pie_b200/csrc/pie_show_payload.cuh compiled by g++, its 32
lanes as fibers: tests/native/payload_host.cpp) against the Python restatement of dispatchShowEvent's schemaVersion 2 body
(oracle/pie_oracle.show_payload_json, reference server/webhookDispatcher.js:545-584): random shows over a hostile alphabet,
cell lengths around the one-round limits (29 / 31 / 32 bytes), 0..40 entries, lists of 0..5 items, every kind of number and
timestamp — each batch emitted through the shared-memory stage AND through the caller's view.

    python scripts/fuzz_payload_cpu.py --seed 1 --minutes 10
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

import pie_oracle as po  # noqa: E402
from test_payload_oracles_cpu import normalised, payload_host_bodies  # noqa: E402

ALPHABET = ['"', ",", "\n", "\r", "\\", "\t", "\b", "\f", "\x00", "\x1f", "\x7f", "|", "é", "漢", "🚁", " ", "a", "B", "7", "'", "/", "{", "]", ":"]
NUMBERS = [None, 0.0, -0.0, 12.5, 1e21, 1e-7, float("nan"), float("inf"), float("-inf"), 1 / 3, 5e-324, 1.7976931348623157e308, 123456.789,
           -3.0, 2.0 ** 53, 0.1 + 0.2]
TIMES = [None, 1704067200000.0, 1704067200000.5, 0, True, False, 1e21, -0.0, float("inf"), float("nan")]


def text(rng):
    kind = rng.random()
    if kind < 0.35:
        return "".join(rng.choice("abcdefgh XYZ0123-_.") for _ in range(rng.choice([0, 1, 5, 12, 27, 28, 29, 30, 31, 32, 33, 40, 64, 65, 100])))
    if kind < 0.5:
        n = rng.choice([28, 29, 30, 31, 32, 33])
        k = rng.randrange(n)
        return "x" * k + rng.choice(ALPHABET) + "x" * (n - k - 1)
    return "".join(rng.choice(ALPHABET) for _ in range(rng.randrange(0, 45)))


def show(rng):
    sh = {"id": text(rng), "date": text(rng), "time": text(rng), "label": text(rng), "crew": [text(rng) for _ in range(rng.randrange(0, 6))],
          "leadPilot": text(rng), "monkeyLead": text(rng), "notes": text(rng) * rng.choice([1, 1, 1, 40]), "entries": []}
    for k in ("createdAt", "updatedAt", "archivedAt", "deletedAt"):
        if rng.random() < 0.7:
            sh[k] = rng.choice(TIMES)
    for _ in range(rng.choice([0, 1, 2, 3, 5, 9, 21, 31, 32, 33, 40])):
        sh["entries"].append({"id": text(rng), "ts": rng.choice([None, 0.0, 1704067200123.0, 1.5e-7, float("nan")]), "unitId": text(rng),
                              "planned": rng.choice(["Yes", "No", ""]), "launched": text(rng),
                              "status": rng.choice(["Completed", "Abort", "completed", "No-launch", text(rng)]), "primaryIssue": text(rng),
                              "subIssue": text(rng), "otherDetail": text(rng), "severity": text(rng), "rootCause": text(rng),
                              "actions": [text(rng) for _ in range(rng.randrange(0, 4))], "operator": text(rng), "batteryId": text(rng),
                              "delaySec": rng.choice(NUMBERS), "commandRx": text(rng), "notes": text(rng)})
    return sh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--minutes", type=float, default=5.0)
    args = ap.parse_args()
    rng = random.Random(args.seed)
    stats = dict(seed=args.seed, shows=0, entries=0, staged=0, bytes=0)
    t0 = time.time()
    frame = ("ev\"ent\n", "2024-07-05T04:00:00.000Z", "https://hooks.example/pie?x=1&y=\"2\"", "POST")
    while time.time() - t0 < args.minutes * 60:
        shows = [show(rng) for _ in range(24)]
        meta = rng.choice([None, {"automation": {"k": [1, "two", None]}}])
        want = [po.show_payload_json(frame[0], normalised(s), *frame[1:], meta if meta is not None else po.UNDEFINED) for s in shows]
        for use_stage in (1, 0):
            bodies, staged, status = payload_host_bodies(shows, *frame, meta, use_stage)
            if status != [0, -1] or bodies != want:
                bad = next((i for i, (b, w) in enumerate(zip(bodies or [], want)) if b != w), -1)
                path = os.path.join(ROOT, "gpurun_out", "fuzz_payload_fail_%d.json" % args.seed)
                json.dump({"show": shows[bad] if bad >= 0 else None, "status": status, "use_stage": use_stage}, open(path, "w"))
                print("DIFFERENCE", status, use_stage, bad, "->", path, flush=True)
                sys.exit(1)
            if use_stage:
                stats["staged"] += staged
        stats["shows"] += len(shows)
        stats["entries"] += sum(len(s["entries"]) for s in shows)
        stats["bytes"] += sum(len(w.encode()) for w in want)
        print(json.dumps(stats), flush=True)
    stats["minutes"] = round((time.time() - t0) / 60, 1)
    print("SUMMARY", json.dumps(stats), flush=True)


if __name__ == "__main__":
    main()
