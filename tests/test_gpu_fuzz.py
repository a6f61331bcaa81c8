"""Randomised tables with a hostile alphabet (quotes, commas, CR/LF, control characters, backslashes, multi-byte
UTF-8), ragged shows (empty, single-entry, hundreds of entries), long cells that push tiles over the staging
limits, empty and many-item lists: every GPU path against the C oracle, bit for bit."""
import random

import pytest
import torch

import oracle_c
from sph_pie_b200 import _lib, ops
from sph_pie_b200.columnar import pack_shows

pytestmark = pytest.mark.gpu

ALPHABET = ['"', ",", "\n", "\r", "\\", "\t", "\x00", "\x01", "\x1f", "\x7f", " ", "|", "a", "b", "Z", "0", "9", "-",
            "é", "ü", "漢", "🚁", " ", "﻿", " "]
STATUS = ["Completed", "No-launch", "Abort", "completed", "Completed ", "", "x"]
YESNO = ["Yes", "No", "yes", " YES ", "\tyes\n", "", "y", "﻿yes", "yes "]
ISSUES = ["", "Battery", "RF link", "Other", "10", "2", "02", "4294967294", "4294967295", "ü", " Battery "]


def rand_text(rng, max_len, p_empty=0.3, p_clean=0.4):
    if rng.random() < p_empty:
        return ""
    n = rng.randint(1, max_len)
    if rng.random() < p_clean:
        return "".join(rng.choice("abcXYZ 0189-") for _ in range(n))
    return "".join(rng.choice(ALPHABET) for _ in range(n))


def rand_number(rng):
    k = rng.random()
    if k < 0.15:
        return None
    if k < 0.5:
        return float(rng.randint(-5, 300))
    if k < 0.7:
        return rng.randint(0, 10 ** 6) / 10 ** rng.randint(0, 6)
    if k < 0.8:
        return rng.random() * 10 ** rng.randint(-8, 22)
    return rng.choice([float("nan"), float("inf"), -0.0, 2.0 ** 60, 5e-324, -1.5e-7, 1e21])


def rand_shows(rng, n_shows, max_entries, long_cell_p):
    shows = []
    for _ in range(n_shows):
        k = rng.random()
        if k < 0.1:
            shows.append(None if rng.random() < 0.3 else {"id": rand_text(rng, 8), "entries": []})
            continue
        n_e = 1 if k < 0.3 else rng.randint(1, max_entries)
        entries = []
        for _ in range(n_e):
            big = rng.random() < long_cell_p
            entries.append({
                "id": rand_text(rng, 40), "unitId": rand_text(rng, 12), "planned": rng.choice(YESNO),
                "launched": rng.choice(YESNO), "status": rng.choice(STATUS), "primaryIssue": rng.choice(ISSUES),
                "subIssue": rand_text(rng, 16), "otherDetail": rand_text(rng, 30), "severity": rand_text(rng, 10),
                "rootCause": rand_text(rng, 10), "actions": [rand_text(rng, 12, 0.1) for _ in range(rng.choice([0, 0, 1, 2, 5]))],
                "operator": rand_text(rng, 14), "batteryId": rand_text(rng, 6), "delaySec": rand_number(rng),
                "commandRx": rng.choice(YESNO), "notes": rand_text(rng, 30000 if big else 60, 0.2)})
        shows.append({"id": rand_text(rng, 36), "date": rand_text(rng, 10), "time": rand_text(rng, 5),
                      "label": rand_text(rng, 20), "crew": [rand_text(rng, 10, 0.1) for _ in range(rng.choice([0, 1, 3, 9]))],
                      "leadPilot": rand_text(rng, 12), "monkeyLead": rand_text(rng, 12),
                      "notes": rand_text(rng, 2000 if rng.random() < long_cell_p else 40), "entries": entries})
    return shows


@pytest.mark.parametrize("seed,n_shows,max_entries,long_cell_p", [
    (1, 400, 21, 0.0), (2, 1500, 8, 0.002), (3, 300, 400, 0.001), (4, 2500, 2, 0.0), (5, 60, 21, 0.05)])
def test_fuzz_all_paths(cuda, seed, n_shows, max_entries, long_cell_p):
    rng = random.Random(seed)
    table = pack_shows(rand_shows(rng, n_shows, max_entries, long_cell_p))
    dev = table.to(cuda)
    lib = _lib.load()
    for name, gpu_fn, ref_fn in (("csv", ops.csv_rows, oracle_c.csv_rows), ("payload", ops.archive_payloads, oracle_c.payload_rows)):
        off, data = ref_fn(table)
        for force in (0, 1):
            lib.pie_debug_csv_force_slow_path(force)
            try:
                got = gpu_fn(dev)
            finally:
                lib.pie_debug_csv_force_slow_path(0)
            assert torch.equal(got.row_offsets.cpu(), off), (name, force, "offsets")
            assert torch.equal(got.data.cpu(), data), (name, force, "bytes")
        got = gpu_fn(table)  # host entry point (chunked pipeline)
        assert torch.equal(got.row_offsets, off) and torch.equal(got.data, data), (name, "host")
    i32, text = oracle_c.compute_metrics(table)
    m = ops.compute_metrics(dev)
    assert torch.equal(m.i32.cpu(), i32) and torch.equal(m.text.cpu(), text)
