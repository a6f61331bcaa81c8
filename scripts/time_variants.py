"""Developer tool: export-row kernel time on differently shaped tables (is it tuned to one distribution?)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
lib = _lib.load()


def timeit(fn, n=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for label, kw, force in (("bench table (dirty vocabularies)", dict(), 0), ("clean vocabularies", dict(dirty=False), 0),
                         ("max 5 entries per show", dict(max_entries=5), 0),
                         ("notes x4 (~100 B of free text)", dict(notes_repeat=4), 0),
                         ("notes x12 (~300 B of free text)", dict(notes_repeat=12), 0),
                         ("notes x40 (~1 KB of free text)", dict(notes_repeat=40), 0),
                         ("bench table, slow path forced", dict(), 1)):
    shows = (1 << 20) if "max 5" not in label else (1 << 21)
    if force or "notes x" in label:
        shows = 1 << 18
    if "x40" in label:
        shows = 1 << 16
    table = synth_archive(shows, seed=1234, device="cuda:0", **kw)
    E = table.n_entries
    lib.pie_debug_csv_force_slow_path(force)
    for name, fn in (("csv", ops.csv_rows_dev), ("payload", ops.archive_payloads_dev)):
        sizing = ops.CsvBuffers(E, 0, "cuda:0")
        fn(table, sizing, size_only=True)
        total = int(sizing.total.cpu())
        bufs = ops.CsvBuffers(E, total, "cuda:0")
        ms = timeit(lambda: fn(table, bufs))
        slow = ops.csv_slow_tiles(table, bufs)
        print(f"{label:34s} {name:8s} rows {E:9d} out {total / 1e9:5.2f} GB  {ms:7.3f} ms  "
              f"{E / ms / 1e6:7.2f} G rows/s  {total / ms / 1e6:7.1f} GB/s out  slow tiles {slow}", flush=True)
        del bufs, sizing
    lib.pie_debug_csv_force_slow_path(0)
    del table
