"""Developer tool: time the export-row kernels (size-only pass vs full pass) on the bench table."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
shows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
table = synth_archive(shows, seed=1234, device="cuda:0")
E = table.n_entries
sizing = ops.CsvBuffers(E, 0, "cuda:0")
ops.csv_rows_dev(table, sizing, size_only=True)
total = int(sizing.total.cpu())
bufs = ops.CsvBuffers(E, total, "cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("rows", E, "csv bytes", total)
print("size-only pass ms", timeit(lambda: ops.csv_rows_dev(table, sizing, size_only=True)))
print("full pass ms     ", timeit(lambda: ops.csv_rows_dev(table, bufs)))
