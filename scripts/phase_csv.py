"""Developer tool: per-phase cycle counts of export_rows_kernel (library built with -DPIE_CSV_PROFILE)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
shows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
table = synth_archive(shows, seed=1234, device="cuda:0")
E = table.n_entries
sizing = ops.CsvBuffers(E, 0, "cuda:0")
ops.csv_rows_dev(table, sizing, size_only=True)
total = int(sizing.total.cpu())
bufs = ops.CsvBuffers(E, total, "cuda:0")
for _ in range(3):
    ops.csv_rows_dev(table, bufs)
torch.cuda.synchronize()
print("rows", E, "full pass: phase cycles summed over all tiles (worker thread 0):", flush=True)
ops.csv_slow_tiles(table, bufs)
print("size-only pass:", flush=True)
ops.csv_rows_dev(table, sizing, size_only=True)
ops.csv_slow_tiles(table, sizing)
