"""Developer tool: hammer the export-row kernel (both row formats, size-only and full passes, several table
sizes, slow path on and off) to flush out protocol races; every result is compared with the first one."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
lib = _lib.load()
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t0 = time.perf_counter()
launches = 0
for shows in (1 << 20, 1 << 17, 5000, 137, 1 << 19):
    table = synth_archive(shows, seed=shows % 97, device="cuda:0")
    E = table.n_entries
    for name, fn in (("csv", ops.csv_rows_dev), ("payload", ops.archive_payloads_dev)):
        sizing = ops.CsvBuffers(E, 0, "cuda:0")
        fn(table, sizing, size_only=True)
        total = int(sizing.total.cpu())
        ref = ops.CsvBuffers(E, total, "cuda:0")
        fn(table, ref)
        torch.cuda.synchronize()
        bufs = ops.CsvBuffers(E, total, "cuda:0")
        n = max(3, rounds * 4000 // max(E, 4000)) if shows < (1 << 19) else rounds // 4
        for i in range(n):
            if i % 3 == 0:
                fn(table, sizing, size_only=True)
            fn(table, bufs)
            launches += 2 if i % 3 == 0 else 1
            if i % 16 == 15 or i == n - 1:
                torch.cuda.synchronize()
                assert int(bufs.total.cpu()) == total
                assert torch.equal(bufs.row_offsets, ref.row_offsets) and torch.equal(bufs.data, ref.data), (name, shows, i)
        print(f"{name:8s} shows {shows:8d} entries {E:9d}: {n} rounds identical ({time.perf_counter() - t0:.1f} s)", flush=True)
        del bufs, ref, sizing
print("launches", launches, "ok")
