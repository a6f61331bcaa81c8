"""Developer tool: time compute_metrics_kernel on the bench table."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
shows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
table = synth_archive(shows, seed=1234, device="cuda:0")
S = table.n_shows
i32 = torch.empty((_lib.PIE_CM_COUNT, S), dtype=torch.int32, device="cuda:0")
text = torch.empty((S, _lib.PIE_CM_TEXT), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    ops.compute_metrics_dev(table, i32, text)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.compute_metrics_dev(table, i32, text)
b.record()
torch.cuda.synchronize()
print("compute_metrics ms", a.elapsed_time(b) / 20, "entries", table.n_entries)
