"""Developer tool: the from-JSON step, single batch vs pipelined, with a phase breakdown by wall clock."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_stored_docs  # noqa: E402

_lib.init(0)
dev = torch.device("cuda:0")
sample = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 128
docs, n_entries, nbytes, _ = synth_stored_docs(sample, copies, dev, seed=4321)
hdocs = docs.to("cpu").pin()
del docs
st, dl, rows, _ = ops.archive_step_from_json(hdocs, 0)
hout = ops.HostOutputs(hdocs.n_docs, pinned=True)
h_off = torch.empty(rows.row_offsets.numel(), dtype=torch.int64, pin_memory=True)
h_csv = torch.empty(rows.data.numel(), dtype=torch.uint8, pin_memory=True)
del st, dl, rows
for name, fn in (("single", ops.archive_step_from_json), ("pipelined", ops.archive_step_from_json_pipelined)):
    for chunk in ((None,) if name == "single" else (65536, 131072, 262144)):
        kw = {} if chunk is None else {"chunk_docs": chunk}
        fn(hdocs, 0, hout, h_off, h_csv, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn(hdocs, 0, hout, h_off, h_csv, **kw)
        dt = (time.perf_counter() - t0) / 3
        print(f"{name} chunk={chunk}: {dt * 1e3:.1f} ms, {n_entries / dt / 1e6:.1f} M entries/s", flush=True)

lib = _lib.load()
for chunk in (1 << 30, 65536, 131072, 262144):
    old = lib.pie_set_json_chunk_docs(chunk)
    ops.archive_step_json_host(hdocs, 0, hout, h_off, h_csv)
    t0 = time.perf_counter()
    for _ in range(3):
        ops.archive_step_json_host(hdocs, 0, hout, h_off, h_csv)
    dt = (time.perf_counter() - t0) / 3
    lib.pie_set_json_chunk_docs(old)
    print(f"pie_archive_step_json_host chunk_docs={chunk if chunk < 1 << 30 else 'single batch'}: {dt * 1e3:.1f} ms, "
          f"{n_entries / dt / 1e6:.1f} M entries/s", flush=True)
