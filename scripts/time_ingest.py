"""Times the JSON-ingest kernels on the GPU: a sample of stored documents (json.dumps of a synthetic archive),
replicated on the device to the bench size, through pie_ingest_measure_dev + pie_ingest_fill_dev.
usage: python scripts/time_ingest.py [sample_shows=16384] [copies=64] [iters=5]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_stored_docs  # noqa: E402


def main():
    sample = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    copies = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    _lib.init(0)
    t0 = time.time()
    docs, n_entries, nbytes, _ = synth_stored_docs(sample, copies, dev)
    print(f"docs={docs.n_docs} entries={n_entries} text={nbytes / 1e9:.3f} GB (built in {time.time() - t0:.1f}s)", flush=True)
    bufs = ops.IngestBuffers(docs.n_docs, dev)
    tables = {}
    for path in ("warp", "walk"):  # the warp-per-document path, then the thread-per-document walk alone
        ops.set_ingest_warp_path(1 if path == "warp" else 0)
        ops.ingest_measure_dev(docs, bufs)
        totals = bufs.totals.cpu().tolist()
        assert bufs.status.cpu().tolist()[0] == 0
        declined = ops.ingest_declined(bufs, docs.n_docs) if path == "warp" else docs.n_docs
        table = ops.alloc_ingest_table(docs.n_docs, totals, dev)
        out_bytes = table.nbytes()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tm = tf = 0.0
        for it in range(iters + 2):
            ev[0].record()
            ops.ingest_measure_dev(docs, bufs)
            ev[1].record()
            ops.ingest_fill_dev(docs, bufs, table)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                tm += ev[0].elapsed_time(ev[1])
                tf += ev[1].elapsed_time(ev[2])
        tm, tf = tm / iters, tf / iters
        tables[path] = table
        print(json.dumps({"path": path, "declined": declined, "docs": docs.n_docs, "entries": n_entries, "text_gb": nbytes / 1e9,
                          "table_gb": out_bytes / 1e9, "measure_ms": tm, "fill_ms": tf, "total_ms": tm + tf,
                          "text_gbs_measure": nbytes / tm / 1e6, "text_gbs_total": nbytes / (tm + tf) / 1e6,
                          "algorithmic_gbs": (nbytes + out_bytes) / (tm + tf) / 1e6,
                          "entries_per_s": n_entries / (tm + tf) * 1e3}), flush=True)
    ops.set_ingest_warp_path(1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ingest_helpers import assert_tables_equal
    assert_tables_equal(tables["warp"], tables["walk"], "warp path vs walk")
    print("tables of both paths equal", flush=True)


if __name__ == "__main__":
    main()
