"""Times the JSON ingest on shows of up to 45 entries (documents of up to ~17 KB): what the roomy configuration of the
warp path is for.  usage: python scripts/time_ingest_long.py [sample_shows=4096] [copies=32] [iters=3]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive, table_to_shows  # noqa: E402


def main():
    sample = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    copies = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device("cuda:0")
    _lib.init(0)
    host = synth_archive(sample, seed=77, max_entries=45)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)
    host.delay_valid[lost] = 0
    texts = [json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in table_to_shows(host)]
    sizes = sorted(len(t.encode()) for t in texts)
    one = ops.JsonDocs.from_texts(texts)
    n, nbytes = one.n_docs, int(one.offsets[-1])
    text = torch.cat([one.data[:nbytes].to(dev).repeat(copies), torch.zeros(8, dtype=torch.uint8, device=dev)])
    lens = (one.offsets[1:] - one.offsets[:-1]).to(dev).repeat(copies)
    offsets = torch.zeros(n * copies + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=offsets[1:])
    docs = ops.JsonDocs(offsets, text)
    print(f"docs={docs.n_docs} text={nbytes * copies / 1e9:.3f} GB; document bytes: median {sizes[len(sizes) // 2]}, "
          f"max {sizes[-1]}, over 8.5 KB: {sum(s > 8500 for s in sizes) / len(sizes):.1%}, over 16 KB: "
          f"{sum(s > 16000 for s in sizes) / len(sizes):.1%}", flush=True)
    bufs = ops.IngestBuffers(docs.n_docs, dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for path in ("warp", "walk"):
        ops.set_ingest_warp_path(1 if path == "warp" else 0)
        ops.ingest_measure_dev(docs, bufs)
        totals = bufs.totals.cpu().tolist()
        assert bufs.status.cpu().tolist()[0] == 0
        declined = ops.ingest_declined(bufs, docs.n_docs) if path == "warp" else docs.n_docs
        table = ops.alloc_ingest_table(docs.n_docs, totals, dev)
        tm = tf = 0.0
        for it in range(iters + 2):
            ev[0].record()
            ops.ingest_measure_dev(docs, bufs)
            ev[1].record()
            ops.ingest_fill_dev(docs, bufs, table)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                tm += ev[0].elapsed_time(ev[1]) / iters
                tf += ev[1].elapsed_time(ev[2]) / iters
        print(json.dumps({"path": path, "declined_to_the_walk": declined, "measure_ms": tm, "fill_ms": tf, "total_ms": tm + tf}),
              flush=True)
    ops.set_ingest_warp_path(1)


if __name__ == "__main__":
    main()
