"""Developer tool: a small run of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle_c  # noqa: E402
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.columnar import pack_shows  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
lib = _lib.load()
shows = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
host = synth_archive(shows, seed=5, shuffle_days=True)
long_show = pack_shows([{"id": "long", "entries": [{"id": "L", "notes": "x" * 70000 + '"' * 9}, {"id": "after"}]}])
for table in (host, long_show):
    dev = table.to("cuda:0")
    for force in (0, 1):
        lib.pie_debug_csv_force_slow_path(force)
        rows = ops.csv_rows(dev)
        off, data = oracle_c.csv_rows(table)
        assert torch.equal(rows.row_offsets.cpu(), off) and torch.equal(rows.data.cpu(), data)
        rows = ops.archive_payloads(dev)
        off, data = oracle_c.payload_rows(table)
        assert torch.equal(rows.row_offsets.cpu(), off) and torch.equal(rows.data.cpu(), data)
    lib.pie_debug_csv_force_slow_path(0)
    ops.archive_analytics(dev, -300)
    ops.compute_metrics(dev)
print("ok")

# JSON ingest: both walks, hostile documents included (escapes, surrogate pairs, damaged texts, ragged sizes)
import json  # noqa: E402

from sph_pie_b200.synth import table_to_shows  # noqa: E402

texts = [json.dumps(s, ensure_ascii=(i % 2 == 0)) for i, s in enumerate(table_to_shows(host.slice_shows(0, min(shows, 500))))]
texts += ['{"id":"\\ud83d\\ude81 \\u00e9","entries":[null,{"delaySec":1e-7,"actions":["a",null]}]}', "", "{", '{"id":"cut',
          "[" * 64 + "]" * 64, json.dumps({"notes": "x" * 70000, "entries": [{}] * 700}), "null"]
docs = ops.JsonDocs.from_texts(texts)
for d in (docs, docs.to("cuda:0")):
    table, status = ops.ingest_json(d)
    assert int(status.sum()) == 4, status.sum()
torch.cuda.synchronize()
print("ingest ok", table.n_shows, table.n_entries)
