"""CUDA-event time of pie_show_payloads_dev (measure pass + scan + write pass) on a synthetic archive resident in HBM, and
a check of a sample of the documents against the Python oracle (reference server/webhookDispatcher.js:545-584).

    python scripts/time_show_payloads.py [log2_shows=18] [runs=3]
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import pie_oracle as po  # noqa: E402
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive, table_to_shows  # noqa: E402
from sph_pie_b200.webhook import payload_frame  # noqa: E402

ENTRY_KEYS = ["id", "ts", "unitId", "planned", "launched", "status", "primaryIssue", "subIssue", "otherDetail", "severity",
              "rootCause", "actions", "operator", "batteryId", "delaySec", "commandRx", "notes"]


def main():
    S = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 18)
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    _lib.init(0)
    table = synth_archive(S, seed=0, device=dev)
    args = ("show.updated", "2024-07-03T12:00:00.000Z", "https://example.invalid/hook", "POST")
    head, tail = payload_frame(*args, None)
    t0 = time.time()
    first = ops.show_payloads(table, head, tail)
    torch.cuda.synchronize()
    first_s = time.time() - t0
    total = int(first.data.numel())
    lib = _lib.load()
    h = torch.tensor(list(head), dtype=torch.uint8, device=dev)
    t = torch.tensor(list(tail), dtype=torch.uint8, device=dev)
    offs = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.pie_show_payloads_scratch_bytes(S)) + 256, dtype=torch.uint8, device=dev)
    sptr = (scratch.data_ptr() + 255) & ~255
    view = table.view()
    stream = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(runs):
        _lib.check(lib.pie_show_payloads_dev(C.byref(view), h.data_ptr(), len(head), t.data_ptr(), len(tail), offs.data_ptr(),
                                             first.data.data_ptr(), total, tot.data_ptr(), status.data_ptr(), sptr, stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / runs
    assert status.cpu().tolist()[0] == 0 and int(tot.cpu()) == total
    in_bytes = table.nbytes()
    print(json.dumps({"shows": S, "entries": table.n_entries, "json_bytes_out": total, "ms_per_call": ms,
                      "first_call_s_incl_alloc": first_s, "algorithmic_bytes": in_bytes + total + 8 * (S + 1),
                      "achieved_gbs": (in_bytes + total + 8 * (S + 1)) / (ms * 1e-3) / 1e9,
                      "shows_per_s": S / (ms * 1e-3)}), flush=True)
    # parity of the same kernels on a SMALL table of its own (slices of the big one share its heaps: a per-show copy to the
    # host would move the whole heap each time — that is what ate the last GPU seconds of round 2)
    small = synth_archive(400, seed=1, device=dev)
    docs = ops.show_payloads(small, head, tail).documents()
    equal = True
    for show, body in zip(table_to_shows(small.to("cpu")), docs):
        show["entries"] = [{k: e.get(k, [] if k == "actions" else "" if k not in ("ts", "delaySec") else None) for k in ENTRY_KEYS}
                           for e in show.get("entries", [])]
        if body != po.show_payload_json(args[0], show, *args[1:]):
            equal = False
            break
    print(json.dumps({"sample_equals_oracle": equal, "sampled": len(docs)}), flush=True)


if __name__ == "__main__":
    main()
