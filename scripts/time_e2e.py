"""Developer tool: where the end-to-end (host buffers in, host results out) time of a step goes."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_pie_b200 import _lib, ops  # noqa: E402
from sph_pie_b200.synth import synth_archive  # noqa: E402

_lib.init(0)
lib = _lib.load()
shows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
host = synth_archive(shows, seed=1234, device="cuda:0").to("cpu").pin()
S, E = host.n_shows, host.n_entries
hout = ops.HostOutputs(S, pinned=True)
view = host.view()
total = C.c_uint64(0)
off = torch.empty(E + 1, dtype=torch.int64, pin_memory=True)
_lib.check(lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), None, 0, C.byref(total)))
csv_total = int(total.value)
data = torch.empty(csv_total, dtype=torch.uint8, pin_memory=True)


def t(fn, n=5):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


def bw_test():
    a = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)

    def h2d():
        d.copy_(a, non_blocking=True)
        torch.cuda.synchronize()

    def d2h():
        b.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()

    def both():
        with torch.cuda.stream(s1):
            d.copy_(a, non_blocking=True)
        with torch.cuda.stream(s2):
            b.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()

    print(f"PCIe: h2d {1.0737 / t(h2d) * 1e3:.1f} GB/s, d2h {1.0737 / t(d2h) * 1e3:.1f} GB/s, "
          f"both at once {2 * 1.0737 / t(both) * 1e3:.1f} GB/s in total")


bw_test()
a = t(lambda: ops.archive_analytics(host, -480, hout))
h2d_a, d2h_a = _lib.last_transfer_bytes()
c = t(lambda: _lib.check(lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), data.data_ptr(), csv_total, C.byref(total))))
h2d_c, d2h_c = _lib.last_transfer_bytes()
q = t(lambda: _lib.check(lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), None, 0, C.byref(total))))
print(f"entries {E}: analytics host {a:.1f} ms (h2d {h2d_a / 1e9:.2f} GB, d2h {d2h_a / 1e9:.2f} GB); "
      f"csv host {c:.1f} ms (h2d {h2d_c / 1e9:.2f} GB, d2h {d2h_c / 1e9:.2f} GB); csv size query {q:.1f} ms")
for rows in (1 << 18, 1 << 19, 1 << 21, 1 << 22):
    lib.pie_set_csv_chunk_rows(rows)
    c = t(lambda: _lib.check(lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), data.data_ptr(), csv_total, C.byref(total))))
    print(f"  chunk rows {rows}: csv host {c:.1f} ms")
