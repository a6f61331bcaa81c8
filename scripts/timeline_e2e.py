import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo") if os.path.exists("/root/repo/bench.py") else None
sys.path.insert(0, os.getcwd())
import torch
from sph_pie_b200 import _lib, ops
from sph_pie_b200.synth import synth_archive
_lib.init(0); lib=_lib.load()
host = synth_archive(1<<20, seed=1234, device="cuda:0").to("cpu").pin()
S,E=host.n_shows,host.n_entries
view=host.view(); total=C.c_uint64(0)
off=torch.empty(E+1,dtype=torch.int64,pin_memory=True)
_lib.check(lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), None, 0, C.byref(total)))
data=torch.empty(int(total.value),dtype=torch.uint8,pin_memory=True)
for i in range(3):
    t0=time.perf_counter()
    _lib.check(lib.pie_csv_rows_host(C.byref(view), off.data_ptr(), data.data_ptr(), data.numel(), C.byref(total)))
    print("call", i, (time.perf_counter()-t0)*1e3, "ms", file=sys.stderr)
