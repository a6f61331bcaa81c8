"""profiles/sass_r02.txt: what the built library's SASS says about each kernel — the Blackwell-native instructions
(1-D TMA bulk copies UBLKCP, mbarrier SYNCS), registers and spills — so that the evidence need not be re-derived by
disassembling the .so.  Run after build():  python scripts/sass_evidence.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

lib = g.LIB
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True)
usage = {}
name = None
for line in (res.stdout + res.stderr).splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and name:
        usage[name] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
counts = collections.defaultdict(collections.Counter)
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["total"] += 1
        for key in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "UTCHMMA", "LDTM", "HMMA", "LDGSTS", "ATOMS", "BAR", "LDL", "STL"):
            if op.startswith(key):
                counts[fn][key] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}  (csrc sha16 {g.csrc_sha16()}); cuobjdump -sass / -res-usage, CUDA 12.9, sm_100a")
print("# UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops; no tensor-core work exists on this byte-streaming path,")
print("# so the absence of UTC*MMA / LDTM is expected (DESIGN.md section 4).")
for mangled, pretty in sorted(zip(counts, demangle), key=lambda kv: -counts[kv[0]]["total"]):
    c = counts[mangled]
    u = usage.get(mangled)
    short = pretty.replace("(anonymous namespace)::", "").replace("pie::", "").replace("jw::", "")
    short = re.sub(r"^void ", "", re.sub(r"\(.*", "", short))
    if short.startswith("cub::"):
        short = "cub::DeviceScanKernel<int64>" if "DeviceScanKernel" in short else "cub::DeviceScanInitKernel" if "Init" in short else short[:60]
    extra = " ".join(f"{k}={c[k]}" for k in ("UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "BAR", "LDL", "STL") if c[k])
    print(f"{short:48s} sass={c['total']:6d} " + (f"regs={u[0]:3d} smem={u[1]:6d} local={u[2]:4d} " if u else "") + extra)
