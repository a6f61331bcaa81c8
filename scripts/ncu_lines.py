"""Developer tool: per-source-line instruction / stall-sample shares of one kernel.

Joins `ncu -i X.ncu-rep --page source --csv` (per-SASS-instruction metrics, no line info in CSV form)
with `nvdisasm -g -c` of the same cubin (line info, no metrics) by instruction offset.

usage: python scripts/ncu_lines.py gpurun_out/prof.ncu-rep csv_rows_kernel [top_n] [csv_rows]
       (last argument: which .cu's cubin inside libsphpie_b200.so holds the kernel)
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
unit = sys.argv[4] if len(sys.argv) > 4 else "csv_rows"

src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
# a report may hold several kernels: one "Kernel Name" line, one header and the instructions for each; take the
# first section whose kernel name contains NCU_KERNEL (default: the first section)
want = os.environ.get("NCU_KERNEL", "")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
sec = next((i for i in starts if want in (rows[i][1] if len(rows[i]) > 1 else "")), starts[0])
end = next((i for i in starts if i > sec), len(rows))
rows = rows[sec:end]
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
insts = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
base = int(insts[0][0], 16)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "sph_pie_b200", "libsphpie_b200.so")], cwd=tmp,
               capture_output=True)
cubin = os.path.join(tmp, f"{unit}.sm_100a.cubin")
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
infn, cur = False, None
for l in sass:
    if l.startswith("//--------------------- .text."):
        infn = kernel in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))

agg = {}
tot_i = tot_s = 0
unknown = 0
for r in insts:
    off = int(r[0], 16) - base
    inst = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    tinst = int(r[col["Thread Instructions Executed"]] or 0)
    key = line_of.get(off, (None, ""))[0]
    if key is None:
        unknown += inst
        key = ("?", 0)
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += inst
    a[1] += samp
    a[2] += tinst
    tot_i += inst
    tot_s += samp

src_cache = {}


def text(key):
    f, n = key
    path = os.path.join(ROOT, "sph_pie_b200", "csrc", f)
    if not os.path.exists(path):
        return ""
    if f not in src_cache:
        src_cache[f] = open(path).read().splitlines()
    return src_cache[f][n - 1].strip()[:100] if 0 < n <= len(src_cache[f]) else ""


print(f"{kernel}: {tot_i} warp-instructions, {tot_s} samples, {len(insts)} SASS lines ({unknown} inst unmapped)")
SORT = 0 if os.environ.get("BY_INST") else 1  # BY_INST=1: sort by instructions instead of samples
print("by instructions:" if SORT == 0 else "by samples:")
for key, (i, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
    print(f"  {key[0]:18s}:{key[1]:4d} inst {100 * i / tot_i:5.1f}%  samp {100 * s / max(tot_s, 1):5.1f}%  "
          f"lanes {t / max(i, 1):4.1f}  {text(key)}")
