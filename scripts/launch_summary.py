"""Developer tool: per-kernel averages from an ncu launch list
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv ...`).
Only launches on the bench workload are averaged (duration within 2x of the kernel's longest launch).
usage: python scripts/launch_summary.py gpurun_out/launches.csv [traffic.json shows]"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    full = r[col["Kernel Name"]].replace("pie::", "").replace("void ", "")
    name = full.split("(")[0].split("<")[0]
    if name == "export_rows_kernel":  # the two row formats are different instantiations
        name += "<json>" if ("<1>" in full or "<(bool)1>" in full or "<true>" in full) else "<csv>"
    agg[name][r[col["Metric Name"]]].append(float(r[col["Metric Value"]].replace(",", "")))
out = {}
for k, m in agg.items():
    t = m["gpu__time_duration.sum"]
    idx = [i for i, x in enumerate(t) if x > 0.5 * max(t)]
    w = m["dram__bytes_write.sum"]
    if k.startswith("export_rows_kernel"):  # leave the size-only passes (no output bytes) out of the average
        idx = [i for i in idx if w[i] > 0.5 * max(w)]
    avg = lambda key: sum(m[key][i] for i in idx) / len(idx)  # noqa: E731
    out[k] = {"dram_bytes_read": round(avg("dram__bytes_read.sum"), -5), "dram_bytes_write": round(avg("dram__bytes_write.sum"), -5),
              "time_us": round(avg("gpu__time_duration.sum") / 1000.0, 1), "launches_averaged": len(idx)}
total = sum(v["time_us"] for v in out.values())
for k, v in sorted(out.items(), key=lambda kv: -kv[1]["time_us"]):
    v["share_pct"] = round(100 * v["time_us"] / total, 1)
    print(f"{k:28s} {v['time_us']:9.1f} us {v['share_pct']:5.1f} %  read {v['dram_bytes_read'] / 1e6:8.1f} MB  "
          f"write {v['dram_bytes_write'] / 1e6:8.1f} MB  ({v['launches_averaged']} launches)")
if len(sys.argv) > 3:
    json.dump({"source": f"{sys.argv[1]}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "--clock-control none, `python bench.py --steps 3 --warmup 3`; per-launch averages over the "
                         "launches on the bench workload; export_rows_kernel: full passes only (size-only passes write nothing). The share "
                         "is of the sum over ALL kernels listed, of which export_rows_kernel<json> and compute_metrics_kernel are "
                         "measured outside the step",
               "shows": int(sys.argv[3]), "kernels": dict(sorted(out.items(), key=lambda kv: -kv[1]["time_us"]))},
              open(sys.argv[2], "w"), indent=1)
