"""profiles/traffic_r02.json from an ncu launch list of `python bench.py --steps 3 --warmup 3`
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
--log-file gpurun_out/launches.csv ...`): per-kernel per-launch averages over the launches on the bench workload, and
the same summed over the kernel GROUPS bench.py times as one unit (roofline.traffic must cover what roofline.kernel
names).  Records the hash of the sources the captured library was built from.
usage: python scripts/traffic_from_launches.py gpurun_out/launches.csv profiles/traffic_r02.json <shows> [trimmed.csv]
(trimmed.csv: the list reduced to this library's kernels, at most 120 launches of each — the copy kept under profiles/)"""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

# kernels launched by one call of the entry point the group stands for (csv_rows.cu launch_rows)
GROUPS = {
    "export_rows_csv": ["expand_entry_show_kernel", "column_sample_kernel", "plan_tile_rows_kernel", "export_rows_kernel<csv>"],
    "export_rows_json": ["export_rows_kernel<json>"],
    # the default path: a warp per document (the walk kernels it launches beside them find an empty list on the bench's
    # documents; their large launches in the list are those of the walk timed alone)
    "ingest": ["ingest_init_kernel", "ingest_route_kernel", "ingest_fast_kernel<measure>", "ingest_scan_sums_kernel", "ingest_scan_blocks_kernel",
               "ingest_scan_apply_kernel", "ingest_fast_kernel<fill>"],
    "ingest_walk_alone": ["ingest_order_count_kernel", "ingest_order_scan_kernel", "ingest_order_place_kernel", "ingest_init_kernel",
                          "ingest_walk_kernel<measure>", "ingest_scan_sums_kernel", "ingest_scan_blocks_kernel",
                          "ingest_scan_apply_kernel", "ingest_walk_kernel<fill>", "ingest_rows_to_columns_kernel"],
}


def kernel_name(full: str) -> str:
    full = (full.replace("pie::", "").replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            .replace("unnamed>::", "").replace("(bool)", ""))
    targ = full.split("(")[0]
    name = targ.split("<")[0]
    first = targ.split("<", 1)[1] if "<" in targ else ""
    truthy = first.startswith(("1", "(bool)1", "true"))
    if name == "export_rows_kernel":
        name += "<json>" if truthy else "<csv>"
    if name in ("ingest_walk_kernel", "ingest_fast_kernel"):
        name += "<fill>" if truthy else "<measure>"
    return name


def trim(src: str, dst: str) -> None:
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if len(r) > 10)
    ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Name")
    seen, out = collections.Counter(), []
    for r in rows:
        if len(r) > 10 and r is not hdr:
            if "pie::" not in r[ki]:
                continue
            key = (r[ki].split("(")[0], r[mi])
            seen[key] += 1
            if seen[key] > 120:
                continue
        out.append(r)
    csv.writer(open(dst, "w", newline=""), quoting=csv.QUOTE_ALL).writerows(out)


def main():
    src, dst, shows = sys.argv[1], sys.argv[2], int(sys.argv[3])
    if len(sys.argv) > 4:
        trim(src, sys.argv[4])
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    col = {h: i for i, h in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[1:]:
        agg[kernel_name(r[col["Kernel Name"]])][r[col["Metric Name"]]].append(float(r[col["Metric Value"]].replace(",", "")))
    out = {}
    for k, m in agg.items():
        t = m["gpu__time_duration.sum"]
        idx = [i for i, x in enumerate(t) if x > 0.5 * max(t)]  # launches on the bench workload
        w = m["dram__bytes_write.sum"]
        if k.startswith("export_rows_kernel"):  # full passes only: the size-only passes write nothing
            idx = [i for i in idx if w[i] > 0.5 * max(w)]
        avg = lambda key: sum(m[key][i] for i in idx) / len(idx)  # noqa: E731
        out[k] = {"dram_bytes_read": round(avg("dram__bytes_read.sum")), "dram_bytes_write": round(avg("dram__bytes_write.sum")),
                  "time_us": round(avg("gpu__time_duration.sum") / 1000.0, 1), "launches_averaged": len(idx)}
    groups = {}
    for g, names in GROUPS.items():
        have = [n for n in names if n in out]
        if have:
            groups[g] = {"kernels": have, "dram_bytes_read": sum(out[n]["dram_bytes_read"] for n in have),
                         "dram_bytes_write": sum(out[n]["dram_bytes_write"] for n in have),
                         "time_us": round(sum(out[n]["time_us"] for n in have), 1)}
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["time_us"]):
        print(f"{k:32s} {v['time_us']:9.1f} us  read {v['dram_bytes_read'] / 1e6:8.1f} MB  write {v['dram_bytes_write'] / 1e6:8.1f} MB"
              f"  ({v['launches_averaged']} launches)")
    for g, v in groups.items():
        print(f"group {g}: {v['time_us']} us, read {v['dram_bytes_read'] / 1e6:.1f} MB + write {v['dram_bytes_write'] / 1e6:.1f} MB")
    json.dump({"source": f"{src}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "--clock-control none on `python bench.py --steps 3 --warmup 3`; per-launch averages over the "
                         "launches on the bench workload (ncu times are cold-cache and serialised: shares, not absolutes)",
               "shows": shows, "csrc_sha16": entry.csrc_sha16(),
               "group_src_sha16": {g: entry.group_src_sha16(g) for g in entry.TRAFFIC_GROUP_SOURCES},
               "kernels": dict(sorted(out.items(), key=lambda kv: -kv[1]["time_us"])), "groups": groups},
              open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
