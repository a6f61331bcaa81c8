#!/usr/bin/env python
"""bench.py — one JSON line for the archive-analytics path on B200.

READ THIS FIRST: BASELINE.json's metric is literally "N/A: no GPU hot path"; the
reference is a web app whose largest possible archive is ~6.5k entries.  The metric below
("archive entries analysed per second") is this repo's own, defined on the SURVEY §7 fallback
scope, and the headline workload is ~1700x beyond what the reference can hold.  the line
carries a measurement at the reference-reachable maximum too (`reference_scale`), where the GPU
end-to-end call is NOT faster than one CPU core.  vs_baseline is null: nothing is published.

A step = one pass of the path over one batch:
  analytics  computeArchiveShowStats for every show, the daily groups and the 19 metric summaries per
             day (public/app.js:3401-3502, :3898-3953)
  export     buildCsvRow(buildTableRow(show, entry)) for every entry (server/webhookDispatcher.js:276-342)
  value : entries/s, table resident in HBM, kernels only (CUDA events on the launch stream)
  e2e   : entries/s through pie_archive_step_host (HOST buffers in, HOST results out, H2D + kernels + D2H inside
          the timed region)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "archive_entries_per_s"
UNIT = "entries/s"
DEFAULT_SHOWS = 1 << 20  # ~11 M entries; the columns one step reads are ~0.45 GB, well above the 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shows", type=int, default=DEFAULT_SHOWS, help="shows per GPU (weak scaling)")
    ap.add_argument("--tz", type=int, default=-480, help="fixed local-zone offset in minutes east of UTC")
    ap.add_argument("--cpu-sample-shows", type=int, default=1 << 18)
    return ap.parse_args()


def note(msg: str) -> None:
    """Progress line on stderr (never stdout: stdout carries exactly one JSON line)."""
    print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # NVML missing: report it, do not invent clocks
            self.nv = None
            self.err = repr(e)

    def sample(self):
        if not self.nv:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {
                "hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap,
                "hw_power_brake_slowdown": nv.nvmlClocksEventReasonHwPowerBrakeSlowdown,
                "applications_clocks_setting": nv.nvmlClocksEventReasonApplicationsClocksSetting,
            }
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def analytics_bytes(table, n_groups: int):
    """Algorithmic bytes of one step: every input byte the path must read once + every output byte
    it must write once.  Returns (show_stats_bytes, daily_bytes)."""
    from sph_pie_b200 import _lib

    S, E = table.n_shows, table.n_entries
    cols = [table.entry_cols[c] for c in ("status", "launched", "primary_issue")]
    stats_in = sum(c.nbytes() for c in cols) + 9 * E + 4 * (S + 1)
    stats_out = (4 * _lib.PIE_SI_COUNT + 8 * _lib.PIE_SF_COUNT) * S
    daily_in = 8 * S + (4 * 4 + 8 * 15) * S  # created_at + the 4 i32 / 15 f64 planes the 19 metrics read
    daily_out = 12 * S + (8 + 4) * n_groups + (3 * 8 + 4) * _lib.PIE_N_METRICS * n_groups
    return stats_in + stats_out, daily_in + daily_out


class CpuStep:
    """One step of the path on the host CPU with the C oracle (oracle/pie_oracle.c), output buffers
    preallocated outside the timed region (like the GPU arm)."""

    def __init__(self, host_table, tz, nthreads):
        import torch

        import oracle_c
        from sph_pie_b200.ops import HostOutputs

        self.o, self.t, self.tz, self.n = oracle_c, host_table, tz, nthreads
        self.hout = HostOutputs(host_table.n_shows)
        self.csv_off, data = oracle_c.csv_rows(host_table, nthreads)
        self.csv_data = torch.empty(max(data.numel(), 1), dtype=torch.uint8)

    def __call__(self):
        _, _, rc, _ = self.o.archive_analytics(self.t, self.tz, self.n, self.hout)
        assert rc == 0
        self.o.csv_rows(self.t, self.n, self.csv_off, self.csv_data)


def cpu_baseline_run(host_table, tz, nthreads, min_seconds=1.0, max_reps=8):
    """Times the C oracle over `host_table`; returns (entries/s, reps, seconds)."""
    step = CpuStep(host_table, tz, nthreads)  # construction = one warm pass
    reps, t0 = 0, time.perf_counter()
    while True:
        step()
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= max_reps:
            return host_table.n_entries * reps / dt, reps, dt


def make_table(args, rank, world, local):
    """The synthetic archive this rank works on, the SAME table in both arms (`--impl b200` and `--impl reference`): it
    is generated with torch's CUDA generator when a GPU is visible (both arms run on the same box, so seed -> table is
    one function) and with the CPU generator otherwise (`data_generator` in the config says which).
    N = 1: one segment of `--shows` shows.  N > 1: ONE archive of N such segments (weak scaling), cut into N day ranges
    balanced by entries (sharding.plan_day_shards on the archive's skeleton); a rank assembles its own range, which
    usually spans two stored segments (sharding.assemble_shard).  No data-path collective either way.
    Returns (table, generator, (s0, s1))."""
    import torch

    from sph_pie_b200.sharding import assemble_shard, segment_skeleton
    from sph_pie_b200.synth import synth_archive

    gen_dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    generator = "torch cuda generator" if gen_dev.type == "cuda" else "torch cpu generator"
    if world == 1:
        return synth_archive(args.shows, seed=1234, device=gen_dev), generator, (0, args.shows)
    seeds = [1234 + k for k in range(world)]
    eo, day = segment_skeleton(args.shows, seeds, gen_dev)
    table, span = assemble_shard(args.shows, seeds, eo, day, rank, world, gen_dev)
    return table, generator, span


def bind_host_side(local: int, world: int):
    """Give this rank its own CPUs before any pinned buffer is allocated: the CPUs of the GPU's NUMA node (sysfs), split
    evenly among the local ranks on that node; all CPUs split evenly when the platform reports no node (a VM).  The
    end-to-end leg is bound by the host side of the PCIe copies, and N ranks sharing one set of cores and one first-touch
    node is the worst case."""
    info = {"numa_node": None, "cpus": None}
    try:
        import pynvml

        pynvml.nvmlInit()
        nodes = []
        for i in range(world):
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
            nodes.append(int(open(path).read()) if os.path.exists(path) else -1)
        allowed = sorted(os.sched_getaffinity(0))
        node = nodes[local]
        cpus = allowed
        if node >= 0 and os.path.exists(f"/sys/devices/system/node/node{node}/cpulist"):
            on_node = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                on_node.update(range(int(lo), int(hi or lo) + 1))
            cpus = [c for c in allowed if c in on_node] or allowed
        peers = [r for r in range(world) if nodes[r] == node]
        k, n = peers.index(local), len(peers)
        mine = cpus[k * len(cpus) // n:(k + 1) * len(cpus) // n] or cpus
        os.sched_setaffinity(0, mine)
        info = {"numa_node": node, "cpus": f"{mine[0]}-{mine[-1]}", "n_cpus": len(mine), "ranks_on_node": n}
    except Exception as e:  # no NVML / no sysfs: stay where the launcher put us, and say so
        info["error"] = repr(e)
    return info


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU.  The reference itself is
    JavaScript and cannot run here (no JS engine), so this is the C port (kind "port").  It runs the repo arm's
    config: the same table (rank 0's), the same number of shows, the same warm-up rule."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import oracle_c

    oracle_c.build()
    threads = oracle_c.max_threads()
    table, generator, _ = make_table(args, 0, world, local)  # at N > 1: rank 0's day range of the one archive
    table = table.to("cpu")
    warmup = max(args.warmup, 3)
    step = CpuStep(table, args.tz, threads)  # construction = one warm pass
    for _ in range(warmup - 1):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = table.n_entries * args.steps / dt
    shows = table.n_shows
    sample = f"{shows} shows / {table.n_entries} entries per step, C port of the path (oracle/pie_oracle.c), " \
             f"show statistics and export rows on {threads} threads, daily grouping on 1"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32/f64", "data": "synthetic",
        "config": workload_config(args, shows, table.n_entries, generator),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, shows, entries, generator):
    return {
        "workload": f"synthetic archive: {shows} shows x 0..21 entries ({entries} entries) per GPU, <=5 shows/day, "
                    "show statistics + daily groups + 19 metric summaries + CSV export rows",
        "shows_per_gpu": shows, "entries_per_gpu": entries, "tz_offset_minutes": args.tz, "data_generator": generator,
        "l2_policy": "inputs larger than L2 (a step reads ~2.2 GB and writes ~3 GB per GPU vs 126 MB L2)",
        "baseline_metric": "N/A: no GPU hot path (BASELINE.json); metric defined by this repo, see DESIGN.md",
    }


def _same(a, b) -> bool:
    """Bit-for-bit equality of two tensors (NaN == NaN, -0.0 != 0.0)."""
    import torch

    a, b = a.cpu(), b.cpu()
    if a.shape != b.shape:
        return False
    if a.dtype.is_floating_point:
        return bool(torch.equal(a.contiguous().view(torch.int64), b.contiguous().view(torch.int64)))
    return bool(torch.equal(a, b))


def parity_check_step(host, bufs, cbufs, csv_total, tz, threads):
    """After the timed steps: every resident output of the TIMED table against the C oracle (oracle/pie_oracle.c) on
    the host — the statistics planes, the daily tables and every CSV byte.  The oracle is the checker, never the
    thing measured."""
    import oracle_c

    t0 = time.perf_counter()
    S, E = host.n_shows, host.n_entries
    ref_stats, ref_daily, rc, _ = oracle_c.archive_analytics(host, tz, threads)
    differs = []
    if rc != 0:
        differs.append(f"oracle daily summary status {rc}")
    else:
        G = int(bufs.n_groups.cpu())
        got = {"stats_i32": bufs.stats_i32[:, :S], "stats_f64": bufs.stats_f64[:, :S], "show_day_start": bufs.show_day_start[:S],
               "show_order": bufs.show_order[:S], "group_day_start": bufs.group_day_start[:G],
               "group_offsets": bufs.group_offsets[:G + 1], "summary_f64": bufs.summary_f64[:, :, :G],
               "summary_count": bufs.summary_count[:, :G]}
        ref = {"stats_i32": ref_stats.i32, "stats_f64": ref_stats.f64, "show_day_start": ref_daily.show_day_start,
               "show_order": ref_daily.show_order, "group_day_start": ref_daily.group_day_start,
               "group_offsets": ref_daily.group_offsets, "summary_f64": ref_daily.summary_f64,
               "summary_count": ref_daily.summary_count}
        if G != ref_daily.n_groups:
            differs.append("n_groups")
        differs += [k for k in got if not _same(got[k], ref[k])]
    ref_off, ref_csv = oracle_c.csv_rows(host, threads)
    if not _same(cbufs.row_offsets, ref_off):
        differs.append("csv row_offsets")
    if ref_csv.numel() != csv_total or not _same(cbufs.data[:csv_total], ref_csv):
        differs.append("csv bytes")
    return {"shows": S, "entries": E, "csv_bytes": int(ref_csv.numel()), "equal": not differs, "differs": differs,
            "compared": "stats_i32/f64 planes, show_day_start, show_order, daily groups, 19-metric summaries, CSV row "
                        "offsets and bytes of the timed table vs oracle/pie_oracle.c",
            "oracle_threads": threads, "seconds": round(time.perf_counter() - t0, 2)}


def table_differences(got, ref):
    """Names of the columns in which two archive tables differ (heaps compared up to their used length; delaySec only
    where it is valid; doubles bit for bit)."""
    import torch

    g, r = got.to("cpu"), ref.to("cpu")
    bad = []
    if g.n_shows != r.n_shows or g.n_entries != r.n_entries:
        return ["row counts"]
    S, E = r.n_shows, r.n_entries

    def col(name, a, b, n):
        used = int(b.offsets[n])
        if not (_same(a.offsets[:n + 1], b.offsets[:n + 1]) and _same(a.data[:used], b.data[:used])):
            bad.append(name)

    if not _same(g.entry_offsets[:S + 1], r.entry_offsets[:S + 1]):
        bad.append("entry_offsets")
    for k in r.show_cols:
        col(k, g.show_cols[k], r.show_cols[k], S)
    for k in r.entry_cols:
        col(k, g.entry_cols[k], r.entry_cols[k], E)
    for name, a, b, n in (("crew", g.crew, r.crew, S), ("actions", g.actions, r.actions, E)):
        if not _same(a.list_offsets[:n + 1], b.list_offsets[:n + 1]):
            bad.append(name + ".list_offsets")
        col(name + ".items", a.items, b.items, int(b.list_offsets[n]))
    for name in ("created_at", "archived_at"):
        if not _same(getattr(g, name)[:S], getattr(r, name)[:S]):
            bad.append(name)
    if not _same(g.entry_ts[:E], r.entry_ts[:E]) or not _same(g.delay_valid[:E], r.delay_valid[:E]):
        bad.append("entry_ts / delay_valid")
    v = r.delay_valid[:E].bool()
    if not _same(g.delay_sec[:E][v], r.delay_sec[:E][v]):
        bad.append("delay_sec")
    return bad


def csv_bytes(table, total_out: int) -> int:
    """Algorithmic bytes of the export rows: every column byte once (show-level columns once per
    SHOW), plus the output bytes and the int64 row offsets."""
    E, S = table.n_entries, table.n_shows
    n = sum(c.nbytes() for c in table.entry_cols.values()) + sum(c.nbytes() for c in table.show_cols.values())
    n += table.crew.nbytes() + table.actions.nbytes() + 9 * E + 4 * (S + 1)
    return n + total_out + 8 * (E + 1)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sph_pie_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from sph_pie_b200 import _lib, ops

    _lib.init(local)
    lib = _lib.load()
    affinity = bind_host_side(local, world)  # before any pinned allocation: first touch decides where host pages live

    table, generator, span = make_table(args, rank, world, local)
    S, E = table.n_shows, table.n_entries
    warmup = max(args.warmup, 3)
    bufs = ops.DailyBuffers(S, E, dev)
    sizing = ops.CsvBuffers(E, 0, dev)
    ops.csv_rows_dev(table, sizing, size_only=True)
    csv_total = int(sizing.total.cpu())
    del sizing
    cbufs = ops.CsvBuffers(E, csv_total, dev)

    def step():
        ops.show_stats_dev(table, bufs)
        ops.daily_summary_dev(table, bufs, args.tz)
        ops.csv_rows_dev(table, cbufs)

    note(f"table resident: {S} shows, {E} entries; warm-up")
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    note("warm-up done; timed steps")
    assert int(bufs.status[0]) == 0 and int(cbufs.total) == csv_total
    n_groups = int(bufs.n_groups)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: resident inputs, CUDA events on torch's current stream (the launch stream)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = int(lib.pie_kernel_launch_count())
    barrier()
    with ClockSampler(local) as clocks:
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for k in range(args.steps):
            ev[k][0].record()
            ops.show_stats_dev(table, bufs)
            ev[k][1].record()
            ops.daily_summary_dev(table, bufs, args.tz)
            ev[k][2].record()
            ops.csv_rows_dev(table, cbufs)
            ev[k][3].record()
        t_end.record()
        clocks.sample()
        barrier()
    launches = int(lib.pie_kernel_launch_count()) - launches0
    total_ms = t_start.elapsed_time(t_end)
    stats_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    daily_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    csv_ms = sum(e[2].elapsed_time(e[3]) for e in ev) / args.steps

    note(f"timed steps done: {total_ms / args.steps:.3f} ms per step; checking the timed table against the C oracle")
    # ---- parity of the TIMED configuration: what the timed steps left in HBM vs the C oracle on the host
    import oracle_c

    host_plain = table.to("cpu")
    parity = parity_check_step(host_plain, bufs, cbufs, csv_total, args.tz, max(1, oracle_c.max_threads() // world))
    note(f"parity of the timed table: equal={parity['equal']} ({parity['seconds']} s)")
    # ---- outside the step: archive entry payloads (JSON Lines) on the same resident table — the next row of
    # the scope table (DESIGN.md §0 f), timed alone with CUDA events on the launch stream
    psizing = ops.CsvBuffers(E, 0, dev)
    ops.archive_payloads_dev(table, psizing, size_only=True)
    payload_total = int(psizing.total.cpu())
    pbufs = ops.CsvBuffers(E, payload_total, dev)
    for _ in range(3):
        ops.archive_payloads_dev(table, pbufs)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_payload_runs = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    p0.record()
    for _ in range(n_payload_runs):
        ops.archive_payloads_dev(table, pbufs)
    p1.record()
    torch.cuda.synchronize()
    payload_ms = p0.elapsed_time(p1) / n_payload_runs
    payload_cols = ([table.show_cols[k] for k in ("show_date", "show_time", "show_label", "lead_pilot", "monkey_lead")]
                    + [table.entry_cols[k] for k in ("operator_name", "unit_id", "planned", "launched", "command_rx",
                                                      "primary_issue", "sub_issue")])
    payload_bytes = sum(c.nbytes() for c in payload_cols) + 4 * (S + 1) + payload_total + 8 * (E + 1)
    del pbufs, psizing
    # ... and computeMetrics(show) for every show (live show header)
    m_i32 = torch.empty((_lib.PIE_CM_COUNT, S), dtype=torch.int32, device=dev)
    m_text = torch.empty((S, _lib.PIE_CM_TEXT), dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.compute_metrics_dev(table, m_i32, m_text)
    torch.cuda.synchronize()
    p0.record()
    for _ in range(n_payload_runs):
        ops.compute_metrics_dev(table, m_i32, m_text)
    p1.record()
    torch.cuda.synchronize()
    metrics_ms = p0.elapsed_time(p1) / n_payload_runs
    metrics_bytes = (sum(table.entry_cols[k].nbytes() for k in ("planned", "status", "primary_issue")) + 9 * E
                     + 4 * (S + 1) + (4 * _lib.PIE_CM_COUNT + _lib.PIE_CM_TEXT) * S)
    del m_i32, m_text

    # ... and the rows widened in round 2: the schemaVersion 2 show payload and the provider's maintenance decisions
    widened = widened_leg(table, dev, S, E, n_payload_runs, args.tz)

    # ... and the JSON ingest of the stored documents: a sample archive written out as
    # show_archive.data texts, repeated on the device to the bench's number of shows
    # (N = 1 only, like the CPU baseline: it is reported by rank 0 and would only add pinned memory on the others)
    ingest = ingest_leg(args, dev, S, n_payload_runs, note) if world == 1 else None
    # N > 1: the device-resident part only (every rank ingests its own replica of the documents: the path partitions by
    # document, no collective), max over ranks
    ingest_multi = ingest_multi_leg(args, dev, S, n_payload_runs, world) if world > 1 else None

    note("payload rows, live metrics and JSON ingest timed; end-to-end leg (host buffers)")
    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    import ctypes as C

    host = host_plain.pin()
    hout = ops.HostOutputs(S, pinned=True)
    h_off = torch.empty(E + 1, dtype=torch.int64, pin_memory=True)
    h_csv = torch.empty(max(csv_total, 1), dtype=torch.uint8, pin_memory=True)
    hview = host.view()
    h_total = C.c_uint64(0)

    h_dout = hout.daily_out()

    def e2e_step():
        # one call: show statistics + daily summaries + CSV rows on one pipelined upload; it returns after every
        # device->host copy has completed
        _lib.check(lib.pie_archive_step_host(C.byref(hview), args.tz, hout.stats_i32.data_ptr(), hout.stats_f64.data_ptr(),
                                             hout.S, C.byref(h_dout), h_off.data_ptr(), h_csv.data_ptr(), csv_total,
                                             C.byref(h_total)))
        return _lib.last_transfer_bytes()

    profiling = bool(os.environ.get("PIE_BENCH_PROFILE"))  # the ncu launch-list pass: a launch costs seconds there
    for _ in range(1 if profiling else 2):
        h2d, d2h = e2e_step()
    e2e_steps = 1 if profiling else max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    note(f"end-to-end leg done: {e2e_s / e2e_steps * 1e3:.1f} ms per step")
    # the same call with PAGEABLE caller buffers in and out — what an N-API ArrayBuffer is; the
    # driver stages pageable copies through its own pinned bounce buffers
    e2e_pageable = None
    if world == 1 and not profiling:
        pout = ops.HostOutputs(S, pinned=False)
        p_off = torch.empty(E + 1, dtype=torch.int64)
        p_csv = torch.empty(max(csv_total, 1), dtype=torch.uint8)
        pview, p_dout = host_plain.view(), pout.daily_out()

        def pageable_step():
            _lib.check(lib.pie_archive_step_host(C.byref(pview), args.tz, pout.stats_i32.data_ptr(), pout.stats_f64.data_ptr(),
                                                 pout.S, C.byref(p_dout), p_off.data_ptr(), p_csv.data_ptr(), csv_total,
                                                 C.byref(h_total)))

        pageable_step()
        p_steps = 3
        t0 = time.perf_counter()
        for _ in range(p_steps):
            pageable_step()
        p_s = (time.perf_counter() - t0) / p_steps
        e2e_pageable = {"value": E / p_s, "unit": UNIT, "ms_per_step": p_s * 1e3, "steps": p_steps,
                        "same_result_as_pinned": bool(torch.equal(p_csv, h_csv) and torch.equal(p_off, h_off)),
                        "what": "pie_archive_step_host with pageable (malloc'd) caller buffers in and out"}
        del pout, p_off, p_csv
        note(f"pageable end-to-end: {p_s * 1e3:.1f} ms per step")
    # the host-side ceiling of that leg: nothing but the same bytes over PCIe — pinned H2D and D2H at once, on every rank
    # at the same time (what N ranks sharing one host can move at best)
    up = torch.empty(max(h2d, 1), dtype=torch.uint8, pin_memory=True)
    down = torch.empty(max(d2h, 1), dtype=torch.uint8, pin_memory=True)
    d_up, d_down = torch.empty_like(up, device=dev), torch.empty_like(down, device=dev)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy_only():
        with torch.cuda.stream(s_up):
            d_up.copy_(up, non_blocking=True)
        with torch.cuda.stream(s_down):
            down.copy_(d_down, non_blocking=True)
        torch.cuda.synchronize()

    copy_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        copy_only()
    ceiling_s = (time.perf_counter() - t0) / 3
    del up, down, d_up, d_down
    my_e2e_ms = e2e_s / e2e_steps * 1e3
    per_rank_ms, per_rank_ceiling_ms, spans = [my_e2e_ms], [ceiling_s * 1e3], [list(span)]
    n_groups_total, days_disjoint = n_groups, True
    if world > 1:
        import torch.distributed as dist

        mine = torch.tensor([my_e2e_ms, ceiling_s * 1e3, float(span[0]), float(span[1]), float(n_groups),
                             float(bufs.group_day_start[0]) if n_groups else 0.0,
                             float(bufs.group_day_start[n_groups - 1]) if n_groups else 0.0], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        every = [x.cpu().tolist() for x in every]
        per_rank_ms = [x[0] for x in every]
        per_rank_ceiling_ms = [x[1] for x in every]
        spans = [[int(x[2]), int(x[3])] for x in every]
        n_groups_total = int(sum(x[4] for x in every))
        # the stitched daily table is the ranks' tables one after the other: their day ranges must not meet
        days_disjoint = all(every[r][6] < every[r + 1][5] for r in range(world - 1) if every[r][4] and every[r + 1][4])
        ceiling_s = max(per_rank_ceiling_ms) / 1e3
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([total_ms, e2e_s, float(E), 0.0 if parity["equal"] else 1.0, float(S)], dtype=torch.float64, device=dev)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, total_entries = float(mx[0]), float(mx[1]), float(sm[2])
        parity = dict(parity, equal=float(mx[3]) == 0.0, shows=int(sm[4]), entries=int(sm[2]),
                      note="every rank checked its own timed table; equal = all ranks equal; differs = rank 0's list")
    else:
        total_entries = float(E)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    value = total_entries * args.steps / (total_ms * 1e-3)
    e2e_value = total_entries * e2e_steps / e2e_s
    stats_bytes, daily_bytes = analytics_bytes(table, n_groups)
    export_bytes = csv_bytes(table, csv_total)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    def gbs(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9

    # DRAM traffic of the kernels the roofline names — the whole export-row group, not the row kernel alone — from
    # the ncu launch list of this same command (scripts/traffic_from_launches.py writes the file from the capture;
    # `traffic_build_current` says whether the capture was taken on the sources this library was built from)
    import __graft_entry__ as entry

    traffic, traffic_kernels, traffic_current, ingest_traffic = None, None, None, None
    tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        grp = tj.get("groups", {}).get("export_rows_csv")
        if tj.get("shows") == S and grp:
            traffic = grp["dram_bytes_read"] + grp["dram_bytes_write"]
            traffic_kernels = grp["kernels"]
            # current = the capture was taken on the sources THIS group's kernels are compiled from (csv_rows.cu and what it
            # includes, same flags); another translation unit may have changed since (csrc_sha16 is the whole tree)
            per_group = tj.get("group_src_sha16", {}).get("export_rows_csv")
            traffic_current = (per_group == entry.group_src_sha16("export_rows_csv")) if per_group else \
                (tj.get("csrc_sha16") == entry.csrc_sha16())
            ig = tj.get("groups", {}).get("ingest")
            ingest_traffic = ig["dram_bytes_read"] + ig["dram_bytes_write"] if ig else None

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32/f64", "data": "synthetic",
        "config": workload_config(args, S, E, generator),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "pie_archive_step_host (statistics + daily summaries + CSV rows; pinned host buffers in and out)",
                "pageable": e2e_pageable, "per_rank_ms": per_rank_ms,
                "copy_ceiling": {"value": total_entries / ceiling_s, "unit": UNIT, "per_rank_ms": per_rank_ceiling_ms,
                                 "what": "the same H2D + D2H bytes per rank as plain pinned copies, both directions at once, "
                                         "all ranks at the same time: the host side's best case for this leg"},
                "host_binding": affinity},
        "sharding": {"archive": f"one archive of {world} segment(s) x {args.shows} shows, day ranges balanced by entries "
                                "(sharding.plan_day_shards); no data-path collective",
                     "show_ranges": spans, "n_groups_total": n_groups_total, "day_ranges_disjoint": days_disjoint},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "parity_checked": parity,
        "roofline": {  # the dominant kernel group of the step: the export rows (pie_csv_rows_dev)
            "bound": "hbm", "kernel": "export rows: every kernel pie_csv_rows_dev launches (export_rows_kernel<csv> and its "
                                      "helper kernels), timed as one group with CUDA events",
            "achieved": gbs(export_bytes, csv_ms), "peak": peak, "unit": "GB/s", "frac": gbs(export_bytes, csv_ms) / peak,
            "traffic": traffic, "traffic_kernels": traffic_kernels, "traffic_build_current": traffic_current,
            "traffic_source": "profiles/traffic_r02.json (ncu dram__bytes_read + dram__bytes_write, summed over the "
                              "group's kernels, per launch of the group; scripts/traffic_from_launches.py)",
            "peak_source": peak_src, "algorithmic_bytes_per_launch": export_bytes,
            "ms_per_launch": csv_ms, "bytes_per_entry": export_bytes / max(E, 1), "csv_bytes_out": csv_total,
            "other_kernels": {
                "show_stats_kernel": {"ms_per_launch": stats_ms, "algorithmic_bytes": stats_bytes,
                                      "achieved_gbs": gbs(stats_bytes, stats_ms), "frac": gbs(stats_bytes, stats_ms) / peak},
                "daily pipeline (7 small kernels)": {"ms_per_step": daily_ms, "algorithmic_bytes": daily_bytes,
                                                     "achieved_gbs": gbs(daily_bytes, daily_ms),
                                                     "frac": gbs(daily_bytes, daily_ms) / peak},
                "archive entry payloads (export_rows_kernel<json>, not part of the step)": {
                    "ms_per_launch": payload_ms, "algorithmic_bytes": payload_bytes, "json_bytes_out": payload_total,
                    "achieved_gbs": gbs(payload_bytes, payload_ms), "frac": gbs(payload_bytes, payload_ms) / peak,
                    "entries_per_s": E / (payload_ms * 1e-3)},
                "JSON ingest of stored documents (ingest_fast_kernel x2 + ingest_walk_kernel x2 + scans, not part of the step)":
                    dict(ingest, frac=ingest["achieved_gbs"] / peak, traffic=ingest_traffic) if ingest else ingest_multi,
                "schemaVersion 2 show payloads (payload_measure_kernel + scan + payload_write_kernel, not part of the step)":
                    dict(widened["show_payloads"], frac=widened["show_payloads"]["achieved_gbs"] / peak),
                "provider maintenance (get_timestamps + _archiveDailyShows + _purgeExpiredArchives decisions, not part of the step)":
                    widened["maintenance"],
                "computeMetrics per show (compute_metrics_kernel, not part of the step)": {
                    "ms_per_launch": metrics_ms, "algorithmic_bytes": metrics_bytes,
                    "achieved_gbs": gbs(metrics_bytes, metrics_ms), "frac": gbs(metrics_bytes, metrics_ms) / peak,
                    "entries_per_s": E / (metrics_ms * 1e-3)},
            },
        },
    }

    # ---- CPU baseline (rank 0, N=1 only): the C port on one core, bounded sample of the same workload
    if world == 1:
        sample_shows = min(S, args.cpu_sample_shows)
        sample = host.slice_shows(0, sample_shows)
        v, reps, secs = cpu_baseline_run(sample, args.tz, 1, min_seconds=2.0)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"first {sample_shows} shows ({sample.n_entries} entries) of the workload, "
                                         f"{reps} passes in {secs:.2f} s, oracle/pie_oracle.c single thread"}
        if not profiling:
            out["reference_scale"] = reference_scale(args, dev)  # not in `config`: both arms print the same config
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


def widened_leg(table, dev, S, E, runs, tz):
    """CUDA-event times of pie_show_payloads_dev (sizing call + writing call, as a caller makes them) and of
    pie_get_timestamps_dev + pie_archive_due_dev + pie_archive_expired_dev on the resident bench table."""
    import ctypes as C

    import torch

    from sph_pie_b200 import _lib, ops
    from sph_pie_b200.webhook import payload_frame

    lib = _lib.load()
    head, tail = payload_frame("show.updated", "2024-07-03T12:00:00.000Z", "https://example.invalid/hook", "POST", None)
    first = ops.show_payloads(table, head, tail)          # sizes, allocates, writes; also the warm-up
    total = int(first.data.numel())
    h = torch.tensor(list(head), dtype=torch.uint8, device=dev)
    t = torch.tensor(list(tail), dtype=torch.uint8, device=dev)
    offs = torch.zeros(S + 1, dtype=torch.int64, device=dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    scratch = torch.empty(int(lib.pie_show_payloads_scratch_bytes(S)) + 256, dtype=torch.uint8, device=dev)
    sptr = (scratch.data_ptr() + 255) & ~255
    view = table.view()
    stream = torch.cuda.current_stream().cuda_stream
    data = first.data

    def payload_call():
        _lib.check(lib.pie_show_payloads_dev(C.byref(view), h.data_ptr(), len(head), t.data_ptr(), len(tail), offs.data_ptr(),
                                             data.data_ptr(), total, tot.data_ptr(), status.data_ptr(), sptr, stream))

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    payload_call()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(runs):
        payload_call()
    e1.record()
    torch.cuda.synchronize()
    payload_ms = e0.elapsed_time(e1) / runs
    assert status.cpu().tolist()[0] == 0 and int(tot.cpu()) == total and torch.equal(offs, first.doc_offsets)
    in_bytes = table.nbytes()                              # every column once
    payload_bytes = in_bytes + total + 8 * (S + 1)
    del first, data, scratch

    # maintenance decisions on the same table
    created = ops.get_timestamps(table, None, tz).created_at
    now_ms = float(created[~torch.isnan(created)].max().cpu()) + 13 * 3600e3 if S else 0.0
    ops.archive_due(table, created, now_ms)
    ops.archive_expired(created, now_ms, tz)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(runs):
        times = ops.get_timestamps(table, None, tz)       # includes its status read-back (a sync), as the caller pays it
        due, _first = ops.archive_due(table, times.created_at, now_ms)
        expired = ops.archive_expired(times.created_at, now_ms, tz)
    e1.record()
    torch.cuda.synchronize()
    maint_ms = e0.elapsed_time(e1) / runs
    return {
        "show_payloads": {"ms_per_call": payload_ms, "algorithmic_bytes": payload_bytes, "json_bytes_out": total,
                          "achieved_gbs": payload_bytes / (payload_ms * 1e-3) / 1e9, "shows_per_s": S / (payload_ms * 1e-3),
                          "entries_per_s": E / (payload_ms * 1e-3),
                          "what": "one call: measure pass + scan + write pass over every column; documents of ~"
                                  f"{total // max(S, 1)} bytes",
                          "note": "kernels rebuilt in the last session of round 2 (a warp stages its show in shared memory, "
                                  "pie_show_payload.cuh); their first version took 2198.8 ms on 2^20 shows "
                                  "(profiles/bench_r02_h_final.json, profiles/show_payload_r02.md)"},
        "maintenance": {"ms_per_round": maint_ms, "shows_per_s": S / (maint_ms * 1e-3), "due": int(due.sum().cpu()),
                        "expired": int(expired.sum().cpu()),
                        "what": "pie_get_timestamps_dev + pie_archive_due_dev + pie_archive_expired_dev, S-sized "
                                "(no per-entry work), through the Python operators incl. their allocations and one status read-back"},
    }


def ingest_multi_leg(args, dev, n_shows, runs, world):
    """N > 1: pie_ingest_measure_dev + pie_ingest_fill_dev on every rank's own copy of the bench's documents (CUDA events
    per rank, max over ranks; the aggregate is N times the documents over that time: replicas, no collective)."""
    import torch
    import torch.distributed as dist

    from sph_pie_b200 import ops
    from sph_pie_b200.synth import synth_stored_docs

    sample = min(n_shows, 8192)
    copies = max(1, n_shows // sample)
    docs, n_entries, text_bytes, _ = synth_stored_docs(sample, copies, dev, seed=4321)
    bufs = ops.IngestBuffers(docs.n_docs, dev)
    ops.ingest_measure_dev(docs, bufs)
    totals = bufs.totals.cpu().tolist()
    assert bufs.status.cpu().tolist()[0] == 0
    table = ops.alloc_ingest_table(docs.n_docs, totals, dev)
    table_bytes = table.nbytes()
    for _ in range(2):
        ops.ingest_measure_dev(docs, bufs)
        ops.ingest_fill_dev(docs, bufs, table)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    dist.barrier()
    ev[0].record()
    for _ in range(runs):
        ops.ingest_measure_dev(docs, bufs)
        ops.ingest_fill_dev(docs, bufs, table)
    ev[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / runs], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    del docs, bufs, table
    return {"ms_per_launch": ms, "documents_per_rank": sample * copies, "entries_per_rank": n_entries, "ranks": world,
            "entries_per_s": world * n_entries / (ms * 1e-3), "text_gbs": world * text_bytes / (ms * 1e-3) / 1e9,
            "achieved_gbs": world * (text_bytes + table_bytes) / (ms * 1e-3) / 1e9,
            "what": "device-resident texts, every rank its own replica of the bench's documents, max over ranks; the host "
                    "legs, the parity check and the second workload are reported at N = 1"}


def ingest_leg(args, dev, n_shows, runs, note):
    """pie_ingest_measure_dev + pie_ingest_fill_dev on device-resident texts (CUDA events), then pie_ingest_host with
    pinned host texts in and the host table out (wall clock), then json.loads of the sample on one core."""
    import ctypes as C

    import torch

    from sph_pie_b200 import _lib, ops
    from sph_pie_b200.synth import synth_stored_docs

    sample = min(n_shows, 8192)
    copies = max(1, n_shows // sample)
    docs, n_entries, text_bytes, texts = synth_stored_docs(sample, copies, dev, seed=4321)
    bufs = ops.IngestBuffers(docs.n_docs, dev)
    ops.ingest_measure_dev(docs, bufs)
    totals = bufs.totals.cpu().tolist()
    assert bufs.status.cpu().tolist()[0] == 0
    table = ops.alloc_ingest_table(docs.n_docs, totals, dev)
    table_bytes = table.nbytes()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def timed(d, b, t, n_runs):
        for _ in range(2):
            ops.ingest_measure_dev(d, b)
            ops.ingest_fill_dev(d, b, t)
        m = f = 0.0
        for _ in range(n_runs):
            ev[0].record()
            ops.ingest_measure_dev(d, b)
            ev[1].record()
            ops.ingest_fill_dev(d, b, t)
            ev[2].record()
            torch.cuda.synchronize()
            m += ev[0].elapsed_time(ev[1]) / n_runs
            f += ev[1].elapsed_time(ev[2]) / n_runs
        return m, f

    # the thread-per-document walk alone (the round-1 pipeline) beside the default (a warp per document, the walk
    # for what it declines): same documents, same table
    # PIE_BENCH_PROFILE=1 (the ncu launch-list pass: every launch costs seconds there) keeps to the default path
    profiling = bool(os.environ.get("PIE_BENCH_PROFILE"))
    old_path = ops.set_ingest_warp_path(0)
    wm, wf = (0.0, 0.0) if profiling else timed(docs, bufs, table, runs)
    ops.set_ingest_warp_path(1)
    tm, tf = timed(docs, bufs, table, runs)
    declined = ops.ingest_declined(bufs, docs.n_docs)
    paths = {"warp_per_document": {"measure_ms": tm, "fill_ms": tf, "ms": tm + tf, "declined_to_the_walk": declined},
             "thread_per_document_walk": {"measure_ms": wm, "fill_ms": wf, "ms": wm + wf},
             "note": "the walk is at its best here: the copies of a document are neighbours in its length order, so a "
                     "warp's 32 lanes walk identical documents; different_documents below is the other case"}
    note(f"JSON ingest on the device: {tm:.2f} + {tf:.2f} ms for {text_bytes / 1e9:.2f} GB of text")
    # parity of the timed ingest: the table the timed walks left in HBM against the C oracle's own parser
    # (oracle_ingest_measure + oracle_ingest_fill, recursive descent + strtod) on ALL the timed documents
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c

    t0 = time.perf_counter()
    hdocs_plain = docs.to("cpu")
    ref_table, ref_status, ref_err = oracle_c.ingest(hdocs_plain, nthreads=oracle_c.max_threads())
    differs = ["oracle status %r" % (ref_err,)] if ref_err != (0, -1) else table_differences(table, ref_table)
    if ref_err == (0, -1) and not torch.equal(bufs.doc_status[:docs.n_docs].cpu(), torch.from_numpy(ref_status)):
        differs.append("doc_status")
    ingest_parity = {"documents": docs.n_docs, "entries": n_entries, "equal": not differs, "differs": differs,
                     "compared": "every column of the ingested table + doc_status vs oracle_ingest_* on the timed documents",
                     "seconds": round(time.perf_counter() - t0, 2)}
    note(f"parity of the timed ingest: equal={ingest_parity['equal']} ({ingest_parity['seconds']} s)")
    del ref_table
    # how much the documents decide: the same number of documents, 8x as many different ones
    different = None
    if n_shows >= 8 * sample and not profiling:
        sample2 = 8 * sample
        docs2, n_entries2, text_bytes2, _ = synth_stored_docs(sample2, max(1, n_shows // sample2), dev, seed=8765)
        bufs2 = ops.IngestBuffers(docs2.n_docs, dev)
        ops.ingest_measure_dev(docs2, bufs2)
        totals2 = bufs2.totals.cpu().tolist()
        assert bufs2.status.cpu().tolist()[0] == 0
        table2 = ops.alloc_ingest_table(docs2.n_docs, totals2, dev)
        ops.set_ingest_warp_path(0)
        wm2, wf2 = timed(docs2, bufs2, table2, max(2, runs // 2))
        ops.set_ingest_warp_path(1)
        tm2, tf2 = timed(docs2, bufs2, table2, max(2, runs // 2))
        different = {"workload": f"{sample2} different documents x{max(1, n_shows // sample2)}", "documents": docs2.n_docs,
                     "entries": n_entries2, "text_bytes": text_bytes2,
                     "warp_per_document": {"measure_ms": tm2, "fill_ms": tf2, "ms": tm2 + tf2,
                                           "declined_to_the_walk": ops.ingest_declined(bufs2, docs2.n_docs)},
                     "thread_per_document_walk": {"measure_ms": wm2, "fill_ms": wf2, "ms": wm2 + wf2}}
        note(f"JSON ingest, {sample2} different documents: warp path {tm2 + tf2:.2f} ms, walk {wm2 + wf2:.2f} ms")
        del docs2, bufs2, table2
    ops.set_ingest_warp_path(old_path)
    # host buffers through the C ABI
    lib = _lib.load()
    hdocs = hdocs_plain.pin()
    del table, docs, hdocs_plain
    d = hdocs.c()
    view = _lib.ArchiveViewC()
    st = torch.empty(hdocs.n_docs, dtype=torch.uint8)
    tot = torch.zeros(_lib.PIE_INGEST_TOTALS, dtype=torch.int64)
    bad = C.c_int64(-1)

    def host_call():
        _lib.check(lib.pie_ingest_host(C.byref(d), C.byref(view), st.data_ptr(), tot.data_ptr(), C.byref(bad)))

    host_call()
    host_call()
    h_runs = max(2, min(runs, 4))
    t0 = time.perf_counter()
    for _ in range(h_runs):
        host_call()
    host_s = (time.perf_counter() - t0) / h_runs
    h2d, d2h = _lib.last_transfer_bytes()
    lib.pie_ingest_host_release()
    # the whole workspace from the stored texts: upload, ingest, statistics + daily summaries + CSV rows on the
    # device-resident table, results back in pinned host buffers (ops.archive_step_from_json)
    st, dl, rows, _ = ops.archive_step_from_json(hdocs, args.tz)
    hout = ops.HostOutputs(hdocs.n_docs, pinned=True)
    h_off = torch.empty(rows.row_offsets.numel(), dtype=torch.int64, pin_memory=True)
    h_csv = torch.empty(rows.data.numel(), dtype=torch.uint8, pin_memory=True)
    del st, dl
    ops.archive_step_from_json(hdocs, args.tz, hout, h_off, h_csv)
    t0 = time.perf_counter()
    for _ in range(h_runs):
        ops.archive_step_from_json(hdocs, args.tz, hout, h_off, h_csv)
    json_step_s = (time.perf_counter() - t0) / h_runs
    ops.archive_step_from_json_pipelined(hdocs, args.tz, hout, h_off, h_csv)
    t0 = time.perf_counter()
    for _ in range(h_runs):
        ops.archive_step_from_json_pipelined(hdocs, args.tz, hout, h_off, h_csv)
    json_pipe_s = (time.perf_counter() - t0) / h_runs
    ops.archive_step_json_host(hdocs, args.tz, hout, h_off, h_csv)
    t0 = time.perf_counter()
    for _ in range(h_runs):
        ops.archive_step_json_host(hdocs, args.tz, hout, h_off, h_csv)
    json_abi_s = (time.perf_counter() - t0) / h_runs
    json_step_d2h = h_csv.numel() + 8 * h_off.numel() + (4 * _lib.PIE_SI_COUNT + 8 * _lib.PIE_SF_COUNT) * hdocs.n_docs
    del rows, hout, h_off, h_csv
    # one core of the host through Python's json module (C accelerated; parse only, no projection on the table)
    t0 = time.perf_counter()
    parsed = 0
    while time.perf_counter() - t0 < 2.0:
        for t in texts:
            json.loads(t)
        parsed += 1
    cpu_s = (time.perf_counter() - t0) / parsed
    sample_bytes = sum(len(t.encode("utf-8")) for t in texts)
    # ... and the C port of the same path (oracle/pie_oracle.c: recursive descent + strtod, both passes, table out)
    sample_docs = ops.JsonDocs.from_texts(texts)
    sample_entries = n_entries // copies
    port = {}
    for label, threads in (("one_thread", 1), ("all_threads", oracle_c.max_threads())):
        oracle_c.ingest(sample_docs, nthreads=threads)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 1.5:
            table_c, _, err = oracle_c.ingest(sample_docs, nthreads=threads)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        assert err == (0, -1) and table_c.n_entries == sample_entries
        port[label] = {"threads": threads, "text_mbs": sample_bytes / dt / 1e6, "entries_per_s": sample_entries / dt}
    # algorithmic bytes: the text ONCE + the table once (that the implementation walks the text twice is its traffic,
    # not the algorithm's)
    alg = text_bytes + table_bytes
    ms = tm + tf
    return {
        "ms_per_launch": ms, "measure_ms": tm, "fill_ms": tf, "documents": hdocs.n_docs, "entries": n_entries,
        "paths": paths, "different_documents": different,
        "text_bytes": text_bytes, "table_bytes": table_bytes, "algorithmic_bytes": alg,
        "two_pass_bytes": 2 * text_bytes + table_bytes + 2 * 4 * 26 * hdocs.n_docs, "parity_checked": ingest_parity,
        "achieved_gbs": alg / (ms * 1e-3) / 1e9, "text_gbs": text_bytes / (ms * 1e-3) / 1e9,
        "entries_per_s": n_entries / (ms * 1e-3),
        "e2e_host": {"api": "pie_ingest_host (pinned texts in, host table out)", "ms": host_s * 1e3,
                     "entries_per_s": n_entries / host_s, "text_gbs": text_bytes / host_s / 1e9,
                     "h2d_bytes": h2d, "d2h_bytes": d2h},
        "e2e_json_to_outputs": {"api": "ops.archive_step_from_json: pinned texts in; ingest, statistics, daily summaries "
                                       "and CSV rows on the device; pinned results out",
                                "ms": json_step_s * 1e3, "entries_per_s": n_entries / json_step_s,
                                "c_abi_ms": json_abi_s * 1e3, "c_abi_entries_per_s": n_entries / json_abi_s,
                                "c_abi": "pie_archive_step_json_host: the same as one C-ABI call (single batch)",
                                "pipelined_ms": json_pipe_s * 1e3, "pipelined_entries_per_s": n_entries / json_pipe_s,
                                "pipelined_api": "ops.archive_step_from_json_pipelined: the same in chunks of 131072 "
                                                 "documents over three streams (upload / kernels / download overlap)",
                                "h2d_bytes": text_bytes + 8 * (hdocs.n_docs + 1), "d2h_bytes_at_least": json_step_d2h},
        "cpu_port": dict(port, what="oracle/pie_oracle.c oracle_ingest_measure + oracle_ingest_fill on the sample's documents "
                                    "(JSON text in, columnar table out), kind: port"),
        "cpu_json_loads": {"what": "json.loads of the sample's documents, one core, parse only", "text_mbs":
                           sample_bytes / cpu_s / 1e6, "sample_documents": len(texts)},
        "workload": f"{sample} synthetic shows written as JSON documents (json.dumps, no whitespace), x{copies} on the device",
    }


def reference_scale(args, dev):
    """The same step at the largest archive the reference's own rules allow (310 shows x <=21
    entries): GPU end-to-end vs one CPU core.  Reported so nobody reads the headline
    as a speed-up of the reference."""
    import torch

    import oracle_c
    from sph_pie_b200 import ops
    from sph_pie_b200.synth import synth_archive

    import ctypes as C

    from sph_pie_b200 import _lib

    small = synth_archive(310, seed=99, device="cpu")
    pinned = small.pin()
    hout = ops.HostOutputs(small.n_shows, pinned=True)
    cpu = CpuStep(small, args.tz, 1)
    off = torch.empty(small.n_entries + 1, dtype=torch.int64, pin_memory=True)
    data = torch.empty(cpu.csv_data.numel(), dtype=torch.uint8, pin_memory=True)
    view, total, lib = pinned.view(), C.c_uint64(0), _lib.load()

    dout = hout.daily_out()

    def gpu_step():
        _lib.check(lib.pie_archive_step_host(C.byref(view), args.tz, hout.stats_i32.data_ptr(), hout.stats_f64.data_ptr(),
                                             hout.S, C.byref(dout), off.data_ptr(), data.data_ptr(), data.numel(),
                                             C.byref(total)))

    for _ in range(5):
        gpu_step()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        gpu_step()
    gpu_us = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for _ in range(n):
        cpu()
    cpu_us = (time.perf_counter() - t0) / n * 1e6
    return {"shows": small.n_shows, "entries": small.n_entries, "gpu_e2e_us_per_call": gpu_us,
            "cpu_port_1core_us_per_call": cpu_us,
            "note": "the largest archive the reference's own rules allow: both are millisecond-scale, once-per-click "
                    "operations; the GPU call is launch/copy-latency bound here (DESIGN.md section 0)"}


if __name__ == "__main__":
    main()
