"""Multi-GPU partitioning of an archive: one process per GPU, shards by DAY RANGE.

The path has no exchange step: show statistics are per show, and a daily group never spans two
calendar days.  So an archive ordered by day is cut at day boundaries into `world` contiguous
show ranges (balanced by entry count); every rank runs the single-GPU operators on its range and the
per-day tables are simply concatenated in rank order.  No collective touches the data path;
`torch.distributed` is used only to hand the (small) result tables to whoever wants them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Tuple

import torch

from .columnar import ArchiveTable

PIE_DAY_NONE = -(2 ** 63)  # include/sph_pie_b200.h


@dataclass
class ShardPlan:
    bounds: List[Tuple[int, int]]  # [s0, s1) per rank; empty ranges allowed

    def range_of(self, rank: int) -> Tuple[int, int]:
        return self.bounds[rank]


def plan_day_shards(entry_offsets: torch.Tensor, show_day: torch.Tensor, world: int) -> ShardPlan:
    """Cut shows [0, S) into `world` contiguous ranges whose borders fall on day changes.

    `show_day`: any per-show key that is constant within a day and non-decreasing over the archive
    (e.g. pie_daily_out.show_day_start of an ordered archive).  Raises if the archive is not ordered
    by day: an unordered archive must be sorted (or processed on one GPU) first."""
    S = show_day.numel()
    if world < 1:
        raise ValueError("world must be >= 1")
    day = show_day.cpu().clone()
    eo = entry_offsets.cpu().to(torch.int64)
    if S == 0:
        return ShardPlan([(0, 0)] * world)
    # A show without a usable timestamp (PIE_DAY_NONE: buildArchiveDailyGroups skips it, public/app.js:3408-3411; a
    # row the JSON ingest dropped is one) belongs to no day: it takes the day of the show before it — it may sit in
    # any shard, its statistics are per show and no group contains it — and neither breaks the order nor makes a cut.
    none = day == PIE_DAY_NONE
    if bool(none.any()):
        idx = torch.arange(S)
        last_valid = torch.cummax(torch.where(none, torch.full_like(idx, -1), idx), 0).values
        first_valid = int((~none).nonzero()[0]) if bool((~none).any()) else 0
        day = day[torch.where(last_valid < 0, torch.full_like(idx, first_valid), last_valid)]
    if S > 1 and bool((day[1:] < day[:-1]).any()):
        raise ValueError("archive is not ordered by day: cannot shard by day range")
    # candidate cut points: show indices where a new day starts
    starts = torch.nonzero(day[1:] != day[:-1]).flatten() + 1
    cuts = torch.cat([torch.zeros(1, dtype=torch.int64), starts, torch.tensor([S], dtype=torch.int64)])
    total = int(eo[S])
    bounds, prev = [], 0
    for r in range(1, world):
        target = total * r // world
        # first day boundary whose entry offset reaches the target, not before the previous cut
        pos = int(torch.searchsorted(eo[cuts], torch.tensor(target)))
        pos = min(max(pos, 0), cuts.numel() - 1)
        cut = max(int(cuts[pos]), prev)
        bounds.append((prev, cut))
        prev = cut
    bounds.append((prev, S))
    return ShardPlan(bounds)


def segment_skeleton(shows_per_segment: int, seeds: List[int], device):
    """(entry_offsets int64[S+1], day int64[S]) of ONE synthetic archive made of len(seeds) stored segments —
    segment k = synth_archive(shows_per_segment, seed=seeds[k]) moved to its own range of days — without building any
    of them (synth.synth_skeleton): what every rank needs to plan the day shards."""
    from .synth import synth_skeleton

    days_per_segment = (shows_per_segment + 4) // 5
    n_per, days = [], []
    for k, seed in enumerate(seeds):
        n, d = synth_skeleton(shows_per_segment, seed, device)
        n_per.append(n)
        days.append(d + k * days_per_segment)
    n_per = torch.cat(n_per)
    eo = torch.zeros(n_per.numel() + 1, dtype=torch.int64, device=n_per.device)
    torch.cumsum(n_per, 0, out=eo[1:])
    return eo, torch.cat(days)


def assemble_shard(shows_per_segment: int, seeds: List[int], entry_offsets: torch.Tensor, day: torch.Tensor, rank: int,
                   world: int, device, start_ms: int = 1704067200000):
    """This rank's day range of the segmented archive as one compact table on `device`: the shard plan comes from the
    skeleton; only the segments the range touches are generated (a range balanced by entries rarely ends on a segment
    border), sliced and joined (columnar.concat_tables).  Returns (table, (s0, s1))."""
    from .columnar import concat_tables
    from .synth import synth_archive

    s0, s1 = plan_day_shards(entry_offsets, day, world).range_of(rank)
    days_per_segment = (shows_per_segment + 4) // 5

    def segment(k):
        return synth_archive(shows_per_segment, seed=seeds[k], device=device, start_ms=start_ms + k * days_per_segment * 86400000)

    parts = []
    for k in range(len(seeds)):
        a, b = max(s0, k * shows_per_segment), min(s1, (k + 1) * shows_per_segment)
        if a < b:
            parts.append(segment(k).slice_shows(a - k * shows_per_segment, b - k * shows_per_segment))
    if not parts:  # an empty range (more ranks than days): an empty table of the right shape
        parts.append(segment(0).slice_shows(0, 0))
    return concat_tables(parts), (s0, s1)


def concat_daily(parts: List[dict]) -> dict:
    """Concatenate per-rank daily tables (rank order == day order)."""
    keys = ("group_day_start", "summary_f64", "summary_count")
    out = {"n_groups": sum(p["n_groups"] for p in parts)}
    out["group_day_start"] = torch.cat([p["group_day_start"] for p in parts])
    out["summary_f64"] = torch.cat([p["summary_f64"] for p in parts], dim=2)
    out["summary_count"] = torch.cat([p["summary_count"] for p in parts], dim=1)
    out["group_sizes"] = torch.cat([p["group_sizes"] for p in parts])
    out["stats_i32"] = torch.cat([p["stats_i32"] for p in parts], dim=1)
    out["stats_f64"] = torch.cat([p["stats_f64"] for p in parts], dim=1)
    for extra in set(parts[0]) - set(out) - {"range"}:  # whatever else the ranks hand over (e.g. their CSV bytes)
        if isinstance(parts[0][extra], torch.Tensor):
            out[extra] = torch.cat([p[extra] for p in parts])
    del keys
    return out


def run_sharded(table: ArchiveTable, show_day: torch.Tensor, rank: int, world: int,
                compute: Callable[[ArchiveTable], tuple]) -> dict:
    """Run `compute` (ArchiveTable -> (ShowStats, DailySummary)) on this rank's day range and return
    the local result tables as CPU tensors."""
    plan = plan_day_shards(table.entry_offsets, show_day, world)
    s0, s1 = plan.range_of(rank)
    local = table.slice_shows(s0, s1)
    st, daily = compute(local)
    go = daily.group_offsets.cpu()
    return {
        "range": (s0, s1),
        "n_groups": daily.n_groups,
        "group_day_start": daily.group_day_start.cpu().clone(),
        "group_sizes": (go[1:] - go[:-1]).clone(),
        "summary_f64": daily.summary_f64.cpu().clone(),
        "summary_count": daily.summary_count.cpu().clone(),
        "stats_i32": st.i32.cpu().clone(),
        "stats_f64": st.f64.cpu().clone(),
    }


def gather_to_rank0(local: dict, rank: int, world: int):
    """Hand the result tables to rank 0 (torch.distributed object gather; gloo or nccl)."""
    import torch.distributed as dist

    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    return concat_daily(gathered)
