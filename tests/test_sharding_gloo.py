"""The N>1 path on CPU: world_size 2 (and 3) over gloo.  Shards are day ranges, there is no
data-path collective; the compute backend here is the C oracle (no GPU in CI), which exercises
the planning, slicing and stitching logic that the GPU ranks use unchanged."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_shows, tz, ret, holes=False):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import oracle_c
    from sph_pie_b200.sharding import gather_to_rank0, run_sharded
    from sph_pie_b200.synth import synth_archive

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    table = synth_archive(n_shows, seed=21)  # same table on every rank (seeded)
    if holes:  # shows without any usable timestamp in the middle of the archive (a dropped row is one): no day at all
        from sph_pie_b200.columnar import pack_shows
        from sph_pie_b200.synth import table_to_shows

        shows = table_to_shows(table)
        for i in (0, 7, 8, n_shows // 2, n_shows - 1):
            shows[i] = None
        shows[11] = dict(shows[11], createdAt=None, archivedAt=None, date="", entries=[])
        table = pack_shows(shows)

    def compute(t):
        st, daily, rc, _ = oracle_c.archive_analytics(t, tz)
        assert rc == 0
        return st, daily

    whole_st, whole_daily = compute(table)
    if holes:
        assert int((whole_daily.show_day_start == -(2 ** 63)).sum()) == 6
    local = run_sharded(table, whole_daily.show_day_start, rank, world, compute)
    merged = gather_to_rank0(local, rank, world)
    if rank == 0:
        go = whole_daily.group_offsets
        ok = (merged["n_groups"] == whole_daily.n_groups
              and torch.equal(merged["group_day_start"], whole_daily.group_day_start)
              and torch.equal(merged["group_sizes"], go[1:] - go[:-1])
              and torch.equal(merged["summary_count"], whole_daily.summary_count)
              and torch.equal(merged["summary_f64"].view(torch.int64), whole_daily.summary_f64.contiguous().view(torch.int64))
              and torch.equal(merged["stats_i32"], whole_st.i32)
              and torch.equal(merged["stats_f64"].view(torch.int64), whole_st.f64.contiguous().view(torch.int64)))
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_shows,tz,holes", [(2, 1000, 0, False), (2, 37, -480, False), (3, 500, 330, False),
                                                    (2, 60, 0, True)])
def test_day_sharded_equals_single_process(built, world, n_shows, tz, holes):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_shows, tz, ret, holes)) for r in range(world)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_plan_cuts_only_at_day_boundaries():
    from sph_pie_b200.sharding import plan_day_shards

    day = torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 2, 5])
    eo = torch.arange(0, 11) * 3
    for world in (1, 2, 3, 4, 8):
        plan = plan_day_shards(eo, day, world)
        assert plan.bounds[0][0] == 0 and plan.bounds[-1][1] == 10 and len(plan.bounds) == world
        for (a, b), (c, _) in zip(plan.bounds, plan.bounds[1:] + [(10, 10)]):
            assert b == c and a <= b
            assert b in (0, 3, 5, 9, 10)  # a new day starts at every cut
    with pytest.raises(ValueError):
        plan_day_shards(eo, torch.tensor([0, 1, 0, 1, 1, 2, 2, 2, 2, 5]), 2)
    # shows without a day (PIE_DAY_NONE) neither break the order nor make a cut
    none = -(2 ** 63)
    plan = plan_day_shards(eo, torch.tensor([none, 0, 0, 1, none, 1, 2, 2, none, 5]), 2)
    assert plan.bounds[0][0] == 0 and plan.bounds[-1][1] == 10 and plan.bounds[0][1] in (3, 6, 9)
    assert plan_day_shards(eo, torch.full((10,), none), 2).bounds == [(0, 10), (10, 10)]  # no day, no cut
    assert plan_day_shards(torch.zeros(1, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 3).bounds == [(0, 0)] * 3
