"""The N > 1 path with the GPU operators: world size 2 (two processes on cuda:0 — the path has no data-path collective, so
the ranks need not sit on different GPUs; gloo hands the small result tables to rank 0).  Each rank plans the day shards
of ONE archive from its skeleton, assembles its own range (which spans two stored segments), runs the single-GPU
operators on it, and rank 0 checks the stitched tables against the single-GPU result on the whole archive."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shows_per_segment, tz, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    from sph_pie_b200 import _lib, ops
    from sph_pie_b200.sharding import assemble_shard, gather_to_rank0, run_sharded, segment_skeleton

    torch.cuda.set_device(0)
    _lib.init(0)
    dev = torch.device("cuda:0")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    seeds = [100 + k for k in range(world)]
    eo, day = segment_skeleton(shows_per_segment, seeds, dev)
    local, (s0, s1) = assemble_shard(shows_per_segment, seeds, eo, day, rank, world, dev)
    st, daily = ops.archive_analytics(local, tz)
    rows = ops.csv_rows(local)
    go = daily.group_offsets.cpu()
    part = {"range": (s0, s1), "n_groups": daily.n_groups, "group_day_start": daily.group_day_start.cpu().clone(),
            "group_sizes": (go[1:] - go[:-1]).clone(), "summary_f64": daily.summary_f64.cpu().clone(),
            "summary_count": daily.summary_count.cpu().clone(), "stats_i32": st.i32.cpu().clone(), "stats_f64": st.f64.cpu().clone(),
            "csv": rows.data.cpu().clone()}
    merged = gather_to_rank0(part, rank, world)
    if rank == 0:
        # the whole archive on one GPU
        whole, _ = assemble_shard(shows_per_segment, seeds, eo, day, 0, 1, dev)
        wst, wdaily = ops.archive_analytics(whole, tz)
        wrows = ops.csv_rows(whole)
        wgo = wdaily.group_offsets.cpu()
        same = lambda a, b: torch.equal(a.contiguous().view(torch.int64) if a.dtype.is_floating_point else a,
                                        b.contiguous().view(torch.int64) if b.dtype.is_floating_point else b)
        ok = (merged["n_groups"] == wdaily.n_groups and same(merged["group_day_start"], wdaily.group_day_start.cpu())
              and same(merged["group_sizes"], wgo[1:] - wgo[:-1]) and same(merged["summary_count"], wdaily.summary_count.cpu())
              and same(merged["summary_f64"], wdaily.summary_f64.cpu()) and same(merged["stats_i32"], wst.i32.cpu())
              and same(merged["stats_f64"], wst.f64.cpu()) and same(merged["csv"], wrows.data.cpu()))
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shows_per_segment,tz", [(4000, -480), (503, 0)])
def test_day_sharded_gpu_ranks_equal_the_single_gpu_result(cuda, shows_per_segment, tz):
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shows_per_segment, tz, ret)) for r in range(world)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
