"""Shared checkers for the parity tests."""
import math

import torch

import pie_oracle as po
from sph_pie_b200 import _lib
from sph_pie_b200.archive import _stats_dict


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    a, b = a.cpu().contiguous(), b.cpu().contiguous()
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == torch.float64:
        return torch.equal(a.view(torch.int64), b.view(torch.int64))
    return torch.equal(a, b)


def assert_analytics_equal(got, ref, what=""):
    (gs, gd), (rs, rd) = got, ref
    assert bits_equal(gs.i32, rs.i32), f"{what} stats_i32"
    assert bits_equal(gs.f64, rs.f64), f"{what} stats_f64"
    if rd is None:
        return
    assert gd.n_groups == rd.n_groups, f"{what} n_groups {gd.n_groups} != {rd.n_groups}"
    for name in ("show_day_start", "show_order", "group_day_start", "group_offsets", "summary_f64", "summary_count"):
        assert bits_equal(getattr(gd, name), getattr(rd, name)), f"{what} {name}"


def same_value(a, b) -> bool:
    """JS-level equality of numbers/None including the sign of zero and NaN."""
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, float) or isinstance(b, float):
        a, b = float(a), float(b)
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        return a == b and math.copysign(1, a) == math.copysign(1, b)
    return a == b


def assert_stats_match_py_oracle(shows, stats):
    """Plane-major stats (from any backend) against the Python oracle run on the JSON documents."""
    i32, f64 = stats.i32.cpu().numpy(), stats.f64.cpu().numpy()
    for s, show in enumerate(shows):
        want = po.compute_archive_show_stats(show)
        got = _stats_dict(i32, f64, s)
        assert list(want["issueCounts"].items()) == list(got["issueCounts"].items()), (s, want, got)
        for k in want:
            if k in ("issueCounts",):
                continue
            if k == "issueRates":
                assert all(same_value(want[k][i], got[k][i]) for i in po.PRIMARY_ISSUES), (s, k, want[k], got[k])
            else:
                assert same_value(want[k], got[k]), (s, k, want[k], got[k])


def assert_daily_match_py_oracle(shows, daily, tz):
    groups = po.build_archive_daily_groups(shows, tz)
    assert len(groups) == daily.n_groups
    order = daily.show_order.cpu().numpy()
    goff = daily.group_offsets.cpu().numpy()
    sf, sc = daily.summary_f64.cpu().numpy(), daily.summary_count.cpu().numpy()
    for g, grp in enumerate(groups):
        assert grp["timestamp"] == int(daily.group_day_start[g])
        members = [int(order[i]) for i in range(goff[g], goff[g + 1])]
        assert [id(shows[s]) for s in members] == [id(it["show"]) for it in grp["shows"]]
        for m, key in enumerate(po.ALL_METRIC_KEYS):
            want = po.group_metric_summary(grp, key)
            n = int(sc[m][g])
            assert want["count"] == n, (g, key)
            for k, name in ((_lib.DF_AVERAGE, "average"), (_lib.DF_MIN, "min"), (_lib.DF_MAX, "max")):
                got = float(sf[k][m][g]) if n else None
                assert same_value(want[name], got), (g, key, name, want[name], got)


def build_c_consumer() -> str:
    """tests/native/c_consumer.c — a plain C99 program over include/sph_pie_b200.h with malloc'd buffers only — built
    with -Wall -Wextra -Werror -pedantic against the in-tree library; returns the path of the executable."""
    import os
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    src, exe = os.path.join(here, "native", "c_consumer.c"), os.path.join(here, "native", "c_consumer")
    lib = os.path.join(root, "sph_pie_b200", "libsphpie_b200.so")
    deps = [src, os.path.join(root, "include", "sph_pie_b200.h"), lib]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["gcc", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1",
                               "-I", os.path.join(root, "include"), "-o", exe, src, "-L", os.path.dirname(lib), "-lsphpie_b200",
                               "-Wl,-rpath," + os.path.dirname(lib)])
    return exe
