"""Host-side logic that needs no GPU: packing, slicing, date keys, the ABI surface."""
import os
import re

import pytest
import torch

from sph_pie_b200 import _lib
from sph_pie_b200.archive import _iso_date_key
from sph_pie_b200.columnar import pack_shows
from sph_pie_b200.synth import synth_archive, table_to_shows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "sph_pie_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pie_[a-z0-9_]+)\s*\(", body))
    assert declared, "no functions found in the header"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sph_pie_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pie_abi_version() == 2


def test_struct_layouts_match_header(built):
    import ctypes as C

    assert C.sizeof(_lib.StrColC) == 16 and C.sizeof(_lib.StrListColC) == 24
    # 3 scalars/pointers + 7 strcols + strlist + 2 ptr + 14 strcols + strlist + 3 ptr + (ABI 2) 3 ptr
    assert C.sizeof(_lib.ArchiveViewC) == 8 * 3 + 16 * 7 + 24 + 16 + 16 * 14 + 24 + 24 + 24
    assert C.sizeof(_lib.DocTimesC) == 32
    assert C.sizeof(_lib.DailyOutC) == 8 * 9


def test_no_gpu_means_loud_failure(built):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.PieError) as e:
        _lib.init(0)
    assert e.value.code == _lib.PIE_ERR_NO_DEVICE
    from sph_pie_b200 import computeArchiveShowStats

    with pytest.raises(_lib.PieError):
        computeArchiveShowStats({"entries": []})


def test_plain_c_caller_builds_and_fails_loudly_without_a_gpu(built, tmp_path):
    """include/sph_pie_b200.h is C99, and a C program with malloc'd buffers links and calls the library (the drop-in
    boundary is the C ABI, not the ctypes binding); without a device it stops with the library's message, exit 3."""
    import subprocess

    from helpers import build_c_consumer

    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c",
                           os.path.join(ROOT, "include", "sph_pie_b200.h")])
    exe = build_c_consumer()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_abi_errors.py runs the program")
    docs = tmp_path / "docs.jsonl"
    docs.write_text('{"id":"a","entries":[{"id":"e"}]}\n')
    r = subprocess.run([exe, str(docs), "0", str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "pie_init(0) = %d" % _lib.PIE_ERR_NO_DEVICE in r.stderr and "no CPU fallback" in r.stderr
    assert not (tmp_path / "out.csv").exists()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sph_pie_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pie_oracle" not in text and "oracle_c" not in text and "libpie_oracle" not in text, f


def test_pack_shows_schema():
    t = pack_shows([{"id": "s", "crew": ["a", None, "b"], "createdAt": 5, "entries": [
        {"status": "Abort", "delaySec": 3, "actions": ["x", "y"], "ts": 7}, {"delaySec": None}]}, None])
    assert (t.n_shows, t.n_entries) == (2, 2)
    assert t.entry_offsets.tolist() == [0, 2, 2]
    assert t.entry_cols["status"].get(0) == "Abort" and t.entry_cols["status"].get(1) == ""
    assert t.delay_valid.tolist() == [1, 0] and t.delay_sec.tolist()[0] == 3.0
    assert t.crew.list_offsets.tolist() == [0, 3, 3] and t.crew.items.get(1) == ""
    assert t.created_at[0] == 5 and torch.isnan(t.created_at[1])
    with pytest.raises(TypeError):
        pack_shows([{"entries": [{"status": 5}]}])
    with pytest.raises(TypeError):
        pack_shows([{"entries": [{"delaySec": "5"}]}])
    with pytest.raises(TypeError):
        pack_shows([{"label": "\ud800"}])


def test_slice_shows_keeps_absolute_string_offsets():
    t = synth_archive(50, seed=3)
    part = t.slice_shows(10, 30)
    assert part.n_shows == 20 and part.entry_offsets[0] == 0
    assert part.n_entries == int(t.entry_offsets[30] - t.entry_offsets[10])
    whole, sl = table_to_shows(t), table_to_shows(part)
    assert sl == whole[10:30]


def test_iso_date_key():
    assert _iso_date_key(0) == "1970-01-01" and _iso_date_key(-1) == "1969-12-31"
    assert _iso_date_key(1719964800000) == "2024-07-03"
    assert _iso_date_key(253402300800000) == "+010000-01"[:10]   # year 10000: toISOString is +010000-01-01T...
    assert _iso_date_key(-62198755200000) == "-000001-01"[:10]


def test_synth_is_deterministic_and_well_formed():
    a, b = synth_archive(100, seed=11), synth_archive(100, seed=11)
    assert torch.equal(a.entry_cols["status"].data, b.entry_cols["status"].data)
    assert torch.equal(a.delay_sec.view(torch.int64), b.delay_sec.view(torch.int64))
    for c in list(a.entry_cols.values()) + list(a.show_cols.values()):
        o = c.offsets
        assert o[0] == 0 and bool((o[1:] >= o[:-1]).all()) and int(o[-1]) == c.data.numel()
    assert int(a.entry_offsets[-1]) == a.n_entries


def test_launch_list_kernel_names():
    """scripts/traffic_from_launches.py groups an ncu launch list by kernel: the names ncu prints for kernels in an
    unnamed namespace, with bool / class template arguments, must land in the groups bench.py reads the traffic of."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("traffic_from_launches", os.path.join(ROOT, "scripts", "traffic_from_launches.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = {
        "void pie::<unnamed>::ingest_fast_kernel<(bool)0, pie::jf::CapsSmall, (bool)0>(const long *, const unsigned char *, long, pie::<unnamed>::IngestScratch, unsigned char *, pie::jw::IngestOut)": "ingest_fast_kernel<measure>",
        "void pie::<unnamed>::ingest_fast_kernel<(bool)1, pie::jf::CapsBig, (bool)1>(const long *, const unsigned char *, long, pie::<unnamed>::IngestScratch, unsigned char *, pie::jw::IngestOut)": "ingest_fast_kernel<fill>",
        "void pie::<unnamed>::ingest_walk_kernel<(bool)1>(const long *, const unsigned char *, long, pie::<unnamed>::IngestScratch, unsigned char *, pie::jw::IngestOut, const int *, const unsigned int *, unsigned long long *)": "ingest_walk_kernel<fill>",
        "pie::<unnamed>::ingest_init_kernel(pie::<unnamed>::IngestScratch)": "ingest_init_kernel",
        "pie::<unnamed>::ingest_route_kernel(const long *, long, pie::<unnamed>::IngestScratch)": "ingest_route_kernel",
        "void pie::(anonymous namespace)::export_rows_kernel<false>(pie_archive_view, pie::RowTable)": "export_rows_kernel<csv>",
        "void pie::<unnamed>::export_rows_kernel<(bool)1>(pie_archive_view, pie::<unnamed>::RowTable)": "export_rows_kernel<json>",
        "pie::<unnamed>::show_stats_kernel(pie_archive_view, int *, double *, long)": "show_stats_kernel",
    }
    for full, want in cases.items():
        assert mod.kernel_name(full) == want, (full, mod.kernel_name(full))
    for group, names in mod.GROUPS.items():
        assert len(set(names)) == len(names), group
    assert "ingest_fast_kernel<measure>" in mod.GROUPS["ingest"] and "ingest_fast_kernel<fill>" in mod.GROUPS["ingest"]
