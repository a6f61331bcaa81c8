"""Shared pieces of the JSON-ingest tests: the oracle side (pie_oracle.map_archive_row + pack_shows), exact table
comparison, document generators, and the HOST build of the kernels' walker (tests/native/ingest_host.cpp)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import torch

import pie_oracle as po
from sph_pie_b200 import _lib
from sph_pie_b200.columnar import ArchiveTable, StrCol, StrListCol, pack_shows

HERE = os.path.dirname(os.path.abspath(__file__))


def oracle_ingest(docs):
    """(table, doc_status) the reference's JSON.parse + the table's projection give; raises TypeError (schema) or
    po.UnsupportedJson with .doc = the first offending document."""
    shows, status = [], []
    for i, d in enumerate(docs):
        try:
            show = po.map_archive_row(d)
        except po.UnsupportedJson as e:
            e.doc = i
            raise
        shows.append(show)
        status.append(0 if show is not None else 1)
    for i, show in enumerate(shows):  # find the first offender the way the table packer would
        try:
            pack_shows([show])
        except TypeError as e:
            e.doc = i
            raise
    return pack_shows(shows), np.array(status, dtype=np.uint8)


def _col_eq(a: StrCol, b: StrCol, n: int, what: str):
    ao, bo = a.offsets.cpu().numpy().astype(np.int64), b.offsets.cpu().numpy().astype(np.int64)
    assert len(ao) >= n + 1 and len(bo) >= n + 1, what
    assert np.array_equal(ao[:n + 1] - ao[0], bo[:n + 1] - bo[0]), f"{what} offsets"
    ad, bd = a.data.cpu().numpy(), b.data.cpu().numpy()
    assert np.array_equal(ad[ao[0]:ao[n]], bd[bo[0]:bo[n]]), f"{what} bytes"


def _list_eq(a: StrListCol, b: StrListCol, n: int, what: str):
    al, bl = a.list_offsets.cpu().numpy().astype(np.int64), b.list_offsets.cpu().numpy().astype(np.int64)
    assert np.array_equal(al[:n + 1], bl[:n + 1]), f"{what} list_offsets"
    _col_eq(a.items, b.items, int(al[n]), what + ".items")


def _f64_eq(a, b, what):
    a, b = a.cpu().numpy().view(np.int64), b.cpu().numpy().view(np.int64)
    # NaN payloads carry no meaning on the path (Number.isFinite): compare NaN-ness, everything else bitwise
    an, bn = np.isnan(a.view(np.float64)), np.isnan(b.view(np.float64))
    assert np.array_equal(an, bn), f"{what} NaN pattern"
    assert np.array_equal(a[~an], b[~bn]), f"{what} bits"


def assert_tables_equal(got: ArchiveTable, ref: ArchiveTable, what=""):
    assert (got.n_shows, got.n_entries) == (ref.n_shows, ref.n_entries), f"{what} shape"
    S, E = ref.n_shows, ref.n_entries
    assert torch.equal(got.entry_offsets.cpu(), ref.entry_offsets.cpu()), f"{what} entry_offsets"
    for k in ref.show_cols:
        _col_eq(got.show_cols[k], ref.show_cols[k], S, f"{what} {k}")
    for k in ref.entry_cols:
        _col_eq(got.entry_cols[k], ref.entry_cols[k], E, f"{what} {k}")
    _list_eq(got.crew, ref.crew, S, f"{what} crew")
    _list_eq(got.actions, ref.actions, E, f"{what} actions")
    _f64_eq(got.created_at, ref.created_at, f"{what} created_at")
    _f64_eq(got.archived_at, ref.archived_at, f"{what} archived_at")
    _f64_eq(got.entry_ts, ref.entry_ts, f"{what} entry_ts")
    assert torch.equal(got.delay_valid.cpu(), ref.delay_valid.cpu()), f"{what} delay_valid"
    valid = ref.delay_valid.cpu().numpy().astype(bool)
    gd, rd = got.delay_sec.cpu().numpy().view(np.int64), ref.delay_sec.cpu().numpy().view(np.int64)
    assert np.array_equal(gd[valid], rd[valid]), f"{what} delay_sec"


def docs_to_buffers(docs):
    """list of bytes/str -> (offsets int64[n+1], data uint8[] with 8 spare bytes)"""
    enc = [d.encode("utf-8") if isinstance(d, str) else bytes(d) for d in docs]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in enc], out=offs[1:])
    data = np.zeros(int(offs[-1]) + 8, dtype=np.uint8)
    data[:offs[-1]] = np.frombuffer(b"".join(enc), dtype=np.uint8)
    return offs, data


def empty_table(n_docs: int, totals, device="cpu") -> ArchiveTable:
    """An archive table with room for what the measure pass counted (what ops.ingest_json allocates)."""
    from sph_pie_b200.ops import alloc_ingest_table

    return alloc_ingest_table(n_docs, totals, device)


# ---- the walker built for the host ------------------------------------------------------------------------------
_host = None


def host_walker():
    global _host
    if _host is None:
        src = os.path.join(HERE, "native", "ingest_host.cpp")
        so = os.path.join(HERE, "native", "libingest_host.so")
        csrc = os.path.join(HERE, "..", "sph_pie_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in ("pie_json_walk.cuh", "pie_numparse.cuh", "pow5_128_table.h")]
        deps.append(os.path.join(HERE, "..", "include", "sph_pie_b200.h"))
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
        _host = C.CDLL(so)
    return _host


def host_ingest(docs):
    """The kernels' walk run on the CPU: (table | None, doc_status, (pie_status, document))."""
    lib = host_walker()
    offs, data = docs_to_buffers(docs)
    n = len(docs)
    planes = np.zeros((_lib.PIE_INGEST_TOTALS, max(n, 1)), dtype=np.uint32)
    doc_status = np.zeros(max(n, 1), dtype=np.uint8)
    totals = np.zeros(_lib.PIE_INGEST_TOTALS, dtype=np.int64)
    status = np.zeros(2, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    lib.ingest_host_measure(p(data), p(offs), C.c_int64(n), p(planes), p(doc_status), p(totals), p(status))
    if status[0] != 0:
        return None, doc_status[:n], (int(status[0]), int(status[1]))
    table = empty_table(n, totals.tolist())
    view = table.view()
    lib.ingest_host_fill(p(data), p(offs), C.c_int64(n), p(planes), p(doc_status), C.byref(view))
    return table, doc_status[:n], (0, -1)


# ---- documents -----------------------------------------------------------------------------------------------
def stored_doc(show, rng=None, style="stringify"):
    """The text a provider would hold for `show` (a dict as table_to_shows / the tests build them).
    stringify: JSON.stringify's exact form; ascii: every non-ASCII character escaped (\\uXXXX, surrogate pairs);
    pretty: indented, with spaces; shuffled: keys of every object in random order."""
    if style == "stringify":
        return po.js_json_stringify(show)

    def jsonable(v):  # NaN / Infinity are written as null, as JSON.stringify does
        if isinstance(v, dict):
            return {k: jsonable(x) for k, x in v.items()}
        if isinstance(v, list):
            return [jsonable(x) for x in v]
        if isinstance(v, float) and not np.isfinite(v):
            return None
        return v

    show = jsonable(show)
    if style == "ascii":
        return json.dumps(show, ensure_ascii=True, separators=(",", ":"))
    if style == "pretty":
        return json.dumps(show, ensure_ascii=False, indent=rng.choice([1, 2, "\t"]) if rng else 2)
    if style == "shuffled":
        def shuffle(v):
            if isinstance(v, dict):
                keys = list(v)
                rng.shuffle(keys)
                return {k: shuffle(v[k]) for k in keys}
            if isinstance(v, list):
                return [shuffle(x) for x in v]
            return v
        return json.dumps(shuffle(show), ensure_ascii=False, separators=(", ", " : "))
    raise ValueError(style)


# ---- the warp-per-document path built for the host (its lanes as fibers) -----------------------------------------
_fast = None
ROUTE_SLOW, ROUTE_FAST, ROUTE_RECORDS, ROUTE_FAST_BIG, ROUTE_LONG = 0, 1, 2, 3, 4


def fast_host():
    global _fast
    if _fast is None:
        src = os.path.join(HERE, "native", "fast_host.cpp")
        so = os.path.join(HERE, "native", "libfast_host.so")
        csrc = os.path.join(HERE, "..", "sph_pie_b200", "csrc")
        deps = [src, os.path.join(HERE, "native", "cuda_shim", "cuda_runtime.h")]
        deps += [os.path.join(csrc, f) for f in ("pie_json_fast.cuh", "pie_json_walk.cuh", "pie_numparse.cuh", "pie_device.cuh",
                                                 "pow5_128_table.h")]
        deps.append(os.path.join(HERE, "..", "include", "sph_pie_b200.h"))
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(HERE, "native", "cuda_shim"),
                                   "-o", so, src])
        _fast = C.CDLL(so)
    return _fast


def _guarded_text(data: np.ndarray, nbytes: int, guard: str, shift: int):
    """The text in pages of its own between two pages that must not be touched: guard = 'end': the aligned 32-byte word
    that holds the last byte ends where the forbidden page starts; 'start': the one that holds the first byte starts
    where the forbidden page ends.  shift (0..31): where in its 32-byte word that byte sits."""
    import mmap

    page = mmap.PAGESIZE
    body = (nbytes + 64 + page - 1) // page * page
    m = mmap.mmap(-1, body + 2 * page)
    addr = C.addressof(C.c_char.from_buffer(m))
    libc = C.CDLL(None, use_errno=True)
    libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    assert libc.mprotect(addr, page, 0) == 0 and libc.mprotect(addr + page + body, page, 0) == 0
    if guard == "end":
        # the last byte sits `shift` bytes before the end of its 32-byte word, and that word ends at the forbidden page
        end = page + body - shift
        start = end - nbytes
    else:
        start = page + shift
    view = np.frombuffer(m, dtype=np.uint8)
    view[start:start + nbytes] = data[:nbytes]
    return m, view[start:]


def fast_host_ingest(docs, pool_units_per_doc=288, guard=None, shift=0):
    """The DEVICE code of the warp-per-document ingest run on the CPU (32 lanes = 32 fibers), driven like json_ingest.cu:
    (table | None, doc_status, (pie_status, document), routes).  guard: see _guarded_text (a read outside the aligned
    32-byte words that hold the documents kills the process)."""
    lib = fast_host()
    offs, data = docs_to_buffers(docs)
    keep = None
    if guard:
        keep, text = _guarded_text(data, int(offs[-1]), guard, shift)
    else:
        # the device code reads the aligned 32-byte words that hold a document: room before and behind the text
        pad = np.zeros(len(data) + 96, dtype=np.uint8)
        base = 64 - (pad.ctypes.data % 32) % 32
        pad[base:base + len(data)] = data
        text = pad[base:]
    n = len(docs)
    rows = np.zeros((max(n, 1), _lib.PIE_INGEST_TOTALS), dtype=np.uint32)
    doc_status = np.zeros(max(n, 1), dtype=np.uint8)
    routes = np.zeros(max(n, 1), dtype=np.uint8)
    totals = np.zeros(_lib.PIE_INGEST_TOTALS, dtype=np.int64)
    status = np.zeros(2, dtype=np.int32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.fast_host_measure(p(text), p(offs), C.c_int64(n), p(rows), p(doc_status), p(totals), p(status), p(routes),
                               C.c_int(pool_units_per_doc))
    assert rc == 0, "the lanes of a warp disagreed on a document"
    if status[0] != 0:
        return None, doc_status[:n], (int(status[0]), int(status[1])), routes[:n]
    table = empty_table(n, totals.tolist())
    view = table.view()
    rc = lib.fast_host_fill(p(text), p(offs), C.c_int64(n), p(rows), p(doc_status), C.byref(view))
    assert rc == 0, f"pass 2 of the warp path failed ({rc})"
    return table, doc_status[:n], (0, -1), routes[:n]
