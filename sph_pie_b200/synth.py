"""Synthetic archive tables, generated column-wise with torch on any device.

The vocabularies are the reference's own enumerations (public/app.js:1-15 ISSUE_MAP / ACTIONS /
STATUS, :5182-5186 severity / root-cause options) plus "dirty" variants (case, padding, unknown
values, empty) so that every branch of the path is exercised.  There is no reference dataset:
`data` in bench.py is "synthetic".
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .columnar import ArchiveTable, StrCol, StrListCol

ISSUE_MAP = {  # public/app.js:1-12
    "Tracking lost": ["occlusion", "calibration", "marker loss", "software", "unknown"],
    "Failed to launch": ["mechanical", "arming", "safety", "unknown"],
    "Command delay": ["network latency", "controller queue", "unknown"],
    "RF link": ["TX fault", "RX fault", "interference", "antenna", "unknown"],
    "Battery": ["low voltage", "BMS fault", "poor contact", "swelling", "unknown"],
    "Motor or prop": ["no spin", "desync", "damage", "unknown"],
    "Sensor or IMU": ["bias", "calibration", "saturation", "unknown"],
    "Software or show control": ["cue timing", "state desync", "crash", "unknown"],
    "Operator input": ["incorrect mode", "early abort", "missed cue", "unknown"],
    "Other": [],
}
PRIMARY_ISSUES = list(ISSUE_MAP)
ACTIONS = ["Reboot", "Swap battery", "Swap drone", "Retry launch", "Abort segment", "Logged only"]
SEVERITIES = ["Critical show stop", "Major visible", "Minor contained"]
ROOT_CAUSES = ["Hardware", "Software", "Ops", "Environment", "Unknown"]

STATUS_VOCAB = ["Completed", "No-launch", "Abort", "completed", "ABORT", "NO-LAUNCH", "", "Scrubbed", "Completed "]
STATUS_P = [0.70, 0.10, 0.10, 0.02, 0.02, 0.02, 0.02, 0.01, 0.01]
YESNO_VOCAB = ["Yes", "No", "", "yes", "YES", "N/A"]
YESNO_P = [0.70, 0.20, 0.04, 0.03, 0.02, 0.01]
ISSUE_VOCAB = [""] + PRIMARY_ISSUES + [" Battery ", "\tRF link\n", " Other", "Weather", "battery", "  ",
                                             " Command delay ", "Propulsión", "Tracking lost "]
NOTES_VOCAB = ["", "Green across the board", "Recovered after reboot", 'Pilot said "hold", then released',
               "Wind gusts, 12 kt; held 30 s", "line one\nline two", "carriage\r\nreturn", "swapped pack, relaunched",
               "Überprüfung nötig — später", "see ticket #4521, follow-up", ","]
OTHER_DETAIL_VOCAB = ["", "", "", "vendor firmware 2.1.4", 'label "B" peeled', "unknown, needs triage"]
LABEL_VOCAB = ["Show 1", "Show 2", "Show 3", "Show 4", "Show 5", "Rehearsal", "Matinee, early", 'The "Late" show']
NAME_VOCAB = [f"Operator {i:02d}" for i in range(1, 22)]  # 21 seeded users (userStore.js:28-50)


def _encode_vocab(vocab: Sequence[str], device):
    enc = [v.encode("utf-8") for v in vocab]
    lens = torch.tensor([len(b) for b in enc], dtype=torch.int64, device=device)
    starts = torch.zeros(len(enc), dtype=torch.int64, device=device)
    if len(enc) > 1:
        starts[1:] = torch.cumsum(lens, 0)[:-1]
    flat = torch.tensor(list(b"".join(enc)) or [0], dtype=torch.uint8, device=device)
    return lens, starts, flat


def strcol_from_codes(codes: torch.Tensor, vocab: Sequence[str]) -> StrCol:
    """String column whose row i is vocab[codes[i]] (vectorised gather of the vocabulary bytes)."""
    device = codes.device
    n = codes.numel()
    lens, starts, flat = _encode_vocab(vocab, device)
    l = lens[codes]
    offs = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(l, 0, out=offs[1:])
    total = int(offs[-1])
    if total >= 2 ** 31:
        raise ValueError("string column exceeds 2 GiB: generate a smaller batch")
    if total == 0:
        return StrCol(offs.to(torch.int32), torch.zeros(0, dtype=torch.uint8, device=device))
    rows = torch.repeat_interleave(torch.arange(n, device=device), l)
    src = starts[codes][rows] + (torch.arange(total, device=device) - offs[:-1][rows])
    return StrCol(offs.to(torch.int32), flat[src])


def _choice(n: int, p: Sequence[float], gen: torch.Generator, device) -> torch.Tensor:
    probs = torch.tensor(p, dtype=torch.float64, device=device)
    cdf = torch.cumsum(probs / probs.sum(), 0)
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    return torch.searchsorted(cdf, u).clamp_(max=len(p) - 1)


def _numbered(prefix: str, num: torch.Tensor, width: int) -> StrCol:
    """Fixed-width `${prefix}${zero-padded num}` strings."""
    device = num.device
    n = num.numel()
    pre = torch.tensor(list(prefix.encode()), dtype=torch.uint8, device=device)
    w = len(pre) + width
    out = torch.empty((n, w), dtype=torch.uint8, device=device)
    out[:, : len(pre)] = pre
    x = num.clone()
    for k in range(width - 1, -1, -1):
        out[:, len(pre) + k] = (x % 10 + 48).to(torch.uint8)
        x = x // 10
    offs = torch.arange(n + 1, device=device, dtype=torch.int64) * w
    if int(offs[-1]) >= 2 ** 31:
        raise ValueError("string column exceeds 2 GiB: generate a smaller batch")
    return StrCol(offs.to(torch.int32), out.reshape(-1))


def _uuid_like(n: int, gen: torch.Generator, device) -> StrCol:
    """36-byte uuid-v4-shaped ids (entry.id is uuidv4(), sqlProvider.js:389)."""
    hexd = torch.tensor(list(b"0123456789abcdef"), dtype=torch.uint8, device=device)
    out = hexd[torch.randint(0, 16, (n, 36), generator=gen, device=device)]
    for pos in (8, 13, 18, 23):
        out[:, pos] = 45
    out[:, 14] = 52
    offs = torch.arange(n + 1, device=device, dtype=torch.int64) * 36
    if int(offs[-1]) >= 2 ** 31:
        raise ValueError("string column exceeds 2 GiB: generate a smaller batch")
    return StrCol(offs.to(torch.int32), out.reshape(-1))


def synth_archive(n_shows: int, seed: int = 0, device="cpu", max_entries: int = 21, shows_per_day: int = 5,
                  shuffle_days: bool = False, dirty: bool = True, start_ms: int = 1704067200000,
                  missing_created_frac: float = 0.0, notes_repeat: int = 1) -> ArchiveTable:
    """`n_shows` shows with 0..max_entries entries each (21 seeded operators, one entry per operator
    per show: sqlProvider.js:434-457), up to `shows_per_day` shows per calendar day
    (sqlProvider.js:427), ascending by createdAt unless `shuffle_days`."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    S = n_shows

    n_per = torch.randint(0, max_entries + 1, (S,), generator=gen, device=device)
    entry_offsets = torch.zeros(S + 1, dtype=torch.int64, device=device)
    torch.cumsum(n_per, 0, out=entry_offsets[1:])
    E = int(entry_offsets[-1])
    if E >= 2 ** 31:
        raise ValueError("too many entries for one batch")

    # show level ------------------------------------------------------------------------------
    sidx = torch.arange(S, device=device)
    day = sidx // shows_per_day
    slot = sidx % shows_per_day
    tod = (17 * 3600 + slot * 5400) * 1000 + torch.randint(0, 600000, (S,), generator=gen, device=device)
    created = (start_ms + day * 86400000 + tod).to(torch.float64)
    if shuffle_days and S > 1:
        perm = torch.randperm(S, generator=gen, device=device)
        created = created[perm]
        day = day[perm]
    archived = created + 12 * 3600 * 1000.0
    if missing_created_frac > 0:
        miss = torch.rand(S, generator=gen, device=device) < missing_created_frac
        created = torch.where(miss, torch.full_like(created, float("nan")), created)
    # YYYY-MM-DD via civil-from-days on the tensor
    z = (start_ms // 86400000 + day) + 719468
    era = torch.div(z, 146097, rounding_mode="floor")
    doe = z - era * 146097
    yoe = torch.div(doe - torch.div(doe, 1460, rounding_mode="floor") + torch.div(doe, 36524, rounding_mode="floor")
                    - torch.div(doe, 146096, rounding_mode="floor"), 365, rounding_mode="floor")
    y = yoe + era * 400
    doy = doe - (365 * yoe + torch.div(yoe, 4, rounding_mode="floor") - torch.div(yoe, 100, rounding_mode="floor"))
    mp = torch.div(5 * doy + 2, 153, rounding_mode="floor")
    dd = doy - torch.div(153 * mp + 2, 5, rounding_mode="floor") + 1
    mm = torch.where(mp < 10, mp + 3, mp - 9)
    y = y + (mm <= 2).to(y.dtype)
    ymd = _numbered("", y * 10000 + mm * 100 + dd, 8).data.reshape(S, 8)
    date_bytes = torch.empty((S, 10), dtype=torch.uint8, device=device)
    date_bytes[:, 0:4] = ymd[:, 0:4]
    date_bytes[:, 4] = 45
    date_bytes[:, 5:7] = ymd[:, 4:6]
    date_bytes[:, 7] = 45
    date_bytes[:, 8:10] = ymd[:, 6:8]
    show_date = StrCol((torch.arange(S + 1, device=device) * 10).to(torch.int32), date_bytes.reshape(-1))
    hh = 17 + torch.div(slot * 90, 60, rounding_mode="floor")
    mi = (slot * 90) % 60
    hm = _numbered("", hh * 100 + mi, 4).data.reshape(S, 4)
    time_bytes = torch.empty((S, 5), dtype=torch.uint8, device=device)
    time_bytes[:, 0:2] = hm[:, 0:2]
    time_bytes[:, 2] = 58
    time_bytes[:, 3:5] = hm[:, 2:4]
    show_time = StrCol((torch.arange(S + 1, device=device) * 5).to(torch.int32), time_bytes.reshape(-1))

    show_cols = {
        "show_id": _uuid_like(S, gen, device),
        "show_date": show_date,
        "show_time": show_time,
        "show_label": strcol_from_codes(torch.randint(0, len(LABEL_VOCAB), (S,), generator=gen, device=device), LABEL_VOCAB),
        "lead_pilot": strcol_from_codes(torch.randint(0, 15, (S,), generator=gen, device=device), NAME_VOCAB),
        "monkey_lead": strcol_from_codes(torch.randint(15, 21, (S,), generator=gen, device=device), NAME_VOCAB),
        "show_notes": strcol_from_codes(torch.randint(0, len(NOTES_VOCAB), (S,), generator=gen, device=device), NOTES_VOCAB),
    }
    crew_n = torch.randint(0, 5, (S,), generator=gen, device=device)
    crew_lo = torch.zeros(S + 1, dtype=torch.int64, device=device)
    torch.cumsum(crew_n, 0, out=crew_lo[1:])
    crew_items = strcol_from_codes(torch.randint(0, 21, (int(crew_lo[-1]),), generator=gen, device=device), NAME_VOCAB)
    crew = StrListCol(crew_lo.to(torch.int32), crew_items)

    # entry level -----------------------------------------------------------------------------
    st_p = STATUS_P if dirty else [0.8, 0.1, 0.1]
    st_vocab = STATUS_VOCAB if dirty else STATUS_VOCAB[:3]
    status_code = _choice(E, st_p, gen, device)
    yn_p = YESNO_P if dirty else [0.8, 0.2]
    yn_vocab = YESNO_VOCAB if dirty else YESNO_VOCAB[:2]
    completed = (status_code == 0)
    n_issue_vocab = len(ISSUE_VOCAB) if dirty else 1 + len(PRIMARY_ISSUES)
    issue_code = torch.randint(1, n_issue_vocab, (E,), generator=gen, device=device)
    keep_issue = torch.rand(E, generator=gen, device=device) < 0.05  # stale issue on a completed entry
    issue_code = torch.where(completed & ~keep_issue, torch.zeros_like(issue_code), issue_code)
    sub_vocab = [""] + sorted({s for v in ISSUE_MAP.values() for s in v})
    sub_code = torch.where(issue_code > 0, torch.randint(0, len(sub_vocab), (E,), generator=gen, device=device),
                           torch.zeros_like(issue_code))
    sev_code = torch.where(issue_code > 0, torch.randint(0, 4, (E,), generator=gen, device=device),
                           torch.zeros_like(issue_code))
    root_code = torch.where(issue_code > 0, torch.randint(0, 6, (E,), generator=gen, device=device),
                            torch.zeros_like(issue_code))

    # operator = k-th seeded user within the show (unique per show)
    show_of_entry = torch.repeat_interleave(torch.arange(S, device=device), n_per)
    within = torch.arange(E, device=device) - entry_offsets[:-1][show_of_entry]
    entry_cols = {
        "entry_id": _uuid_like(E, gen, device),
        "unit_id": _numbered("Drone-", torch.randint(1, 100, (E,), generator=gen, device=device), 2),
        "planned": strcol_from_codes(_choice(E, yn_p, gen, device), yn_vocab),
        "launched": strcol_from_codes(_choice(E, yn_p, gen, device), yn_vocab),
        "status": strcol_from_codes(status_code, st_vocab),
        "primary_issue": strcol_from_codes(issue_code, ISSUE_VOCAB),
        "sub_issue": strcol_from_codes(sub_code, sub_vocab),
        "other_detail": strcol_from_codes(
            torch.where(issue_code > 0, torch.randint(0, len(OTHER_DETAIL_VOCAB), (E,), generator=gen, device=device),
                        torch.zeros_like(issue_code)), OTHER_DETAIL_VOCAB),
        "severity": strcol_from_codes(sev_code, [""] + SEVERITIES),
        "root_cause": strcol_from_codes(root_code, [""] + ROOT_CAUSES),
        "operator_name": strcol_from_codes(within % 21, NAME_VOCAB),
        "battery_id": _numbered("B-", torch.randint(1, 400, (E,), generator=gen, device=device), 3),
        "command_rx": strcol_from_codes(_choice(E, yn_p, gen, device), yn_vocab),
        "notes": strcol_from_codes(torch.randint(0, len(NOTES_VOCAB), (E,), generator=gen, device=device),
                                   [n * notes_repeat for n in NOTES_VOCAB]),  # notes_repeat > 1: long free text
    }
    act_n = torch.randint(0, 3, (E,), generator=gen, device=device)
    act_lo = torch.zeros(E + 1, dtype=torch.int64, device=device)
    torch.cumsum(act_n, 0, out=act_lo[1:])
    act_items = strcol_from_codes(torch.randint(0, len(ACTIONS), (int(act_lo[-1]),), generator=gen, device=device), ACTIONS)
    actions = StrListCol(act_lo.to(torch.int32), act_items)

    # delaySec: mostly small integers, some halves / decimals, a few nulls and (dirty) non-finite
    base = torch.randint(0, 121, (E,), generator=gen, device=device).to(torch.float64)
    kind = torch.rand(E, generator=gen, device=device)
    delay = torch.where(kind < 0.15, base + 0.5, base)
    delay = torch.where((kind >= 0.15) & (kind < 0.25), base + torch.rand(E, generator=gen, device=device, dtype=torch.float64), delay)
    delay = torch.where(completed, torch.where(kind < 0.5, torch.zeros_like(delay), delay), delay)
    valid = (torch.rand(E, generator=gen, device=device) >= 0.08)
    if dirty:
        bad = torch.rand(E, generator=gen, device=device)
        delay = torch.where(bad < 0.002, torch.full_like(delay, float("nan")), delay)
        delay = torch.where((bad >= 0.002) & (bad < 0.003), torch.full_like(delay, float("inf")), delay)
        delay = torch.where((bad >= 0.003) & (bad < 0.004), -delay, delay)  # includes -0.0
    delay = torch.where(valid, delay, torch.zeros_like(delay))
    entry_ts = (created[show_of_entry] if S else torch.zeros(0, dtype=torch.float64, device=device))
    entry_ts = torch.nan_to_num(entry_ts, nan=float(start_ms)) + within.to(torch.float64) * 1000.0

    return ArchiveTable(
        n_shows=S, n_entries=E, entry_offsets=entry_offsets.to(torch.int32),
        show_cols=show_cols, crew=crew, created_at=created, archived_at=archived,
        entry_cols=entry_cols, actions=actions,
        delay_sec=delay, delay_valid=valid.to(torch.uint8), entry_ts=entry_ts)


def synth_skeleton(n_shows: int, seed: int = 0, device="cpu", max_entries: int = 21, shows_per_day: int = 5):
    """(entries per show int64[n], day index int64[n]) of synth_archive(n_shows, seed, device) without building it: the
    generator's first draw is the entry counts, and an unshuffled archive has `shows_per_day` shows on each day.  What
    a rank needs of the OTHER segments of a sharded archive to plan the day ranges (sharding.plan_day_shards)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_per = torch.randint(0, max_entries + 1, (n_shows,), generator=gen, device=device)
    return n_per, torch.arange(n_shows, device=device) // shows_per_day


def table_to_shows(table: ArchiveTable) -> List[dict]:
    """Inverse of pack_shows for SMALL tables: the JSON documents the table stands for."""
    import math

    t = table.to("cpu")

    def col_values(c: StrCol):
        o = c.offsets.tolist()
        b = bytes(c.data.numpy())
        return [b[o[i]:o[i + 1]].decode("utf-8") for i in range(len(o) - 1)]

    from .columnar import ENTRY_KEY_TO_COL, SHOW_KEY_TO_COL

    sv = {k: col_values(t.show_cols[c]) for k, c in SHOW_KEY_TO_COL.items()}
    ev = {k: col_values(t.entry_cols[c]) for k, c in ENTRY_KEY_TO_COL.items()}
    crew_items, act_items = col_values(t.crew.items), col_values(t.actions.items)
    clo, alo = t.crew.list_offsets.tolist(), t.actions.list_offsets.tolist()
    eo = t.entry_offsets.tolist()
    created, archived = t.created_at.tolist(), t.archived_at.tolist()
    delay, valid, ets = t.delay_sec.tolist(), t.delay_valid.tolist(), t.entry_ts.tolist()
    shows = []
    for s in range(t.n_shows):
        show = {k: sv[k][s] for k in sv}
        show["crew"] = crew_items[clo[s]:clo[s + 1]]
        show["createdAt"] = None if math.isnan(created[s]) else created[s]
        show["archivedAt"] = None if math.isnan(archived[s]) else archived[s]
        show["entries"] = []
        for e in range(eo[s], eo[s + 1]):
            entry = {k: ev[k][e] for k in ev}
            entry["actions"] = act_items[alo[e - eo[0]]:alo[e - eo[0] + 1]]
            entry["delaySec"] = delay[e] if valid[e] else None
            entry["ts"] = None if math.isnan(ets[e]) else ets[e]
            show["entries"].append(entry)
        shows.append(show)
    return shows


def synth_stored_docs(sample_shows: int, copies: int, device, seed: int = 0):
    """Synthetic `show_archive.data` texts at bench size: a sample archive written out as JSON documents (one per
    show, no whitespace, keys in provider order) and repeated `copies` times on `device`.
    Returns (JsonDocs, n_entries, text_bytes, sample_docs: List[str])."""
    import json

    from .ops import JsonDocs

    host = synth_archive(sample_shows, seed=seed)
    lost = host.delay_valid.bool() & ~torch.isfinite(host.delay_sec)  # JSON has no NaN / Infinity: stored as null
    host.delay_valid[lost] = 0
    texts = [json.dumps(s, ensure_ascii=False, separators=(",", ":")) for s in table_to_shows(host)]
    one = JsonDocs.from_texts(texts)
    n, nbytes = one.n_docs, int(one.offsets[-1])
    text = torch.cat([one.data[:nbytes].to(device).repeat(copies), torch.zeros(8, dtype=torch.uint8, device=device)])
    lens = (one.offsets[1:] - one.offsets[:-1]).to(device).repeat(copies)
    offsets = torch.zeros(n * copies + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=offsets[1:])
    return JsonDocs(offsets, text), host.n_entries * copies, nbytes * copies, texts
