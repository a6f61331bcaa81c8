"""Developer tool: build show_stats_kernel variants (-D tunables) and time each on the bench table.
Usage (GPU box):  python scripts/sweep_stats.py build   (on the CPU box, cross-compiles variants)
                  python scripts/sweep_stats.py run     (on the GPU box)"""
import glob
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "scripts", "_variants")

VARIANTS = [dict(T=t, C=c, R=r, B=b) for t, c, r, b in [
    (256, 4096, 2, 4), (256, 2048, 2, 4), (256, 4096, 1, 4), (128, 2048, 2, 8), (128, 4096, 2, 8), (256, 4096, 2, 5),
    (256, 6144, 2, 3), (512, 8192, 2, 2), (128, 2048, 1, 8),
]]


def name(v):
    return f"T{v['T']}_C{v['C']}_R{v['R']}_B{v['B']}"


def build():
    import __graft_entry__ as g

    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(g.CSRC, "*.cu")))
    for v in VARIANTS:
        so = os.path.join(OUT, name(v) + ".so")
        cmd = ["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + [
            f"-DPIE_STATS_THREADS={v['T']}", f"-DPIE_STATS_CHUNK={v['C']}", f"-DPIE_STATS_ROWS={v['R']}",
            f"-DPIE_STATS_MIN_BLOCKS={v['B']}", "-Xptxas", "-v", "-o", so] + srcs
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        regs = [l for l in r.stderr.splitlines() if "show_stats_kernel" in l or "registers" in l]
        k = [i for i, l in enumerate(r.stderr.splitlines()) if "Compiling entry function" in l and "show_stats_kernel" in l]
        info = r.stderr.splitlines()[k[0] + 2: k[0] + 4] if k else []
        print(name(v), r.returncode, " | ".join(x.strip() for x in info))


def run_one():
    import torch

    from sph_pie_b200 import _lib, ops
    from sph_pie_b200.synth import synth_archive

    _lib.init(0)
    table = synth_archive(1 << 20, seed=1234, device="cuda:0")
    bufs = ops.DailyBuffers(table.n_shows, table.n_entries, "cuda:0")
    for _ in range(5):
        ops.show_stats_dev(table, bufs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30):
        ops.show_stats_dev(table, bufs)
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"variant": os.path.basename(os.environ.get("PIE_LIB_PATH", "default")), "ms": a.elapsed_time(b) / 30,
                      "checksum": int(bufs.stats_i32.sum())}))


def run():
    for v in VARIANTS:
        env = dict(os.environ, PIE_LIB_PATH=os.path.join(OUT, name(v) + ".so"))
        subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, cwd=ROOT)


if __name__ == "__main__":
    {"build": build, "run": run, "one": run_one}[sys.argv[1]]()
