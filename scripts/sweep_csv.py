"""Developer tool: build export_rows_kernel variants (-D tunables) and time each on the bench table.
Usage:  python scripts/sweep_csv.py build   (CPU box: cross-compiles the variants into scripts/_variants)
        python scripts/sweep_csv.py run     (GPU box: times each variant in its own process)"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "scripts", "_variants")

# rows per tile, output tile KB, stage KB, min CTAs per SM
VARIANTS = [dict(R=r, O=o, S=s, B=b) for r, o, s, b in [
    (128, 42, 50, 1), (64, 22, 27, 2), (32, 11, 14, 4), (160, 52, 58, 1),
]]


def name(v):
    return f"csv_R{v['R']}_O{v['O']}_S{v['S']}_B{v['B']}"


def build():
    import __graft_entry__ as g

    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(g.CSRC, "*.cu")))
    for v in VARIANTS:
        so = os.path.join(OUT, name(v) + ".so")
        cmd = ["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + [
            f"-DPIE_CSV_ROWS={v['R']}", f"-DPIE_CSV_OUT_KB={v['O']}", f"-DPIE_CSV_STAGE_KB={v['S']}",
            f"-DPIE_CSV_MIN_BLOCKS={v['B']}", "-Xptxas", "-v", "-o", so] + srcs
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
        lines = r.stderr.splitlines()
        k = [i for i, l in enumerate(lines) if "Compiling entry function" in l and "export_rows_kernel" in l]
        info = lines[k[0] + 2: k[0] + 4] if k else lines[-5:]
        print(name(v), r.returncode, " | ".join(x.strip() for x in info))


def run():
    shows = sys.argv[2] if len(sys.argv) > 2 else str(1 << 20)
    for v in VARIANTS:
        so = os.path.join(OUT, name(v) + ".so")
        if not os.path.exists(so):
            continue
        env = dict(os.environ, PIE_LIB_PATH=so)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_csv.py"), shows], env=env,
                           capture_output=True, text=True, timeout=300)
        out = " | ".join(l for l in r.stdout.splitlines() if "ms" in l)
        print(name(v), out if r.returncode == 0 else f"FAILED rc={r.returncode} {r.stderr[-300:]}", flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
