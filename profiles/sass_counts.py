"""Static SASS mnemonic counts per kernel of libspateo_b200.so (cuobjdump -sass; runs without a GPU).
Usage: python profiles/sass_counts.py > profiles/sass_r02_final.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spateo_release_b200", "libspateo_b200.so")
COLS = ["UBLKCP", "UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "LDS", "SHFL",
        "DFMA", "ATOMG", "RED", "LDL", "STL"]
KEEP = ("estep_sweep", "col_select", "build_col_lists", "col_finalize", "row_stats_p2p", "gene_cost_tc", "gram_tc_kernel",
        "field_apply_lowrank", "nonrigid_solve", "gram_small", "weighted_gram", "field_geometry")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, total, name = {}, {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name], total[name] = collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    counts[name][c] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS mnemonic counts per kernel, final round-2 build (`python profiles/sass_counts.py`; cuobjdump -sass, sm_100a, nvcc 12.9)\n")
    print("Static instruction counts. The sweep kernels' main loop is one fully unrolled 8-column stage; template arguments of the")
    print("sweeps: <columns per stage, stages, CTAs per SM, [sparse, debug mode,] dimension>. Each product sweep holds TWO unrolled stage")
    print("bodies: the full one and the one for spatially dead columns (no spatial posterior: half the MUFU.EX2) — sweep 2 in 3-D:")
    print("144 + 128 FFMA2, 64 + 48 FMUL2, 64 + 64 FADD2, 64 + 32 MUFU.EX2 (34 resp. 30 packed fp32x2 operations per column and row")
    print("quad); tcgen05 = UTCHMMA (mma) / UTCBAR (commit) / LDTM (tcgen05.ld) / UTCATOMSWS (TMEM alloc), TMA = UTMALDG (tensor) /")
    print("UBLKCP (1-D bulk copy), mbarrier = SYNCS.\n")
    print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for mangled, nice in sorted(zip(counts, demangle), key=lambda t: t[1]):
        if not any(k in nice for k in KEEP):
            continue
        short = re.sub(r"\(anonymous namespace\)::|void |\(.*$", "", nice)
        print(f"| `{short}` | {total[mangled]} | " + " | ".join(str(counts[mangled][c]) for c in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
