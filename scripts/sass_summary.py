"""Counts of the Blackwell-specific SASS instructions per kernel of libbsub_b200.so (cuobjdump -sass) -> markdown table.
   python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "background-subtraction_b200", "libbsub_b200.so")
COLS = [("UTCIMMA (tcgen05.mma i8)", r"\bUTCIMMA\b"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("UTMALDG (TMA load)", r"\bUTMALDG(?!\S*MULTICAST)"),
        ("UTMALDG.MULTICAST", r"\bUTMALDG\S*MULTICAST"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("UBLKCP (bulk copy)", r"\bUBLKCP"),
        ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("DMMA", r"\bDMMA"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"\bUCGABAR_ARV"),
        ("R2UR", r"\bR2UR")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[name] = collections.Counter()
            continue
        if name:
            for col, pat in COLS:
                if re.search(pat, line):
                    counts[name][col] += 1
    print("# SASS instruction counts per kernel of libbsub_b200.so (cuobjdump -sass, sm_100a) -- round 2, scripts/sass_summary.py\n")
    print("| kernel | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for k in sorted(counts):
        c = counts[k]
        if any(c[col] for col, _ in COLS[:8]) or "eig" in k or "rpca" in k or "prox_graph3" in k:
            print("| %s | " % k + " | ".join(str(c[col]) for col, _ in COLS) + " |")
    print("\nThe MMA warps of the two gram_i8 kernels issue from uniform registers (no R2UR.BROADCAST / BRA.U.ANY waterfall around UTCIMMA): "
          "see DESIGN.md 4.1.")


if __name__ == "__main__":
    main()
