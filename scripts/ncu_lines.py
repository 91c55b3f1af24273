"""Attribute the warp-stall samples of an ncu report (--import-source on) to CUDA source lines.

usage: ncu_lines.py report.ncu-rep file.cubin kernel_substring [top]
The SASS page of the report carries the samples per instruction; nvdisasm -g on the cubin carries the line of every
instruction; both list the kernel's instructions in the same order.
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
# the page holds one section per profiled launch: "Kernel Name",<name> / header / instructions; take the first match
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = next(i for i in starts if kern in rows[i][1])
end = next((i for i in starts if i > sec), len(rows))
hi = sec + 1
H = rows[hi]
si, src_i = H.index("# Samples"), H.index("Source")
inst = [(r[src_i].strip(), int(r[si] or 0)) for r in rows[hi + 1:end] if len(r) > si]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
sections, cur, on = {}, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)", ln)
    ms = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if ms:
        on = ms.group(1) if kern in ms.group(1) else None
        if on:
            sections[on] = []
        continue
    if not on:
        continue
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m2:
        sections[on].append((cur, m2.group(1).strip()))
# several template instantiations may match: take the one with the report's instruction count
name, lines = min(sections.items(), key=lambda kv: abs(len(kv[1]) - len(inst)))
print("cubin function:", name)
print("instructions: report %d, cubin %d" % (len(inst), len(lines)))
agg = defaultdict(int)
n = min(len(inst), len(lines))
for (txt, smp), (loc, dtxt) in zip(inst[:n], lines[:n]):
    agg[loc] += smp
tot = sum(agg.values())
srcs = {}
for loc, smp in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if loc is None:
        print("%6d %5.1f%%  <no line>" % (smp, 100.0 * smp / tot)); continue
    f, l = loc
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/background-subtraction_b200/csrc/" + f).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
    print("%6d %5.1f%%  %s:%d  %s" % (smp, 100.0 * smp / tot, f, l, text[:110]))
