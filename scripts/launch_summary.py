"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): per-kernel count / sum / mean and the sequences."""
import collections
import csv
import sys

path = sys.argv[1]
keys = sys.argv[2:]
lines = [l for l in open(path) if not l.startswith('==')]
agg, seq = collections.OrderedDict(), []
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')) / 1000.0
    short = row['Kernel Name'].split('(')[0][-50:]
    agg.setdefault(short, []).append(v)
    seq.append((short, v))
tot = sum(sum(v) for v in agg.values())
print("| kernel | launches | sum us | mean us | min | max | share |\n|---|---:|---:|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
    print("| %s | %d | %.1f | %.1f | %.1f | %.1f | %.3f |" % (k, len(v), sum(v), sum(v) / len(v), min(v), max(v), sum(v) / tot))
for key in keys:
    print(key, [round(us) for s, us in seq if key in s][:24])
