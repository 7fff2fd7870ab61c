"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/aggregate_launches.py in.csv > out.txt"""
import collections, csv, re, sys
d = collections.defaultdict(lambda: [0, 0.0])
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
for r in csv.DictReader(lines):
    n = re.sub(r"\(.*", "", r["Kernel Name"])
    d[n][0] += 1
    d[n][1] += float(r["Metric Value"])
tot = sum(v[1] for v in d.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in d.values())} launches, {tot / 1e6:.2f} ms summed device time (ncu: cold-cache, serialised -> compare shares)")
for k, v in sorted(d.items(), key=lambda x: -x[1][1]):
    print(f"{v[1] / 1e6:10.3f} ms {v[0]:5d} launches {100 * v[1] / tot:5.1f}%  {k}")
