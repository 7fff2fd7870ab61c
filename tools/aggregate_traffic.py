"""Per-kernel-family time and DRAM traffic from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
log of one training step (tools/profile_step.py):  python tools/aggregate_traffic.py launches.csv [out.json]"""
import collections, csv, json, re, sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
per = collections.OrderedDict()
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    else:
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(u, 1)
    per.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = v
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    n = re.sub(r"\(.*", "", d["name"]); n = re.sub(r"^void ", "", n); n = n.split("::")[-1]; n = re.sub(r"<.*", "", n)
    f = fam[n]; f[0] += 1; f[1] += d.get("gpu__time_duration.sum", 0); f[2] += d.get("dram__bytes_read.sum", 0); f[3] += d.get("dram__bytes_write.sum", 0)
tot_t = sum(f[1] for f in fam.values()); tot_b = sum(f[2] + f[3] for f in fam.values())
print(f"# {sys.argv[1]}: {len(per)} launches, {tot_t:.2f} ms summed device time, {tot_b / 1e9:.1f} GB DRAM traffic ({tot_b / 1e9 / (tot_t * 1e-3):.0f} GB/s average)")
print("# ncu serialises the launches (cold caches): compare shares, not absolutes")
out = {}
for n, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    out[n] = {"launches": f[0], "ms": round(f[1], 3), "dram_read_bytes_per_launch": round(f[2] / f[0]), "dram_write_bytes_per_launch": round(f[3] / f[0]),
              "gbs": round((f[2] + f[3]) / 1e9 / (f[1] * 1e-3))}
    print(f"{f[1]:8.3f} ms {f[0]:5d} launches {100 * f[1] / tot_t:5.1f}%   read {f[2] / 1e9:6.2f} GB  write {f[3] / 1e9:6.2f} GB  {out[n]['gbs']:5d} GB/s  {n}")
if len(sys.argv) > 2:
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on tools/profile_step.py "
                         "(one G+D training step, B = 32, 64x128 grid)", "step_ms_summed": round(tot_t, 2), "step_dram_bytes": round(tot_b), "families": out},
              open(sys.argv[2], "w"), indent=1)
