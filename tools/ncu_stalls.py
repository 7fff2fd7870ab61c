"""Top stall sites of an .ncu-rep (source page, SASS level):  python tools/ncu_stalls.py report.ncu-rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d["Kernel Name"][:70], "time", d.get("gpu__time_duration.sum"), "cycles", d.get("sm__cycles_active.avg"), "regs", d.get("launch__registers_per_thread"))
    for k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
        if k in d:
            print("   ", k, d[k])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for n, r in enumerate(data):
    r.append(n)
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:topn]:
    s = int(r[idx["# Samples"]])
    st = {k.replace("stall_", ""): int(r[idx[k]] or 0) for k in stalls}
    st = {k: v for k, v in st.items() if v > 0.15 * s}
    print(f"{r[-1]:5d} {s:6d} {100.0 * s / tot:5.1f}% x{r[idx['Instructions Executed']]:>9s}  {r[idx['Source']][:64]:64s} {st}")
