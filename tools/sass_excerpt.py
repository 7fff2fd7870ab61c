"""Per-kernel count of the Blackwell-only SASS mnemonics in the shipped library (cuobjdump -sass): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG = TMA tensor loads, UBLKCP = bulk copies, plus HMMA (legacy mma.sync: must be absent).  CPU only.

    python tools/sass_excerpt.py > profiles/r02_sass_tcgen05.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gan_danet_b200", "libgandanet_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
pat = {"UTCHMMA": r"\bUTCHMMA", "UTC*MMA (all)": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP",
       "HMMA (legacy)": r"\bHMMA", "MUFU.EX2": r"\bMUFU\.EX2"}
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in pat.items():
        if re.search(p, line):
            counts[cur][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(counts)} kernels, arch from `cuobjdump -lelf`: sm_100a; counts of SASS mnemonics per kernel (only kernels with tensor-core / TMA instructions)")
print(f"# {'kernel':100s} " + " ".join(f"{k:>14s}" for k in pat))
tot = collections.Counter()
for fn, c in counts.items():
    tot.update(c)
    if c["UTC*MMA (all)"] or c["UTMALDG"] or c["LDTM"]:
        name = demangle(fn)
        name = re.sub(r"\(.*", "", name)[:100]
        print(f"  {name:100s} " + " ".join(f"{c[k]:14d}" for k in pat))
print(f"  {'TOTAL (whole library)':100s} " + " ".join(f"{tot[k]:14d}" for k in pat))
