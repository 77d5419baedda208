"""Summarise an ncu report's source page: top stall-sample instructions and every mbarrier wait.
usage: ncu_src_summary.py report.ncu-rep [ntop]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = rows[2:]
base = int(data[0][0], 16)
tot = sum(int(r[2]) for r in data)
print("total samples", tot)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -int(r[2]))[:ntop]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[0], 16) - base:#7x} {r[1].strip()[:70]:70s} {int(r[2]):6d} {100 * int(r[2]) / tot:5.1f}% exec {r[5]:>9s} {st}")
print("--- waits")
for i, r in enumerate(data):
    if "TRYWAIT" in r[1]:
        s = int(r[2]) + int(data[i + 1][2])
        print(f"{int(r[0], 16) - base:#7x} {r[1].strip()[:70]:70s} {s:6d} {100 * s / tot:5.1f}% exec {r[5]}")
