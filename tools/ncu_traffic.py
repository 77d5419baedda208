"""DRAM traffic of the HBM-bound kernels against their algorithmic bytes.
usage: ncu_traffic.py report.ncu-rep rows.json out.txt
`rows.json` is the last line printed by tools/ncu_hbm_kernels.py ([kernel-name prefix, what, algorithmic bytes] per op);
the report is its `ncu --set full` capture (one launch per kernel, kernel replay: cold caches, serialised)."""
import csv
import json
import subprocess
import sys

rep, rows_json, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
spec = json.loads([ln for ln in open(rows_json) if ln.startswith("{")][-1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(suffix):
    ks = [h for h in hdr if h == suffix] or [h for h in hdr if h.endswith(suffix)]
    return ks[0] if ks else None


K = {k: col(k) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
                         "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3,
         "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "s": 1.0}
u = dict(zip(hdr, units))


def val(d, key):
    k = K[key]
    if k is None or d.get(k, "") == "":
        return float("nan")
    return float(d[k].replace(",", "")) * SCALE.get(u[k], 1.0)


out = [f"ncu --set full --clock-control none, one launch per kernel at batch {spec['B']} (64x64, C = 128); "
       "traffic = dram__bytes_read.sum + dram__bytes_write.sum",
       f"{'kernel':34s} {'us':>8s} {'read MB':>9s} {'write MB':>9s} {'algorithmic MB':>15s} {'traffic/alg':>11s} "
       f"{'alg GB/s':>9s} {'DRAM %':>7s} {'L2 hit %':>8s} {'regs':>5s}  what"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "")
    for pref, what, by in spec["rows"]:
        if pref.split("<")[0] in name and (("<" not in pref) or pref.split("<")[1].replace(" ", "") in name.replace(" ", "")):
            t = val(d, "gpu__time_duration.sum")
            rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
            out.append(f"{pref:34s} {t * 1e6:8.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {by / 1e6:15.1f} {(rd + wr) / by:11.2f} "
                       f"{by / t / 1e9:9.0f} {val(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f} "
                       f"{val(d, 'lts__t_sector_hit_rate.pct'):8.1f} {val(d, 'launch__registers_per_thread'):5.0f}  {what}")
            break
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
