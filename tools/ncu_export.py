"""Export an ncu report for profiles/: the details page and a selection of raw metrics per kernel.
usage: ncu_export.py report.ncu-rep out_prefix "header line"   ->  <out_prefix>_details.txt, <out_prefix>_selected.txt"""
import csv
import subprocess
import sys

rep, prefix, header = sys.argv[1], sys.argv[2], sys.argv[3]
SEL = [
    "gpu__time_duration.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes.sum.per_second",
    "launch__registers_per_thread",
]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(prefix + "_details.txt", "w").write(det)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
out = [header, ""]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    out.append(f"---- {d.get('Kernel Name', '')[:80]}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}")
    for m in SEL:
        keys = [h for h in hdr if h == m or h.endswith("." + m) or h.endswith(m)]
        if keys:
            out.append(f"   {keys[0]:110s} {d[keys[0]]:>14s} {u[keys[0]]}")
open(prefix + "_selected.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
