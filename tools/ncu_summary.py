"""Summarise one kernel of an .ncu-rep (ncu --set full capture) into the JSON kept under profiles/.
usage: python tools/ncu_summary.py REPORT.ncu-rep OUT.json --signals N --channels C --hw H --K K [--command "..."]
(signals = signals processed by the captured launch; algorithmic bytes = 4*H*W in + 8*K features out per signal)"""
import argparse
import csv
import json
import subprocess

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

ap = argparse.ArgumentParser()
ap.add_argument("report"); ap.add_argument("out")
ap.add_argument("--signals", type=int, required=True); ap.add_argument("--channels", type=int, required=True)
ap.add_argument("--hw", type=int, required=True); ap.add_argument("--K", type=int, required=True)
ap.add_argument("--command", default="")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
metrics = {}
for h in hdr:
    if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h):
        metrics[h] = {"value": vals[col[h]], "unit": units[col[h]]}
dram = sum(float(vals[col[m]].replace(",", "")) * TO_BYTES[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
out = {"command": a.command, "kernel": vals[col["Kernel Name"]], "signals_in_launch": a.signals,
       "patches_per_launch": a.signals // a.channels, "dram_bytes_per_launch": dram,
       "algorithmic_bytes_per_launch": a.signals * (4 * a.hw * a.hw + 8 * a.K), "metrics": metrics}
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps({k: out[k] for k in out if k != "metrics"}, indent=1))
