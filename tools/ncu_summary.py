"""Summarise an ncu report (CSV of `ncu -i X.ncu-rep --page raw --csv`) into the per-kernel numbers the roofline blocks
cite: duration, DRAM bytes, DRAM / L2 / SM throughput, achieved occupancy, registers.  usage: ncu_summary.py raw.csv out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = {"Kernel Name": "kernel", "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_bytes_read",
        "dram__bytes_write.sum": "dram_bytes_write", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
        "smsp__inst_executed.sum": "warp_instructions", "lts__t_sectors_op_atom.sum": "l2_atom_sectors",
        "lts__t_sectors_op_red.sum": "l2_red_sectors", "l1tex__data_bank_conflicts_pipe_lsu.sum": "smem_bank_conflicts",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct"}
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
out = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    d = {}
    for k, name in want.items():
        if k in idx:
            v = r[idx[k]]
            try:
                v = float(v.replace(",", ""))
            except Exception:
                pass
            d[name] = v
            if k in idx and units[idx[k]] and name not in ("kernel",):
                d[name + "_unit"] = units[idx[k]]
    out.append(d)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for d in out:
    print(d.get("kernel", "?")[:60], d.get("duration"), d.get("duration_unit"), "dram r/w", d.get("dram_bytes_read"), d.get("dram_bytes_read_unit"),
          d.get("dram_bytes_write"), "dram%", d.get("dram_pct"), "l2%", d.get("l2_pct"), "sm%", d.get("sm_pct"))
