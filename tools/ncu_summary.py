#!/usr/bin/env python
"""Summarise an Nsight Compute report (developer tool): python tools/ncu_summary.py report.ncu-rep [out.json]
Prints, per profiled launch, the metrics DESIGN.md / bench.py quote (duration, DRAM bytes, throughput fractions, issue
utilisation, registers, occupancy, top stall reasons) and optionally writes them as JSON for profiles/."""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
    "sm__inst_executed.sum": "warp_instructions",
    "smsp__inst_executed.sum": "warp_instructions_smsp",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "smsp__inst_executed_op_local_ld.sum": "local_loads",
    "smsp__inst_executed_op_local_st.sum": "local_stores",
}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = dict(zip(head, row))
        u = dict(zip(head, units))
        rec = {"kernel": d.get("Kernel Name", "?")}
        for k, name in KEYS.items():
            if k in d and d[k] != "":
                if name in ("dram_read", "dram_write"):
                    rec[name + "_bytes"] = to_bytes(d[k], u[k])
                elif name == "duration":
                    rec["duration_us"] = float(d[k].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u[k], 1)
                else:
                    try:
                        rec[name] = float(d[k].replace(",", ""))
                    except ValueError:
                        rec[name] = d[k]
        stalls = {}
        for k in d:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                try:
                    stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(d[k])
                except ValueError:
                    pass
        rec["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
            rec["traffic_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
            if rec.get("duration_us"):
                rec["dram_GBps_under_ncu"] = rec["traffic_bytes"] / rec["duration_us"] / 1e3
        res.append(rec)
    for r in res:
        print(json.dumps(r, indent=1))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
