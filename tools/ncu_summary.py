"""ncu report (.ncu-rep) -> small JSON summary of the metrics the bench and DESIGN cite (run here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [...] > profiles/ncu_full_r02.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_inst",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
         "nsecond": 1e-9, "second": 1.0}
out = {"source": sys.argv[1:], "kernels": []}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = {"report": rep}
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        rec["kernel"] = d.get("Kernel Name", "")
        for k, name in WANT.items():
            if k in d and d[k] != "":
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                rec[name] = v * SCALE.get(u.get(k, ""), 1.0)
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_bytes_per_launch"] = rec["dram_read"] + rec["dram_write"]
        out["kernels"].append(rec)
print(json.dumps(out, indent=1))
