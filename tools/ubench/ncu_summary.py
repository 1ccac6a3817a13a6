"""profiles/rNN_ncu_summary.json + rNN_traffic.json from one `ncu --set full` report (tools/ubench/collect_profiles.sh).
Usage: python tools/ubench/ncu_summary.py gpurun_out/r02_prof.ncu-rep profiles/r02"""
import csv
import io
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
kern = {}
for r in rows[2:]:
    name = re.sub(r"^void |\(.*$", "", r[hdr.index("Kernel Name")])
    d = {}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if k.startswith("dram__bytes") or k == "gpu__time_duration.sum":
                v *= SCALE.get(units[i], 1.0)
            d[k] = v
    kern.setdefault(name, d)  # first launch of each kernel
summary = {"what": "ncu --set full --clock-control none --import-source on, `python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-e2e` "
                   "(BASELINE configs[1]: -e utf-8 -n 10, 4 GiB), one launch per kernel; bytes in bytes, gpu__time_duration in microseconds "
                   "(cold, serialised: shares, not absolutes)", "kernels": kern}
json.dump(summary, open(out + "_ncu_summary.json", "w"), indent=1)


def find(prefix):
    for k, v in kern.items():
        if k.startswith(prefix):
            return v
    return {}


pf, hd, ga = find("sx_prefilter_kernel"), find("sx_sp_heads_kernel"), find("sx_sp_gather_kernel")
traffic = {"workload": "-e utf-8 -n 10 over 4 GiB random buffer, 1xB200", "source": out + "_ncu_summary.json (ncu --set full, one launch)",
           "sx_prefilter_kernel": {"dram_bytes_read": pf.get("dram__bytes_read.sum"), "dram_bytes_write": pf.get("dram__bytes_write.sum"), "algorithmic_bytes": 4 << 30},
           "sx_sp_scan+gather_kernels": {"dram_bytes_read": ga.get("dram__bytes_read.sum"), "dram_bytes_write": ga.get("dram__bytes_write.sum")},
           "sx_sp_heads_kernel": {"dram_bytes_read": hd.get("dram__bytes_read.sum"), "dram_bytes_write": hd.get("dram__bytes_write.sum")}}
json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)
for k, v in kern.items():
    print(k, {a: round(b, 2) for a, b in v.items() if a in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")})
