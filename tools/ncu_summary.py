#!/usr/bin/env python
"""ncu report -> the small JSON summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_<what>_ncu_full_summary.json \
        --command "<the command that was profiled>" --workload "<text>" [--alg-bytes N]

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps, per profiled
launch, the metrics the roofline discussion needs.
"""
import argparse
import csv
import io
import json
import subprocess

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__cycles_active.avg", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "l1tex__m_xbar2l1tex_read_sectors.sum", "lts__t_requests.sum", "lts__t_sectors.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.max",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--command", default="")
    ap.add_argument("--workload", default="")
    ap.add_argument("--alg-bytes", type=float, default=None)
    ap.add_argument("--alg-flops", type=float, default=None)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"Kernel Name": {"value": r[header.index("Kernel Name")], "unit": ""}}
        for i, name in enumerate(header):
            if name in KEEP or any(name.startswith(k) for k in ("sm__pipe_tensor", )):
                try:
                    val = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[name] = {"value": val, "unit": units[i]}

        def get(name, scale_units):
            if name not in d:
                return None
            v, u = d[name]["value"], d[name]["unit"].lower()
            return v * scale_units.get(u, 1.0)
        byte_units = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
        rd, wr = get("dram__bytes_read.sum", byte_units), get("dram__bytes_write.sum", byte_units)
        if rd is not None and wr is not None:
            d["traffic_bytes_per_launch"] = rd + wr
        if a.alg_bytes:
            d["algorithmic_bytes_per_launch"] = a.alg_bytes
        t = get("gpu__time_duration.sum", {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3,
                                          "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0})
        if t and a.alg_bytes:
            d["algorithmic_GBps"] = a.alg_bytes / t / 1e9
        if t and a.alg_flops:
            d["algorithmic_TFLOPs"] = a.alg_flops / t / 1e12
        launches.append(d)
    with open(a.out, "w") as f:
        json.dump({"command": a.command, "workload": a.workload, "launches": launches}, f, indent=1)
    print("wrote %s (%d launches)" % (a.out, len(launches)))


if __name__ == "__main__":
    main()
