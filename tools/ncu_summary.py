#!/usr/bin/env python
"""Summarise `ncu --set full` reports (one or more .ncu-rep files) as JSON: per captured launch the kernel name, grid, duration,
registers, DRAM bytes, L2 -> SM bytes, tensor-pipe / XU-pipe / L2 / DRAM utilisation and the top warp-stall reasons.

    python tools/ncu_summary.py label=path.ncu-rep [label=path ...] > profiles/r02_ncu_full_summary.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "time_us": ("gpu__time_duration.sum", 1.0),
    "sm_mhz": ("gpc__cycles_elapsed.max.per_second", 1.0),
    "regs": ("launch__registers_per_thread", 1.0),
    "grid": ("launch__grid_size", 1.0),
    "cluster": ("launch__cluster_size", 1.0),
    "dram_read_mb": ("dram__bytes_read.sum", 1.0),
    "dram_write_mb": ("dram__bytes_write.sum", 1.0),
    "l2_to_sm_mb": ("l1tex__m_xbar2l1tex_read_bytes.sum", 1.0),
    "dram_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "lts_pct": ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l1tex_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "tensor_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "tensor_pct_rt": ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    "xu_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    "fma_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "issue_pct": ("sm__inst_executed.avg.pct_of_peak_sustained_active", 1.0),
    "sm_active_cycles": ("sm__cycles_active.avg", 1.0),
    "elapsed_cycles": ("gpc__cycles_elapsed.max", 1.0),
}
UNIT_SCALE = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "Ghz": 1e3, "Mhz": 1.0}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    res = {}
    for arg in sys.argv[1:]:
        label, rep = arg.split("=", 1)
        hdr, units, data = rows_of(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        entries = []
        for d in data:
            e = {"kernel": d[ix["Kernel Name"]][:90]}
            for k, (metric, _) in KEYS.items():
                if metric in ix and d[ix[metric]] not in ("", "n/a"):
                    try:
                        v = float(d[ix[metric]].replace(",", ""))
                    except ValueError:
                        continue
                    u = units[ix[metric]]
                    e[k] = round(v * UNIT_SCALE.get(u, 1.0), 3)
            if "dram_read_mb" in e and "time_us" in e:
                e["dram_gbs"] = round((e["dram_read_mb"] + e.get("dram_write_mb", 0.0)) / e["time_us"] * 1e3, 1)
            stalls = {h: float(d[i]) for h, i in ix.items() if h.startswith("smsp__average_warp") and "issue_stalled" in h
                      and h.endswith("_per_warp_active.pct") and d[i] not in ("", "n/a")}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
            e["top_stalls_pct_of_warp_cycles"] = {k.split("issue_stalled_")[1].split("_per_warp")[0]: round(v, 1) for k, v in top}
            entries.append(e)
        res[label] = entries
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
