#!/usr/bin/env python3
"""Summarise .ncu-rep captures (ncu --set full) into a markdown table + profiles/traffic.json.

    python tools/ncu_summary.py OUT.md NAME=report.ncu-rep[:ALGORITHMIC_BYTES] ...

Reads each report with `ncu -i ... --page raw --csv` (works without a GPU) and prints, per kernel
launch in it: duration, DRAM read/write bytes (the `traffic` of bench.py's roofline object), DRAM
throughput, registers, occupancy limiters, shared-memory bank conflicts, executed instructions.
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__bytes.sum.per_second", "dram_bw"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e-6, "ms": 1e-3, "ns": 1e-9,
         "s": 1, "byte/s": 1, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def read_report(path):
    if path.endswith(".csv"):  # the raw page already exported on the GPU box (`ncu -i rep --page raw --csv`)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m, short in METRICS:
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[short] = v * SCALE.get(units[i], 1)
        launches.append(d)
    return launches


def main():
    out_md = sys.argv[1]
    traffic_path = os.path.join(os.path.dirname(out_md), "traffic.json")
    traffic = {}
    if os.path.exists(traffic_path):
        traffic = json.load(open(traffic_path))
    lines = ["| capture | kernel | time (us) | DRAM read (MB) | DRAM write (MB) | DRAM traffic / algorithmic | DRAM GB/s | "
             "DRAM % | L2 hit % | SM % | regs | CTA/SM limit (regs, smem) | grid x block | smem bank conflicts | warp insts |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for spec in sys.argv[2:]:
        name, rest = spec.split("=", 1)
        path, _, alg = rest.partition(":")
        if name.startswith("plan:"):
            # every launch of ONE exec of a plan (captured with --cache-control none, so what one pass leaves in L2 is there
            # for the next): their DRAM bytes summed = the traffic of the whole transform; bench.py's shapes[] rows read it
            ls = read_report(path)
            tot = sum(d.get("dram_read", 0) + d.get("dram_write", 0) for d in ls)
            traffic[name] = {"dram_bytes_per_exec": tot, "algorithmic_bytes_per_exec": float(alg) if alg else None,
                             "launches": len(ls), "kernels": [d["kernel"].strip()[:160] for d in ls],
                             "time_us_under_ncu": sum(d.get("time", 0) for d in ls) * 1e6, "source": os.path.basename(path)}
        for d in read_report(path):
            tr = d.get("dram_read", 0) + d.get("dram_write", 0)
            ratio = "%.3f" % (tr / float(alg)) if alg else "-"
            if not name.startswith("plan:"):  # (a plan entry holds the sum over its launches, written above)
                traffic[name] = {"dram_bytes_per_launch": tr, "algorithmic_bytes_per_launch": float(alg) if alg else None,
                                 "kernel": d["kernel"].strip(), "source": os.path.basename(path)}
            k = d["kernel"].strip().replace("|", "/")
            lines.append("| %s | `%s` | %.1f | %.1f | %.1f | %s | %.0f | %.1f | %.1f | %d | %d, %d | %d x %d | %d | %.3g |" % (
                name, k, d.get("time", 0) * 1e6, d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6, ratio,
                d.get("dram_bw", 0) / 1e9, d.get("l2_hit_pct", 0), d.get("sm_pct", 0), d.get("regs", 0),
                d.get("occ_lim_regs", 0), d.get("occ_lim_smem", 0), d.get("grid", 0), d.get("block", 0),
                d.get("smem_bank_conflicts", 0), d.get("warp_insts", 0)))
    with open(out_md, "a") as f:
        f.write("\n".join(lines) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
