#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into the per-kernel lines we commit under
profiles/: usage  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "sm__sass_inst_executed_op_local_ld.sum", "sm__sass_inst_executed_op_local_st.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(r[col["Kernel Name"]][:110])
        for m in METRICS:
            if m in col:
                print(f"  {m:78s} {r[col[m]]:>16s} {units[col[m]]}")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") \
                    and "not_issued" not in h:
                try:
                    stalls.append((float(r[col[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top warp stall reasons (warps per issue-active cycle):")
        for v, nm in stalls[:6]:
            print(f"    {nm:30s} {v:.2f}")
        print()


def traffic(path, n, out_json, tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the sphere step's kernel families
    -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic)"""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    fam = {"force": "sph_force", "deposit": "sph_deposit_reg", "sort": "sort_rank"}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        for key, pat in fam.items():
            if pat in name:
                b = sum(float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
                        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                acc.setdefault((key, name.split("(")[0][-60:]), []).append(b)
    tab = json.load(open(out_json)) if os.path.exists(out_json) else {}
    for (key, name), v in acc.items():
        # the deposit is several sweeps per step: report the sum over the distinct sweeps of one step
        tab.setdefault(key, [])
    for key in fam:
        per = [sum(v) / len(v) for (k, _), v in acc.items() if k == key]
        if per:
            tab[key] = [e for e in tab.get(key, []) if e["n"] != n] + [
                {"n": n, "dram_bytes_per_launch": sum(per), "launch_kinds": len(per), "source": tag}]
    json.dump(tab, open(out_json, "w"), indent=1)
    print(json.dumps(tab, indent=1))


if __name__ == "__main__":
    if len(sys.argv) >= 6 and sys.argv[2] == "--traffic":
        traffic(sys.argv[1], int(float(sys.argv[3])), sys.argv[4], sys.argv[5])
    else:
        main(sys.argv[1])
