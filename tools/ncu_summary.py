"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [ndofs_per_launch] > profiles/<name>.json
"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def main(path, ndofs=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        stalls = {}
        for i, k in enumerate(h):
            if k in KEYS:
                d[k] = f"{r[i]} {u[i]}".strip()
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k:
                try:
                    stalls[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        d["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        if ndofs:
            tr = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                val, unit = d[k].split()
                tr += float(val.replace(",", "")) * SCALE[unit]
            d["dram_bytes_per_launch"] = tr
            d["dram_bytes_per_dof"] = tr / ndofs
            inst = float(d["smsp__inst_executed.sum"].split()[0].replace(",", ""))
            d["warp_inst_per_dof"] = inst / ndofs
            wf = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].split()[0].replace(",", ""))
            d["smem_wavefronts_per_element64"] = wf / (ndofs / 64)
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
