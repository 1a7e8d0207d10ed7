#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the small JSON that profiles/ keeps.

    python tools/ncu_summary.py gpurun_out/k1_prof.ncu-rep profiles/r1_k1_ncu_full.json [--note "..."]

Per profiled launch: duration, DRAM bytes read / written (traffic), DRAM / L2 / SM throughput as % of peak, occupancy,
registers, L1/L2 hit rates, atomic / reduction sectors at L2, executed warp instructions and thread efficiency.
"""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__maximum_warps_per_active_cycle_pct": "theoretical_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_sectors_op_atom.sum": "l2_atom_sectors",
    "lts__t_sectors_op_red.sum": "l2_red_sectors",
    # (this ncu build's full set carries the L1 -> crossbar side of the same traffic: sectors of global atomics / reductions sent to L2)
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_atom.sum": "l2_atom_sectors",
    "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum": "l2_red_sectors",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors_from_sm",
    "lts__t_sectors_srcunit_tex_op_write.sum": "l2_write_sectors_from_sm",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_warp_inst",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_per_issue",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_per_issue",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for name, key in WANT.items():
            if name not in hdr:
                continue
            i = hdr.index(name)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if u in SCALE and key in ("duration", "dram_read", "dram_write"):
                v *= SCALE[u]
                u = "s" if key == "duration" else "B"
            d[key] = v
            if key in ("duration", "dram_read", "dram_write"):
                d[key + "_unit"] = u
        if "dram_read" in d and "dram_write" in d:
            d["traffic_bytes"] = d["dram_read"] + d["dram_write"]
            if d.get("duration"):
                d["dram_gbs_under_ncu"] = d["traffic_bytes"] / d["duration"] / 1e9
        launches.append(d)
    return launches


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    res = {"source": rep, "note": note, "how": "ncu --set full --clock-control none --import-source on (one GPU, under gpurun); "
           "read with ncu -i --page raw --csv. Durations under ncu are cold-cache and serialised: not bench numbers.",
           "launches": summarise(rep)}
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res["launches"], indent=1))


if __name__ == "__main__":
    main()
