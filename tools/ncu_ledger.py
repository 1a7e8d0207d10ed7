#!/usr/bin/env python
"""Fold the per-family ncu captures of tools/gpu_r2_ncu.sh (gpurun_out/r2_<family>.ncu-rep) into ONE small JSON,
profiles/r2_ncu_ledger.json: per profiled launch the duration, DRAM bytes read / written, L2 atomic / reduction sectors,
occupancy, issue utilisation, instruction count -- what DESIGN.md's kernel table cites.

    python tools/ncu_ledger.py gpurun_out profiles/r2_ncu_ledger.json
"""
import glob
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import summarise  # noqa: E402

KEEP = ("kernel", "duration", "dram_read", "dram_write", "traffic_bytes", "dram_pct_of_peak", "l2_pct_of_peak", "sm_pct_of_peak", "achieved_occupancy_pct",
        "theoretical_occupancy_pct", "registers_per_thread", "grid", "block", "dyn_smem_per_block", "l1_hit_pct", "l2_hit_pct", "l2_atom_sectors",
        "l2_red_sectors", "l2_read_sectors_from_sm", "l2_write_sectors_from_sm", "warp_instructions", "active_threads_per_warp_inst", "fp64_pipe_pct", "issue_active_pct", "stall_long_scoreboard_per_issue")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    out = {"how": "ncu --set full --clock-control none --import-source on, one capture per kernel family, bench.py --frames 1024 --steps 2 --octomap-scans 32 "
                  "(tools/gpu_r2_ncu.sh); durations under ncu are cold-cache and serialised: they explain, they are not bench values",
           "families": {}}
    for rep in sorted(glob.glob(os.path.join(src, "r2_*.ncu-rep"))):
        fam = os.path.basename(rep)[3:-8]
        try:
            launches = summarise(rep)
        except Exception as exc:
            out["families"][fam] = {"error": str(exc)[:200]}
            continue
        out["families"][fam] = [{k: l[k] for k in KEEP if k in l} for l in launches]
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    for fam, ls in out["families"].items():
        if isinstance(ls, list):
            for l in ls:
                print("%-14s %-60s %8.3f ms  dram %8.1f MB  red %10.0f atom %9.0f  issue %5.1f%%" % (
                    fam, l["kernel"][:60], 1e3 * l.get("duration", 0), l.get("traffic_bytes", 0) / 1e6, l.get("l2_red_sectors", 0), l.get("l2_atom_sectors", 0),
                    l.get("issue_active_pct", 0)))


if __name__ == "__main__":
    main()
