"""ncu report -> text summary + per-kernel DRAM traffic (profiles/).

  python tools/ncu_summary.py gpurun_out/prof_v5.ncu-rep profiles/r01_v5_ncu_full_c2.txt c2 "comment"

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed), keeps the metrics the
hot-path rooflines are argued with, and merges the DRAM bytes per launch into
profiles/dram_traffic.json under "<workload>:<bench kernel name>"."""
import csv
import json
import os
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]

# report kernel name -> key of bench.py's roofline table ("<key>" = first word of the entry).
# `state` carries what the launch order tells: the radix passes before the first gather of a
# step sort the vertex-face list, those after it the edge list.
def bench_name(k, state):
    vf = "<(bool)1" in k or "<1," in k or "<1>" in k
    if "sweep_count" in k:
        return "sweep_count_vf" if vf else "sweep_count_ee"
    if "sweep_place_kernel" in k:
        return "sweep_place_vf" if vf else "sweep_place_ee"
    if "narrow_cull_kernel" in k:
        return "narrow_cull_vf" if vf else "narrow_cull_ee"
    if "narrow_round_kernel" in k or "narrow_coop_kernel" in k or "narrow_group_kernel" in k:
        # all solver launches of a pass (scout, bulk, later rounds) add up under one key: the
        # kernel name does not tell the rounds apart
        return "narrow_solve_vf" if vf else "narrow_solve_ee"
    if "gather_sorted" in k or "gather_rebuild" in k:
        state["list"] = 1 - state.get("list", 0)
        return "gather"
    if "radix_pass_kernel" in k or "radix_hist_kernel" in k or "digit_hist_kernel" in k:
        if "DestDigit" in k:
            return "partition"
        return "radix_sort_vf" if state.get("list", 0) == 0 else "radix_sort_ee"
    if "boxes_kernel" in k:
        return "boxes"
    if "expand_" in k:
        return "expand"
    return None


def main():
    rep, out, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    comment = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    traffic = {}
    state = {}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, workload {workload}. {comment}\n")
        f.write("# one block per profiled launch, in launch order; times are cold-cache and serialised\n")
        for r in rows[2:]:
            k = r[name_i]
            f.write(f"\n## {k[:160]}\n")
            for m in KEEP:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m:90s} {r[i]:>18s} {units[i]}\n")
            if "vertex_boxes_kernel" in k:      # a step begins; traffic is summed over the first
                state["step"] = state.get("step", 0) + 1   # COMPLETE step of the capture
                state["list"] = 0
            b = bench_name(k, state) if state.get("step", 0) == 1 else None
            if b and "dram__bytes_read.sum" in hdr:
                def val(m):
                    i = hdr.index(m)
                    v = float(r[i].replace(",", ""))
                    u = units[i].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                # a bench "kernel" may be several launches (narrow rounds, two box kernels,
                # two gathers): their traffic adds up per step
                traffic[b] = traffic.get(b, 0.0) + val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    tj = os.path.join(os.path.dirname(out), "dram_traffic.json")
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    for b, v in traffic.items():
        cur[f"{workload}:{b}"] = v
    json.dump(cur, open(tj, "w"), indent=1)
    print("wrote", out, "and", tj, traffic)


if __name__ == "__main__":
    main()
