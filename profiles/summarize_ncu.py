"""Summarise an `ncu --set full` report of bf_refine_kernel: one column per captured launch + DRAM traffic JSON.

    python profiles/summarize_ncu.py gpurun_out/X.ncu-rep profiles/NAME "command line that was profiled"
writes profiles/NAME_ncu_summary.txt and profiles/NAME_traffic.json."""
import csv
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
           "sm__icc_requests.sum", "sm__icc_request_hit_rate.pct",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main(rep, out_prefix, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"ncu --set full --clock-control none --import-source on -k regex:bf_refine_kernel  {cmd}",
             "(one column per captured launch)", ""]
    traffic = [0.0] * len(data)
    for m in METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        lines.append(f"{m:88s} [{units[i]}] " + "  ".join(r[i] for r in data))
        if m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            for k, r in enumerate(data):
                traffic[k] += to_bytes(r[i], units[i])
    lines.append("")
    # the engine's captured keyframe launches the kernel every keyframe; with no box to refine it returns at once (a few us):
    # those launches are not what the roofline is about
    it = hdr.index("gpu__time_duration.sum")
    work = [k for k, r in enumerate(data) if float(r[it]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[it], 1.0) > 20.0] or list(range(len(data)))
    mean_work = sum(traffic[k] for k in work) / len(work)
    lines.append(f"DRAM traffic per launch (read+write): {[round(t) for t in traffic]} bytes; mean over the {len(work)} launches that refined boxes: {round(mean_work)}")
    open(out_prefix + "_ncu_summary.txt", "w").write("\n".join(lines) + "\n")
    json.dump({"kernel": "bf_refine_kernel", "source": f"{out_prefix}_ncu_summary.txt (ncu --set full, {cmd}, {len(work)} launches with boxes to refine of {len(data)} captured)",
               "dram_bytes_per_launch_mean": mean_work, "per_launch": traffic},
              open(out_prefix + "_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
