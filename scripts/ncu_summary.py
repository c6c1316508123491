"""key counters of an `ncu --set full` report as a markdown table: python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
cols = []
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    cols.append((rep.split("/")[-1], d))
print("| metric | unit | " + " | ".join(n for n, _ in cols) + " |\n|---|---|" + "---:|" * len(cols))
kn = " / ".join(d.get("Kernel Name", ("", "?"))[1][:60] for _, d in cols)
print(f"| kernel | | {' | '.join(d.get('Kernel Name', ('', '?'))[1].replace('|', '/')[:70] for _, d in cols)} |")
for k in KEYS:
    if any(k in d for _, d in cols):
        u = next((d[k][0] for _, d in cols if k in d), "")
        print(f"| `{k}` | {u} | " + " | ".join(d.get(k, ("", "-"))[1] for _, d in cols) + " |")
