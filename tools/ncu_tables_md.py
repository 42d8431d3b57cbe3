"""Markdown table of the judged metrics from `ncu -i <rep> --page raw --csv` dumps (one column per dump).

    python tools/ncu_tables_md.py label=gpurun_out/x_raw.csv [label=...] > profiles/rN_ncu_raw_tables.md
"""
import csv
import statistics
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name = body[0][col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    out = {}
    for m in METRICS:
        if m in col:
            vals = [float(r[col[m]].replace(",", "")) for r in body if r[col[m]] not in ("", "n/a")]
            if vals:
                out[m] = (statistics.median(vals), units[col[m]])
    return name, len(body), out


def fmt(v):
    return f"{v:,.0f}" if abs(v) >= 1e4 else f"{v:.4g}"


def main():
    cols = [a.split("=", 1) for a in sys.argv[1:]]
    data = [(lab, *load(p)) for lab, p in cols]
    print("| metric | " + " | ".join(f"{lab}: `{name}` ({n} launches, median)" for lab, name, n, _ in data) + " | unit |")
    print("|---|" + "---|" * (len(data) + 1))
    for m in METRICS:
        if any(m in d[3] for d in data):
            unit = next(d[3][m][1] for d in data if m in d[3])
            print(f"| `{m}` | " + " | ".join(fmt(d[3][m][0]) if m in d[3] else "-" for d in data) + f" | {unit} |")


if __name__ == "__main__":
    main()
