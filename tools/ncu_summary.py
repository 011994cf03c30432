"""Condense `ncu --page raw --csv` exports into the handful of counters the design discussion uses.
   python tools/ncu_summary.py profiles/r2_decode3_ncu_raw.csv [...] > profiles/r2_ncu_summary.md"""
import csv, sys

KEYS = [("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall LG throttle")]

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, units = rows[0], rows[1]
    print(f"### {path}\n")
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "")
        cells = []
        for k, _ in KEYS:
            v = d.get(k, "")
            try:
                v = f"{float(v.replace(',', '')):.3g}"
            except ValueError:
                pass
            unit = u.get(k, "")
            cells.append(f"{v} {unit}".strip() if unit not in ("%", "", "ratio") else v)
        print(f"| `{name}` | " + " | ".join(cells) + " |")
    print()
