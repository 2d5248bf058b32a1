"""Text summary of one `ncu --set full` report for profiles/: key metrics from the raw page plus the stall profile of
the hottest SASS lines from the source page.

    python tools/ncu_extract.py gpurun_out/screen_full2.ncu-rep "title" [KERNEL_REGEX [SKIP]] > profiles/r01_ncu_xxx.txt

With KERNEL_REGEX (matched against the function name without template arguments) launch number SKIP (default 0) among
the matching launches of a multi-kernel report is extracted (ncu -i ... -k regex:KERNEL_REGEX -s SKIP -c 1).
"""
import csv
import io
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors.sum.per_second",
        "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__bytes.sum.per_second", "smsp__pcsamp_sample_count",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle")


KERNEL = []


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"] + KERNEL, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    if len(sys.argv) > 3:
        KERNEL.extend(["-k", "regex:" + sys.argv[3], "-s", sys.argv[4] if len(sys.argv) > 4 else "0", "-c", "1"])
    print(f"ncu --set full --clock-control none --import-source on (one launch) -- extracted from {rep}")
    print(title)
    raw = page(rep, "raw")
    head, units, vals = raw[0], raw[1], raw[2]
    print("Kernel:", vals[head.index("Kernel Name")])
    for h, u, v in zip(head, units, vals):
        if any(h == k or h.endswith("." + k) for k in KEYS):
            print(f"{h:100s} {u:14s} {v}")
    src = page(rep, "source")
    h = src[1]
    i_src, i_smp, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    rows = []
    for i, r in enumerate(src[2:]):
        if len(r) <= i_ex:
            continue
        if r[i_smp] == "# Samples":   # a multi-kernel report repeats the header per launch: keep the first launch only
            break
        rows.append((int(r[i_smp] or 0), i, int(r[i_ex] or 0), r[i_src].strip()))
    total = sum(r[0] for r in rows)
    print(f"\nSASS lines: {len(rows)}; warp-state samples: {total}.  Hottest lines (line, samples, % of all, executions, SASS):")
    for smp, i, ex, s in sorted(sorted(rows, reverse=True)[:16], key=lambda x: x[1]):
        print(f"{i:6d} {smp:8d} {100.0 * smp / max(total, 1):5.1f} {ex:12d}  {s[:90]}")


if __name__ == "__main__":
    main()
