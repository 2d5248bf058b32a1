"""Whole-chain ncu summary of the encoder: per kernel duration, DRAM bytes and tensor-pipe utilisation, plus the two
numbers bench.py reads (DRAM bytes per pattern, time-weighted tensor-pipe %).

Capture (one GPU, after the same command has exited 0 without ncu; ONE chunk of 1184 patterns = 12 launches, the
first `--launch-skip` launches are the warm-up encodes of tools/encode_once.py and the weight packing):

    python tools/encode_once.py 1184 1 > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none -k regex:'conv3x3_fused|conv0_stats|heads_norm' -s 12 -c 12 \
        -o gpurun_out/r02_chain python tools/encode_once.py 1184 1

Then here (no GPU needed):

    python tools/ncu_chain_summary.py gpurun_out/r02_chain.ncu-rep 1184 profiles/r02_encoder_chain_ncu
        -> profiles/r02_encoder_chain_ncu.txt (table) and .json (read by bench.py)
"""
import csv
import io
import json
import re
import subprocess
import sys

# Calibration (profiles/r02_tensor_metric_calibration.txt: tools/mma_rate.cu under `ncu --set full`): a back-to-back
# M128 x N256 x K16 stream reads 96 % on sm__pipe_tensor_cycles_active (kind::f16 AND kind::f8f6f4), while the
# `_realtime` variant reads 8 / 33 / 13 / 21 % for launches of that same full-rate kernel -- it is not usable.
TENSOR = ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "smsp__issue_active.avg.pct_of_peak_sustained_active") + TENSOR


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def scale(value, unit, kind):
    """Normalise to microseconds / bytes."""
    if kind == "time":
        return value * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6,
                        "second": 1e6}.get(unit, 1.0)
    return value * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    rep, n_patterns, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(head):
        for w in WANT + ("Kernel Name",):
            if h == w or h.endswith("." + w):
                col.setdefault(w, i)
    tensor_metric = next((t for t in TENSOR if t in col), None)
    table, tot_us, tot_bytes, tensor_weighted = [], 0.0, 0.0, 0.0
    for r in rows[2:]:
        if len(r) != len(head):
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("ebsd::", "")
        us = scale(to_float(r[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]], "time")
        rd = scale(to_float(r[col["dram__bytes_read.sum"]]), units[col["dram__bytes_read.sum"]], "bytes")
        wr = scale(to_float(r[col["dram__bytes_write.sum"]]), units[col["dram__bytes_write.sum"]], "bytes")
        tp = to_float(r[col[tensor_metric]]) if tensor_metric else float("nan")
        mt = to_float(r[col["sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]) \
            if "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in col else float("nan")
        l2 = to_float(r[col["lts__t_sector_hit_rate.pct"]]) if "lts__t_sector_hit_rate.pct" in col else float("nan")
        mhz = to_float(r[col["sm__cycles_elapsed.avg.per_second"]]) if "sm__cycles_elapsed.avg.per_second" in col else float("nan")
        table.append((name, us, rd, wr, tp, mt, l2, mhz))
        tot_us += us
        tot_bytes += rd + wr
        if tp == tp:
            tensor_weighted += tp * us
    lines = [f"ncu --set full --clock-control none, one chunk of {n_patterns} patterns ({len(table)} launches), from {rep}",
             f"tensor metric: {tensor_metric}",
             f"{'kernel':58s} {'us':>9s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'tensor %':>9s} {'mem-tensor %':>12s} {'L2 hit %':>9s} {'SM MHz':>7s}"]
    for name, us, rd, wr, tp, mt, l2, mhz in table:
        lines.append(f"{name[:58]:58s} {us:9.1f} {rd / 1e6:11.1f} {wr / 1e6:11.1f} {tp:9.1f} {mt:12.1f} {l2:9.1f} {mhz / 1e6 if mhz > 1e5 else (mhz * 1e3 if mhz < 10 else mhz):7.0f}")
    tw = tensor_weighted / tot_us if tot_us else float("nan")
    lines.append(f"{'chain total':58s} {tot_us:9.1f} {'':11s} {tot_bytes / 1e6:11.1f} {tw:9.1f}   (tensor % weighted by kernel time)")
    lines.append(f"DRAM bytes per pattern: {tot_bytes / n_patterns:,.0f}   (algorithmic: 16 384 in + 128 out)")
    open(out + ".txt", "w").write("\n".join(lines) + "\n")
    json.dump({"dram_bytes_per_pattern": tot_bytes / n_patterns, "tensor_pipe_pct_time_weighted": tw,
               "tensor_metric": tensor_metric, "chain_us_under_ncu": tot_us, "patterns": n_patterns,
               "source": f"{out.split('/')[-1]}.txt (ncu --set full, one chunk of {n_patterns} patterns, this build)",
               "per_kernel": [{"kernel": n, "us": u, "dram_read": a, "dram_write": b, "tensor_pct": t, "l2_hit_pct": l}
                              for n, u, a, b, t, _, l, _ in table]}, open(out + ".json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
