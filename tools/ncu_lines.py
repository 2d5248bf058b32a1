"""Warp-state samples of one kernel of an ncu report by SOURCE LINE (needs -lineinfo; run where the sources live):

    python tools/ncu_lines.py gpurun_out/r02_chain.ncu-rep front_u8 [top=40] [file-filter] [skip=0]
(skip = launches of the matching base name to skip: template arguments are not part of the name ncu matches)

Prints per file:line the samples, the share of all samples, the dominant stall reasons and the executed instructions;
then the totals per stall reason.  The "cuda,sass" source view lists every SASS instruction under the line it came from.
"""
import collections
import csv
import io
import os
import subprocess
import sys


def to_int(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0


def main():
    rep, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    filt = sys.argv[4] if len(sys.argv) > 4 else ""
    skip = sys.argv[5] if len(sys.argv) > 5 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k",
                          "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True, cwd="/tmp").stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, head = None, None
    per_line = collections.defaultdict(lambda: collections.Counter())
    text = {}
    stalls_total = collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            head = r
            continue
        if head is None or len(r) < len(head) - 2:
            continue
        d = dict(zip(head, r))
        # a CUDA line row carries the aggregate of the SASS rows listed under it (their Address column is set)
        line = d.get("Line No", "")
        if not line:
            continue
        key = (cur_file, int(line))
        text[key] = r[1].strip()
        smp = to_int(d.get("# Samples"))
        per_line[key]["samples"] += smp
        per_line[key]["inst"] += to_int(d.get("Instructions Executed"))
        for h in head:
            if h.startswith("stall_") and "Not Issued" not in h:
                v = to_int(d.get(h))
                if v:
                    per_line[key][h] += v
                    stalls_total[h] += v
    total = sum(v["samples"] for v in per_line.values()) or 1
    print(f"kernel ~ {kern}: {total} warp-state samples over {len(per_line)} source lines")
    items = [(k, v) for k, v in per_line.items() if filt in k[0]]
    for (f, ln), v in sorted(items, key=lambda kv: -kv[1]["samples"])[:top]:
        st = ", ".join(f"{h[6:]} {100 * c // max(v['samples'], 1)}%" for h, c in
                       sorted(((h, c) for h, c in v.items() if h.startswith("stall_")), key=lambda x: -x[1])[:3])
        print(f"{f}:{ln:<5d} {v['samples']:7d} {100.0 * v['samples'] / total:5.1f}%  inst {v['inst']:10d}  [{st}]  {text.get((f, ln), '')[:70]}")
    print("stall totals:", ", ".join(f"{h[6:]} {100.0 * c / total:.1f}%" for h, c in stalls_total.most_common(10)))


if __name__ == "__main__":
    main()
