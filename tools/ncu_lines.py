#!/usr/bin/env python
"""Per-source-line executed-instruction / stall-sample breakdown from an .ncu-rep captured with
--import-source on (kernels compiled with -lineinfo).
    python tools/ncu_lines.py rep.ncu-rep kernel_regex [top] [rows_per_launch]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows_per = float(sys.argv[4]) if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur = None, "?"
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split("/")[-1]; continue
    if "Instructions Executed" in r:
        hdr = r; iex = r.index("Instructions Executed"); ism = r.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[iex].isdigit():
        key = (cur, int(r[0]) if r[0].isdigit() else -1)
        agg[key][0] += int(r[iex]); agg[key][1] += int(r[ism]) if r[ism].isdigit() else 0
        if r[1].strip(): agg[key][2] = r[1].strip()
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot}, stall samples {tots}")
for (f, ln), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    extra = f" {n / rows_per:8.1f}/row" if rows_per else ""
    print(f"{f}:{ln:<5d} {100 * n / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% stall{extra}  {src[:90]}")
