#!/usr/bin/env python
"""Hot SASS regions of one kernel in an .ncu-rep: consecutive instructions with the same executed
count are grouped; prints share of executed instructions, opcode mix and stall reasons per group.
    python tools/ncu_regions.py rep.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if "Instructions Executed" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
iex, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iex]) for r in data)
alls = {}
for r in data:
    for i in st:
        alls[hdr[i]] = alls.get(hdr[i], 0) + (int(r[i]) if r[i].isdigit() else 0)
ts = sum(alls.values())
print(f"{len(data)} SASS instructions, {tot} warp-instructions executed, {ts} stall samples")
print("all stalls:", [(k[6:], round(100 * v / ts, 1)) for k, v in sorted(alls.items(), key=lambda x: -x[1])[:9]])
segs, prev, start = [], None, 0
for i, r in enumerate(data):
    n = int(r[iex])
    if prev is None or abs(n - prev) > 0.02 * max(n, prev, 1):
        if prev is not None: segs.append((start, i, prev))
        start = i
    prev = n
segs.append((start, len(data), prev))
for s, e, n in sorted(segs, key=lambda x: -(x[1] - x[0]) * x[2])[:top]:
    ops, d = {}, {}
    for r in data[s:e]:
        t = r[isrc].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
        for i in st:
            d[hdr[i]] = d.get(hdr[i], 0) + (int(r[i]) if r[i].isdigit() else 0)
    sm = sum(d.values())
    print(f"[{s}:{e}] {e - s} instr x {n} = {100 * (e - s) * n / tot:.1f}% inst, {100 * sm / ts:.1f}% samples; "
          f"ops {sorted(ops.items(), key=lambda x: -x[1])[:7]}; stalls {[(k[6:], v) for k, v in sorted(d.items(), key=lambda x: -x[1])[:4]]}")
