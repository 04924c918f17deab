#!/usr/bin/env python
"""Executed-instruction and stall-sample shares of a kernel's phases from an .ncu-rep (source page, SASS): the SASS is
cut at barriers, MMA issues, first TMEM loads and first TMA loads (as tools/sass_regions.py does statically), each
segment is reported with its share of executed warp instructions and of stall samples; with --detail S E the runs of
equal execution count inside [S, E) are listed, with --hot S E N the instructions with >= N samples.
    python tools/ncu_phases.py rep.ncu-rep kernel_regex [--detail S E] [--hot S E N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = None, []
for r in rows:
    if "Instructions Executed" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
iex, isrc, ism = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
st = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iex]) for r in data); ts = sum(int(r[ism]) for r in data)
def op(s):
    t = s.split(); return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
def top(r):
    d = {hdr[i][6:]: int(r[i]) for i in st if r[i].isdigit() and int(r[i]) > 0}
    return sorted(d.items(), key=lambda x: -x[1])[:3]
print(f"{len(data)} SASS instructions, {tot} warp-instructions executed, {ts} samples")
if "--detail" in sys.argv:
    a = sys.argv.index("--detail"); s, e = int(sys.argv[a + 1]), int(sys.argv[a + 2])
    i = s
    while i < e:
        j = i
        while j + 1 < e and abs(int(data[j + 1][iex]) - int(data[i][iex])) <= 0.03 * int(data[i][iex]): j += 1
        ops = {}; n = 0
        for r in data[i:j + 1]:
            ops[op(r[isrc])] = ops.get(op(r[isrc]), 0) + 1; n += int(r[iex])
        sm = sum(int(r[ism]) for r in data[i:j + 1])
        if n / tot >= 0.001: print(f"[{i}:{j+1}] n={j+1-i} x{data[i][iex]} {100*n/tot:.1f}% samp={sm} {sorted(ops.items(), key=lambda x:-x[1])[:9]}")
        i = j + 1
elif "--hot" in sys.argv:
    a = sys.argv.index("--hot"); s, e, n = int(sys.argv[a + 1]), int(sys.argv[a + 2]), int(sys.argv[a + 3])
    for i in range(s, min(e, len(data))):
        if int(data[i][ism]) >= n: print(i, data[i][iex], data[i][ism], data[i][isrc][:70], top(data[i]))
elif "--dump" in sys.argv:
    a = sys.argv.index("--dump"); s, e = int(sys.argv[a + 1]), int(sys.argv[a + 2])
    for i in range(s, min(e, len(data))): print(i, data[i][iex], data[i][ism], data[i][isrc][:90])
else:
    segs, start = [], 0
    for i, r in enumerate(data):
        src = r[isrc]; mark = None
        if 'BAR.SYNC' in src or 'BAR.ARV' in src: mark = 'BAR'
        elif 'UTCHMMA' in src and 'UTCHMMA' not in data[i - 1][isrc]: mark = 'MMA'
        elif 'LDTM' in src and not any('LDTM' in data[j][isrc] for j in range(max(0, i - 30), i)): mark = 'LDTM'
        elif 'UTMALDG' in src and not any('UTMALDG' in data[j][isrc] for j in range(max(0, i - 30), i)): mark = 'TMA'
        if mark:
            segs.append((start, i + 1, mark)); start = i + 1
    segs.append((start, len(data), 'END'))
    for s, e, m in segs:
        n = sum(int(r[iex]) for r in data[s:e]); sm = sum(int(r[ism]) for r in data[s:e])
        if n / tot > 0.004 or sm / ts > 0.004:
            print(f"[{s:5d}:{e:5d}] {100*n/tot:5.1f}% inst {100*sm/ts:5.1f}% samp  exec~{data[s][iex]:>8s} end={m}")
