#!/usr/bin/env python
"""Static SASS outline of one kernel in an object file: instruction count and opcode mix of the segments
between barriers / MMA issues / first TMEM loads / first TMA loads (the same split tools/ncu_regions.py
applies to executed counts), so that a change in a phase's code size can be read without a GPU.
    python tools/sass_regions.py build/mxprune_fused.o 'k_fused_pruned_attentionILi7ELi13ELb0' [dump.txt]"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append(m.group(2).strip())
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(f"{i} {s}" for i, s in enumerate(ins)))
def op(s):
    t = s.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
segs, start = [], 0
for i, s in enumerate(ins):
    mark = None
    if "BAR.SYNC" in s: mark = "BAR"
    elif "UTCHMMA" in s and "UTCHMMA" not in ins[i - 1]: mark = "MMA"
    elif "LDTM" in s and not any("LDTM" in x for x in ins[max(0, i - 30):i]): mark = "LDTM"
    elif "UTMALDG" in s and not any("UTMALDG" in x for x in ins[max(0, i - 30):i]): mark = "TMA"
    if mark:
        segs.append((start, i + 1, mark)); start = i + 1
segs.append((start, len(ins), "END"))
print(f"{len(ins)} SASS instructions")
for s, e, m in segs:
    if e - s < 12: continue
    ops = {}
    for x in ins[s:e]: ops[op(x)] = ops.get(op(x), 0) + 1
    print(f"[{s:5d}:{e:5d}] {e - s:5d} end={m:4s} {sorted(ops.items(), key=lambda kv: -kv[1])[:10]}")
