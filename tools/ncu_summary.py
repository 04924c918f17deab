#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full --import-source on`) into a small text file for
profiles/: headline metrics per kernel + executed-instruction mix by SASS opcode.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt [rows_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
    'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows_per_launch = float(sys.argv[3]) if len(sys.argv) > 3 else None
    lines = [f"# summary of {rep} (ncu --set full --clock-control none --import-source on)"]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    names = []
    for r in raw[2:]:
        kn = r[hdr.index('Kernel Name')]
        names.append(kn)
        lines.append(f"\n== {kn}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:72s} {r[i]} {units[i]}")
    for kn in dict.fromkeys(names):
        import re
        mm = re.search(r'([A-Za-z_][A-Za-z_0-9]*)\s*(<|\()', kn)
        short = mm.group(1) if mm else kn
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name",
                                                "regex:" + short]))))
        if len(src) < 3:
            continue
        h = src[1]
        data = [r for r in src[2:] if len(r) == len(h)]
        isrc, iex = h.index('Source'), h.index('Instructions Executed')
        by = collections.Counter()
        for r in data:
            if not r[iex].isdigit():
                continue
            t = r[isrc].split()
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            by[op] += int(r[iex])
        tot = sum(by.values())
        lines.append(f"\n== SASS mix of {short}: {len(data)} SASS lines, {tot} warp-instructions executed"
                     " (source page; sums every captured launch of this kernel)")
        for op, c in by.most_common(22):
            extra = f"  {c / rows_per_launch:9.1f} per row-launch" if rows_per_launch else ""
            lines.append(f"  {op:10s} {c:14d} {100 * c / tot:5.1f}%{extra}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
