#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, opcode mix and the
hottest source lines (SASS samples mapped through nvdisasm line info).

    python profiles/analyze_ncu.py gpurun_out/prof.ncu-rep <kernel-substr> <cubin-substr> [out.md]
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter

rep, kname, cubin_sub = sys.argv[1], sys.argv[2], sys.argv[3]
out = open(sys.argv[4], "w") if len(sys.argv) > 4 else sys.stdout
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vsmartmom.jl_b200", "lib", "libvsmom_b200.so")

def sh(cmd, **kw):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True, **kw).stdout

raw = list(csv.reader(io.StringIO(sh(f"ncu -i {rep} --page raw --csv"))))
hdr, units = raw[0], raw[1]
iK = hdr.index("Kernel Name")
vals = next((r for r in raw[2:] if kname in r[iK]), raw[2])      # the launch of the requested kernel
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]
print(f"# ncu summary: {os.path.basename(rep)}\n", file=out)
for k in want:
    if k in M:
        print(f"- `{k}` = {M[k][0]} {M[k][1]}", file=out)

src = list(csv.reader(io.StringIO(sh(f"ncu -i {rep} --page source --csv --kernel-name regex:{kname}"))))
# the export holds one section per launch / view: "Kernel Name" line, header line, rows; take the first (SASS) one
starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
end = starts[1] if len(starts) > 1 else len(src)
h = src[starts[0] + 1]; data = [r for r in src[starts[0] + 2:end] if len(r) == len(h)]
iS, iI, iA, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Address"), h.index("Source")
tot = sum(int(r[iS]) for r in data) or 1
toti = sum(int(r[iI]) for r in data) or 1
ops, opsS = Counter(), Counter()
for r in data:
    t = r[iSrc].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    ops[op] += int(r[iI]); opsS[op] += int(r[iS])
print("\n## opcode mix (executed warp instructions / stall samples)\n", file=out)
for op, v in ops.most_common(14):
    print(f"- {op:10s} {100*v/toti:5.1f}% instr  {100*opsS[op]/tot:5.1f}% samples", file=out)

with tempfile.TemporaryDirectory() as td:
    sh(f"cuobjdump -xelf all {LIB}", cwd=td)
    cub = [f for f in os.listdir(td) if cubin_sub in f]
    off2line = {}
    if cub:
        dis = sh(f"nvdisasm -g -c {os.path.join(td, cub[0])}")
        infunc, cur = False, None
        for ln in dis.split("\n"):
            if ln.startswith(".text."):
                infunc = kname in ln; continue
            if not infunc: continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", ln)
            if m: off2line[int(m.group(1), 16)] = cur
base = int(data[0][iA], 16)
cs, ci = Counter(), Counter()
for r in data:
    l = off2line.get(int(r[iA], 16) - base, ("?", 0))
    cs[l] += int(r[iS]); ci[l] += int(r[iI])
print("\n## hottest source lines (stall samples / executed instructions)\n", file=out)
srcs = {}
for (f, ln), v in cs.most_common(28):
    if f not in srcs:
        p = os.path.join(ROOT, "vsmartmom.jl_b200", "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print(f"- {f}:{ln:<4d} {100*v/tot:5.1f}% samples {100*ci[(f,ln)]/toti:5.1f}% instr   `{text}`", file=out)
