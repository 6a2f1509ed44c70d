#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics of every captured launch + executed SASS opcode mix of the first kernel."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"{k} = {r[i][:90]} {units[i]}")
    for i, k in enumerate(hdr):
        if 'issue_stalled' in k and 'per_issue_active' in k and r[i] and float(r[i]) > 0.08:
            print("  stall", k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i][:6])
    if cells:
        i = hdr.index('smsp__inst_executed.sum'); print("warp-instr per cell x32 =", float(r[i]) * 32 / cells)
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    i0 = hi[0]; h = rows[i0]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
    body = [r for r in rows[i0 + 1:end] if len(r) == len(h)]
    ci, ce = h.index('Source'), h.index('Instructions Executed')
    ops = collections.Counter(); tot = 0
    for r in body:
        parts = r[ci].split(); op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
        n = int(r[ce] or 0); ops[op] += n; tot += n
    print("static", len(body), "executed", tot)
    for op, n in ops.most_common(28):
        print(f"  {op:8s} {100*n/tot:5.1f}%  per-cell(thread-instr) {n*32/cells if cells else 0:7.1f}")
