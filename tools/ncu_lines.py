#!/usr/bin/env python3
"""Attribute executed SASS instructions of one kernel in an .ncu-rep to source lines (nvdisasm -g line markers).
usage: ncu_lines.py report.ncu-rep build/csrc/sc_19_1.o 'k_stream_collide_pipeILi19ELi0ELi1ELb0ELi0E' [cells]"""
import csv, subprocess, sys, collections, io, re, os, tempfile, glob
rep, lib, kname = sys.argv[1:4]
cells = int(sys.argv[4]) if len(sys.argv) > 4 else 512**3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = hi[0]; h = rows[i0]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[i0 + 1:end] if len(r) == len(h)]
ci, ce = h.index('Source'), h.index('Instructions Executed')
counts = [int(r[ce] or 0) for r in body]
ops = [r[ci] for r in body]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
best = None
for cub in glob.glob(tmp + "/*.cubin"):
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if kname not in out: continue
    # isolate the function
    m = re.search(r"\.text\.[^\n]*" + re.escape(kname) + r"[^\n]*:\n", out)
    if not m: continue
    seg = out[m.end():]
    nxt = re.search(r"\n\s*\.section", seg)
    seg = seg[:nxt.start()] if nxt else seg
    best = seg; break
assert best, "kernel not found"
line = None; instr = []
for l in best.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: instr.append((line, m.group(2)))
print("sass rows", len(body), "disasm instrs", len(instr))
n = min(len(body), len(instr))
per = collections.Counter(); perop = collections.defaultdict(collections.Counter)
for k in range(n):
    per[instr[k][0]] += counts[k]
    parts = ops[k].split(); op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
    perop[instr[k][0]][op] += counts[k]
tot = sum(counts)
for ln, c in per.most_common(45):
    top = ", ".join(f"{o}:{v*32/cells:.1f}" for o, v in perop[ln].most_common(5))
    print(f"{str(ln):34s} {100*c/tot:5.1f}%  {c*32/cells:6.1f}/cell   {top}")
