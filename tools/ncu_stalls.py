#!/usr/bin/env python3
"""Top stalled SASS instructions of the first kernel in an .ncu-rep, with the dominant stall reason and the source line
(from nvdisasm -g of the object file). usage: ncu_stalls.py report.ncu-rep build/csrc/sc_19_1.o kernel_mangled_substring [reason]"""
import csv, subprocess, sys, io, re, os, tempfile, glob
rep, obj, kname = sys.argv[1:4]
reason = sys.argv[4] if len(sys.argv) > 4 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = hi[0]; h = rows[i0]; end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[i0 + 1:end] if len(r) == len(h)]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
seg = None
for cub in glob.glob(tmp + "/*.cubin"):
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    m = re.search(r"\.text\.[^\n]*" + re.escape(kname) + r"[^\n]*:\n", out)
    if m:
        seg = out[m.end():]; n = re.search(r"\n\s*\.section", seg); seg = seg[:n.start()] if n else seg; break
lines = []; line = None
for l in (seg or "").splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = f"{os.path.basename(m.group(1))}:{m.group(2)}"; continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l): lines.append(line)
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
cs = h.index('# Samples')
tot = sum(int(r[cs] or 0) for r in body)
per_reason = {h[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
print("samples", tot, {k: round(100 * v / tot, 1) for k, v in sorted(per_reason.items(), key=lambda kv: -kv[1]) if v * 50 > tot})
key = (lambda r: int(r[h.index(reason)] or 0)) if reason else (lambda r: int(r[cs] or 0))
order = sorted(range(len(body)), key=lambda k: -key(body[k]))[:40]
for k in order:
    r = body[k]
    top = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100*int(r[cs] or 0)/tot:5.2f}%  {lines[k] if k < len(lines) else '?':24s} {r[1][:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
