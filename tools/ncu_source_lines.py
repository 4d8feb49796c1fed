"""Attribute ncu warp-stall samples to SOURCE LINES of a kernel (run here, no GPU needed).

ncu's `--page source --csv` lists SASS instructions with their sample counts but no line numbers; `nvdisasm -g` lists the
same instructions of the same cubin with `//## File "...", line N` markers.  Both are in program order, so zipping them
gives samples per (file, line).  Needs the kernel to be compiled with -lineinfo (build.py does).

usage: python tools/ncu_source_lines.py <report.ncu-rep> <object.o> <mangled-name-substring> [<ncu kernel regex>] [top N]
  e.g. python tools/ncu_source_lines.py gpurun_out/r01c_fused_4096.ncu-rep \
           passivetracerflows.jl_b200/build/fused_inst_4096.o k_fused_xILi4096ELi0ELi256E k_fused_x 40
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, func = sys.argv[1], sys.argv[2], sys.argv[3]
kregex = sys.argv[4] if len(sys.argv) > 4 else func
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
    sass = ""
    for c in cubins:
        sass += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, c)], capture_output=True, text=True).stdout
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and func in l)
insts, cur = [], ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("//--------------------- .text"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(2)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if "# Samples" in r)
data = rows[rows.index(hdr) + 1:][: len(insts)]
si = hdr.index("# Samples")
per_line, per_file, per_inst, tot = collections.Counter(), collections.Counter(), [], 0
for (cur, txt), r in zip(insts, data):
    s = int(r[si] or 0)
    tot += s
    per_line[cur] += s
    per_file[cur[0]] += s
    per_inst.append((s, cur, txt))
print(f"{func}: {len(insts)} SASS instructions, {tot} stall samples")
for k, v in per_file.most_common():
    print(f"  {100 * v / tot:5.1f} %  {k}")
print("top source lines:")
for (f, n), v in per_line.most_common(top):
    print(f"  {100 * v / tot:5.1f} %  {f}:{n}")
print("top instructions:")
for s, cur, txt in sorted(per_inst, reverse=True)[:15]:
    print(f"  {100 * s / tot:5.1f} %  {cur[0]}:{cur[1]:<5d} {txt.strip()[:90]}")
