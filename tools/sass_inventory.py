"""Static inventory of the built library (no GPU needed): per kernel family the registers / stack / shared memory ptxas
assigned and the Blackwell-specific SASS mnemonics it contains (LDTM/STTM = tcgen05.ld/st, CCTL/prefetch, LDG.E.ENL2.256 =
256-bit loads, DFMA/DADD/DMUL = the fp64 pipe).  usage: python tools/sass_inventory.py [lib.so] > profiles/rNN_sass_inventory.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "passivetracerflows.jl_b200", "libptf_b200.so")
# the default instantiations of BASELINE configs[1] (4096^2) and of the partitioned 1024^3 leg, plus everything at 1024
WANT = re.compile(r"k_fused_[xy]ILi(4096|1024)E|k_yinv3ILi1024E|k_yfwd3ILi1024E|k_xbarrier|k_sep_fill|k_fused1d")
MNEMONICS = ["LDTM", "STTM", "UTCBAR", "CCTL", "LDG.E.ENL2.256", "LDG.E.128", "LDG.E.64", "STG.E.ENL2.256", "STG.E.128",
             "LDS.128", "STS.128", "LDL", "STL", "DFMA", "DADD", "DMUL", "BAR.SYNC", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
fn = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and fn:
        usage[fn] = tuple(int(v) for v in m.groups())
        fn = None
keep = [f for f in usage if WANT.search(f)]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1) if m.group(1) in usage and WANT.search(m.group(1)) else None
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["instructions"] += 1
    for mn in MNEMONICS:
        if op == mn or op.startswith(mn + ".") or (mn.count(".") and op.startswith(mn)):
            counts[cur][mn] += 1
names = demangle(keep)
print(f"# {os.path.relpath(lib, ROOT)}: {len(usage)} kernels in the library, {len(keep)} listed (4096 / 1024-point instantiations)")
print("# static counts of SASS instructions (not executed counts); stack > 0 = spills or local arrays")
seen = set()
for f in sorted(keep, key=lambda f: names[f]):
    r, st, sh = usage[f]
    if (names[f], usage[f]) in seen:      # file-local kernels appear once per translation unit that includes them
        continue
    seen.add((names[f], usage[f]))
    c = counts[f]
    short = re.sub(r"ptf::\(anonymous namespace\)::|void ", "", names[f])
    short = re.sub(r"\((ptf::)?\(?.*$", "", short)
    print(f"{short}\n    regs {r}  stack {st} B  static smem {sh} B  SASS {c['instructions']}  " +
          "  ".join(f"{mn} {c[mn]}" for mn in MNEMONICS if c[mn]))
