#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel (no GPU needed).
    python tools/sass_lines.py [--lib pdp_solver_b200/csrc/libpdp_b200.so] [--kernel k_sp_runILb1ELb0ELi2E] [--top 40]
Extracts the cubins of the library (cuobjdump -xelf), disassembles with line info (nvdisasm -g) and prints, for the
kernel whose mangled name contains --kernel, the number of instructions attributed to each file:line.  Inside a loop
body the static count is the per-iteration count, so the table shows what one edge costs in each phase."""
import argparse
import collections
import os
import re
import subprocess
import tempfile

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                "pdp_solver_b200", "csrc", "libpdp_b200.so"))
ap.add_argument("--kernel", default="k_sp_runILb1ELb0ELi2E")
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--range", default="", help="file:lo-hi : list opcodes of the instructions attributed to these lines")
a = ap.parse_args()

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
counts = collections.Counter()
ops = collections.defaultdict(collections.Counter)
total = 0
for cubin in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    inside = False
    cur = None
    for line in out.splitlines():
        if line.startswith(".text."):
            inside = a.kernel in line
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            counts[cur] += 1
            ops[cur][m.group(2).split(".")[0]] += 1
            total += 1
print("kernel *%s*: %d instructions" % (a.kernel, total))
for (f, l), n in counts.most_common(a.top):
    print("%-22s %5d   %s" % ("%s:%d" % (f, l), n, " ".join("%s=%d" % kv for kv in ops[(f, l)].most_common(6))))
if a.range:
    f, r = a.range.split(":")
    lo, hi = [int(x) for x in r.split("-")]
    s = 0
    agg = collections.Counter()
    for (ff, l), n in sorted(counts.items()):
        if ff == f and lo <= l <= hi:
            s += n
            agg.update(ops[(ff, l)])
            print("  %s:%d  %d" % (ff, l, n))
    print("range total %d: %s" % (s, " ".join("%s=%d" % kv for kv in agg.most_common(12))))
