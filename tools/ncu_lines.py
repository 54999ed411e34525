#!/usr/bin/env python3
"""Per-source-line totals of an ncu capture: joins `ncu --page source --csv` (SASS rows) with the line
table of `nvdisasm --print-line-info` of the same cubin.
    tools/ncu_lines.py <report.ncu-rep> <object.o|.so> <kernel-substring> [top]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
lines = {}
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
    inside, cur = False, ("?", 0)
    for ln in txt:
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            lines[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ci = {n: H.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed",
                              "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Global")}
stall_cols = [(i, n) for i, n in enumerate(H) if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[hdr + 1:] if len(r) > ci["Thread Instructions Executed"]]
base = int(body[0][ci["Address"]], 16)
agg = {}
tot = [0, 0, 0]
for r in body:
    off = int(r[ci["Address"]], 16) - base
    key = lines.get(off, ("?", 0))
    a = agg.setdefault(key, {"inst": 0, "tinst": 0, "samp": 0, "sw": 0, "swi": 0, "l2": 0, "stalls": {}})
    a["inst"] += int(r[ci["Instructions Executed"]] or 0)
    a["tinst"] += int(r[ci["Thread Instructions Executed"]] or 0)
    a["samp"] += int(r[ci["# Samples"]] or 0)
    a["sw"] += int(r[ci["L1 Wavefronts Shared"]] or 0)
    a["swi"] += int(r[ci["L1 Wavefronts Shared Ideal"]] or 0)
    a["l2"] += int(r[ci["L2 Theoretical Sectors Global"]] or 0)
    for i, n in stall_cols:
        v = int(r[i] or 0)
        if v:
            a["stalls"][n] = a["stalls"].get(n, 0) + v
    tot[0] += int(r[ci["Instructions Executed"]] or 0)
    tot[1] += int(r[ci["# Samples"]] or 0)
    tot[2] += int(r[ci["Thread Instructions Executed"]] or 0)
print("total warp-instructions %d, thread-instructions %d, samples %d" % (tot[0], tot[2], tot[1]))
print("%-28s %6s %6s %5s %9s %9s %9s  top stalls" % ("file:line", "inst%", "samp%", "thr", "smem_wf", "wf_ideal", "l2_sect"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    print("%-28s %6.2f %6.2f %5.1f %9d %9d %9d  %s" % ("%s:%d" % key, 100.0 * a["inst"] / max(tot[0], 1), 100.0 * a["samp"] / max(tot[1], 1),
                                                   a["tinst"] / max(a["inst"], 1), a["sw"], a["swi"], a["l2"],
                                                   " ".join("%s=%d" % (n[6:], v) for n, v in st)))

if os.environ.get("NCU_LINES_ALL"):
    print("\nall lines by thread instructions:")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["tinst"])[:int(os.environ["NCU_LINES_ALL"])]:
        print("%-28s tinst%% %6.2f inst%% %6.2f thr %5.1f" % ("%s:%d" % key, 100.0 * a["tinst"] / tot[2], 100.0 * a["inst"] / tot[0], a["tinst"] / max(a["inst"], 1)))
