#!/usr/bin/env python3
"""Turns the raw artefacts of tools/make_profiles.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:
   <tag>_sp_run_ncu.md (key metrics + per-source-line hot spots of k_sp_run), k_sp_run_traffic.json,
   <tag>_launches.csv (copy) + <tag>_launches_summary.md, <tag>_bench.json.
       python tools/summarize_profiles.py r1 <edge_updates_in_the_ncu_launch>"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
edge_updates = float(sys.argv[2])
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
rep = os.path.join(G, tag + "_sp_run.ncu-rep")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
H, U, Vv = rows[0], rows[1], rows[2]
m = {n: (Vv[i], U[i]) for i, n in enumerate(H)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def gb(name):
    v, u = m[name]
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[u]


dram = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
per = dram / edge_updates
json.dump({"dram_bytes_per_edge_update": per, "dram_bytes_read": gb("dram__bytes_read.sum"),
           "dram_bytes_write": gb("dram__bytes_write.sum"), "edge_updates_in_profiled_launch": edge_updates,
           "source": "ncu --set full --clock-control none, tools/make_profiles.sh, %s_sp_run.ncu-rep" % tag},
          open(os.path.join(P, "k_sp_run_traffic.json"), "w"), indent=1)
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep,
                        os.path.join(ROOT, "pdp_solver_b200", "csrc", "pdp_loop.o"), "k_sp_runILb1ELb0ELi2", "30"],
                       capture_output=True, text=True).stdout
with open(os.path.join(P, tag + "_sp_run_ncu.md"), "w") as f:
    f.write("# ncu --set full: k_sp_run<blocked> (persistent SP propagate / decimate loop)\n\n")
    f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:k_sp_run -c 1 python tools/prof_sweep.py "
            "--problems 8 --iterations 12` (8 x random 3-SAT n = 1 000 000, alpha = 4.2: E = 100.8 M edges, 12 iterations = "
            "%.4g edge-updates in the launch).  Numbers under the profiler are not bench values.\n\n" % edge_updates)
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for n in want:
        if n in m:
            f.write("| %s | %s | %s |\n" % (n, m[n][0], m[n][1]))
    f.write("\nDRAM bytes per edge-update: **%.2f B** (algorithmic 20 B; messages 24 B incl. the second survey buffer of the "
            "convergence statistic + 2 x 2 B of node-order position tables + run tables, block descriptors, degree-sorted variable list).\n" % per)
    f.write("Warp instructions per edge-update: %.2f (thread instructions: see the first line of the table below); issue slots busy %s %%.\n\n" % (
        float(m["smsp__inst_executed.sum"][0].replace(",", "")) / edge_updates,
        m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]))
    f.write("## hot spots by source line (tools/ncu_lines.py: SASS rows of the source page joined with nvdisasm line info)\n\n```\n")
    f.write(lines)
    f.write("```\n")

# launch list
src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, tag + "_launches.csv"))
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    Hh = rows[hdr]
    ki, vi = Hh.index("Kernel Name"), Hh.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", "")) / 1e6) for r in rows[hdr + 1:] if len(r) > vi]
    idx = [i for i, (n, v) in enumerate(seq) if "k_sp_run" in n]
    # the last complete step = launches after the second-to-last k_walksat
    ws = [i for i, (n, v) in enumerate(seq) if "k_walksat" in n]
    start = ws[-2] + 1 if len(ws) >= 2 else 0
    step = seq[start:]
    # drop the trailing part after the final cnf_eval of the step
    agg = {}
    for n, v in step:
        a = agg.setdefault(n.split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, tag + "_launches_summary.md"), "w") as f:
        f.write("# launch list of one bench step (ncu --metrics gpu__time_duration.sum, `bench.py --steps 2 --warmup 1 --problems 2`)\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare shares, not absolutes.  %d launches, %.2f ms in total.\n\n" % (len(step), tot))
        f.write("| kernel | launches | ms | share |\n|---|---|---|---|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (n, a[0], a[1], 100 * a[1] / tot))
for name in [tag + "_bench.json", tag + "_bench_reference.json", tag + "_edge_nn_timing.log"] + [tag + "_bench_config%d.json" % i for i in (0, 1, 2, 4)]:
    if os.path.exists(os.path.join(G, name)) and os.path.getsize(os.path.join(G, name)) > 0:
        shutil.copy(os.path.join(G, name), os.path.join(P, name))

# ---- the tensor-core GRU kernel
rep2 = os.path.join(G, tag + "_edge_gru.ncu-rep")
if os.path.exists(rep2):
    raw2 = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r2 = list(csv.reader(raw2.split("\n")))
    m2 = {n: (r2[2][i], r2[1][i]) for i, n in enumerate(r2[0])}
    want2 = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
             "launch__shared_mem_per_block_dynamic", "sm__inst_issued.avg.per_cycle_active",
             "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
             "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
             "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
             "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
             "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
             "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
    rows_gru = 300000
    lines2 = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep2,
                             os.path.join(ROOT, "pdp_solver_b200", "csrc", "pdp_edge_nn.o"), "k_edge_nnILi1", "24"],
                            capture_output=True, text=True).stdout
    with open(os.path.join(P, tag + "_edge_gru_ncu.md"), "w") as f:
        f.write("# ncu --set full: k_edge_nn<GRU> (pdp_edge_gru_forward: tcgen05 kind::tf32, three-term split, TMEM accumulators)\n\n")
        f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:k_edge_nn -c 1 python tools/prof_edge_nn.py %d` "
                "(GRU cell 151 | 150 -> 150 over %d rows: 2 passes x 19 K-chunks x 3 tf32 terms of M128 x N(256+48) x K8 MMAs per tile of 128 rows = "
                "%.3g tensor flop in the launch).  Numbers under the profiler are not bench values; `%s_edge_nn_timing.log` has the CUDA-event times.\n\n"
                % (rows_gru, rows_gru, 2.0 * rows_gru * 304 * 608 * 3, tag))
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for n in want2:
            if n in m2:
                f.write("| %s | %s | %s |\n" % (n, m2[n][0], m2[n][1]))
        f.write("\n## hot spots by source line\n\n```\n" + lines2 + "```\n")

# ---- SASS evidence: the Blackwell-specific instructions of the two hot kernels
def mnemonics(obj, pats):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cnt = {}
    import re
    for ln in out.split("\n"):
        mm = re.search(r"\s(" + "|".join(pats) + r")[A-Z0-9_.]*", ln)
        if mm:
            k = mm.group(0).strip()
            cnt[k] = cnt.get(k, 0) + 1
    return cnt
with open(os.path.join(P, tag + "_blackwell_sass.txt"), "w") as f:
    f.write("cuobjdump -sass, instruction mnemonics that only exist on sm_100 (counts over the object file)\n\n")
    for obj, what in (("pdp_edge_nn.o", "tcgen05 edge layers (k_edge_nn): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, UBLKCP = cp.async.bulk"),
                      ("pdp_loop.o", "persistent SP loop (k_sp_run): UBLKCP = cp.async.bulk, SYNCS = mbarrier")):
        c = mnemonics(os.path.join(ROOT, "pdp_solver_b200", "csrc", obj), ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UBLKPF", "SYNCS", "UTMALDG"])
        f.write("%s -- %s\n" % (obj, what))
        for k in sorted(c):
            f.write("    %-40s %d\n" % (k, c[k]))
        f.write("\n")
print("profiles written; DRAM bytes per edge-update %.2f" % per)
