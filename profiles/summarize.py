"""Turn the ncu artefacts in gpurun_out/ into the text summaries committed under profiles/<round>/.
usage: python profiles/summarize.py <tag> <outdir>"""
import csv, os, subprocess, sys
from collections import defaultdict

tag, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
G = "gpurun_out"

# ---- launch list
rows = [r for r in csv.reader(l for l in open(f"{G}/launches_{tag}.csv") if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    tot[r[ki]] += v
    cnt[r[ki]] += 1
T = sum(tot.values())
with open(f"{out}/launches_{tag}.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    f.write("# command: python bench.py --steps 2 --warmup 1 --no-cpu --sweeps 50 --em-n 200000 --em-maxit 3 --em-steps 1\n")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        f.write(f"{v:12.1f} us {100 * v / T:5.1f}%  x{cnt[k]:3d}  {k[:110]}\n")

KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "launch__shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg ",
        "smsp__average_warps_issue_stalled", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__sass_inst_executed_op_shared"]
for kern in ("rj", "em"):
    rep = f"{G}/prof_{kern}_{tag}.ncu-rep"
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    with open(f"{out}/ncu_{kern}_{tag}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; kernel: {v[h.index('Kernel Name')]}\n")
        for a, b, c in zip(h, u, v):
            if any(a.startswith(k.strip()) for k in KEEP):
                f.write(f"{a:95s} {c} {b}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    s = list(csv.reader(src.splitlines()))
    sh, data = s[1], s[2:]
    si, ii = sh.index("# Samples"), sh.index("Instructions Executed")
    ts = sum(int(x[si]) for x in data) or 1
    ti = sum(int(x[ii]) for x in data) or 1
    mix, smp = defaultdict(int), defaultdict(int)
    for x in data:
        op = [o for o in x[1].split() if not o.startswith("@")]
        name = op[0].split(".")[0] if op else "?"
        mix[name] += int(x[ii])
        smp[name] += int(x[si])
    with open(f"{out}/ncu_{kern}_{tag}.txt", "a") as f:
        f.write(f"\n# SASS instruction mix ({ti} warp-instructions, {ts} stall samples)\n")
        for k, val in sorted(mix.items(), key=lambda z: -z[1])[:18]:
            f.write(f"{k:12s} {100 * val / ti:5.1f}% of instructions  {100 * smp[k] / ts:5.1f}% of samples\n")
        f.write("\n# hottest SASS lines by stall samples\n")
        for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:12]):
            x = data[i]
            st = {q: int(x[sh.index(q)]) for q in sh if q.startswith("stall_") and "Not Issued" not in q}
            f.write(f"{100 * int(x[si]) / ts:5.1f}%  exec={x[ii]:>11s}  {max(st, key=st.get):20s} {x[1].strip()[:80]}\n")
print("wrote", os.listdir(out))
