#!/usr/bin/env python
"""Turns ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: make_profile_summary.py <round tag> <launches.csv> <kernel.ncu-rep> [<kernel.ncu-rep> ...]"""
import csv, io, os, subprocess, sys

tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and share of the step
rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu].split("second")[0].strip() if False else r[iu], 1e-6 if r[iu] in ("ns", "nsecond") else 1.0)
    name = r[ik].split("(")[0].split("<")[0].strip()
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
allms = sum(t[1] for t in tot.values())
with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: ncu launch list of `python bench.py --steps 2 --warmup 1` (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
    f.write("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for name, (n, ms) in sorted(tot.items(), key=lambda x: -x[1][1]):
        f.write("| %s | %d | %.3f | %.1f %% |\n" % (name, n, ms, 100 * ms / allms))
    step = {k: v[1] for k, v in tot.items() if any(x in k for x in ("sweep_kernel", "sweep_group_kernel", "traceback_kernel", "gather_kernel", "hw_distance_kernel"))}
    if step:
        f.write("\nThe timed step of bench.py is sweep + traceback (+ gather); `identity_kernel` (the `rescoring` extra), `int_peak_kernel` "
                "(the roofline peak probe) and torch's memset run outside it.  Shares within the step: " +
                ", ".join("%s %.1f %%" % (k.split("::")[-1], 100 * v / sum(step.values())) for k, v in sorted(step.items(), key=lambda x: -x[1])) + ".\n")
print(open(os.path.join(out_dir, "%s_launches.md" % tag)).read())

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__inst_executed.max", "smsp__inst_executed.min",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.max.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, u, v = rr[0], rr[1], rr[2]
    name = v[h.index("Kernel Name")]
    base = os.path.splitext(os.path.basename(rep))[0]
    with open(os.path.join(out_dir, "%s_%s.md" % (tag, base)), "w") as f:
        f.write("# %s: ncu --set full --clock-control none, `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % (tag, name))
        for k in KEYS:
            if k in h:
                f.write("| %s | %s | %s |\n" % (k, v[h.index(k)], u[h.index(k)]))
        f.write("\nWarp stall reasons (average warps stalled per issue-active cycle):\n\n| reason | ratio |\n|---|---|\n")
        for i, k in enumerate(h):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                f.write("| %s | %.3f |\n" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v[i])))
        # hottest SASS lines
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        sr = list(csv.reader(io.StringIO(src)))
        sh = sr[1]
        isrc, isamp = sh.index("Source"), sh.index("# Samples")
        stall_cols = [i for i, x in enumerate(sh) if x.startswith("stall_") and "Not Issued" not in x]
        lines = [(int(r[isamp]), r[isrc].strip(), sorted(((int(r[i] or 0), sh[i][6:]) for i in stall_cols), reverse=True)[0]) for r in sr[2:] if len(r) > isamp and r[isamp].isdigit()]
        total = sum(x[0] for x in lines) or 1
        f.write("\nHottest SASS instructions (warp-stall samples; compiled with -lineinfo):\n\n| samples | share | top stall | instruction |\n|---|---|---|---|\n")
        for smp, txt, st in sorted(lines, key=lambda x: -x[0])[:25]:
            f.write("| %d | %.1f %% | %s | `%s` |\n" % (smp, 100.0 * smp / total, st[1], txt[:90]))
    print("wrote", base)
