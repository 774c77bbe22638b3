#!/usr/bin/env python
"""Turn gpurun_out/{launches,prof_z,prof_e}_<tag> into text summaries under profiles/ (run where ncu is installed)."""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum"]

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return {h: (data[0][i], units[i]) for i, h in enumerate(hdr)}

# 1. launch list: share of each kernel in the step
lf = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf)) if len(r) > 10]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        name = r[ki].split("(")[0]
        tot[name][0] += 1; tot[name][1] += v
    s = sum(v[1] for v in tot.values())
    with open(os.path.join(P, "%s_launch_list.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --no-extras\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("%-40s %8s %14s %12s %8s\n" % ("kernel", "launches", "total_ns", "mean_us", "share"))
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s %8d %14.0f %12.2f %7.1f%%\n" % (k, n, t, t / n / 1e3, 100 * t / s))
    print(open(os.path.join(P, "%s_launch_list.txt" % tag)).read())
# 2. full captures
for short, kname in (("z", "z_alloc_fused"), ("e", "e_family"), ("p", "p_family")):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (short, tag))
    if not os.path.exists(rep): continue
    m = raw(rep)
    with open(os.path.join(P, "%s_%s_ncu.txt" % (tag, kname)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:%s -s 20 -c 1 (one launch, config 3)\n" % kname)
        for k in KEYS:
            if k in m: f.write("%-90s %16s %s\n" % (k, m[k][0], m[k][1]))
        sass = os.path.join(G, "prof_%s_%s_sass.csv" % (short, tag))
        open(sass, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
        subprocess.run("cd /tmp && rm -f sg_*.cubin; cuobjdump -xelf all %s/signer_b200/libsigner_gibbs.so > /dev/null; nvdisasm -g -c signer_gibbs*.cubin > /tmp/sass_lines.txt" % ROOT, shell=True)
        br = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_by_line.py"), sass, "/tmp/sass_lines.txt", kname, "30"], capture_output=True, text=True).stdout
        f.write("\n# per source line (SASS joined with nvdisasm line info): share of warp instructions, active threads per instruction, share of stall samples\n")
        f.write(br)
    if short == "z":
        try:
            rd = float(m["dram__bytes_read.sum"][0].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[m["dram__bytes_read.sum"][1]]
            wr = float(m["dram__bytes_write.sum"][0].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[m["dram__bytes_write.sum"][1]]
            json.dump(dict(kernel=kname, dram_bytes_per_launch=rd + wr, source="%s_%s_ncu.txt" % (tag, kname)), open(os.path.join(P, "z_alloc_fused_traffic.json"), "w"))
        except Exception as e:
            print("traffic:", e)
    print(open(os.path.join(P, "%s_%s_ncu.txt" % (tag, kname))).read()[:3000])
