"""Copies the evidence of `bash tools_round2_profiles.sh` (gpurun_out/r02_*) into profiles/ and re-stamps
profiles/traffic.json to the library the capture was taken on.  usage: python tools_install_profiles.py"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for src, dst in [("r02_fwd_ncu_summary.txt", "r02_fwd_core_kernel_ncu.txt"), ("r02_bwdA_summary.txt", "r02_bwd_sweep_kernel_ncu.txt"),
                 ("r02_bwdB_summary.txt", "r02_bwd_contraction_kernel_ncu.txt"), ("r02_fwd_launches.csv", "r02_fwd_launches.csv"),
                 ("r02_bwd_launches.csv", "r02_bwd_launches.csv"), ("r02_sanitizer.txt", "r02_sanitizer.txt")]:
    shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
line = open(os.path.join(G, "r02_bench.json")).read().strip().splitlines()[-1]
bench = json.loads(line)
open(os.path.join(P, "r02_bench_line.json"), "w").write(line + "\n")
json.dump(bench["train_step"], open(os.path.join(P, "r02_train_step.json"), "w"), indent=1)

# launch list of the training-shaped step -> share per kernel
rows = [r for r in csv.reader(l for l in open(os.path.join(G, "r02_train_launches.csv")) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki])[:90]
    agg[name][0] += us
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
with open(os.path.join(P, "r02_train_step_launches.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none python tools_train_step.py --shape cfg5 --steps 1\n"
            "(3 warm-up + 1 synchronised + 1 free-running step = 5 training-shaped steps, bs=4 x 64x64 x 64; cold-cache, "
            "serialised: compare shares)\n")
    f.write(f"total {tot / 1e3:.2f} ms in {sum(v[1] for v in agg.values())} launches\n\n share   us/launch  launches  kernel\n")
    for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:50]:
        f.write(f"{100 * us / tot:6.2f}  {us / n:10.1f}  {n:8d}  {name}\n")

# traffic stamp: DRAM bytes of the forward core from the full-set capture, stamped with the library of this tree
txt = open(os.path.join(G, "r02_fwd_ncu_summary.txt")).read()


def gb(metric):
    m = re.search(metric + r" = ([0-9.]+) (\w+)", txt)
    return float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m.group(2)]


tj = json.load(open(os.path.join(P, "traffic.json")))
stamp = open(os.path.join(ROOT, "object_intrinsics_b200", "lib", "liboi_b200.stamp")).read().strip()
if bench.get("lib_stamp") != stamp:
    sys.exit(f"bench line was produced by another build ({bench.get('lib_stamp')} vs {stamp}): not stamping")
tj["render_tc_kernel"]["dram_bytes_per_launch"] = int(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
tj["render_tc_kernel"]["stamp"] = stamp
json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=1)
subprocess.run([sys.executable, os.path.join(ROOT, "tools_sass_summary.py")], stdout=open(os.path.join(P, "r02_sass_summary.txt"), "w"),
               check=True)
print("installed; forward core DRAM bytes per launch:", tj["render_tc_kernel"]["dram_bytes_per_launch"])
