#!/bin/bash
# usage (GPU box): bash tools_ncu_backward.sh <tag>  -> launch list + full-set captures of the two backward kernels
# (two steps; per step both operand-format variants of each kernel are launched and the first step is the TF32 probe:
#  the 4th launch of each kernel is the working fp16 variant of step 2)
tag=$1; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_bwd_launches.csv python tools_step_backward.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^bwd_tc_kernel -s 3 -c 1 -o gpurun_out/${tag}_bwdA -f python tools_step_backward.py 2 > gpurun_out/${tag}_ncuA.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 3 -c 1 -o gpurun_out/${tag}_bwdB -f python tools_step_backward.py 2 > gpurun_out/${tag}_ncuB.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_bwd_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg={}
for r in rows[1:]:
    k=r[ki][:60]; v=float(r[vi].replace(",",""))
    if r[ui]=="ns": v/=1e3
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:8]: print(f"{t/n:10.1f} us/launch  x{n:3d}  {k}")
PY
for k in A B; do python tools_ncu_summary.py gpurun_out/${tag}_bwd$k.ncu-rep > gpurun_out/${tag}_bwd${k}_summary.txt 2>&1; done
ls -la gpurun_out | head -20
