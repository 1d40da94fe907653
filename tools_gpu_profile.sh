#!/bin/bash
# usage (on the GPU box): bash tools_gpu_profile.sh <tag> [bench args...]
# writes gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_prof.ncu-rep, gpurun_out/<tag>_bench.json
tag=$1; shift
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -1 gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 1 -o gpurun_out/${tag}_prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/
