#!/bin/bash
# Round-2 evidence run (GPU box, one GPU): bash tools_round2_profiles.sh
# Writes gpurun_out/r02_*: bench line, launch lists (forward bench, training-shaped step), full-set ncu captures of the
# forward core and the two backward kernels + their summaries, compute-sanitizer memcheck of the hot path.
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 600 gpurun_out/r02_bench.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_fwd_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_tc -s 3 -c 1 -o gpurun_out/r02_fwd -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-legs > gpurun_out/r02_fwd_ncu.log 2>&1
python tools_ncu_summary.py gpurun_out/r02_fwd.ncu-rep > gpurun_out/r02_fwd_ncu_summary.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv \
    python tools_train_step.py --shape cfg5 --steps 1 > /dev/null 2>&1
bash tools_ncu_backward.sh r02 > gpurun_out/r02_bwd_top.txt 2>&1
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py tests/test_generator_gpu.py \
    -q -m gpu -k "cfg1_n16_m4 or gen_rays or render_maps" > gpurun_out/r02_sanitizer.txt 2>&1
tail -3 gpurun_out/r02_sanitizer.txt
ls -la gpurun_out | tail -20
