#!/bin/bash
# usage (GPU box): bash tools_ab.sh variant...   -- A/B timing of experiment builds (build.py --variant); "base" = product library
for v in "$@"; do
  if [ "$v" = base ]; then unset OI_LIB_PATH; else export OI_LIB_PATH=$PWD/object_intrinsics_b200/lib/variants/$v/liboi_b200.so; fi
  python bench.py --no-cpu-baseline --no-side-legs --steps 10 --warmup 3 2>/dev/null > /tmp/ab_$v.txt
  grep tcprof /tmp/ab_$v.txt | tail -2
  tail -1 /tmp/ab_$v.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'core_ms %.4f'%d['roofline']['core_kernel_ms'], 'step_ms %.4f'%d['ms_per_step'], 'value %.3fM'%(d['value']/1e6))"
done
