#!/bin/bash
# Compare library builds under variants/ on the replay workload: bash scripts/variant_sweep.sh tag name1 name2 ...
tag=$1; shift
for name in "$@"; do
  lib=variants/libcavgym_$name.so; [ "$name" = base ] && lib=cavgym_b200/libcavgym_sm100.so
  echo "=== $name"
  CAVGYM_LIB=$lib python scripts/replay_fixed_cost.py --start 300 --steps 20,80 --chain 4 2>&1 | grep -v starved | tail -6
  CAVGYM_LIB=$lib python bench.py --steps 20 --warmup 5 --skip-configs --skip-cpu --skip-hbm 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('bench20 value %.2f G  frac %.4f  avg_launch_us %.1f  region median %.1f' % (d['value']/1e9, d['roofline']['frac'], d['roofline']['avg_launch_ms']*1e3, d['timed_region']['region_ms_median']*1e3))"
done > gpurun_out/${tag}_sweep.log 2>&1
cat gpurun_out/${tag}_sweep.log
