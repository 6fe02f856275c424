#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/e2e_breakdown.py 200 > $out/exp3_e2e.log 2>&1; cat $out/exp3_e2e.log
for name in base dsync; do
  lib=$PWD/variants/libcavgym_$name.so; [ "$name" = base ] && lib=$PWD/cavgym_b200/libcavgym_sm100.so
  echo "=== $name"; CAVGYM_LIB=$lib timeout 300 python scripts/bench_dense.py --envs 100000 --steps 4 --chunk 50 --warm 100 2>&1 | tail -2 | cut -c1-400
done > $out/exp3_dense.log 2>&1; cat $out/exp3_dense.log
timeout 900 python -m pytest tests -m gpu -q -x > $out/exp3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/exp3_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/exp3_bench.json 2> $out/exp3_bench.err; echo "bench rc=$?"; tail -3 $out/exp3_bench.err
python -c "
import json; d=json.loads(open('$out/exp3_bench.json').read()); print('value %.2f G frac %.4f avg_launch_us %.1f isolated %.1f e2e %.3f G' % (d['value']/1e9, d['roofline']['frac'], d['roofline']['avg_launch_ms']*1e3, d['roofline']['isolated_launch_ms']*1e3, d['e2e']['value']/1e9)); print({k: round(v['env_steps_per_sec']/1e9,2) for k,v in d['scenario_configs']['scenarios'].items()})"
