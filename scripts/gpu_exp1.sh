#!/bin/bash
# replay kernel variants + rollout sync variant
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_replay.py -q -x > $out/exp1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/exp1_pytest.log
for name in base pdl0 out3 in8; do
  lib=$PWD/variants/libcavgym_$name.so; [ "$name" = base ] && lib=$PWD/cavgym_b200/libcavgym_sm100.so
  echo "=== $name"
  CAVGYM_LIB=$lib timeout 300 python scripts/replay_fixed_cost.py --steps 1,5,20,80 2>&1 | grep -v starved | tail -6
done > $out/exp1_replay.log 2>&1
cat $out/exp1_replay.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-configs --skip-cpu --skip-hbm > $out/exp1_bench.json 2> $out/exp1_bench.err; echo "bench rc=$?"; tail -3 $out/exp1_bench.err
python -c "
import json; d=json.loads(open('$out/exp1_bench.json').read()); print('value %.2f G frac %.4f avg_launch_us %.1f isolated %.1f e2e %.3f G' % (d['value']/1e9, d['roofline']['frac'], d['roofline']['avg_launch_ms']*1e3, d['roofline']['isolated_launch_ms']*1e3, d['e2e']['value']/1e9))"
for name in base sync1; do
  lib=$PWD/variants/libcavgym_$name.so; [ "$name" = base ] && lib=$PWD/cavgym_b200/libcavgym_sm100.so
  for sc in pedestrians crossroads bus-stop pelican-crossing; do
    echo "=== $name $sc"; CAVGYM_LIB=$lib timeout 300 python scripts/profile_rollout.py --scenario $sc --envs 1048576 --launches 4 2>&1 | tail -1 | cut -c1-200
  done
done > $out/exp1_rollout.log 2>&1
cat $out/exp1_rollout.log
