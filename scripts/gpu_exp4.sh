#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python scripts/e2e_breakdown.py 200 > $out/exp4_e2e.log 2>&1; cat $out/exp4_e2e.log
for name in base dsync2; do
  lib=$PWD/variants/libcavgym_$name.so; [ "$name" = base ] && lib=$PWD/cavgym_b200/libcavgym_sm100.so
  echo "=== $name"; CAVGYM_LIB=$lib timeout 300 python scripts/bench_dense.py --envs 100000 --steps 4 --chunk 50 --warm 100 2>&1 | tail -1 | cut -c150-330
  CAVGYM_LIB=$lib timeout 300 python scripts/bench_dense.py --envs 100000 --steps 4 --chunk 50 --warm 100 --epsilon 0.01 2>&1 | tail -1 | cut -c150-330
done > $out/exp4_dense.log 2>&1; cat $out/exp4_dense.log
CAVGYM_LIB=$PWD/variants/libcavgym_dsync2.so timeout 900 python -m pytest tests/test_gpu_dense.py -q -x > $out/exp4_pytest.log 2>&1; echo "pytest dsync2 rc=$?"; tail -3 $out/exp4_pytest.log
