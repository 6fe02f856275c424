#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 150 python -m pytest tests/test_gpu_dense.py -q -x -k "ragged or rollout_with_device_agents" > $out/exp5_pytest_a.log 2>&1; rc=$?; echo "pytest a rc=$rc"; tail -3 $out/exp5_pytest_a.log
[ $rc -ne 0 ] && exit 0
for sc in crossroads bus-stop pelican-crossing; do
  timeout 120 python scripts/profile_rollout.py --scenario $sc --envs 1048576 --launches 4 2>&1 | tail -1 | cut -c1-160
done > $out/exp5_rollout.log 2>&1; cat $out/exp5_rollout.log
timeout 600 python -m pytest tests -m gpu -q -x > $out/exp5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/exp5_pytest.log
