#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python scripts/profile_rollout.py --scenario bus-stop --envs 131072 --launches 3 > $out/exp2_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rollout_kernel<double' -s 2 -c 1 -f -o $out/exp2_busstop \
   python scripts/profile_rollout.py --scenario bus-stop --envs 131072 --launches 3 > $out/exp2_ncu.log 2>&1
echo "ncu rc=$?"; tail -1 $out/exp2_plain.log | cut -c1-150
timeout 300 python scripts/profile_rollout.py --scenario crossroads --envs 131072 --launches 3 > $out/exp2_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:rollout_kernel<double' -s 2 -c 1 -f -o $out/exp2_crossroads \
   python scripts/profile_rollout.py --scenario crossroads --envs 131072 --launches 3 > $out/exp2_ncu2.log 2>&1
echo "ncu rc=$?"; tail -1 $out/exp2_plain2.log | cut -c1-150
