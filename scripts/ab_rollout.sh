#!/bin/bash
# A/B of tuning builds on the heterogeneous rollout kernels: bash scripts/ab_rollout.sh "lib1.so lib2.so" "bus-stop crossroads" [envs]
libs=${1}; scen=${2:-"bus-stop crossroads pelican-crossing"}; envs=${3:-1048576}
for lib in $libs; do for s in $scen; do
  echo -n "$lib "; CAVGYM_LIB=$PWD/$lib timeout 300 python scripts/profile_rollout.py --scenario $s --envs $envs --launches 5 2>&1 | tail -1 | cut -c1-150
done; done
