#!/bin/bash
# A/B of tuning builds (CAVGYM_LIB) on the heterogeneous rollout: bash scripts/ab_libs.sh "lib1.so lib2.so" [ab_team.py args]
libs=$1; shift
for lib in $libs; do echo "== $lib"; CAVGYM_LIB=$PWD/$lib timeout 300 python scripts/ab_team.py --paths 1 "$@" 2>&1 | tail -3 | cut -c1-150; done
