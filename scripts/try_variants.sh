#!/bin/bash
# usage: scripts/try_variants.sh lib1.so lib2.so ...   (tuning builds under variants/)
for lib in "$@"; do
  echo "== $lib"
  CAVGYM_LIB=$PWD/$lib python scripts/quick_parity.py 2>&1 | tail -1
  CAVGYM_LIB=$PWD/$lib python scripts/profile_kernels.py --mode step --envs 4194304 --launches 6 --advance 300 2>&1 | tail -1
done
