#!/bin/bash
# usage: scripts/try_variants.sh lib1.so lib2.so ...   (tuning builds under variants/)
for lib in "$@"; do
  echo "== $lib"
  CAVGYM_LIB=$PWD/$lib python scripts/quick_parity.py 2>&1 | tail -1
  CAVGYM_LIB=$PWD/$lib python scripts/replay_scaling.py 2>&1 | grep -E "steps   (10|50):"
done
