#!/bin/bash
# Round-2 GPU visit: parity tests, parity report, driver-shaped bench (both arms), sanitizer, ncu of the 20-step replay launch.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round2.sh <tag> "test report bench san ncu"'
tag=${1:-r2}
parts=${2:-"test report bench"}
has() { [[ " $parts " == *" $1 "* ]]; }
out=gpurun_out
mkdir -p $out
if has test; then
  timeout 1500 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
  tail -15 $out/${tag}_pytest.log
fi
if has report; then
  timeout 600 python scripts/parity_report.py > $out/${tag}_parity_report.txt 2> $out/${tag}_parity_report.err; echo "report rc=$?"
  tail -4 $out/${tag}_parity_report.txt; tail -3 $out/${tag}_parity_report.err
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
  tail -3 $out/${tag}_bench.err; cat $out/${tag}_bench.json
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
  cat $out/${tag}_bench_ref.json
fi
if has san; then
  timeout 300 python scripts/sanitize_small.py > $out/${tag}_san_plain.log 2>&1; echo "plain rc=$?"
  for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > $out/${tag}_sanitizer_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $out/${tag}_sanitizer_$tool.log | tail -3
  done
fi
if has ncu; then
  timeout 300 python scripts/profile_kernels.py --mode replay --launches 2 --replay-steps 20 > $out/${tag}_prof_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:replay_tma_kernel<double' -s 1 -c 1 -f -o $out/${tag}_replay20 \
      python scripts/profile_kernels.py --mode replay --launches 2 --replay-steps 20 > $out/${tag}_ncu_replay.log 2>&1
  echo "ncu replay rc=$?"; tail -1 $out/${tag}_prof_plain.log
fi
exit 0
