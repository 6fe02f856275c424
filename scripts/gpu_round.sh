#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list of the timed region, full ncu captures.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag> "test bench list prof dense"'
# Everything lands in gpurun_out/<tag>_*.  ncu runs only after the same command exited 0 without it.
tag=${1:-r1}
parts=${2:-"test bench list prof dense"}
has() { [[ " $parts " == *" $1 "* ]]; }
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $out/${tag}_clocks.csv &
smi=$!
if has test; then
  timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
  tail -3 $out/${tag}_pytest.log
fi
if has bench; then
  timeout 900 python bench.py --steps 1000 --warmup 100 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
  tail -3 $out/${tag}_bench.err; cat $out/${tag}_bench.json
  timeout 600 python bench.py --impl reference --steps 50 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
  cat $out/${tag}_bench_ref.json
fi
kill $smi
if has list; then   # launch list of the bench command's TIMED REGION (cudaProfilerStart/Stop around it; same command line otherwise)
  timeout 300 python bench.py --steps 1000 --warmup 100 --skip-cpu --skip-hbm --skip-configs > $out/${tag}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 1000 --warmup 100 --skip-cpu --skip-hbm --skip-configs --profile-region > $out/${tag}_ncu_list.log 2>&1
  echo "launch list rc=$?"
fi
if has prof; then   # full captures of the per-step kernel at 4M envs and of the fused replay kernel
  timeout 300 python scripts/profile_kernels.py --mode step --envs 4194304 --launches 3 --advance 300 > $out/${tag}_prof_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:step_tma_kernel<double' -s 4 -c 1 -f -o $out/${tag}_step \
      python scripts/profile_kernels.py --mode step --envs 4194304 --launches 3 --advance 300 > $out/${tag}_ncu_step.log 2>&1
  echo "ncu step rc=$?"; tail -2 $out/${tag}_prof_plain.log
  timeout 300 python scripts/profile_kernels.py --mode replay --launches 2 > $out/${tag}_prof_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:replay_tma_kernel<double' -s 1 -c 1 -f -o $out/${tag}_replay \
      python scripts/profile_kernels.py --mode replay --launches 2 > $out/${tag}_ncu_replay.log 2>&1
  echo "ncu replay rc=$?"; tail -1 $out/${tag}_prof_plain2.log
fi
if has dense; then   # full capture of the warp-per-env dense kernel (config C4 shape, 20,000 envs, 10 fused steps)
  timeout 300 python scripts/bench_dense.py --envs 20000 --steps 3 --chunk 10 --warm 40 > $out/${tag}_dense_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dense_kernel<double' -s 6 -c 1 -f -o $out/${tag}_dense \
      python scripts/bench_dense.py --envs 20000 --steps 3 --chunk 10 --warm 40 > $out/${tag}_ncu_dense.log 2>&1
  echo "ncu dense rc=$?"; tail -1 $out/${tag}_dense_plain.log
fi
exit 0
