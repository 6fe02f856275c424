#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list and one full capture of the per-step kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
# Everything lands in gpurun_out/<tag>_*.  ncu runs only after the same command exited 0 without it.
tag=${1:-r1}
parts=${2:-"test bench list prof"}
has() { [[ " $parts " == *" $1 "* ]]; }
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $out/${tag}_clocks.csv &
smi=$!
has test && timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
has test && tail -5 $out/${tag}_pytest.log
has bench && timeout 600 python bench.py --steps 1000 --warmup 100 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
has bench && { tail -3 $out/${tag}_bench.err; cat $out/${tag}_bench.json; }
has bench && timeout 600 python bench.py --impl reference --steps 50 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"
has bench && cat $out/${tag}_bench_ref.json
kill $smi
# launch list of the bench command's TIMED REGION (cudaProfilerStart/Stop around it; same command line otherwise)
has list && timeout 300 python bench.py --steps 1000 --warmup 100 --skip-cpu --skip-hbm --skip-configs > $out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1000 --warmup 100 --skip-cpu --skip-hbm --skip-configs --profile-region > $out/${tag}_ncu_list.log 2>&1
echo "launch list rc=$?"
# full capture of the per-step kernel at 4M envs and of the fused replay kernel
has prof && timeout 300 python scripts/profile_kernels.py --mode step --envs 4194304 --launches 3 --advance 300 > $out/${tag}_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:step_tma_kernel<double' -s 4 -c 1 -f -o $out/${tag}_step \
    python scripts/profile_kernels.py --mode step --envs 4194304 --launches 3 --advance 300 > $out/${tag}_ncu_step.log 2>&1
echo "ncu step rc=$?"; cat $out/${tag}_prof_plain.log | tail -2
has prof && timeout 300 python scripts/profile_kernels.py --mode replay --launches 2 > $out/${tag}_prof_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:replay_tma_kernel<double' -s 1 -c 1 -f -o $out/${tag}_replay \
    python scripts/profile_kernels.py --mode replay --launches 2 > $out/${tag}_ncu_replay.log 2>&1
echo "ncu replay rc=$?"; tail -1 $out/${tag}_prof_plain2.log

# full capture of the warp-per-env dense kernel (config C4 shape, 20,000 envs, 10 fused steps)
has dense && timeout 300 python scripts/bench_dense.py --envs 20000 --steps 3 --chunk 10 --warm 40 > $out/${tag}_dense_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dense_kernel<double' -s 6 -c 1 -f -o $out/${tag}_dense \
    python scripts/bench_dense.py --envs 20000 --steps 3 --chunk 10 --warm 40 > $out/${tag}_ncu_dense.log 2>&1
echo "ncu dense rc=$?"; tail -1 $out/${tag}_dense_plain.log
