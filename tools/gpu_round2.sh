#!/bin/bash
# One GPU visit of round 2: quick check first (a hang shows up in seconds, under a hard timeout), then the GPU suite, smoke,
# the bench line, and optionally (second argument "ncu") the launch list + one full capture of the step kernel.
tag=${1:-r02_x}
mkdir -p gpurun_out
timeout -s KILL 180 python tools/quick_check.py > gpurun_out/qc_$tag.log 2>&1; echo "quick_check rc=$?"; tail -12 gpurun_out/qc_$tag.log
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_$tag.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke_$tag.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
if [ "$2" = "ncu" ]; then
  python tools/prof_step.py tcgen05 0 4 > gpurun_out/plain_prof_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pinn_step_tc -s 2 -c 2 -f -o gpurun_out/prof_$tag \
      python tools/prof_step.py tcgen05 0 4 > gpurun_out/ncu_prof_$tag.log 2>&1
  python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/plain_bench_$tag.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/ncu_bench_$tag.log 2>&1
fi
echo round done
