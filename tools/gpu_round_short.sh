#!/bin/bash
# gpu_round.sh without the CPU-only reference arm (for a visit with few GPU-minutes left); the essentials come first.
tag=${1:-r01_x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke_$tag.log
python bench.py --cpu-seconds 8 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
python tools/quick_check.py > gpurun_out/qc_$tag.log 2>&1
python tools/prof_step.py tcgen05 0 4 > gpurun_out/plain_prof_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pinn_step_tc -s 2 -c 2 -f -o gpurun_out/prof_$tag \
    python tools/prof_step.py tcgen05 0 4 > gpurun_out/ncu_prof_$tag.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
echo round done
