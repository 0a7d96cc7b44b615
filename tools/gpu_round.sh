#!/bin/bash
# One GPU-box visit: parity tests, the bench (both arms), the ncu launch list of the bench command and one full
# capture of the step kernel.  Usage (CPU box): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01_d'
tag=${1:-r01_x}
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi_$tag.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$tag.log
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/prof_step.py tcgen05 0 4 > gpurun_out/plain_prof_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pinn_step_tc -s 2 -c 2 -f -o gpurun_out/prof_$tag \
    python tools/prof_step.py tcgen05 0 4 > gpurun_out/ncu_prof_$tag.log 2>&1
cat gpurun_out/plain_prof_$tag.log
