#!/bin/bash
# One-box scaling table: N = 1, 2, 4, 8 back to back on ONE 8-GPU box (gpurun --gpus 8 -- 'bash tools/scale_one_box.sh tag').
tag=${1:-r02_s}
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py --steps 20 --warmup 3 --scaling-legs-only > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; echo "n1 rc=$?"
for n in 2 4 8; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n${n}_$tag.json 2> gpurun_out/bench_n${n}_$tag.err; echo "n$n rc=$?"
done
python - <<'PY'
import json, glob, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
PY
for n in 1 2 4 8; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n${n}_$tag.json').read().strip().splitlines()[-1])
print($n, '%.4g'%d['value'], '%.4g'%d['value_steady']['value'], '%.4g'%d['e2e']['value'], '%.4g'%d['config4_global_batch']['value'], '%.4g'%d['device_train_loop']['value'], d['ms_per_step'], d.get('dp_parity',{}).get('ok'), d['e2e']['host_link_probe']['h2d_GBps_per_gpu_min'])
"; done
