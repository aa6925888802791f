#!/bin/bash
# Multi-GPU round (gpurun --gpus N): halo check over both transports, weak-scaling bench at 1..N, hori_diff too.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1
echo "mgpu_check exit $?"; tail -4 gpurun_out/mgpu_check_$N.log
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 20 --no-extras > gpurun_out/scale_va_1.json 2>> gpurun_out/scale.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    for st in vert_adv hori_diff; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
          bench.py --gpus $n --steps 200 --warmup 20 --stencil $st > gpurun_out/scale_${st}_$n.json 2>> gpurun_out/scale.err
      echo "bench $st $n exit $?"
    done
  fi
done
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 20 --no-extras --stencil hori_diff > gpurun_out/scale_hori_diff_1.json 2>> gpurun_out/scale.err
for f in gpurun_out/scale_*.json; do echo $f; python3 -c "
import json,sys
for l in open('$f'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['metric'], round(d['value']), 'Mpts/s', round(d['ms_per_step']*1e3,2), 'us/step', 'e2e', d.get('e2e',{}).get('value'))
"; done
tail -5 gpurun_out/scale.err
