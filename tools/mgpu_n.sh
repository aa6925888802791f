#!/bin/bash
# Multi-GPU spot check at exactly N ranks (gpurun --gpus N): halo stamp check + one bench line per stencil.
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1
echo "mgpu_check exit $?"; grep "MGPU\|mismatch\|Error" gpurun_out/mgpu_check_$N.log | tail -4
for st in vert_adv hori_diff; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 200 --warmup 20 --stencil $st --no-extras > gpurun_out/scale_${st}_$N.json 2>> gpurun_out/scale.err
  echo "bench $st $N exit $?"
  python3 - <<PY
import json
for l in open('gpurun_out/scale_${st}_$N.json'):
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], d['metric'], round(d['value']), 'Mpts/s', round(d['ms_per_step']*1e3, 2), 'us/step', d['config']['decomposition'][:40])
PY
done
tail -3 gpurun_out/scale.err
