#!/bin/bash
# Two weak-scaling bench lines (vert_adv, hori_diff; 200 steps) at N ranks: the short form of tools/mgpu_scale.sh.
N=${1:-2}
mkdir -p gpurun_out
for st in vert_adv hori_diff; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 200 --warmup 20 --stencil $st --no-extras > gpurun_out/r02_scale_${st}_$N.json 2>> gpurun_out/scale_$N.err
  python3 - gpurun_out/r02_scale_${st}_$N.json <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['n_gpus'], d['metric'], round(d['ms_per_step'] * 1e3, 2), 'us/step', 'ranks', ' '.join('%.1f' % (x * 1e3) for x in d.get('rank_ms_per_step', [])))
PY
done
tail -2 gpurun_out/scale_$N.err
