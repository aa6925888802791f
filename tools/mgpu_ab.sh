#!/bin/bash
# A/B of the exchange options at N ranks: GTB_HALO_FUSED x GTB_RESERVE_SMS, both stencils, --no-extras.
N=${1:-4}
mkdir -p gpurun_out
: > gpurun_out/mgpu_ab_$N.txt
for fused in 0 1; do for rs in 4 8; do for st in vert_adv hori_diff; do
  GTB_HALO_FUSED=$fused GTB_RESERVE_SMS=$rs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29514 bench.py --gpus $N --steps 300 --warmup 20 --stencil $st --no-extras 2>> gpurun_out/mgpu_ab.err | grep "^{" | \
      python3 -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N fused=$fused reserve=$rs', d['metric'], round(d['ms_per_step']*1e3,2), 'us/step')" | tee -a gpurun_out/mgpu_ab_$N.txt
done; done; done
tail -3 gpurun_out/mgpu_ab.err
