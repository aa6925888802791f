#!/bin/bash
# Device timeline of stencil launches and exchange kernels (GTB_TIMELINE=1) for a few settings, N GPUs.
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/timeline_$N.txt
: > $OUT
run() {
  local name=$1 st=$2; shift 2
  echo "=== $name $st" >> $OUT
  env "$@" GTB_TIMELINE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29541 bench.py --gpus $N --steps 60 --warmup 10 --stencil $st --no-extras 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | \
      python3 -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step %.2f us' % (d['ms_per_step'] * 1e3))
    else:
        print(l.rstrip())" >> $OUT
}
for st in vert_adv hori_diff; do
  run "periodic dma r4" $st GTB_PERIODIC=1
  run "periodic nodma r4" $st GTB_PERIODIC=1 GTB_HALO_DMA=0
done
cat $OUT
