#!/bin/bash
# 8-GPU A/B of the default bench (vert_adv weak scaling): what the last 15 % go to.
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/mgpu_ab8.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29516 bench.py --gpus $N --steps 300 --warmup 20 --stencil ${ST:-vert_adv} --no-extras 2>> gpurun_out/mgpu_ab8.err | grep "^{" | \
      python3 -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N ${ST:-vert_adv} $name', round(d['ms_per_step']*1e3,2), 'us/step')" | tee -a gpurun_out/mgpu_ab8.txt
}
run baseline A=1
run no_sampler GTB_NO_SAMPLER=1
run nsets3 GTB_NSETS=3 GTB_NO_SAMPLER=1
run reserve8 GTB_RESERVE_SMS=8 GTB_NO_SAMPLER=1
run reserve2 GTB_RESERVE_SMS=2 GTB_NO_SAMPLER=1
ST=hori_diff run hd_no_sampler GTB_NO_SAMPLER=1
ST=hori_diff run hd_nsets4 GTB_NSETS=4 GTB_NO_SAMPLER=1
tail -3 gpurun_out/mgpu_ab8.err
