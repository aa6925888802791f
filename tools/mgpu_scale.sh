#!/bin/bash
# Scaling round at exactly N ranks (gpurun --gpus N): stamp check of the exchange, the weak-scaling bench lines of both
# stencils (as the driver runs them and with more steps), and BASELINE.json configs 4-5 (4096^2 strong / weak, chain).
# The first argument may be a list ("2 4" on a 4-GPU box).
MODE=${2:-full}
mkdir -p gpurun_out
for N in ${1:-8}; do
run() { # tag, script, args...
  local tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      "$@" > gpurun_out/$tag.json 2>> gpurun_out/scale_$N.err
  echo "$tag exit $?"
  python3 - "$tag" <<'PY'
import json, sys
for l in open('gpurun_out/%s.json' % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('   ', d['n_gpus'], d['metric'], round(d['value']), d['unit'], round(d['ms_per_step'] * 1e3, 2), 'us/step',
              'ranks', ' '.join('%.1f' % (x * 1e3) for x in d.get('rank_ms_per_step', [])))
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1
echo "mgpu_check exit $?"; grep "MGPU\|mismatch\|Error" gpurun_out/mgpu_check_$N.log | tail -3
run r02_scale_vert_adv_${N}          bench.py --gpus $N --steps 200 --warmup 20 --no-extras
run r02_scale_vert_adv_${N}_steps20  bench.py --gpus $N --steps 20 --warmup 3
run r02_scale_hori_diff_${N}         bench.py --gpus $N --steps 200 --warmup 20 --stencil hori_diff --no-extras
if [ "$MODE" = full ]; then
  run r02_hd4096_strong_${N}         bench.py --gpus $N --steps 20 --warmup 3 --stencil hori_diff --ni 4096 --nj 4096 --scaling strong --no-extras
  run r02_hd4096_weak_${N}           bench.py --gpus $N --steps 10 --warmup 3 --stencil hori_diff --ni 4096 --nj 4096 --scaling weak --no-extras
  run r02_chain_1024_${N}            bench_chain.py --steps 20 --warmup 3
fi
tail -3 gpurun_out/scale_$N.err
done
