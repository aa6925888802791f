#!/bin/bash
# Exchange/overlap variants of the weak-scaling bench on N GPUs (gpurun --gpus N): one short bench run per variant.
# GTB_PERIODIC=1 gives every rank all 8 IJ neighbours, the exchange load of an interior rank of a large process grid.
N=${1:-2}
STEPS=${2:-100}
OUT=gpurun_out/variants_$N.txt
mkdir -p gpurun_out
: > $OUT
run() { # name stencil env...
  local name=$1 st=$2; shift 2
  local line
  line=$(env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29531 bench.py --gpus $N --steps $STEPS --warmup 10 --stencil $st --no-extras 2>> gpurun_out/variants.err | tail -1)
  python3 - "$name" "$st" "$line" >> $OUT <<'PY'
import json, sys
name, st, line = sys.argv[1:4]
try:
    d = json.loads(line)
    print("%-34s %-9s %7.2f us/step  ranks %s" % (name, st, d["ms_per_step"] * 1e3,
          " ".join("%.1f" % (x * 1e3) for x in d["rank_ms_per_step"])))
except Exception as e:
    print("%-34s %-9s FAILED %s %r" % (name, st, e, line[:200]))
PY
  tail -1 $OUT
}
n1() {
  timeout 300 python bench.py --gpus 1 --steps $STEPS --warmup 10 --stencil $1 --no-extras 2>> gpurun_out/variants.err | tail -1 | \
    python3 -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-34s %-9s %7.2f us/step' % ('N=1', '$1', d['ms_per_step']*1e3))" | tee -a $OUT
}
# round 2, last sweep: pattern objects taking the steps in turn (GTB_PIPE) against one object
n1 hori_diff
run "plain r8 pipe 1"                 hori_diff GTB_RESERVE_SMS=8 GTB_PIPE=1
run "plain r8 pipe 2"                 hori_diff GTB_RESERVE_SMS=8 GTB_PIPE=2
run "periodic r8 pipe 1"              hori_diff GTB_PERIODIC=1 GTB_RESERVE_SMS=8 GTB_PIPE=1
run "periodic r8 pipe 2"              hori_diff GTB_PERIODIC=1 GTB_RESERVE_SMS=8 GTB_PIPE=2
run "periodic r8 pipe 3"              hori_diff GTB_PERIODIC=1 GTB_RESERVE_SMS=8 GTB_PIPE=3
run "periodic r12 pipe 2"             hori_diff GTB_PERIODIC=1 GTB_RESERVE_SMS=12 GTB_PIPE=2
run "periodic r6 pipe 2"              vert_adv GTB_PERIODIC=1 GTB_RESERVE_SMS=6 GTB_PIPE=2
cat $OUT
