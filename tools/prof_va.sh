#!/bin/bash
# ncu full capture of the vertical advection kernel (one GPU); output in gpurun_out/prof_va.ncu-rep
# usage: tools/prof_va.sh [variant] [unroll] [stages] [nk]
mkdir -p gpurun_out
cat > /tmp/prof_va.py <<PY
import sys
sys.path.insert(0, ".")
import torch, bench
from gridtools_b200 import _lib, stencil, storage
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
_lib.set_option("va.variant", ${1:-0}); _lib.set_option("va.unroll", ${2:-0}); _lib.set_option("va.stages", ${3:-0})
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, ${4:-80})
    sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
for s in range(6):
    stencil.vertical_advection_dycore(*sets[s % 2], 0.15)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:va_ -s 4 -c 1 -f -o gpurun_out/prof_va \
    python /tmp/prof_va.py > gpurun_out/ncu_va.log 2>&1
tail -3 gpurun_out/ncu_va.log
