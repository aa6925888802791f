#!/bin/bash
# ncu full capture of a vertical advection variant in STEADY STATE (no cache flush between launches, 2 rotating
# field sets): usage tools/prof_va2.sh <variant> <tag> [extra va.option=value ...]
mkdir -p gpurun_out
V=${1:-5}; TAG=${2:-va$V}; shift; shift
cat > /tmp/prof_va2.py <<PY
import sys
sys.path.insert(0, ".")
import torch, bench
from gridtools_b200 import _lib, stencil, storage
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
_lib.set_option("va.variant", $V)
for kv in "$*".split():
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, 80)
    sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
for s in range(8):
    stencil.vertical_advection_dycore(*sets[s % 2], 0.15)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:va_ -s 5 -c 2 -f -o gpurun_out/prof_$TAG \
    python /tmp/prof_va2.py > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
