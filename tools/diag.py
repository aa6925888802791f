"""Diagnosis runs for the vertical advection kernel (phase timing with the va.debug knob)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
from tune import timeit
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, 80)
    sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
for st in sets:
    for f in st:
        f.const_target_tensor()
_lib.set_option("va.variant", variant)
for wps in (7, 14):
  for dbg in (0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 14):
    _lib.set_option("va.debug", dbg)
    _lib.set_option("va.ctas_per_sm", wps)
    if wps == 14:
        _lib.set_option("va.threads", 64)
        _lib.set_option("va.stages", 3)
    med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=20)
    print("variant %d debug=%2d (1 skip backward, 2 skip forward math, 4 skip slab stores, 8 skip out stores) wps=%d: median %.2f us min %.2f" % (
        variant, dbg, wps, med * 1e3, mn * 1e3), flush=True)
