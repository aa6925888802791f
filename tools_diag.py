"""Diagnosis runs for the vertical advection kernel (phase timing)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
from tools_tune import timeit
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, 80)
    sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
for st in sets:
    for f in st:
        f.const_target_tensor()
_lib.set_option("va.variant", 2)
for wps in (7, 14):
    for stages in (0, 6):
        for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
            _lib.set_option("va.ctas_per_sm", wps)
            _lib.set_option("va.stages", stages)
            _lib.set_option("va.debug", dbg)
            med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=20)
            print("wps=%d stages=%d debug=%d (skip: %s%s%s): median %.2f us min %.2f" % (
                wps, stages, dbg, "bwd " if dbg & 1 else "", "fwdmath " if dbg & 2 else "", "stores" if dbg & 4 else "",
                med * 1e3, mn * 1e3), flush=True)
