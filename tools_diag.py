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
_lib.set_option("va.variant", 4)
for kc in (4, 8):
    for wps in (1, 4, 6, 7, 8, 10):
        _lib.set_option("va.unroll", kc)
        _lib.set_option("va.ctas_per_sm", wps)
        med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=10)
        strips = 2048 / (148 * wps)
        print("variant 4 kc=%d wps=%d: median %.2f us min %.2f -> %.2f us per strip" % (kc, wps, med * 1e3, mn * 1e3,
                                                                                   med * 1e3 / strips), flush=True)
