"""Variant comparison at a column height that fits the register tier (nk = 48): no shared-memory slab, deep rings."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
from tune import timeit
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
for nk in (48, 80):
    sets = []
    for _ in range(3):
        arrs, dtr = bench.repo_vert_adv(256, 256, nk)
        sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
    for st in sets:
        for f in st:
            f.const_target_tensor()
    for cfg in (dict(variant=3, threads=128, stages=4), dict(variant=4, unroll=4, stages=4), dict(variant=4, unroll=4, stages=3),
                dict(variant=4, unroll=4, stages=2), dict(variant=5), dict(variant=5, stages=3), dict(variant=5, stages=5), dict(variant=5, ctas_per_sm=6), dict(variant=5, stages=3, ctas_per_sm=8)):
        for k in ("variant", "threads", "unroll", "stages", "ctas_per_sm"):
            _lib.set_option("va." + k, cfg.get(k, 0))
        try:
            med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=20)
        except Exception as e:
            print(nk, cfg, "FAILED", e)
            continue
        print("nk=%d %s: median %.2f us min %.2f -> %.0f GB/s" % (nk, cfg, med * 1e3, mn * 1e3, 48 * 256 * 256 * nk / med / 1e6), flush=True)
