"""One round of strips: 14 warps per SM with shallow rings (variant 3, stages=2)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
from tune import timeit
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, 80)
    sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
for st in sets:
    for f in st:
        f.const_target_tensor()
for cfg in (dict(variant=3, threads=128, stages=4, ctas_per_sm=7), dict(variant=3, threads=32, stages=2, ctas_per_sm=14, save_upos=2),
            dict(variant=3, threads=64, stages=2, ctas_per_sm=14, save_upos=2), dict(variant=3, threads=32, stages=2, ctas_per_sm=14, save_upos=1),
            dict(variant=3, threads=64, stages=2, ctas_per_sm=14, save_upos=1), dict(variant=3, threads=64, stages=2, ctas_per_sm=12, save_upos=2),
            dict(variant=3, threads=64, stages=2, ctas_per_sm=10, save_upos=2), dict(variant=3, threads=64, stages=2, ctas_per_sm=7, save_upos=2)):
    for dbg in (0, 1, 8):
        for k in ("variant", "threads", "unroll", "stages", "ctas_per_sm", "save_upos"):
            _lib.set_option("va." + k, cfg.get(k, 0))
        _lib.set_option("va.debug", dbg)
        med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=20)
        print("%s debug=%d: median %.2f us min %.2f" % (cfg, dbg, med * 1e3, mn * 1e3), flush=True)
