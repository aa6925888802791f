"""Does the relative placement of the five same-shaped fields in HBM matter?  (storage.SKEW_BYTES)"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
from tools_tune import timeit
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
for skew in (0, 4352, 16384 + 1152, 65536 + 4352, 262144 + 17408, 1048576 + 69632, 128, 2176):
    storage.SKEW_BYTES = skew
    storage._n_allocated = 0
    for name in ("va", "hd"):
        sets = []
        for _ in range(2 if name == "va" else 3):
            if name == "va":
                arrs, dtr = bench.repo_vert_adv(256, 256, 80)
                sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
            else:
                inp, coeff = bench.repo_hori_diff(256, 256, 80)
                sets.append([storage.from_numpy(inp, (2, 2, 0)), storage.from_numpy(coeff, (2, 2, 0)),
                             storage.from_numpy(np.zeros_like(inp), (2, 2, 0))])
        for st in sets:
            for f in st:
                f.const_target_tensor()
        for dbg in ((0, 7) if name == "va" else (0,)):
            _lib.set_option("va.debug", dbg)
            if name == "va":
                med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=30)
            else:
                med, mn = timeit(lambda st: stencil.horizontal_diffusion(*st), sets, n=30)
            print("skew %8d %s debug=%d: median %.2f us min %.2f  (addresses mod 2MiB: %s)" % (
                skew, name, dbg, med * 1e3, mn * 1e3, [hex(f.raw_ptr(True) % (1 << 21)) for f in sets[0]]), flush=True)
        del sets
        torch.cuda.empty_cache()
