import sys
import numpy as np, torch
sys.path.insert(0, ".")
from gridtools_b200 import _lib, stencil, storage
torch.cuda.set_device(0); _lib.check(_lib.lib().gtb_init(0))
rng = np.random.default_rng(0)
for cfg in (dict(variant=7, ctas_per_sm=-2), dict(variant=7, ctas_per_sm=3), dict(variant=7)):
    for k in ("variant", "ctas_per_sm"):
        _lib.set_option("va." + k, cfg.get(k, 0))
    for nk in (12, 83):
        arrs = [rng.uniform(5, 9, (nk, 11, 73)) for _ in range(5)]
        st = [storage.from_numpy(a, (3, 3, 0)) for a in arrs]
        for rep in range(2):
            stencil.vertical_advection_dycore(*st, 0.15)
        torch.cuda.synchronize()
    print(cfg, "ok", flush=True)
