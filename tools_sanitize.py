"""Small run of every kernel variant for compute-sanitizer (memcheck / racecheck) on the GPU box."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gridtools_b200 import _lib, stencil, storage  # noqa: E402

torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
rng = np.random.default_rng(0)
for dtype in (np.float64, np.float32):
    for variant in (1, 2, 3):
        _lib.set_option("hd.variant", variant)
        inp = rng.standard_normal((3, 23, 135)).astype(dtype)
        st = [storage.from_numpy(inp, (2, 2, 0)), storage.from_numpy(inp, (2, 2, 0)),
              storage.from_numpy(np.zeros_like(inp), (2, 2, 0))]
        stencil.horizontal_diffusion(*st)
        torch.cuda.synchronize()
        print("hd", dtype.__name__, variant, "ok", flush=True)
    for cfg in (dict(), dict(variant=3, ctas_per_sm=-2), dict(variant=2, ctas_per_sm=-2, unroll=2), dict(variant=1, scratch=2, threads=32),
                dict(variant=1, ctas_per_sm=-2, threads=32)):
        for k in ("variant", "scratch", "threads", "ctas_per_sm", "save_upos", "unroll"):
            _lib.set_option("va." + k, cfg.get(k, 0))
        arrs = [rng.uniform(5, 9, (12, 11, 41)).astype(dtype) for _ in range(5)]
        st = [storage.from_numpy(a, (3, 3, 0)) for a in arrs]
        stencil.vertical_advection_dycore(*st, 0.15)
        torch.cuda.synchronize()
        print("va", dtype.__name__, cfg, "ok", flush=True)
print("done")
