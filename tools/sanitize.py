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
    for cfg in (dict(), dict(variant=7, ctas_per_sm=-2), dict(variant=7, ctas_per_sm=3), dict(variant=6, ctas_per_sm=-1), dict(variant=5, ctas_per_sm=-2),
                dict(variant=3, ctas_per_sm=-2), dict(variant=2, ctas_per_sm=-2, unroll=2), dict(variant=1, scratch=2, threads=32),
                dict(variant=1, ctas_per_sm=-2, threads=32)):
        for k in ("variant", "scratch", "threads", "ctas_per_sm", "save_upos", "unroll"):
            _lib.set_option("va." + k, cfg.get(k, 0))
        for nk in (12, 83):  # 83 levels: TMEM tier + shared-memory slab tier
            arrs = [rng.uniform(5, 9, (nk, 11, 73)).astype(dtype) for _ in range(5)]
            st = [storage.from_numpy(a, (3, 3, 0)) for a in arrs]
            for rep in range(2):
                stencil.vertical_advection_dycore(*st, 0.15)
            torch.cuda.synchronize()
        print("va", dtype.__name__, cfg, "ok", flush=True)
    # simple_hori_diff, boundary conditions
    inp = rng.standard_normal((3, 23, 135)).astype(dtype)
    si, sc = storage.from_numpy(inp, (2, 2, 0)), storage.from_numpy(inp, (2, 2, 0))
    so = storage.from_numpy(np.zeros_like(inp), (2, 2, 0))
    jb = storage.builder.type(dtype).dimensions(135, 23, 3).halos(2, 2, 0).selector(0, 1, 0)
    stencil.simple_hori_diff(sc, si, so, jb.value(1.0).build(), jb.value(0.5).build())
    from gridtools_b200 import boundaries as bd
    halos = [(2, 2, 2, 132, si.padded_lengths[0]), (2, 2, 2, 20, 23), (0, 0, 0, 2, 3)]
    bd.boundary(halos, bd.value_boundary(1.5)).apply(si, so)
    bd.boundary(halos, bd.copy_boundary()).apply(so, si)
    torch.cuda.synchronize()
    print("simple_hori_diff + boundaries", dtype.__name__, "ok", flush=True)
print("done")
