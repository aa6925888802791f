"""Vertical advection, TMEM variants (va.variant = 5, 6, 7): parity against variant 3 / the oracle at the bench size,
and a timing table (CUDA events around 100 back-to-back launches over two rotating field sets > L2)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
KEYS = ("variant", "ctas_per_sm", "stages", "debug", "unroll", "threads", "save_upos", "stagger")
which = [int(a) for a in sys.argv[1:]] or [5, 6, 7]


def setopt(**cfg):
    for k in KEYS:
        _lib.set_option("va." + k, cfg.get(k, 0))


def fields(dtype, nk=80):
    arrs, dtr = bench.repo_vert_adv(256, 256, nk)
    return [a.astype(dtype) for a in arrs], dtr


PARITY = {5: [dict(variant=5), dict(variant=5, ctas_per_sm=7), dict(variant=5, ctas_per_sm=-40), dict(variant=5, ctas_per_sm=4)],
          6: [dict(variant=6), dict(variant=6, ctas_per_sm=7), dict(variant=6, ctas_per_sm=-40),
              dict(variant=6, ctas_per_sm=-3, stages=2, unroll=2)],
          7: [dict(variant=7), dict(variant=7, ctas_per_sm=7), dict(variant=7, ctas_per_sm=3), dict(variant=7, ctas_per_sm=-40), dict(variant=7, stages=2, unroll=2),
              dict(variant=7, ctas_per_sm=-3, stages=3, unroll=4), dict(variant=7, ctas_per_sm=-1)]}
for dtype in (np.float64, np.float32):
    for nk in (80, 61, 137):
        arrs, dtr = fields(dtype, nk)
        outs = {}
        cfgs = [dict(variant=3)] + [c for v in which for c in PARITY[v]]
        for cfg in cfgs:
            setopt(**cfg)
            st = [storage.from_numpy(x, (3, 3, 0)) for x in arrs]
            for rep in range(2):  # twice: the second launch starts from the ticket counters the first one left
                st[0] = storage.from_numpy(arrs[0], (3, 3, 0))
                stencil.vertical_advection_dycore(*st, dtr)
            torch.cuda.synchronize()
            outs[str(cfg)] = st[0].to_numpy()
        ref = outs[str(dict(variant=3))]
        bad = [k for k, v in outs.items() if not np.array_equal(v, ref, equal_nan=True)]
        print("parity", np.dtype(dtype).name, "nk=%d" % nk, "%d configs identical to variant 3" % (len(outs) - len(bad)),
              "MISMATCH: %s" % bad if bad else "", flush=True)
        if nk == 80:
            from oracle import pyoracle as o
            want = o.vert_adv(*arrs, dtr)
            inner = (slice(None), slice(3, -3), slice(3, -3))
            print("parity", np.dtype(dtype).name, "variant 3 vs oracle:", np.array_equal(ref[inner], want[inner]), flush=True)

TIMING = {5: [dict(variant=5), dict(variant=5, ctas_per_sm=7)],
          6: [dict(variant=6), dict(variant=6, ctas_per_sm=7)],
          7: [dict(variant=7), dict(variant=7, unroll=2), dict(variant=7, ctas_per_sm=7)]}
for dtype in (np.float64, np.float32):
    sets = []
    for _ in range(2):
        arrs, dtr = fields(dtype)
        sets.append([storage.from_numpy(x, (3, 3, 0)) for x in arrs])
    for st in sets:
        for f in st:
            f.const_target_tensor()
    b = 6 * np.dtype(dtype).itemsize * 256 * 256 * 80
    for cfg in [dict(variant=3)] + [c for v in which for c in TIMING[v]]:
        setopt(**cfg)
        try:
            for s in range(6):
                stencil.vertical_advection_dycore(*sets[s % 2], dtr)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(100):
                stencil.vertical_advection_dycore(*sets[s % 2], dtr)
            e1.record()
            torch.cuda.synchronize()
        except Exception as e:
            print("time", np.dtype(dtype).name, cfg, "FAILED", e, flush=True)
            continue
        us = e0.elapsed_time(e1) * 10
        print("time %s %s: %.2f us per launch -> %.0f GB/s" % (np.dtype(dtype).name, cfg, us, b / us / 1e3), flush=True)
