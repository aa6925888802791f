"""Sweeps the knobs of the default kernels on the GPU box (used to pick the defaults in csrc/*.cu).

    python tools/tune.py [va|hd] > gpurun_out/tune.txt

Timing: K launches back to back between one pair of events over four rotating field sets (what bench.py measures).
"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gridtools_b200 import _lib, stencil, storage  # noqa: E402

NI = NJ = 256
NK = 80


def timeit(plans, n=120):
    import ctypes as C
    h = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for s in range(8):
        plans[s % len(plans)](h)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(n):
        plans[s % len(plans)](h)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.cuda.set_device(0)
    _lib.check(_lib.lib().gtb_init(0))
    if "GTB_PDL" in os.environ:  # programmatic dependent launch: 0 off, 1 on (default)
        _lib.set_option("pdl", int(os.environ["GTB_PDL"]))
        print("pdl =", os.environ["GTB_PDL"])
    if only in ("all", "va"):
        for dtype in (np.float64, np.float32):
            sets = []
            for _ in range(4):
                arrs, dtr = bench.repo_vert_adv(NI, NJ, NK)
                sets.append([storage.from_numpy(x.astype(dtype), (3, 3, 0)) for x in arrs])
            for st in sets:
                for f in st:
                    f.const_target_tensor()
            plans = [stencil.plan("vertical_advection_dycore", *st, dtr_stage=0.15) for st in sets]
            res = []
            combos = [dict(variant=7, bldg=b, ctas_per_sm=p, stages=s, unroll=u) for b, p, s, u in
                      itertools.product((1, 2), (6, 7, 8), (0, 3, 4, 5), (2, 3))] if dtype == np.float64 else \
                     [dict(variant=7, bldg=b) for b in (1, 2)]
            combos += [dict(variant=3, ctas_per_sm=w, threads=32 * nb, stages=4) for w, nb in itertools.product((6, 7, 8), (1, 2, 4))]
            for cfg in combos:
                for k in ("variant", "threads", "unroll", "ctas_per_sm", "stages", "bldg"):
                    _lib.set_option("va." + k, cfg.get(k, 0))
                try:
                    ms = timeit(plans)
                except Exception as e:
                    print("va", dtype.__name__, cfg, "FAILED", e)
                    continue
                b = 6 * np.dtype(dtype).itemsize * NI * NJ * NK
                res.append((ms, "va %s %s: %.2f us -> %.0f GB/s" % (dtype.__name__, " ".join("%s=%d" % kv for kv in cfg.items()),
                                                                  ms * 1e3, b / ms / 1e6)))
            for _, line in sorted(res):
                print(line)
            for k in ("variant", "threads", "unroll", "ctas_per_sm", "stages", "bldg"):
                _lib.set_option("va." + k, 0)
    if only in ("all", "hd"):
        for dtype, n in ((np.float64, 256), (np.float32, 256), (np.float64, 512), (np.float32, 512)):
            sets = []
            for _ in range(4):
                inp, coeff = bench.repo_hori_diff(n, n, NK)
                sets.append([storage.from_numpy(inp.astype(dtype), (2, 2, 0)), storage.from_numpy(coeff.astype(dtype), (2, 2, 0)),
                             storage.from_numpy(np.zeros_like(inp, dtype=dtype), (2, 2, 0))])
            for st in sets:
                for f in st:
                    f.const_target_tensor()
            plans = [stencil.plan("horizontal_diffusion", *st) for st in sets]
            for stages, ctas in list(itertools.product((2, 3, 4, 5), (1, 2))) + [(2, 0), (3, 0), (4, 0), (5, 0)]:
                # ctas = 0: two pipelines in one CTA per SM (hd.variant 4)
                for k, v in (("hd.variant", 2 if ctas else 4), ("hd.stages", stages), ("hd.ctas_per_sm", ctas)):
                    _lib.set_option(k, v)
                try:
                    ms = timeit(plans)
                except Exception as e:
                    print("hd", dtype.__name__, n, stages, ctas, "FAILED", e)
                    continue
                b = 3 * np.dtype(dtype).itemsize * n * n * NK
                print("hd %s %d stages=%d ctas=%d: %.2f us -> %.0f GB/s" % (dtype.__name__, n, stages, ctas, ms * 1e3, b / ms / 1e6))
            for k in ("hd.variant", "hd.stages", "hd.ctas_per_sm"):
                _lib.set_option(k, 0)
            del sets, plans


if __name__ == "__main__":
    main()
