"""Sweeps the kernel variants on the GPU box and prints a table (used to pick the defaults in csrc/*.cu).

    python tools/tune.py > gpurun_out/tune.txt
"""
import itertools
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gridtools_b200 import _lib, stencil, storage  # noqa: E402

NI = NJ = 256
NK = 80


def timeit(run, sets, n=60):
    for s in range(6):
        run(sets[s % len(sets)])
    torch.cuda.synchronize()
    evs = []
    for s in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(sets[s % len(sets)])
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in evs]
    return statistics.median(t), min(t)


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.cuda.set_device(0)
    _lib.check(_lib.lib().gtb_init(0))
    print("device", _lib.device_info())
    # ---- copy as a bandwidth yardstick
    a = np.zeros((NK, NJ + 4, NI + 4))
    cs = [[storage.from_numpy(a, (2, 2, 0)), storage.from_numpy(a, (2, 2, 0))] for _ in range(4)]
    for st in cs:
        for f in st:
            f.const_target_tensor()
    med, mn = timeit(lambda st: stencil.copy(*st), cs)
    print("copy 256x256x80 f64: median %.2f us min %.2f us -> %.0f GB/s" % (med * 1e3, mn * 1e3, 16 * NI * NJ * NK / med / 1e6))
    # ---- hori_diff
    for dtype, n in ((np.float64, 256), (np.float32, 256), (np.float64, 512)) if only in ("all", "hd") else ():
        sets = []
        for _ in range(3):
            inp, coeff = bench.repo_hori_diff(n, n, NK)
            inp, coeff = inp.astype(dtype), coeff.astype(dtype)
            sets.append([storage.from_numpy(inp, (2, 2, 0)), storage.from_numpy(coeff, (2, 2, 0)),
                         storage.from_numpy(np.zeros_like(inp), (2, 2, 0))])
        for st in sets:
            for f in st:
                f.const_target_tensor()
        for variant, stages, ctas in itertools.product((1, 2, 3), (2, 3, 4, 5), (1, 2)):
            if stages == 5 and ctas == 2 and dtype == np.float64:
                continue
            for k, v in (("hd.variant", variant), ("hd.stages", stages), ("hd.ctas_per_sm", ctas)):
                _lib.set_option(k, v)
            try:
                med, mn = timeit(lambda st: stencil.horizontal_diffusion(*st), sets)
            except Exception as e:
                print("hd", dtype.__name__, n, variant, stages, ctas, "FAILED", e)
                continue
            b = 3 * np.dtype(dtype).itemsize * n * n * NK
            print("hd %s %d variant=%d stages=%d ctas=%d: median %.2f us min %.2f us -> %.0f GB/s" % (
                dtype.__name__, n, variant, stages, ctas, med * 1e3, mn * 1e3, b / med / 1e6))
        del sets
    # ---- vert_adv
    for dtype in (np.float64, np.float32) if only in ("all", "va") else ():
        sets = []
        for _ in range(2):
            arrs, dtr = bench.repo_vert_adv(NI, NJ, NK)
            sets.append([storage.from_numpy(x.astype(dtype), (3, 3, 0)) for x in arrs])
        for st in sets:
            for f in st:
                f.const_target_tensor()
        combos = []
        for wps, nb, stages, save in itertools.product((6, 7, 8), (1, 2, 4), (4,), (1, 2)):
            combos.append(dict(variant=3, unroll=4, ctas_per_sm=wps, threads=32 * nb, stages=stages, save_upos=save, persist=-1))
        for stages in (3, 6):
            combos.append(dict(variant=3, unroll=4, ctas_per_sm=7, threads=128, stages=stages, save_upos=1, persist=-1))
        combos.append(dict(variant=3, unroll=8, ctas_per_sm=5, stages=0, save_upos=1, persist=-1))
        combos.append(dict(variant=3, unroll=8, ctas_per_sm=7, stages=0, save_upos=1, persist=-1))
        combos.append(dict(variant=3, unroll=4, ctas_per_sm=7, threads=128, stages=4, save_upos=1, persist=0))
        combos.append(dict(variant=3, unroll=4, ctas_per_sm=14, threads=32, stages=2, save_upos=2, persist=-1))
        for kc, stages in ((2, 5), (2, 4), (4, 3), (4, 2)):
            combos.append(dict(variant=4, unroll=kc, ctas_per_sm=7, stages=stages, persist=-1))
        combos.append(dict(variant=2, unroll=4, ctas_per_sm=7, save_upos=1, persist=-1))
        combos.append(dict(variant=2, unroll=4, ctas_per_sm=7, save_upos=2, persist=-1))
        combos.append(dict(variant=1, scratch=1, threads=64, unroll=8, ctas_per_sm=0, save_upos=1, persist=-1))
        results = []
        for cfg in combos:
            for k in ("variant", "scratch", "threads", "unroll", "ctas_per_sm", "save_upos", "stages"):
                _lib.set_option("va." + k, cfg.get(k, 0))
            _lib.set_option("l2.persist_mb", cfg["persist"])
            try:
                med, mn = timeit(lambda st: stencil.vertical_advection_dycore(*st, 0.15), sets, n=20)
            except Exception as e:
                print("va", dtype.__name__, cfg, "FAILED", e)
                continue
            b = 6 * np.dtype(dtype).itemsize * NI * NJ * NK
            results.append((med, "va %s %s: median %.2f us min %.2f us -> %.0f GB/s" % (
                dtype.__name__, " ".join("%s=%d" % kv for kv in cfg.items()), med * 1e3, mn * 1e3, b / med / 1e6)))
        for _, line in sorted(results):
            print(line)


if __name__ == "__main__":
    main()
