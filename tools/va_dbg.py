"""Phase timing of the TMEM variants with the va.debug knob (1 skip backward, 2 skip forward math, 8 skip output
stores, 16 no dependent chain in the backward sweep) and a read-only / copy bandwidth yardstick."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
from tune import timeit
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
KEYS = ("variant", "ctas_per_sm", "stages", "debug", "unroll", "threads", "save_upos", "stagger")
def setopt(**cfg):
    for k in KEYS:
        _lib.set_option("va." + k, cfg.get(k, 0))
# yardsticks: torch copy (read+write), torch sum (read only) on 1 GiB
x = torch.empty(2**27, dtype=torch.float64, device="cuda").normal_()
y = torch.empty_like(x)
for name, fn, nbytes in (("copy", lambda: y.copy_(x), 2 * x.numel() * 8), ("sum(read only)", lambda: x.sum(), x.numel() * 8),
                         ("fill(write only)", lambda: y.fill_(1.0), x.numel() * 8)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("yardstick %s: %.0f GB/s" % (name, nbytes * 10 / (a.elapsed_time(b) * 1e-3) / 1e9), flush=True)
del x, y
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, 80)
    sets.append([storage.from_numpy(a, (3, 3, 0)) for a in arrs])
for st in sets:
    for f in st:
        f.const_target_tensor()
b = 48 * 256 * 256 * 80
for w in (8, 7):
    for dbg in (0, 1, 2, 3, 8, 16, 18, 24, 26, 10):
        setopt(variant=5, ctas_per_sm=w, debug=dbg)
        # average over back-to-back launches (event resolution is ~2 us on single launches)
        for s in range(6):
            stencil.vertical_advection_dycore(*sets[s % 2], dtr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(100):
            stencil.vertical_advection_dycore(*sets[s % 2], dtr)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 10
        print("variant 5 warps=%d debug=%2d: %.2f us per launch back to back -> %.0f GB/s algorithmic" % (w, dbg, us, b / us / 1e3), flush=True)
