"""Timeline of the paired-warp vertical advection kernel (va.debug & 128): per pair, when the forward / backward
passes begin and end, relative to the earliest start stamp of the launch.

    python tools/va_trace.py [bldg=1] [ctas_per_sm=7] ...      (va.* options)"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, stencil, storage
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
sets = []
for _ in range(2):
    arrs, dtr = bench.repo_vert_adv(256, 256, 80)
    sets.append([storage.from_numpy(a, (3, 3, 0)) for a in arrs])
for st in sets:
    for f in st:
        f.const_target_tensor()
_lib.set_option("va.variant", 7)
extra_debug = 0
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "debug":
        extra_debug = int(v)
    _lib.set_option(k if k in ("pdl", "reserve_sms") else "va." + k, int(v))
print("options:", " ".join(sys.argv[1:]))
buf = np.zeros((2048, 32), np.int64)
for s in range(6):
    stencil.vertical_advection_dycore(*sets[s % 2], dtr)
torch.cuda.synchronize()
_lib.set_option("va.debug", 128 | extra_debug)
_lib.check(_lib.lib().gtb_debug_trace(buf.ctypes.data, buf.nbytes))  # clear
for rep in range(3):
    for s in range(4):  # back to back; the stamps of the LAST launch survive
        stencil.vertical_advection_dycore(*sets[s % 2], dtr)
    _lib.check(_lib.lib().gtb_debug_trace(buf.ctypes.data, buf.nbytes))
    used = buf[buf[:, 0] > 0].copy()
    t0 = used[:, 0].min()
    used[used < t0] = 0  # stamps of earlier launches (pairs that drew fewer strips this time)
    names = ["start", "F0 begin", "F0 end", "B0 begin", "B0 end", "F1 begin", "F1 end", "B1 begin", "B1 end", "F2 begin", "F2 end",
             "B2 begin", "B2 end"]
    print("launch %d: %d pairs" % (rep, len(used)))
    for e, nm in enumerate(names):
        v = used[:, e][used[:, e] > 0]
        if len(v):
            r = (v - t0) / 1e3
            print("  %-9s n=%4d  min %6.2f  median %6.2f  p90 %6.2f  max %6.2f us" % (nm, len(v), r.min(), np.median(r),
                                                                                     np.percentile(r, 90), r.max()))
    d = used[:, 2] - used[:, 1]
    print("  F0 duration median %.2f us; F1 duration median %.2f us; B0 %.2f us; B1 %.2f us" % (
        np.median(d) / 1e3, np.median((used[:, 6] - used[:, 5])[used[:, 6] > 0]) / 1e3,
        np.median((used[:, 4] - used[:, 3])[used[:, 4] > 0]) / 1e3, np.median((used[:, 8] - used[:, 7])[used[:, 8] > 0]) / 1e3))

_lib.set_option("va.debug", extra_debug)
import ctypes as C
plans = [stencil.plan("vertical_advection_dycore", *st, dtr_stage=dtr) for st in sets]
h = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for s in range(10):
    plans[s % 2](h)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in range(200):
    plans[s % 2](h)
b.record()
torch.cuda.synchronize()
print("back to back, no stamps: %.2f us per launch" % (a.elapsed_time(b) / 200 * 1e3))
