"""Timeline of the paired-warp vertical advection kernel (va.debug & 128): per pair, when the forward / backward
passes begin and end, relative to the earliest start stamp of the launch."""
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
buf = np.zeros((2048, 32), np.int64)
for s in range(6):
    stencil.vertical_advection_dycore(*sets[s % 2], dtr)
torch.cuda.synchronize()
_lib.set_option("va.debug", 128)
_lib.check(_lib.lib().gtb_debug_trace(buf.ctypes.data, buf.nbytes))  # clear
for rep in range(3):
    for s in range(4):  # back to back; the stamps of the LAST launch survive
        stencil.vertical_advection_dycore(*sets[s % 2], dtr)
    _lib.check(_lib.lib().gtb_debug_trace(buf.ctypes.data, buf.nbytes))
    used = buf[buf[:, 0] > 0]
    t0 = used[:, 0].min()
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
