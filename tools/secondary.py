"""Times the secondary kernels (simple_hori_diff, tridiagonal, copy, prepare_tracers, boundary value fill) at 256x256x80
fp64 with CUDA events around 100 back-to-back launches over rotating field sets; algorithmic bytes as in SURVEY.md 8d."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gridtools_b200 import _lib, boundaries as bd, stencil, storage
torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
ni = nj = 256
nk = 80
pts = ni * nj * nk
rng = np.random.default_rng(0)


def timeit(run, n_sets, reps=100):
    for s in range(10):
        run(s % n_sets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(reps):
        run(s % n_sets)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def report(name, us, bytes_per_pt):
    print("%-28s %8.2f us  %9.0f Mpts/s  %6.0f GB/s algorithmic" % (name, us, pts / us, pts * bytes_per_pt / us / 1e3), flush=True)


def field(h, value=None):
    box = rng.standard_normal((nk, nj + 2 * h, ni + 2 * h)) if value is None else np.full((nk, nj + 2 * h, ni + 2 * h), value)
    ds = storage.from_numpy(box, (h, h, 0))
    ds.const_target_tensor()
    return ds


S = 4
# simple_hori_diff (simple_hori_diff.cpp): in, coeff, out = 24 B/pt
sets = []
for _ in range(S):
    jb = storage.builder.type(np.float64).dimensions(ni + 4, nj + 4, nk).halos(2, 2, 0).selector(0, 1, 0)
    cro, cru = jb.value(1.0).build(), jb.value(0.9).build()
    cro.const_target_tensor(), cru.const_target_tensor()
    sets.append((field(2, 0.025), field(2), field(2, 0.0), cro, cru))
report("simple_hori_diff", timeit(lambda s: stencil.simple_hori_diff(*sets[s]), S), 24)
# tridiagonal (tridiagonal.cpp): 5 fields = 40 B/pt
sets = [[field(0, -1.0), field(0, 3.0), field(0, 1.0), field(0, 3.0), field(0, 0.0)] for _ in range(S)]
report("tridiagonal", timeit(lambda s: stencil.tridiagonal(*sets[s]), S), 40)
# copy: 16 B/pt
sets = [[field(0), field(0, 0.0)] for _ in range(8)]
report("copy", timeit(lambda s: stencil.copy(*sets[s]), 8), 16)
# prepare_tracers x11: (2*11+1)*8 B per point
tr = [([field(0, 0.0) for _ in range(11)], [field(0) for _ in range(11)], field(0, 1.1)) for _ in range(2)]
report("prepare_tracers x11", timeit(lambda s: stencil.prepare_tracers(*tr[s]), 2), 8 * 23)
# boundary value fill of a halo-3 field (moves only the halo: report us)
f = field(3)
p0 = f.padded_lengths[0]
halos = [(3, 3, 3, 3 + ni - 1, p0), (3, 3, 3, 3 + nj - 1, nj + 6), (0, 0, 0, nk - 1, nk)]
b = bd.boundary(halos, bd.value_boundary(1.0))
us = timeit(lambda s: b.apply(f), 1)
print("%-28s %8.2f us  (halo of 3 around 256x256x80: %.1f MB written)" % ("boundary value (8 directions)", us,
      ((ni + 6) * (nj + 6) - ni * nj) * nk * 8 / 1e6), flush=True)
