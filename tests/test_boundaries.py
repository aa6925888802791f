"""Boundary conditions (boundaries/boundary.hpp:57-72 with value / zero / copy conditions).
CPU: the oracle's restatement against the committed golden fixture and -- in the build container -- against the
reference's own boundary<..., gcl::cpu, predicate>::apply.  GPU: gtb_boundary_apply and the condition fused into the
halo unpack, bit-for-bit against the oracle."""
import os

import numpy as np
import pytest

HALOS = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (1, 1, 1, 4, 6)]
HALOS_IJ = [(2, 2, 2, 33, 36), (2, 2, 2, 17, 20), (0, 0, 0, 4, 5)]
needs_ref = pytest.mark.skipif(not os.path.exists("/root/reference/include/gridtools"), reason="reference tree absent")


def shape_of(h):
    return (h[2][4], h[1][4], h[0][4])


def random_mask(seed):
    m = [int(x) for x in np.random.default_rng(seed).integers(0, 2, 27)]
    m[13] = 0
    return m


def test_boundary_golden(oracle, golden):
    g = golden("boundary_14x10x6.npz")
    h = [tuple(int(x) for x in r) for r in g["halos"]]
    for case in ("value_all", "value_masked", "copy_masked"):
        f = [a.copy() for a in g[case + "_in"]]
        mask = None if case == "value_all" else [int(x) for x in g["mask"]]
        oracle.boundary_apply(h, mask, 1 if case.startswith("copy") else 0, float(g["value"]), f)
        for a, b in zip(f, g[case + "_ref"]):
            assert np.array_equal(a, b), case


@needs_ref
@pytest.mark.parametrize("kind,nf", [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3)])
@pytest.mark.parametrize("halos", [HALOS, HALOS_IJ])
def test_boundary_oracle_vs_reference(oracle, kind, nf, halos):
    rng = np.random.default_rng(kind * 10 + nf)
    for mask in (None, random_mask(nf), random_mask(7 * nf + 1)):
        f = [rng.standard_normal(shape_of(halos)) for _ in range(nf)]
        a, b = [x.copy() for x in f], [x.copy() for x in f]
        oracle.boundary_apply(halos, mask, kind, -2.5, a)
        oracle.ref_boundary(halos, mask, kind, -2.5, b)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        changed = sum(int((x != y).sum()) for x, y in zip(a, f))
        assert changed > 0 or (mask is not None and sum(mask) == 0)
        # the compute domain is never touched
        (m0, p0, b0, e0, t0), (m1, p1, b1, e1, t1), (m2, p2, b2, e2, t2) = halos
        for x, y in zip(a, f):
            assert np.array_equal(x[b2:e2 + 1, b1:e1 + 1, b0:e0 + 1], y[b2:e2 + 1, b1:e1 + 1, b0:e0 + 1])


def test_direction_mask_and_predicates():
    from gridtools_b200 import boundaries as bd, gcl
    assert sum(bd.direction_mask(bd.default_predicate)) == 26
    grid = gcl.ProcGrid((2, 2, 1), (False, False, False), 0)  # rank (0, 0): neighbours only towards +i / +j
    m = bd.direction_mask(bd.proc_grid_predicate(grid))
    n = lambda e0, e1, e2: (e0 + 1) + 3 * (e1 + 1) + 9 * (e2 + 1)  # noqa: E731
    assert m[n(1, 0, 0)] == 0 and m[n(0, 1, 0)] == 0 and m[n(1, 1, 0)] == 0
    assert m[n(-1, 0, 0)] == 1 and m[n(0, -1, 0)] == 1 and m[n(-1, 1, 0)] == 1 and m[n(0, 0, 1)] == 1 and m[13] == 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gt():
    import torch
    from gridtools_b200 import _lib, boundaries, gcl, storage
    _lib.check(_lib.lib().gtb_init(0))

    class NS:
        pass
    ns = NS()
    ns.lib, ns.bd, ns.gcl, ns.storage, ns.torch = _lib, boundaries, gcl, storage, torch
    return ns


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("halos", [HALOS, HALOS_IJ])
def test_boundary_apply_gpu(gt, oracle, halos, dtype):
    rng = np.random.default_rng(3)
    for kind, nf, mask in ((0, 1, None), (0, 3, random_mask(1)), (1, 2, None), (1, 3, random_mask(2)), (0, 2, [0] * 27)):
        f = [rng.standard_normal(shape_of(halos)).astype(dtype) for _ in range(nf)]
        want = [x.copy() for x in f]
        oracle.boundary_apply(halos, mask, kind, 4.75, want)
        dev = [gt.torch.from_numpy(x.copy()).cuda() for x in f]
        cond = gt.bd.copy_boundary() if kind == 1 else gt.bd.value_boundary(4.75)
        pred = gt.bd.default_predicate if mask is None else (lambda d, m=mask: m[(d[0] + 1) + 3 * (d[1] + 1) + 9 * (d[2] + 1)])
        gt.bd.boundary(halos, cond, pred).apply(*[t.data_ptr() for t in dev]) if dtype == np.float64 else \
            _apply_f32(gt, halos, cond, pred, dev)
        gt.torch.cuda.synchronize()
        for t, w in zip(dev, want):
            assert np.array_equal(t.cpu().numpy(), w), (kind, nf, mask is None)


def _apply_f32(gt, halos, cond, pred, dev):
    import ctypes as C
    b = gt.bd.boundary(halos, cond, pred)
    arr = (C.c_void_p * len(dev))(*[t.data_ptr() for t in dev])
    gt.lib.check(gt.lib.lib().gtb_boundary_apply(b.desc, b.mask, cond.kind, cond.value, arr, len(dev), 4,
                                                 C.c_void_p(gt.torch.cuda.current_stream().cuda_stream)))


@pytest.mark.gpu
def test_boundary_on_data_stores(gt, oracle):
    """zero_boundary through the storage API: halo 2 in i / j of a padded storage::gpu-like store."""
    ni, nj, nk, H = 33, 9, 4, 2
    rng = np.random.default_rng(9)
    box = rng.standard_normal((nk, nj + 2 * H, ni + 2 * H))
    ds = gt.storage.from_numpy(box, (H, H, 0))
    p0 = ds.padded_lengths[0]
    halos = [(H, H, H, H + ni - 1, p0), (H, H, H, H + nj - 1, nj + 2 * H), (0, 0, 0, nk - 1, nk)]
    gt.bd.boundary(halos, gt.bd.zero_boundary()).apply(ds)
    gt.torch.cuda.synchronize()
    got = ds.to_numpy()
    want = np.zeros_like(box)
    want[:, H:-H, H:-H] = box[:, H:-H, H:-H]
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,periodic", [((2, 2, 1), (False, False, False)), ((1, 2, 1), (True, False, False)),
                                           ((1, 1, 1), (False, False, False))])
def test_boundary_fused_into_unpack(gt, oracle, dims, periodic):
    """distributed_boundaries: exchange, then value_boundary where the process grid has no neighbour -- here written by
    the unpack launch itself.  Expected = oracle exchange followed by oracle boundary_apply with proc_grid_predicate."""
    import ctypes as C
    halos, dtype, n_fields, value = HALOS_IJ, np.float64, 2, -9.5
    size = dims[0] * dims[1] * dims[2]
    hes, fields, dev, grids = [], [], [], []
    rng = np.random.default_rng(4)
    for r in range(size):
        grid = gt.gcl.ProcGrid(dims, periodic, r)
        he = gt.gcl.halo_exchange_dynamic_ut(periodic, grid, dtype, comm=None, transport="p2p")
        for d in range(3):
            he.add_halo(d, *halos[d])
        he.setup(n_fields)
        he.set_boundary(value)
        hes.append(he)
        grids.append(grid)
        fields.append([rng.standard_normal(shape_of(halos)) for _ in range(n_fields)])
        dev.append([gt.torch.from_numpy(a.copy()).cuda() for a in fields[-1]])
    gt.gcl.connect_local(hes)
    ptrs = [[t.data_ptr() for t in d] for d in dev]
    for he, p in zip(hes, ptrs):
        he.pack(p)
    for he, p in zip(hes, ptrs):
        he.wait()
        he.unpack(p)
    gt.torch.cuda.synchronize()
    want = [[a.copy() for a in fs] for fs in fields]
    oracle.halo_exchange_all(halos, dims, periodic, want, 8)
    for r in range(size):
        oracle.boundary_apply(halos, gt.bd.direction_mask(gt.bd.proc_grid_predicate(grids[r])), 0, value, want[r])
    for r in range(size):
        assert hes[r].check() == 0
        for t, w in zip(dev[r], want[r]):
            assert np.array_equal(t.cpu().numpy(), w), (dims, r)
    for he in hes:
        he.set_boundary(None)
        he.close()
