"""Halo exchange kernels on one GPU: several ranks of a process grid live in this process (each with its own
gtb_halo object and arena), messages travel through the peer-to-peer path (same code as over NVLink, the "peer"
pointer simply is local), results are compared bit-for-bit with the oracle's restatement of gcl's
pack -> exchange -> unpack (test pattern of regression/gcl/test_halo_exchange_3D.cpp:66-123)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HALOS = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (0, 1, 0, 4, 6)]
HALOS_IJ = [(2, 2, 2, 33, 36), (2, 2, 2, 17, 20), (0, 0, 0, 4, 5)]  # hori_diff-like: halo 2 in i/j, none in k


@pytest.fixture(scope="module")
def gt():
    import torch
    from gridtools_b200 import _lib, gcl, storage
    _lib.check(_lib.lib().gtb_init(0))

    class NS:
        pass
    ns = NS()
    ns.lib, ns.gcl, ns.storage, ns.torch = _lib, gcl, storage, torch
    return ns


def stamp(plan, grid, field_id, dtype):
    shape = plan.storage_shape()
    a = -np.ones(shape, dtype)
    (m0, p0, b0, e0, t0), (m1, p1, b1, e1, t1), (m2, p2, b2, e2, t2) = plan.halos
    n0, n1, n2 = e0 - b0 + 1, e1 - b1 + 1, e2 - b2 + 1
    k, j, i = np.meshgrid(np.arange(n2), np.arange(n1), np.arange(n0), indexing="ij")
    gi, gj, gk = i + n0 * grid.coords[0], j + n1 * grid.coords[1], k + n2 * grid.coords[2]
    a[b2:e2 + 1, b1:e1 + 1, b0:e0 + 1] = field_id * 1e5 + gi * 1e3 + gj * 10 + gk
    return a


def exchange_in_process(gt, halos, dims, periodic, n_fields, dtype, fused, epochs=1):
    size = dims[0] * dims[1] * dims[2]
    hes, fields, dev = [], [], []
    for r in range(size):
        grid = gt.gcl.ProcGrid(dims, periodic, r)
        he = gt.gcl.halo_exchange_dynamic_ut(periodic, grid, dtype, comm=None, transport="p2p")
        for d in range(3):
            he.add_halo(d, *halos[d])
        he.setup(n_fields)
        hes.append(he)
        fields.append([stamp(he.plan, grid, f, dtype) for f in range(n_fields)])
        dev.append([gt.torch.from_numpy(a.copy()).cuda() for a in fields[-1]])
    gt.gcl.connect_local(hes)
    L = gt.lib.lib()
    import ctypes as C
    stream = C.c_void_p(gt.torch.cuda.current_stream().cuda_stream)
    for _ in range(epochs):
        ptrs = [[t.data_ptr() for t in d] for d in dev]
        # all ranks send first, then all wait: the ranks share one stream in this test
        for he, p in zip(hes, ptrs):
            if fused:
                he.pack(p)
            else:
                arr = (C.c_void_p * n_fields)(*p)
                gt.lib.check(L.gtb_halo_pack(he._h, arr, n_fields, stream))
                gt.lib.check(L.gtb_halo_send(he._h, n_fields, stream))
        for he, p in zip(hes, ptrs):
            he.wait()
            he.unpack(p)
    gt.torch.cuda.synchronize()
    for he in hes:
        assert he.check() == 0
    got = [[t.cpu().numpy() for t in d] for d in dev]
    for he in hes:
        he.close()
    return fields, got


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dims,periodic", [((1, 1, 1), (1, 1, 1)), ((2, 1, 1), (0, 0, 0)), ((2, 2, 1), (1, 0, 1)),
                                           ((2, 2, 2), (0, 1, 0)), ((3, 2, 1), (1, 1, 1))])
def test_exchange_matches_oracle(gt, oracle, dims, periodic, dtype, fused):
    n_fields = 3
    fields, got = exchange_in_process(gt, HALOS, dims, periodic, n_fields, dtype, fused)
    oracle.halo_exchange_all(HALOS, dims, periodic, fields, np.dtype(dtype).itemsize)
    for r, (want, have) in enumerate(zip(fields, got)):
        for f in range(n_fields):
            assert np.array_equal(want[f], have[f]), "rank %d field %d" % (r, f)


def test_many_fields_and_epochs(gt, oracle):
    """More fields than one launch carries (16) and three back-to-back epochs (double-buffered arenas)."""
    dims, periodic, n_fields = (2, 2, 1), (1, 1, 0), 19
    fields, got = exchange_in_process(gt, HALOS_IJ, dims, periodic, n_fields, np.float64, True, epochs=3)
    for _ in range(3):
        oracle.halo_exchange_all(HALOS_IJ, dims, periodic, fields, 8)
    for want, have in zip(fields, got):
        for f in range(n_fields):
            assert np.array_equal(want[f], have[f])


def test_non_periodic_border_untouched(gt, oracle):
    fields, got = exchange_in_process(gt, HALOS_IJ, (1, 2, 1), (0, 0, 0), 1, np.float64, True)
    assert (got[0][0] == -1).any() and (got[1][0] == -1).any()
    oracle.halo_exchange_all(HALOS_IJ, (1, 2, 1), (0, 0, 0), fields, 8)
    assert np.array_equal(fields[0][0], got[0][0]) and np.array_equal(fields[1][0], got[1][0])


def test_too_many_fields_is_an_error(gt):
    grid = gt.gcl.ProcGrid((1, 1, 1), (1, 1, 1), 0)
    he = gt.gcl.halo_exchange_dynamic_ut((1, 1, 1), grid, np.float64, comm=None, transport="p2p")
    for d in range(3):
        he.add_halo(d, *HALOS[d])
    he.setup(1)
    with pytest.raises(ValueError):
        he.pack([1, 2])
    import ctypes as C
    arr = (C.c_void_p * 2)(8, 16)
    assert gt.lib.lib().gtb_halo_pack(he._h, arr, 2, None) == gt.lib.GTB_ERR_ARG
    assert gt.lib.lib().gtb_halo_pack_send(he._h, arr, 1, None) == gt.lib.GTB_ERR_STATE  # not connected yet
    he.close()


def test_bound_exchange_single_call(gt, oracle):
    """halo_exchange.bind(): pack + exchange + unpack as ONE pre-marshalled C call (gtb_halo_exchange).  A periodic
    1x1x1 grid is its own neighbour in every direction, so the whole exchange can run in a single stream."""
    import ctypes as C
    dims, periodic, n_fields = (1, 1, 1), (1, 1, 1), 3
    grid = gt.gcl.ProcGrid(dims, periodic, 0)
    he = gt.gcl.halo_exchange_dynamic_ut(periodic, grid, np.float64, comm=None, transport="p2p")
    for d in range(3):
        he.add_halo(d, *HALOS[d])
    he.setup(n_fields)
    gt.gcl.connect_local([he])
    fields = [stamp(he.plan, grid, f, np.float64) for f in range(n_fields)]
    dev = [gt.torch.from_numpy(a.copy()).cuda() for a in fields]
    run = he.bind(*[t.data_ptr() for t in dev])
    stream = C.c_void_p(gt.torch.cuda.current_stream().cuda_stream)
    for _ in range(3):  # three epochs through the double-buffered arenas
        run(stream)
    gt.torch.cuda.synchronize()
    assert he.check() == 0
    for _ in range(3):
        oracle.halo_exchange_all(HALOS, dims, periodic, [fields], 8)
    for f in range(n_fields):
        assert np.array_equal(fields[f], dev[f].cpu().numpy())
    he.close()


def test_device_side_gates_order_stencil_and_exchange(gt):
    """gtb_stencil_gate / gtb_halo_gate: a periodic rank that is its own neighbour runs `exchange -> hori_diff` steps on two
    streams with no stream events; every step must see the halos of ITS exchange (the field is rewritten in between)."""
    import ctypes as C
    from gridtools_b200 import stencil
    torch = gt.torch
    ni, nj, nk, H = 64, 32, 4, 2
    per = (True, True, False)
    grid = gt.gcl.ProcGrid((1, 1, 1), per, 0)
    he = gt.gcl.halo_exchange_dynamic_ut(per, grid, np.float64, comm=None, transport="p2p")
    rng = np.random.default_rng(12)
    box = rng.standard_normal((nk, nj + 2 * H, ni + 2 * H))
    inp = gt.storage.from_numpy(box, (H, H, 0))
    coeff = gt.storage.from_numpy(np.full_like(box, 0.025), (H, H, 0))
    outs = [gt.storage.from_numpy(np.zeros_like(box), (H, H, 0)) for _ in range(3)]
    p0 = inp.padded_lengths[0]
    he.add_halo(0, H, H, H, H + ni - 1, p0)
    he.add_halo(1, H, H, H, H + nj - 1, nj + 2 * H)
    he.add_halo(2, 0, 0, 0, nk - 1, nk)
    he.setup(1)
    he._connect([he.blob])
    for f in [inp, coeff] + outs:
        f.const_target_tensor()
    gt.lib.set_option("reserve_sms", 4)
    try:
        comp, comm = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
        comp_h, comm_h = C.c_void_p(comp.cuda_stream), C.c_void_p(comm.cuda_stream)
        done = torch.zeros(1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        flag, e0 = he.unpacked_flag(), he.epoch()
        seq = stencil.Sequence()
        steps = 3
        for t in range(steps):
            if t >= 1:
                seq.halo_gate(he, done.data_ptr(), t)  # the previous stencil has read the halos
            seq.halo_exchange(he, [inp], comm_h)
            seq.stencil_gate(flag, e0 + t, done.data_ptr())
            seq.horizontal_diffusion(inp, coeff, outs[t], stream=comp_h)
        seq.run()
        torch.cuda.synchronize()
        assert he.check() == 0 and int(done.item()) == steps and gt.lib.gate_timeouts() == 0
        # expected: periodic halo fill, then the stencil (the field itself never changes)
        from oracle import pyoracle as o
        want_in = box.copy()
        o.halo_exchange_all([(H, H, H, H + ni - 1, ni + 2 * H), (H, H, H, H + nj - 1, nj + 2 * H), (0, 0, 0, nk - 1, nk)],
                            (1, 1, 1), per, [[want_in]], 8)
        want = o.hori_diff(want_in, np.full_like(box, 0.025))
        inner = (slice(None), slice(H, -H), slice(H, -H))
        for t in range(steps):
            outs[t]._host_stale = True
            assert np.array_equal(outs[t].to_numpy()[inner], want[inner]), t
        with pytest.raises(gt.lib.GtbError):  # a gate without reserved SMs could deadlock: refused
            gt.lib.set_option("reserve_sms", 0)
            gt.lib.check(gt.lib.lib().gtb_stencil_gate(flag, 1, None))
    finally:
        gt.lib.set_option("reserve_sms", 0)
        he.close()


def test_generic_exchange_two_field_shapes(gt, oracle):
    """halo_exchange_generic on the device: two field shapes with their own halo descriptors (and element types) in one
    pack / exchange / unpack, 2x2 ranks in this process."""
    HB = [(1, 1, 1, 7, 9), (2, 2, 2, 9, 12), (0, 0, 0, 2, 3)]
    dims, periodic = (2, 2, 1), (True, False, False)
    size = 4
    hgs, host, dev = [], [], []
    for r in range(size):
        grid = gt.gcl.ProcGrid(dims, periodic, r)
        hg = gt.gcl.halo_exchange_generic(periodic, grid, comm=None, transport="p2p")
        hg.setup(3)
        pa, pb = gt.gcl.HaloPlan(HALOS, grid), gt.gcl.HaloPlan(HB, grid)
        a = [stamp(pa, grid, f, np.float64) for f in range(2)]
        b = [stamp(pb, grid, 5, np.float32)]
        da = [gt.torch.from_numpy(x.copy()).cuda() for x in a]
        db = [gt.torch.from_numpy(x.copy()).cuda() for x in b]
        fo = [gt.gcl.field_on_the_fly(da[0].data_ptr(), HALOS, np.float64), gt.gcl.field_on_the_fly(db[0].data_ptr(), HB, np.float32),
              gt.gcl.field_on_the_fly(da[1].data_ptr(), HALOS, np.float64)]
        hg.prepare(*fo)
        hgs.append(hg), host.append((a, b)), dev.append((da, db, fo))
    gt.gcl.connect_local_generic(hgs)
    for hg, (_, _, fo) in zip(hgs, dev):
        hg.pack(*fo)
    for hg, (_, _, fo) in zip(hgs, dev):
        hg.exchange()
        hg.unpack(*fo)
    gt.torch.cuda.synchronize()
    ea, eb = [[x.copy() for x in h[0]] for h in host], [[x.copy() for x in h[1]] for h in host]
    oracle.halo_exchange_all(HALOS, dims, periodic, ea, 8)
    oracle.halo_exchange_all(HB, dims, periodic, eb, 4)
    for r in range(size):
        assert hgs[r].check() == 0
        for t, w in zip(dev[r][0], ea[r]):
            assert np.array_equal(t.cpu().numpy(), w), r
        assert np.array_equal(dev[r][1][0].cpu().numpy(), eb[r][0]), r
    for hg in hgs:
        hg.close()


# ---------------------------------------------------------------------- pinned on the reference's own gcl
LAYOUTS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]  # T_layout_map, 2 = unit stride


def _exchange_layout(gt, user_halos, proc_dims, per_user, layout_gt, proc_layout, start, use_vector):
    """The CUDA exchange for in-process ranks with a data layout / process layout; `start` = [rank][field] host
    arrays in storage order.  Returns the exchanged copies."""
    inc = tuple(2 - v for v in layout_gt)
    per_grid = [0, 0, 0]
    for d in range(3):
        per_grid[proc_layout[d]] = per_user[d]
    n = proc_dims[0] * proc_dims[1] * proc_dims[2]
    hes, dev = [], []
    for r in range(n):
        grid = gt.gcl.ProcGrid(proc_dims, per_grid, r)
        he = gt.gcl.halo_exchange_dynamic_ut(per_user, grid, start[r][0].dtype, layout=inc, proc_layout=proc_layout,
                                             comm=None, transport="p2p")
        for d in range(3):
            he.add_halo(d, *user_halos[d])
        he.setup(len(start[r]))
        hes.append(he)
        dev.append([gt.torch.from_numpy(a.view(np.uint8).copy()).cuda() for a in start[r]])
    gt.gcl.connect_local(hes)
    for he, d in zip(hes, dev):
        ptrs = [t.data_ptr() for t in d]
        he.pack(ptrs) if use_vector else he.pack(*ptrs)
    for he, d in zip(hes, dev):
        ptrs = [t.data_ptr() for t in d]
        he.exchange()
        he.unpack(ptrs) if use_vector else he.unpack(*ptrs)
    gt.torch.cuda.synchronize()
    for he in hes:
        assert he.check() == 0
        he.close()
    return [[t.cpu().numpy().view(a.dtype).reshape(a.shape) for t, a in zip(d, s)] for d, s in zip(dev, start)]


@pytest.mark.parametrize("layout", LAYOUTS)
def test_all_layouts_interfaces_periodicities_equal_reference_gcl(gt, oracle, layout):
    """The sweep of tests/regression/gcl/test_halo_exchange_3D.cpp:150-172,195-208: 6 layouts x {vector, variadic} x
    2^3 periodicities, 16-byte elements (array<int, 4>) and doubles; the expectation is the REFERENCE's gcl run by
    oracle/_ref/libgtref.so (threads as ranks over oracle/mpi_shim/mpi.h)."""
    import itertools
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libgtref.so missing")
    rng = np.random.default_rng(17)
    user_halos = [(2, 3, 2, 24, 28), (4, 4, 4, 15, 20), (3, 3, 3, 9, 13)]  # spec 2 of the reference test, padded in i
    order = np.argsort(layout)
    shape = tuple(user_halos[d][4] for d in order)
    for proc_dims in [(2, 1, 2), (2, 2, 1)]:
        n = proc_dims[0] * proc_dims[1] * proc_dims[2]
        for per in itertools.product((1, 0), repeat=3):
            for use_vector, dtype in ((True, "V16"), (False, np.float64)):
                if dtype == "V16":
                    start = [[rng.integers(0, 1 << 30, shape + (4,)).astype(np.int32).view("V16").reshape(shape)
                              for _ in range(3)] for _ in range(n)]
                else:
                    start = [[rng.standard_normal(shape) for _ in range(3)] for _ in range(n)]
                want = [[a.copy() for a in r] for r in start]
                oracle.ref_gcl_exchange(user_halos, proc_dims, per, want, layout=layout, use_vector=use_vector)
                got = _exchange_layout(gt, user_halos, proc_dims, per, layout, (0, 1, 2), start, use_vector)
                for r in range(n):
                    for f in range(3):
                        assert got[r][f].tobytes() == want[r][f].tobytes(), (proc_dims, per, dtype, r, f)


@pytest.mark.parametrize("proc_layout", [(1, 0, 2), (2, 1, 0), (1, 2, 0)])
def test_process_layouts_equal_reference_gcl(gt, oracle, proc_layout):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libgtref.so missing")
    rng = np.random.default_rng(19)
    user_halos = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (0, 1, 0, 4, 6)]
    for layout in [(2, 1, 0), (0, 1, 2), (1, 2, 0)]:
        order = np.argsort(layout)
        shape = tuple(user_halos[d][4] for d in order)
        for proc_dims, per in [((2, 2, 1), (1, 0, 1)), ((1, 2, 2), (0, 1, 0)), ((3, 1, 2), (1, 1, 1))]:
            n = proc_dims[0] * proc_dims[1] * proc_dims[2]
            start = [[rng.standard_normal(shape).astype(np.float32) for _ in range(2)] for _ in range(n)]
            want = [[a.copy() for a in r] for r in start]
            oracle.ref_gcl_exchange(user_halos, proc_dims, per, want, layout=layout, proc_layout=proc_layout)
            got = _exchange_layout(gt, user_halos, proc_dims, per, layout, proc_layout, start, True)
            for r in range(n):
                for f in range(2):
                    assert got[r][f].tobytes() == want[r][f].tobytes(), (layout, proc_dims, per, r, f)


@pytest.mark.parametrize("name", ["halo_2x2x1_p101", "halo_2x4x1_p000", "halo_1x2x2_p010_l021"])
def test_golden_halo_vectors_on_the_device(gt, golden, name):
    """tests/golden/halo_*.npz: written by make_golden.py from the reference's gcl."""
    g = golden(name + ".npz")
    halos, dims, per = [tuple(int(x) for x in h) for h in g["halos"]], tuple(g["proc_dims"]), tuple(g["periodic"])
    layout, n_fields = tuple(int(x) for x in g["layout"]), int(g["n_fields"])
    n = dims[0] * dims[1] * dims[2]
    start = [[g["start"][r, f].copy() for f in range(n_fields)] for r in range(n)]
    got = _exchange_layout(gt, halos, dims, per, layout, (0, 1, 2), start, True)
    for r in range(n):
        for f in range(n_fields):
            assert np.array_equal(got[r][f], g["result"][r, f]), (r, f)


@pytest.mark.parametrize("layout", [(2, 1, 0), (0, 1, 2), (1, 2, 0)])
def test_generic_one_message_per_neighbour_equals_reference_generic(gt, oracle, layout):
    """gcl::halo_exchange_generic + field_on_the_fly (test_halo_exchange_3D.cpp:241-266): three fields with their own
    halo descriptors in ONE pack launch and ONE unpack launch; expectation = the reference's halo_exchange_generic."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libgtref.so missing")
    rng = np.random.default_rng(23)
    hs = [[(0, 1, 0, 8, 10), (2, 3, 2, 8, 12), (2, 1, 2, 7, 9)], [(1, 1, 1, 9, 11), (2, 2, 2, 8, 11), (0, 0, 0, 5, 6)],
          [(0, 1, 0, 8, 10), (2, 3, 2, 8, 12), (0, 1, 0, 5, 7)]]
    inc = tuple(2 - v for v in layout)
    order = np.argsort(layout)
    for proc_dims, per in [((2, 2, 1), (1, 0, 1)), ((1, 2, 2), (0, 1, 0)), ((2, 1, 1), (0, 0, 0))]:
        n = proc_dims[0] * proc_dims[1] * proc_dims[2]
        start = [[rng.standard_normal(tuple(hs[f][d][4] for d in order)) for f in range(3)] for _ in range(n)]
        want = [[a.copy() for a in r] for r in start]
        oracle.ref_gcl_exchange(hs, proc_dims, per, want, layout=layout, generic=True)
        hgs, dev, fotf = [], [], []
        for r in range(n):
            grid = gt.gcl.ProcGrid(proc_dims, per, r)
            hg = gt.gcl.halo_exchange_generic(per, grid, comm=None, transport="p2p", layout=inc)
            hg.setup(3)
            d = [gt.torch.from_numpy(a.copy()).cuda() for a in start[r]]
            fo = [gt.gcl.field_on_the_fly(d[f].data_ptr(), hs[f], np.float64) for f in range(3)]
            hg.prepare(*fo)
            hgs.append(hg), dev.append(d), fotf.append(fo)
        gt.gcl.connect_local_generic(hgs)
        launches0 = gt.lib.launch_count()
        for hg, fo in zip(hgs, fotf):
            hg.pack(*fo)
        for hg, fo in zip(hgs, fotf):
            hg.exchange()
            hg.unpack(fo)
        gt.torch.cuda.synchronize()
        assert gt.lib.launch_count() - launches0 <= 2 * n  # one pack and one unpack launch per rank
        for r in range(n):
            assert hgs[r].check() == 0
            for f in range(3):
                assert np.array_equal(dev[r][f].cpu().numpy(), want[r][f]), (proc_dims, per, r, f)
        for hg in hgs:
            hg.close()


@pytest.mark.parametrize("dma", [0, 1])
def test_asymmetric_halos_many_epochs(gt, oracle, dma):
    """minus = 1, plus = 0 in i on a non-periodic 3 x 1 x 1 grid: every rank sends towards +i but receives nothing from
    there.  The double-buffered receive slots rely on flags travelling BOTH ways (a segment without payload), five
    epochs with new data each; also through the copy-engine transport (halo.dma)."""
    import ctypes as C
    halos = [(1, 0, 1, 24, 32), (0, 0, 0, 9, 10), (0, 0, 0, 3, 4)]
    dims, periodic, n = (3, 1, 1), (0, 0, 0), 3
    gt.lib.set_option("halo.dma", dma)
    try:
        hes, host, dev = [], [], []
        rng = np.random.default_rng(41)
        for r in range(n):
            grid = gt.gcl.ProcGrid(dims, periodic, r)
            he = gt.gcl.halo_exchange_dynamic_ut(periodic, grid, np.float64, comm=None, transport="p2p")
            for d in range(3):
                he.add_halo(d, *halos[d])
            he.setup(2)
            hes.append(he)
            host.append([rng.standard_normal(he.plan.storage_shape()) for _ in range(2)])
            dev.append([gt.torch.from_numpy(a.copy()).cuda() for a in host[-1]])
        gt.gcl.connect_local(hes)
        stream = C.c_void_p(gt.torch.cuda.current_stream().cuda_stream)
        L = gt.lib.lib()
        for epoch in range(5):
            for r in range(n):  # new interior data every epoch
                for f in range(2):
                    host[r][f] += epoch + 1
                    dev[r][f] += epoch + 1
            for he, d in zip(hes, dev):
                arr = (C.c_void_p * 2)(*[t.data_ptr() for t in d])
                gt.lib.check(L.gtb_halo_pack(he._h, arr, 2, stream) if dma else L.gtb_halo_pack_send(he._h, arr, 2, stream))
                if dma:
                    gt.lib.check(L.gtb_halo_send(he._h, 2, stream))
            for he, d in zip(hes, dev):
                he.unpack([t.data_ptr() for t in d])
            oracle.halo_exchange_all(halos, dims, periodic, host, 8)
        gt.torch.cuda.synchronize()
        for r in range(n):
            assert hes[r].check() == 0
            for f in range(2):
                assert np.array_equal(dev[r][f].cpu().numpy(), host[r][f]), (r, f)
        for he in hes:
            he.close()
    finally:
        gt.lib.set_option("halo.dma", 0)
