"""gcl parity, pinned on the reference itself: the UNMODIFIED gcl/halo_exchange.hpp (and descriptors.hpp,
Halo_Exchange_3D.hpp, proc_grids_3D.hpp below it) runs in-process with threads as ranks over oracle/mpi_shim/mpi.h
(oracle/ref_gcl.cpp in oracle/_ref/libgtref.so) and checks

  1. itself, with the reference test's own expectation (tests/regression/gcl/test_halo_exchange_3D.cpp:66-123,
     150-172: coordinate stamps, all 6 layouts x vector/variadic interface x 2^3 periodicities);
  2. the C restatement gto_halo_* of oracle/gt_oracle.c;
  3. the host plan of the product (gridtools_b200/gcl.py: HaloPlan + ProcGrid feed gtb_halo_create) for every data
     layout x process layout, dynamic_ut and generic;
  4. the committed golden vectors tests/golden/halo_*.npz (generated from the reference by make_golden.py).
"""
import itertools

import numpy as np
import pytest

from gridtools_b200.gcl import HaloPlan, ProcGrid

LAYOUTS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]  # T_layout_map, 2 = unit stride
PERIODICITIES = list(itertools.product((1, 0), repeat=3))


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libgtref.so not built (reference tree absent)")
    return oracle


def storage(layout, sizes, dtype, inner=()):
    """C-contiguous array holding a field whose USER dimension d has GridTools layout value layout[d], plus its
    user-indexed view (view[i, j, k] = element (i, j, k))."""
    order = np.argsort(layout)  # slowest ... fastest
    a = np.zeros(tuple(sizes[d] for d in order) + tuple(inner), dtype=dtype)
    view = a.transpose(tuple(np.argsort(order)) + tuple(range(3, 3 + len(inner))))
    assert view.shape[:3] == tuple(sizes)
    return a, view


def coords_of(rank, dims):
    return (rank // (dims[1] * dims[2]), (rank // dims[2]) % dims[1], rank % dims[2])


def stamp_state(spec_dims, halos, proc_dims, coords, field_no):
    """initial_state(i, j, k, field_no) of test_halo_exchange_3D.cpp:66-78 on the whole storage box."""
    axes = []
    for d in range(3):
        size = spec_dims[d]
        i = np.arange(size + halos[d][0] + halos[d][1]) - halos[d][0]
        c = np.full(i.shape, coords[d])
        if coords[d] == 0:
            c = np.where(i < 0, proc_dims[d], c)
        if coords[d] == proc_dims[d] - 1:
            c = np.where((i >= size) & ~((coords[d] == 0) & (i < 0)), -1, c)
        axes.append(c * size + i)
    gi, gj, gk = np.meshgrid(*axes, indexing="ij")
    return np.stack([gi, gj, gk, np.full(gi.shape, field_no)], axis=-1).astype(np.int32)


def in_halo_mask(spec_dims, halos):
    m = []
    for d in range(3):
        i = np.arange(spec_dims[d] + halos[d][0] + halos[d][1]) - halos[d][0]
        m.append((i < 0) | (i >= spec_dims[d]))
    a, b, c = np.meshgrid(*m, indexing="ij")
    return a | b | c


def border_mask(spec_dims, halos, proc_dims, coords, periodicity):
    m = []
    for d in range(3):
        i = np.arange(spec_dims[d] + halos[d][0] + halos[d][1]) - halos[d][0]
        if periodicity[d]:
            m.append(np.zeros(i.shape, bool))
        else:
            m.append(((i < 0) & (coords[d] == 0)) | ((i >= spec_dims[d]) & (coords[d] + 1 == proc_dims[d])))
    a, b, c = np.meshgrid(*m, indexing="ij")
    return a | b | c


def descriptors(spec_dims, halos, totals=None):
    """make_halo_descriptors of test_halo_exchange_3D.cpp:139-148."""
    return [(halos[d][0], halos[d][1], halos[d][0], spec_dims[d] + halos[d][0] - 1,
             totals[d] if totals else spec_dims[d] + halos[d][0] + halos[d][1]) for d in range(3)]


def run_reference_stamp_case(ref, spec_dims, field_halos, proc_dims, layout, periodicity, use_vector, generic):
    n_ranks = proc_dims[0] * proc_dims[1] * proc_dims[2]
    arrays, views = [], []
    for r in range(n_ranks):
        c = coords_of(r, proc_dims)
        fa, fv = [], []
        for f in range(3):
            sizes = [spec_dims[d] + field_halos[f][d][0] + field_halos[f][d][1] for d in range(3)]
            a, v = storage(layout, sizes, np.int32, inner=(4,))
            v[...] = np.where(in_halo_mask(spec_dims, field_halos[f])[..., None], -1,
                              stamp_state(spec_dims, field_halos[f], proc_dims, c, f))
            fa.append(a)
            fv.append(v)
        arrays.append(fa)
        views.append(fv)
    h = [descriptors(spec_dims, field_halos[f]) for f in range(3)]
    flat16 = [[a.reshape(-1, 4).view(np.dtype("V16")).reshape(a.shape[:3]) for a in fa] for fa in arrays]
    ref.ref_gcl_exchange(h if generic else h[0], proc_dims, periodicity, flat16, layout=layout, use_vector=use_vector,
                         generic=generic)
    for r in range(n_ranks):
        c = coords_of(r, proc_dims)
        for f in range(3):
            want = np.where(border_mask(spec_dims, field_halos[f], proc_dims, c, periodicity)[..., None], -1,
                            stamp_state(spec_dims, field_halos[f], proc_dims, c, f))
            assert np.array_equal(views[r][f], want), (r, f, layout, periodicity)


SAME = lambda h: [h, h, h]
# the reference's own parameter sets (test_halo_exchange_3D.cpp:150-172), the first one shrunk 123x56x76 -> 41x19x26
ALL_SPECS = [((41, 19, 26), SAME([(2, 3), (1, 2), (2, 1)])), ((23, 12, 7), SAME([(2, 2), (4, 4), (3, 3)])),
             ((12, 12, 12), SAME([(2, 2), (2, 2), (2, 2)]))]
GENERIC_SPECS = [((33, 18, 29), [[(0, 1), (2, 3), (2, 1)], [(0, 1), (2, 3), (2, 1)], [(0, 1), (2, 3), (0, 1)]]),
                 ((30, 15, 35), SAME([(3, 3), (1, 1), (2, 2)]))]


@pytest.mark.parametrize("proc_dims", [(2, 2, 1), (2, 1, 2), (1, 2, 1)])
@pytest.mark.parametrize("spec", ALL_SPECS)
def test_reference_gcl_runs_its_own_test_through_the_shim(ref, spec, proc_dims):
    """halo_exchange_3D_all: 6 layouts x {vector, variadic} x 2^3 periodicities, expectation = the reference test's."""
    for layout in LAYOUTS:
        for use_vector in (True, False):
            for per in PERIODICITIES:
                run_reference_stamp_case(ref, spec[0], spec[1], proc_dims, layout, per, use_vector, False)


@pytest.mark.parametrize("proc_dims", [(2, 2, 1), (1, 2, 2)])
@pytest.mark.parametrize("spec", GENERIC_SPECS)
def test_reference_generic_runs_its_own_test_through_the_shim(ref, spec, proc_dims):
    """halo_exchange_3D_generic (test_halo_exchange_3D.cpp:241-266)."""
    for layout in LAYOUTS:
        for use_vector in (True, False):
            for per in PERIODICITIES[::3]:
                run_reference_stamp_case(ref, spec[0], spec[1], proc_dims, layout, per, use_vector, True)


# ------------------------------------------------------------------------------ 2. the C restatement
HALOS = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (0, 1, 0, 4, 6)]   # minus != plus, padding in i, k has only a plus halo
HALOS_IJ = [(2, 2, 2, 33, 40), (2, 2, 2, 17, 20), (0, 0, 0, 4, 5)]


def random_fields(rng, halos, n_ranks, n_fields, dtype):
    shape = (halos[2][4], halos[1][4], halos[0][4])
    if np.dtype(dtype).itemsize == 16:
        return [[rng.integers(0, 1 << 30, shape + (4,)).astype(np.int32).view("V16").reshape(shape)
                 for _ in range(n_fields)] for _ in range(n_ranks)]
    return [[rng.standard_normal(shape).astype(dtype) for _ in range(n_fields)] for _ in range(n_ranks)]


def clone(fields):
    return [[a.copy() for a in r] for r in fields]


def same(a, b):
    return all(x.tobytes() == y.tobytes() for ra, rb in zip(a, b) for x, y in zip(ra, rb))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, "V16"])
@pytest.mark.parametrize("dims,periodic", [((1, 1, 1), (1, 1, 1)), ((2, 1, 1), (0, 0, 0)), ((2, 2, 1), (1, 0, 1)),
                                           ((2, 2, 2), (0, 1, 0)), ((3, 2, 1), (1, 1, 1)), ((1, 4, 1), (0, 1, 0)),
                                           ((2, 4, 1), (0, 0, 0))])
@pytest.mark.parametrize("halos", [HALOS, HALOS_IJ])
def test_restated_oracle_equals_reference_gcl(ref, halos, dims, periodic, dtype):
    n = dims[0] * dims[1] * dims[2]
    start = random_fields(np.random.default_rng(3), halos, n, 3, dtype)
    a, b = clone(start), clone(start)
    ref.halo_exchange_all(halos, dims, periodic, a, np.dtype(dtype).itemsize)
    # identity layouts: layout_map<2,1,0> = user dimension 0 has unit stride, proc layout <0,1,2>
    ref.ref_gcl_exchange(halos, dims, periodic, b, layout=(2, 1, 0))
    assert same(a, b)
    assert not same(a, start)
    c = clone(start)  # split-phase calls (gcl/halo_exchange.hpp:286-304) and the variadic interface
    ref.ref_gcl_exchange(halos, dims, periodic, c, layout=(2, 1, 0), use_vector=False, split_phase=True)
    assert same(a, c)


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 4, 1), (3, 2, 2)])
@pytest.mark.parametrize("periodic", [(0, 0, 0), (1, 0, 1), (1, 1, 1)])
def test_proc_grid_equals_reference(ref, dims, periodic):
    """MPI_3D_process_grid_t::proc (proc_grids_3D.hpp:179-211) vs gto_proc_neighbour vs the product's ProcGrid."""
    import ctypes as C
    cd, cp = (C.c_int * 3)(*dims), (C.c_int * 3)(*periodic)
    for rank in range(dims[0] * dims[1] * dims[2]):
        g = ProcGrid(dims, periodic, rank)
        for d in [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (1, -1, 0), (-1, 1, 1), (1, 1, -1)]:
            want = ref.ref_gcl_proc(dims, periodic, rank, *d)
            assert g.proc(*d) == want and ref.lib().gto_proc_neighbour(cd, cp, *g.coords, *d) == want, (rank, d)


# ------------------------------------------------------------------------------ 3. the product's host plan
def plan_exchange(halos_user, proc_dims, periodic_user, fields, layout_gt, proc_layout, per_field_halos=None):
    """pack -> deliver -> unpack of all ranks with the product's HaloPlan / ProcGrid (numpy codec): what the CUDA
    path executes, without a device."""
    n_ranks = len(fields)
    inc = tuple(2 - v for v in layout_gt)  # reverse_map: position in increasing-stride order
    per_grid = [0, 0, 0]
    for d in range(3):
        per_grid[proc_layout[d]] = periodic_user[d]
    n_fields = len(fields[0])
    plans = [[HaloPlan(per_field_halos[f] if per_field_halos else halos_user, ProcGrid(proc_dims, per_grid, r), inc,
                       proc_layout) for f in range(n_fields)] for r in range(n_ranks)]
    msgs = {}
    for r in range(n_ranks):
        for f in range(n_fields):
            p = plans[r][f]
            for n in range(27):
                if p.send_count(n):
                    msgs[(r, n, f)] = p.pack_numpy(n, [fields[r][f]])
    for r in range(n_ranks):
        for f in range(n_fields):
            p = plans[r][f]
            for n in range(27):
                if p.recv_count(n):
                    p.unpack_numpy(n, [fields[r][f]], msgs[(p.neighbour[n], 26 - n, f)])


@pytest.mark.parametrize("proc_layout", [(0, 1, 2), (1, 0, 2), (2, 1, 0), (1, 2, 0), (2, 0, 1)])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_host_plan_equals_reference_gcl_for_every_layout(ref, layout, proc_layout):
    rng = np.random.default_rng(11)
    user_halos = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (0, 1, 0, 4, 6)]
    for proc_dims, per in [((2, 2, 1), (1, 0, 1)), ((1, 2, 2), (0, 1, 0)), ((2, 1, 2), (1, 1, 1)), ((3, 1, 2), (0, 0, 0))]:
        n = proc_dims[0] * proc_dims[1] * proc_dims[2]
        sizes = [h[4] for h in user_halos]
        start = []
        for r in range(n):
            start.append([storage(layout, sizes, np.float64)[0] + rng.standard_normal(1) for _ in range(3)])
            for a in start[-1]:
                a += rng.standard_normal(a.shape)
        a, b = clone(start), clone(start)
        ref.ref_gcl_exchange(user_halos, proc_dims, per, a, layout=layout, proc_layout=proc_layout)
        plan_exchange(user_halos, proc_dims, per, b, layout, proc_layout)
        assert same(a, b), (layout, proc_layout, proc_dims, per)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_host_plan_equals_reference_generic(ref, layout):
    """halo_exchange_generic: every field_on_the_fly brings its own halo descriptors (descriptor_generic_manual.hpp)."""
    rng = np.random.default_rng(5)
    hs = [[(0, 1, 0, 8, 10), (2, 3, 2, 8, 12), (2, 1, 2, 7, 9)], [(1, 1, 1, 9, 11), (2, 2, 2, 8, 11), (0, 0, 0, 5, 6)],
          [(0, 1, 0, 8, 10), (2, 3, 2, 8, 12), (0, 1, 0, 5, 7)]]
    for proc_dims, per in [((2, 2, 1), (1, 0, 1)), ((1, 2, 2), (0, 1, 0))]:
        n = proc_dims[0] * proc_dims[1] * proc_dims[2]
        start = [[storage(layout, [h[4] for h in hs[f]], np.float64)[0] for f in range(3)] for _ in range(n)]
        for r in start:
            for a in r:
                a += rng.standard_normal(a.shape)
        a, b = clone(start), clone(start)
        ref.ref_gcl_exchange(hs, proc_dims, per, a, layout=layout, generic=True)
        plan_exchange(None, proc_dims, per, b, layout, (0, 1, 2), per_field_halos=hs)
        assert same(a, b), (layout, proc_dims, per)


# ------------------------------------------------------------------------------ 4. golden vectors
@pytest.mark.parametrize("name", ["halo_2x2x1_p101", "halo_2x4x1_p000", "halo_1x2x2_p010_l021"])
def test_golden_halo_vectors(oracle, golden, name):
    """tests/golden/halo_*.npz were written by make_golden.py from the reference's gcl; the restatement (identity
    layouts) and the product's host plan must reproduce them -- also where /root/reference does not exist."""
    g = golden(name + ".npz")
    halos, dims, per = [tuple(h) for h in g["halos"]], tuple(g["proc_dims"]), tuple(g["periodic"])
    layout, n_fields = tuple(g["layout"]), int(g["n_fields"])
    n = dims[0] * dims[1] * dims[2]
    start = [[g["start"][r, f].copy() for f in range(n_fields)] for r in range(n)]
    want = [[g["result"][r, f] for f in range(n_fields)] for r in range(n)]
    b = clone(start)
    plan_exchange(halos, dims, per, b, layout, (0, 1, 2))
    assert same(b, want)
    if layout == (2, 1, 0):
        a = clone(start)
        oracle.halo_exchange_all(halos, dims, per, a, 8)
        assert same(a, want)
    if oracle.have_ref():
        c = clone(start)
        oracle.ref_gcl_exchange(halos, dims, per, c, layout=layout)
        assert same(c, want)
