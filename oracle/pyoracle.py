"""ctypes bindings of the CPU checkers (TEST INFRASTRUCTURE ONLY -- see gt_oracle.h).

* ``libgtoracle.so``  : plain-C restatement (``gt_oracle.c``)
* ``_ref/libgtref.so``: the unmodified reference headers driven by ``ref_driver.cpp``

Arrays are numpy, shape ``(d2, d1, d0)`` C-contiguous, i.e. i is the fastest index and the box includes the halo
(``d0 = ni + 2H``).  Importing this module from the product package is a bug (tests/test_boundary.py greps for it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Field(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_i", C.c_int64), ("stride_j", C.c_int64), ("stride_k", C.c_int64)]


class Halo(C.Structure):
    _fields_ = [("minus", C.c_int), ("plus", C.c_int), ("begin", C.c_int), ("end", C.c_int), ("total", C.c_int)]


def build(ref=True):
    """(Re)build the checkers with oracle/Makefile (idempotent)."""
    subprocess.run(["make", "-C", _HERE, "libgtoracle.so"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libgtoracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _lib = C.CDLL(path)
        _lib.gto_halo_send_count.restype = C.c_int64
        _lib.gto_halo_recv_count.restype = C.c_int64
        _lib.gto_halo_pack.restype = C.c_int64
        _lib.gto_halo_unpack.restype = C.c_int64
    return _lib


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libgtref.so"))


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libgtref.so"))
    return _ref


def field(a, halo):
    """Field descriptor of the interior origin of box array ``a`` (shape (d2, d1, d0)) with IJ halo ``halo``."""
    assert a.ndim == 3 and a.flags.c_contiguous
    d2, d1, d0 = a.shape
    hi, hj = (halo, halo) if np.isscalar(halo) else halo
    ptr = a.ctypes.data + a.itemsize * (hi + d0 * hj)
    return Field(ptr, 1, d0, d0 * d1)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with status {rc}")


# ------------------------------------------------------------------------------- C restatement
def copy(src, halo=0):
    out = np.zeros_like(src)
    d2, d1, d0 = src.shape
    _chk(lib().gto_copy(C.byref(field(src, halo)), C.byref(field(out, halo)), d0 - 2 * halo, d1 - 2 * halo, d2,
                        src.itemsize), "gto_copy")
    return out


def hori_diff(inp, coeff, out=None, halo=2):
    d2, d1, d0 = inp.shape
    out = np.zeros_like(inp) if out is None else out
    fn = {np.dtype("f8"): lib().gto_hori_diff_f64, np.dtype("f4"): lib().gto_hori_diff_f32}[inp.dtype]
    _chk(fn(C.byref(field(inp, halo)), C.byref(field(coeff, halo)), C.byref(field(out, halo)), d0 - 2 * halo,
            d1 - 2 * halo, d2), "gto_hori_diff")
    return out


def simple_hori_diff(inp, coeff, crlato, crlatu, halo=2):
    """simple_hori_diff.cpp:25-61.  crlato / crlatu: 1-d arrays over the storage's j (length d1, halo included)."""
    d2, d1, d0 = inp.shape
    out = np.zeros_like(inp)
    crlato, crlatu = np.ascontiguousarray(crlato, inp.dtype), np.ascontiguousarray(crlatu, inp.dtype)
    assert crlato.shape == (d1,) and crlatu.shape == (d1,)
    fn = {np.dtype("f8"): lib().gto_simple_hori_diff_f64, np.dtype("f4"): lib().gto_simple_hori_diff_f32}[inp.dtype]
    _chk(fn(C.byref(field(inp, halo)), C.byref(field(coeff, halo)), C.c_void_p(crlato.ctypes.data + halo * inp.itemsize),
            C.c_void_p(crlatu.ctypes.data + halo * inp.itemsize), C.byref(field(out, halo)), d0 - 2 * halo,
            d1 - 2 * halo, d2), "gto_simple_hori_diff")
    return out


def vert_adv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage, halo=3):
    """Returns the updated utens_stage (input is not modified)."""
    d2, d1, d0 = utens_stage.shape
    res = utens_stage.copy()
    if utens_stage.dtype == np.dtype("f8"):
        fn, sc = lib().gto_vert_adv_f64, C.c_double(dtr_stage)
    else:
        fn, sc = lib().gto_vert_adv_f32, C.c_float(dtr_stage)
    _chk(fn(C.byref(field(res, halo)), C.byref(field(u_stage, halo)), C.byref(field(wcon, halo)),
            C.byref(field(u_pos, halo)), C.byref(field(utens, halo)), sc, d0 - 2 * halo, d1 - 2 * halo, d2),
         "gto_vert_adv")
    return res


def tridiagonal(inf, diag, sup, rhs):
    """Returns (out, sup', rhs') -- sup and rhs are overwritten by the forward sweep like in the reference."""
    d2, d1, d0 = inf.shape
    sup, rhs = sup.copy(), rhs.copy()
    out = np.zeros_like(inf)
    _chk(lib().gto_tridiagonal_f64(*[C.byref(field(a, 0)) for a in (inf, diag, sup, rhs, out)], d0, d1, d2),
         "gto_tridiagonal")
    return out, sup, rhs


def prepare_tracers(ins, rho):
    d2, d1, d0 = rho.shape
    outs = [np.zeros_like(a) for a in ins]
    fo = (Field * len(ins))(*[field(a, 0) for a in outs])
    fi = (Field * len(ins))(*[field(a, 0) for a in ins])
    _chk(lib().gto_prepare_tracers_f64(fo, fi, len(ins), C.byref(field(rho, 0)), d0, d1, d2), "gto_prepare_tracers")
    return outs


def halos3(h):
    """h: three (minus, plus, begin, end, total) tuples in increasing-stride order."""
    return (Halo * 3)(*[Halo(*x) for x in h])


def halo_exchange_all(h, dims, periodic, fields, elem_size):
    """fields: list over ranks of lists of flat/ND numpy arrays (modified in place)."""
    n_fields = len(fields[0])
    flat = [f for r in fields for f in r]
    ptrs = (C.c_void_p * len(flat))(*[f.ctypes.data for f in flat])
    _chk(lib().gto_halo_exchange_all(halos3(h), (C.c_int * 3)(*dims), (C.c_int * 3)(*[int(p) for p in periodic]),
                                     ptrs, n_fields, elem_size), "gto_halo_exchange_all")


def ref_gcl_exchange(halos, proc_dims, periodic, fields, layout=(2, 1, 0), proc_layout=(0, 1, 2), use_vector=True,
                     generic=False, split_phase=False):
    """The REFERENCE's own gcl (oracle/ref_gcl.cpp over oracle/mpi_shim/mpi.h, threads as ranks): one pack / exchange /
    unpack of gcl::halo_exchange_dynamic_ut<layout_map<*layout>, layout_map<*proc_layout>, T, gcl::cpu> or, with
    generic=True, gcl::halo_exchange_generic<layout_map<*proc_layout>, gcl::cpu> with one field_on_the_fly per field.

    layout: the reference's T_layout_map values (GridTools convention: the dimension with value 2 has unit stride);
    halos: (minus, plus, begin, end, total) per USER dimension -- three for dynamic_ut, n_fields x three for generic;
    periodic: USER dimension order; fields: list over ranks (row-major Cartesian ranks) of lists of C-contiguous numpy
    arrays addressed as storage memory (modified in place); element size 4, 8 or 16 bytes."""
    n_fields = len(fields[0])
    flat = [f for r in fields for f in r]
    assert all(f.flags.c_contiguous for f in flat)
    esz = flat[0].dtype.itemsize
    hs = [int(x) for h in halos for x in (h if np.isscalar(h[0]) else [y for t in h for y in t])]
    assert len(hs) == (15 * n_fields if generic else 15), len(hs)
    ptrs = (C.c_void_p * len(flat))(*[f.ctypes.data for f in flat])
    i3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])
    _chk(ref().gtref_gcl_exchange(i3(layout), i3(proc_layout), i3(proc_dims), i3(periodic), (C.c_int * len(hs))(*hs),
                                  esz, n_fields, ptrs, int(use_vector), int(generic), int(split_phase)),
         "gtref_gcl_exchange")


def ref_gcl_proc(proc_dims, periodic, rank, di, dj, dk):
    """MPI_3D_process_grid_t<3>::proc(di, dj, dk) of `rank` (proc_grids_3D.hpp:179-211), periodic in grid order."""
    i3 = lambda v: (C.c_int * 3)(*[int(x) for x in v])
    return ref().gtref_gcl_proc(i3(proc_dims), i3(periodic), rank, di, dj, dk)


def boundary_apply(h, mask, kind, value, fields):
    """boundaries/apply.hpp:44-56; fields: numpy arrays [d2, d1, d0] (modified in place); mask: 27 ints or None."""
    ptrs = (C.c_void_p * len(fields))(*[f.ctypes.data for f in fields])
    m = (C.c_int * 27)(*[int(x) for x in mask]) if mask is not None else None
    _chk(lib().gto_boundary_apply(halos3(h), m, int(kind), C.c_double(value), ptrs, len(fields), fields[0].itemsize),
         "gto_boundary_apply")


def ref_boundary(h, mask, kind, value, fields):
    """The reference's boundary<value_boundary / copy_boundary, gcl::cpu, predicate>::apply on cpu_ifirst stores
    (oracle/_ref/libgtref.so); double precision."""
    d2, d1, d0 = fields[0].shape
    hh = (C.c_int * 15)(*[int(x) for t in h for x in t])
    m = (C.c_int * 27)(*[int(x) for x in (mask if mask is not None else [1] * 27)])
    ptrs = (C.c_void_p * 3)(*([f.ctypes.data for f in fields] + [None] * (3 - len(fields))))
    _chk(ref().gtref_boundary(int(kind), C.c_double(value), hh, m, d0, d1, d2, ptrs, len(fields)), "gtref_boundary")


# ------------------------------------------------------------------------------- reference build
COPY, HORI_DIFF, VERT_ADV, TRIDIAGONAL, SIMPLE_HORI_DIFF = 0, 1, 2, 3, 4
CPU_IFIRST, CPU_KFIRST, NAIVE = 0, 1, 2
BACKENDS = {"cpu_ifirst": CPU_IFIRST, "cpu_kfirst": CPU_KFIRST, "naive": NAIVE}


def ref_run(stencil, backend, ins, outs, ni, nj, nk, scalar=0.0, nrep=0, flush=False):
    """Run the reference backend.  ``outs`` are box arrays that receive the results.  Returns the list of timings."""
    pin = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * max(3, len(outs)))(*([a.ctypes.data for a in outs] + [None] * (3 - len(outs))))
    times = (C.c_double * max(nrep, 1))()
    be = BACKENDS[backend] if isinstance(backend, str) else backend
    _chk(ref().gtref_run(stencil, be, ins[0].itemsize, ni, nj, nk, pin, pout, C.c_double(scalar), nrep, int(flush),
                         times), "gtref_run")
    return list(times)[:nrep]


def ref_num_threads():
    return ref().gtref_num_threads()


def repo_hori_diff(ni, nj, nk):
    """(in, coeff, out) of horizontal_diffusion_repository on the (nk, nj+4, ni+4) box."""
    d0, d1 = ni + 4, nj + 4
    arrs = [np.zeros((nk, d1, d0)) for _ in range(3)]
    ref().gtref_repo_hori_diff(d0, d1, nk, *[C.c_void_p(a.ctypes.data) for a in arrs])
    return arrs


def repo_simple_hori_diff(ni, nj, nk):
    """(in, coeff, crlato[d1], crlatu[d1], out_simple) of horizontal_diffusion_repository on the (nk, nj+4, ni+4) box."""
    d0, d1 = ni + 4, nj + 4
    arrs = [np.zeros((nk, d1, d0)) for _ in range(3)]
    cro, cru = np.zeros(d1), np.zeros(d1)
    ref().gtref_repo_simple_hori_diff(d0, d1, nk, *[C.c_void_p(a.ctypes.data) for a in arrs[:2]],
                                      C.c_void_p(cro.ctypes.data), C.c_void_p(cru.ctypes.data),
                                      C.c_void_p(arrs[2].ctypes.data))
    return arrs[0], arrs[1], cro, cru, arrs[2]


def repo_vert_adv(ni, nj, nk, want_out=True):
    """([utens_stage_in, u_stage, wcon, u_pos, utens], utens_stage_out, dtr_stage) on the (nk, nj+6, ni+6) box."""
    d0, d1 = ni + 6, nj + 6
    arrs = [np.zeros((nk, d1, d0)) for _ in range(5)]
    out = np.zeros((nk, d1, d0))
    dtr = C.c_double()
    ptrs = (C.c_void_p * 5)(*[a.ctypes.data for a in arrs])
    ref().gtref_repo_vert_adv(d0, d1, nk, ptrs, C.c_void_p(out.ctypes.data) if want_out else None, C.byref(dtr))
    return arrs, out, dtr.value
