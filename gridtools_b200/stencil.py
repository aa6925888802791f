"""Host-side entry points of the B200 stencil backend: one function per spec of the reference's regression suite,
each a thin argument marshaller over the C ABI (include/gtb200.h).  What `stencil::run(spec, backend, grid,
fields...)` (frontend/run.hpp:242-250) is to the reference, these functions are here; the C++ equivalent that plugs
into GridTools' own `run` is include/gtb200/stencil/b200.hpp.

All functions enqueue on torch's current CUDA stream and return without synchronising, like the reference GPU
backend (common/cuda_util.hpp:79-96).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .storage import DataStore


class Grid:
    """Compute domain, the part of core::grid (stencil/core/grid.hpp:91-127) the named kernels need:
    i/j start (= halo) and sizes, k size.  `make_grid(halo_descriptor_i, halo_descriptor_j, nk)`."""

    def __init__(self, i_start, j_start, ni, nj, nk, k_start=0):
        self.i_start, self.j_start, self.k_start = int(i_start), int(j_start), int(k_start)
        self.ni, self.nj, self.nk = int(ni), int(nj), int(nk)

    @property
    def origin(self):
        return (self.i_start, self.j_start, self.k_start)


def make_grid(halo_i, halo_j, nk):
    """halo_i / halo_j: (minus, plus, begin, end, total) like halo_descriptor; or plain ints = domain sizes
    (frontend/make_grid.hpp)."""
    def rng(h):
        if isinstance(h, int):
            return 0, h
        _, _, begin, end, _ = h
        return begin, end - begin + 1
    i0, ni = rng(halo_i)
    j0, nj = rng(halo_j)
    return Grid(i0, j0, ni, nj, nk)


def _grid_of(ds, grid):
    if grid is not None:
        return grid
    ni, nj, nk = ds.compute_domain()
    return Grid(ds.halos[0], ds.halos[1], ni, nj, nk, ds.halos[2])


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _same_dtype(*stores):
    dt = stores[0].dtype
    for s in stores:
        if s.dtype != dt:
            raise TypeError("all fields of a stencil must have the same float type")
    return dt


def copy(src: DataStore, dst: DataStore, grid: Grid = None):
    """copy_stencil.cpp:24-36 via run_single_stage(copy_functor(), backend, grid, in, out)."""
    g = _grid_of(src, grid)
    dt = _same_dtype(src, dst)
    fi, fo = src.field(const=True, origin=g.origin), dst.field(origin=g.origin)
    _lib.check(_lib.lib().gtb_copy(C.byref(fi), C.byref(fo), g.ni, g.nj, g.nk, dt.itemsize, _stream()))


def horizontal_diffusion(inp: DataStore, coeff: DataStore, out: DataStore, grid: Grid = None):
    """horizontal_diffusion.cpp:98-106 : run(spec, backend, grid, in, coeff, out)."""
    g = _grid_of(inp, grid)
    dt = _same_dtype(inp, coeff, out)
    fn = _lib.lib().gtb_hori_diff_f64 if dt.itemsize == 8 else _lib.lib().gtb_hori_diff_f32
    fi, fc, fo = inp.field(True, g.origin), coeff.field(True, g.origin), out.field(False, g.origin)
    _lib.check(fn(C.byref(fi), C.byref(fc), C.byref(fo), g.ni, g.nj, g.nk, _stream()))


def simple_hori_diff(coeff: DataStore, inp: DataStore, out: DataStore, crlato: DataStore, crlatu: DataStore,
                     grid: Grid = None):
    """simple_hori_diff.cpp:63-88 : run(spec, backend, grid, coeff, in, out, crlato, crlatu); crlato / crlatu are
    j-only stores (`storage.builder...selector(0, 1, 0)`)."""
    g = _grid_of(inp, grid)
    dt = _same_dtype(inp, coeff, out, crlato, crlatu)
    fn = _lib.lib().gtb_simple_hori_diff_f64 if dt.itemsize == 8 else _lib.lib().gtb_simple_hori_diff_f32
    fi, fc, fo = inp.field(True, g.origin), coeff.field(True, g.origin), out.field(False, g.origin)
    fro, fru = crlato.field(True), crlatu.field(True)  # origin = their own j halo
    _lib.check(fn(C.byref(fi), C.byref(fc), C.byref(fro), C.byref(fru), C.byref(fo), g.ni, g.nj, g.nk, _stream()))


def vertical_advection_dycore(utens_stage: DataStore, u_stage: DataStore, wcon: DataStore, u_pos: DataStore,
                              utens: DataStore, dtr_stage: float, grid: Grid = None):
    """vertical_advection_dycore.cpp:140-149 : run(spec, backend, grid, utens_stage, u_stage, wcon, u_pos, utens,
    dtr_stage); utens_stage is updated in place, dtr_stage is the global_parameter."""
    g = _grid_of(utens_stage, grid)
    dt = _same_dtype(utens_stage, u_stage, wcon, u_pos, utens)
    if dt.itemsize == 8:
        fn, sc = _lib.lib().gtb_vert_adv_f64, C.c_double(dtr_stage)
    else:
        fn, sc = _lib.lib().gtb_vert_adv_f32, C.c_float(dtr_stage)
    f = [utens_stage.field(False, g.origin)] + [s.field(True, g.origin) for s in (u_stage, wcon, u_pos, utens)]
    _lib.check(fn(*[C.byref(x) for x in f], sc, g.ni, g.nj, g.nk, _stream()))


def tridiagonal(inf: DataStore, diag: DataStore, sup: DataStore, rhs: DataStore, out: DataStore, grid: Grid = None):
    """tridiagonal.cpp:76-98 : forward_thomas + backward_thomas; sup and rhs are overwritten."""
    g = _grid_of(inf, grid)
    dt = _same_dtype(inf, diag, sup, rhs, out)
    if dt.itemsize != 8:
        raise TypeError("tridiagonal is double precision only")
    f = [inf.field(True, g.origin), diag.field(True, g.origin), sup.field(False, g.origin),
         rhs.field(False, g.origin), out.field(False, g.origin)]
    _lib.check(_lib.lib().gtb_tridiagonal_f64(*[C.byref(x) for x in f], g.ni, g.nj, g.nk, _stream()))


def prepare_tracers(outs, ins, rho: DataStore, grid: Grid = None):
    """advection_pdbott_prepare_tracers.cpp:36-58 : expandable_run<2>(spec, backend, grid, outs, ins, rho) -- all
    tracers in one launch."""
    if len(outs) != len(ins):
        raise ValueError("prepare_tracers: outs and ins must have the same length")
    g = _grid_of(rho, grid)
    dt = _same_dtype(rho, *outs, *ins)
    if dt.itemsize != 8:
        raise TypeError("prepare_tracers is double precision only")
    n = len(outs)
    fo = (_lib.Field * max(n, 1))(*[s.field(False, g.origin) for s in outs])
    fi = (_lib.Field * max(n, 1))(*[s.field(True, g.origin) for s in ins])
    fr = rho.field(True, g.origin)
    _lib.check(_lib.lib().gtb_prepare_tracers_f64(fo, fi, n, C.byref(fr), g.ni, g.nj, g.nk, _stream()))


def plan(name, *stores, grid: Grid = None, **scalars):
    """Pre-marshalled call: `f = plan("horizontal_diffusion", inp, coeff, out); f()` enqueues the same launch as
    `horizontal_diffusion(inp, coeff, out)` on torch's current stream (or `f(stream_handle)`), with the argument
    structures built once -- for time loops whose step is a few tens of microseconds, where ctypes marshalling would
    otherwise bound the step time.  The stores must stay alive and keep their device buffers."""
    if name == "horizontal_diffusion":
        inp, coeff, out = stores
        g = _grid_of(inp, grid)
        dt = _same_dtype(inp, coeff, out)
        fn = _lib.lib().gtb_hori_diff_f64 if dt.itemsize == 8 else _lib.lib().gtb_hori_diff_f32
        f = [inp.field(True, g.origin), coeff.field(True, g.origin), out.field(False, g.origin)]
        args = tuple(C.byref(x) for x in f) + (g.ni, g.nj, g.nk)
    elif name == "vertical_advection_dycore":
        utens_stage, u_stage, wcon, u_pos, utens = stores
        g = _grid_of(utens_stage, grid)
        dt = _same_dtype(*stores)
        dtr = scalars["dtr_stage"]
        if dt.itemsize == 8:
            fn, sc = _lib.lib().gtb_vert_adv_f64, C.c_double(dtr)
        else:
            fn, sc = _lib.lib().gtb_vert_adv_f32, C.c_float(dtr)
        f = [utens_stage.field(False, g.origin)] + [s.field(True, g.origin) for s in (u_stage, wcon, u_pos, utens)]
        args = tuple(C.byref(x) for x in f) + (sc, g.ni, g.nj, g.nk)
    else:
        raise ValueError("plan(): unknown spec %r" % (name,))
    chk = _lib.check

    def run(stream_handle=None):
        chk(fn(*args, stream_handle if stream_handle is not None else _stream()))
    run.keepalive = (f, stores)
    return run


class Sequence:
    """Recorded time loop (gtb_seq, include/gtb200.h): stencil launches, halo exchanges and the event record / wait
    operations between streams are recorded once and replayed slice by slice with one native call each -- the
    per-call cost of the reference's own C++ driver loop (copy_stencil_parallel.cpp:126-145) instead of ctypes'.

        seq = Sequence()
        seq.wait(comm, 0); seq.halo_exchange(he, [field], comm); seq.record(1, comm)
        seq.wait(comp, 1); seq.horizontal_diffusion(inp, coeff, out, stream=comp); seq.record(0, comp)
        seq.run()            # or seq.run(first, count)

    Streams are raw cudaStream_t handles (ints / c_void_p).  Stores and halo objects must outlive the sequence."""

    def __init__(self):
        h = C.c_void_p()
        _lib.check(_lib.lib().gtb_seq_create(C.byref(h)))
        self._h = h
        self._keep = []

    def __len__(self):
        return _lib.lib().gtb_seq_size(self._h)

    def horizontal_diffusion(self, inp, coeff, out, grid: Grid = None, stream=None):
        g = _grid_of(inp, grid)
        dt = _same_dtype(inp, coeff, out)
        f = [inp.field(True, g.origin), coeff.field(True, g.origin), out.field(False, g.origin)]
        _lib.check(_lib.lib().gtb_seq_add_hori_diff(self._h, dt.itemsize, *[C.byref(x) for x in f], g.ni, g.nj, g.nk,
                                                    stream))
        self._keep.append((inp, coeff, out))

    def vertical_advection_dycore(self, utens_stage, u_stage, wcon, u_pos, utens, dtr_stage, grid: Grid = None,
                                  stream=None):
        g = _grid_of(utens_stage, grid)
        dt = _same_dtype(utens_stage, u_stage, wcon, u_pos, utens)
        f = [utens_stage.field(False, g.origin)] + [s.field(True, g.origin) for s in (u_stage, wcon, u_pos, utens)]
        _lib.check(_lib.lib().gtb_seq_add_vert_adv(self._h, dt.itemsize, *[C.byref(x) for x in f], float(dtr_stage),
                                                   g.ni, g.nj, g.nk, stream))
        self._keep.append((utens_stage, u_stage, wcon, u_pos, utens))

    def prepare_tracers(self, outs, ins, rho, grid: Grid = None, stream=None):
        g = _grid_of(rho, grid)
        n = len(outs)
        fo = (_lib.Field * max(n, 1))(*[s.field(False, g.origin) for s in outs])
        fi = (_lib.Field * max(n, 1))(*[s.field(True, g.origin) for s in ins])
        fr = rho.field(True, g.origin)
        _lib.check(_lib.lib().gtb_seq_add_prepare_tracers(self._h, fo, fi, n, C.byref(fr), g.ni, g.nj, g.nk, stream))
        self._keep.append((outs, ins, rho))

    def halo_exchange(self, he, fields, stream=None):
        """pack + exchange + unpack of `fields` through the halo_exchange_dynamic_ut `he` (p2p transport)."""
        arr, n = he._ptrs(list(fields))
        _lib.check(_lib.lib().gtb_seq_add_halo_exchange(self._h, he._h, arr, n, stream))
        self._keep.append((he, fields))

    def stencil_gate(self, wait_flag=None, wait_value=0, post_counter=None):
        """Device-side gate for the stencil recorded next (gtb_stencil_gate): wait until *wait_flag >= wait_value, add 1
        to *post_counter when done.  Pointers are raw device addresses (ints) or None."""
        _lib.check(_lib.lib().gtb_seq_add_stencil_gate(self._h, wait_flag, int(wait_value), post_counter))

    def halo_gate(self, he, counter, value):
        """The unpack of the exchange recorded next waits on the device until *counter >= value (gtb_halo_gate)."""
        _lib.check(_lib.lib().gtb_seq_add_halo_gate(self._h, he._h, counter, int(value)))

    def record(self, event, stream=None):
        _lib.check(_lib.lib().gtb_seq_add_record(self._h, int(event), stream))

    def wait(self, stream, event):
        _lib.check(_lib.lib().gtb_seq_add_wait(self._h, stream, int(event)))

    def mark(self, mark, stream=None):
        """Timing mark (gtb_seq_add_mark): recorded in `stream` when the sequence reaches this point."""
        _lib.check(_lib.lib().gtb_seq_add_mark(self._h, int(mark), stream))

    def stamp(self, device_u64, stream=None):
        """Diagnosis: a one-thread kernel stores %globaltimer at `device_u64` when the sequence reaches this point."""
        _lib.check(_lib.lib().gtb_seq_add_stamp(self._h, device_u64, stream))

    def elapsed_ms(self, mark_a, mark_b):
        ms = C.c_float()
        _lib.check(_lib.lib().gtb_seq_elapsed_ms(self._h, int(mark_a), int(mark_b), C.byref(ms)))
        return float(ms.value)

    def run(self, first=0, count=None):
        n = len(self) - first if count is None else count
        _lib.check(_lib.lib().gtb_seq_run(self._h, int(first), int(n)))

    def close(self):
        if self._h is not None:
            _lib.lib().gtb_seq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def as_numpy_interior(ds: DataStore, grid: Grid = None):
    g = _grid_of(ds, grid)
    a = ds.const_host_view()
    return np.array(a[g.k_start:g.k_start + g.nk, g.j_start:g.j_start + g.nj, g.i_start:g.i_start + g.ni])
