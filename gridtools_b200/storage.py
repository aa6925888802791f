"""Host-side mirror of the reference's storage layer for the `storage::gpu` traits.

Follows storage/builder.hpp (type / dimensions / halos / value / initializer / name / build), storage/info.hpp:39-53
(unit-stride dimension padded to the alignment), storage/data_store.hpp:64-72 (the first non-halo element is
128-byte aligned) and the host<->target state machine of data_store.hpp:86-147.  Layout is `layout_map<2,1,0>`
(storage/gpu.hpp:69-105): i has stride 1.

Device memory, streams and copies are torch's (plumbing); kernels never are.
"""
import math

import numpy as np
import torch

from . import _lib

BYTE_ALIGNMENT = 128  # storage_alignment(gpu), storage/gpu.hpp:78


class DataStore:
    """3-d field with halo, i-first layout, host mirror with lazy synchronisation."""

    def __init__(self, dtype, lengths, halos, alignment=BYTE_ALIGNMENT, device=None, name=""):
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype("f8"), np.dtype("f4")), "float and double fields only"
        assert len(lengths) == 3 and len(halos) == 3
        self.name = name
        self.lengths = tuple(int(x) for x in lengths)
        self.halos = tuple(int(x) for x in halos)
        isz = self.dtype.itemsize
        ea = max(1, alignment // math.gcd(isz, alignment))  # traits.hpp:32 elem_alignment
        d0, d1, d2 = self.lengths
        p0 = (d0 + ea - 1) // ea * ea
        self.padded_lengths = (p0, d1, d2)
        self.strides = (1, p0, p0 * d1)
        self.length = p0 * d1 * d2
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tdtype = torch.float64 if isz == 8 else torch.float32
        self._raw = torch.zeros(self.length + ea, dtype=tdtype, device=self.device)
        off = sum(h * s for h, s in zip(self.halos, self.strides)) * isz
        addr = self._raw.data_ptr() + off
        byte_align = max(alignment, isz)
        self._base = (addr + byte_align - 1) // byte_align * byte_align - off  # data_store.hpp:67-71
        shift = (self._base - self._raw.data_ptr()) // isz
        self._dev = self._raw[shift:shift + self.length].view(d2, d1, p0)
        # the pinned host mirror is allocated on first host access (a 4096x4096x80 field is 10.7 GB): a store that is
        # only ever filled and read on the device (target_tensor()) never pays for it
        self._host_shape, self._tdtype = (d2, d1, p0), tdtype
        self._host_buf = None
        self._host_np = None
        self._host_stale = False
        self._dev_stale = False

    @property
    def _host_t(self):
        if self._host_buf is None:
            t = torch.zeros(self._host_shape, dtype=self._tdtype)
            self._host_buf = t.pin_memory() if self.device.type == "cuda" else t
            self._host_np = self._host_buf.numpy()
        return self._host_buf

    @property
    def _host(self):
        self._host_t
        return self._host_np

    # ------------------------------------------------------------------ host / target access (data_store.hpp:86-147)
    def host_view(self):
        """Writable numpy view [k, j, i] (halo included); the device copy becomes stale."""
        self._sync_host()
        self._dev_stale = True
        return self._host[:, :, :self.lengths[0]]

    def const_host_view(self):
        self._sync_host()
        return self._host[:, :, :self.lengths[0]]

    def target_tensor(self):
        """torch view [k, j, i_padded] of the device copy; the host copy becomes stale (get_target_ptr)."""
        self._sync_dev()
        self._host_stale = True
        return self._dev

    def const_target_tensor(self):
        self._sync_dev()
        return self._dev

    def _sync_host(self):
        if self._host_stale:
            self._host_t.copy_(self._dev)
            self._host_stale = False

    def _sync_dev(self):
        if self._dev_stale and self._host_buf is not None:
            self._dev.copy_(self._host_t)
            self._dev_stale = False

    def update_target_async(self):
        """Host -> device on the current stream from pinned memory (for end-to-end timing)."""
        self._dev.copy_(self._host_t, non_blocking=True)
        self._dev_stale = False

    def update_host_async(self):
        self._host_t.copy_(self._dev, non_blocking=True)
        self._host_stale = False

    def _box_copy(self, lo, hi, to_device):
        """Sub-box [lo, hi) (storage coordinates i, j, k) between the pinned host mirror and the device on torch's
        current stream (gtb_copy_box_async)."""
        import ctypes as C
        isz = self.dtype.itemsize
        off = sum(o * s for o, s in zip(lo, self.strides)) * isz
        n = [h - o for o, h in zip(lo, hi)]
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().gtb_copy_box_async(self._base + off, self._host_t.data_ptr() + off, isz, self.strides[1],
                                                 self.strides[2], n[0], n[1], n[2], int(to_device), stream))
        return n[0] * n[1] * n[2] * isz

    def update_target_box_async(self, lo, hi):
        """Host -> device of the sub-box only (e.g. the compute domain plus the halo a stencil reads)."""
        nbytes = self._box_copy(lo, hi, True)
        self._dev_stale = False
        return nbytes

    def update_host_box_async(self, lo, hi):
        nbytes = self._box_copy(lo, hi, False)
        self._host_stale = False
        return nbytes

    @property
    def nbytes_host(self):
        return self.length * self.dtype.itemsize

    # ------------------------------------------------------------------ descriptors for the C ABI
    def raw_ptr(self, const=False):
        """Pointer to storage element (0,0,0), halo included (what gcl takes)."""
        self._sync_dev()
        if not const:
            self._host_stale = True
        return self._base

    def field(self, const=False, origin=None):
        """gtb_field at the first compute-domain point (origin-shifted SID, stencil/core/backend.hpp:29-34)."""
        org = self.halos if origin is None else origin
        off = sum(o * s for o, s in zip(org, self.strides)) * self.dtype.itemsize
        return _lib.Field(self.raw_ptr(const) + off, *self.strides)

    def compute_domain(self):
        return tuple(d - 2 * h for d, h in zip(self.lengths, self.halos))

    # ------------------------------------------------------------------ numpy convenience (dense [k, j, i] boxes)
    def assign(self, box):
        box = np.asarray(box)
        assert box.shape == tuple(reversed(self.lengths)), (box.shape, self.lengths)
        self.host_view()[...] = box

    def to_numpy(self):
        return np.array(self.const_host_view())


class _Builder:
    """storage::builder<traits> look-alike: every setter returns a new builder (builder.hpp)."""

    def __init__(self, **kw):
        self._kw = kw

    def _with(self, **kw):
        d = dict(self._kw)
        d.update(kw)
        return _Builder(**d)

    def type(self, dtype):
        return self._with(dtype=dtype)

    def dimensions(self, *d):
        return self._with(lengths=d)

    def halos(self, *h):
        return self._with(halos=h)

    def value(self, v):
        return self._with(value=v, initializer=None)

    def initializer(self, fn):
        return self._with(initializer=fn, value=None)

    def name(self, n):
        return self._with(name=n)

    def selector(self, *mask):
        """builder.selector<0,1,0>() (storage/builder.hpp): masked-out dimensions have extent 1 and no halo, e.g. the
        j-only coefficient fields of simple_hori_diff.cpp:66; the kernels ignore their strides."""
        return self._with(selector=tuple(bool(m) for m in mask))

    def alignment(self, a):
        """Not in the reference builder: lets tests build unaligned / unpadded layouts (alignment=1)."""
        return self._with(alignment=a)

    def build(self):
        kw = self._kw
        if "dtype" not in kw or "lengths" not in kw:
            raise ValueError("builder needs .type() and .dimensions() before .build()")  # static_assert in C++
        lengths = kw["lengths"]
        halos = kw.get("halos", (0,) * len(lengths))
        if kw.get("selector") is not None:
            lengths = tuple(n if m else 1 for n, m in zip(lengths, kw["selector"]))
            halos = tuple(h if m else 0 for h, m in zip(halos, kw["selector"]))
        ds = DataStore(kw["dtype"], lengths, halos,
                       alignment=kw.get("alignment", BYTE_ALIGNMENT), name=kw.get("name", ""))
        if kw.get("value") is not None:
            ds.host_view()[...] = kw["value"]
        elif kw.get("initializer") is not None:
            fn = kw["initializer"]
            d0, d1, d2 = lengths
            k, j, i = np.meshgrid(np.arange(d2), np.arange(d1), np.arange(d0), indexing="ij")
            ds.host_view()[...] = np.vectorize(fn)(i, j, k) if not getattr(fn, "vectorized", False) else fn(i, j, k)
        return ds


builder = _Builder()


def from_numpy(box, halos, alignment=BYTE_ALIGNMENT):
    """DataStore initialised from a dense [k, j, i] numpy box that includes the halo."""
    box = np.asarray(box)
    ds = DataStore(box.dtype, tuple(reversed(box.shape)), halos, alignment=alignment)
    ds.assign(box)
    return ds
