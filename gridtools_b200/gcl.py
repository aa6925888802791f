"""Host side of the halo exchange: a mirror of `gcl::halo_exchange_dynamic_ut` (gcl/halo_exchange.hpp:163-306) and of
the Cartesian process grid `MPI_3D_process_grid_t` (gcl/low_level/proc_grids_3D.hpp:34-245) for ranks that live on
one NVLink/NVSwitch box, one process per GPU.

Layers
  ProcGrid      pure host logic: coordinates, neighbour lookup with periodicity (proc_grids_3D.hpp:179-211)
  HaloPlan      pure host logic: which storage region goes to / comes from which neighbour, message sizes
                (common/halo_descriptor.hpp:90-201, gcl/high_level/empty_field_base.hpp:170-176)
  transports    "p2p"  : fused pack + NVLink store into the neighbour's receive buffer, flag handshake on the device
                         (libgtb200: gtb_halo_pack_send / wait / unpack) -- the product path on a B200 box
                "nccl" : pack -> torch.distributed send/recv of the staging buffers -> unpack (NCCL point-to-point;
                         also runs with gloo + an injected host codec, which is how the CPU tests exercise the
                         choreography without a GPU)
`torch.distributed` is plumbing only: it carries the 512-byte connection blobs and, for the "nccl" transport, the
packed messages.
"""
import ctypes as C
import itertools

import numpy as np

from . import _lib

DIRECTIONS = [e for e in itertools.product((-1, 0, 1), repeat=3)]  # (e2, e1, e0) order, see dir_index


def dir_index(e0, e1, e2):
    """n = (e0+1) + 3*(e1+1) + 9*(e2+1), e_d = offset along storage dimension d (include/gtb200.h)."""
    return (e0 + 1) + 3 * (e1 + 1) + 9 * (e2 + 1)


def dir_of(n):
    return (n % 3 - 1, (n // 3) % 3 - 1, n // 9 - 1)


class ProcGrid:
    """3-d Cartesian process grid, row-major ranks like MPI_Cart_create (proc_grids_3D.hpp:34-245)."""

    def __init__(self, dims, periodic, rank):
        self.dims = tuple(int(d) for d in dims)
        self.periodic = tuple(bool(p) for p in periodic)
        self.size = self.dims[0] * self.dims[1] * self.dims[2]
        if not 0 <= rank < self.size:
            raise ValueError("rank %d outside a %s process grid" % (rank, self.dims))
        self.rank = int(rank)
        self.coords = (rank // (self.dims[1] * self.dims[2]), (rank // self.dims[2]) % self.dims[1],
                       rank % self.dims[2])

    def proc(self, di, dj, dk):
        """Rank of the neighbour at offset (di, dj, dk) in process-grid coordinates, -1 if outside
        (proc_grids_3D.hpp:179-211)."""
        c = [self.coords[0] + di, self.coords[1] + dj, self.coords[2] + dk]
        for d in range(3):
            if self.periodic[d]:
                c[d] %= self.dims[d]
            elif c[d] < 0 or c[d] >= self.dims[d]:
                return -1
        return (c[0] * self.dims[1] + c[1]) * self.dims[2] + c[2]

    @staticmethod
    def dims_create(nranks, ndims=2):
        """Balanced factorisation like MPI_Dims_create; the trailing (3 - ndims) dimensions are 1.  J is split
        first (i is the contiguous axis: J faces are contiguous slabs, I faces are strided strips)."""
        dims = [1, 1, 1]
        n, f, factors = nranks, 2, []
        while n > 1:
            while n % f == 0:
                factors.append(f)
                n //= f
            f += 1
        for f in sorted(factors, reverse=True):
            d = min(range(ndims), key=lambda x: (dims[x], -x))
            dims[d] *= f
        dims[:ndims] = sorted(dims[:ndims])  # larger factor on the later (j) dimension
        return tuple(dims)


class HaloPlan:
    """Regions and message sizes of one rank.  `halos[d]` = (minus, plus, begin, end, total) for USER dimension d;
    `layout[d]` = position of user dimension d in increasing-stride order (0 = unit stride); `proc_layout[d]` =
    process-grid dimension user dimension d is distributed over."""

    def __init__(self, halos, grid: ProcGrid, layout=(0, 1, 2), proc_layout=(0, 1, 2)):
        if sorted(layout) != [0, 1, 2] or sorted(proc_layout) != [0, 1, 2]:
            raise ValueError("layout and proc_layout must be permutations of (0, 1, 2)")
        self.grid = grid
        self.layout, self.proc_layout = tuple(layout), tuple(proc_layout)
        self.halos_user = [tuple(int(x) for x in h) for h in halos]
        self.halos = [None, None, None]  # storage order
        for d in range(3):
            self.halos[layout[d]] = self.halos_user[d]
        for m, p, b, e, t in self.halos:
            if m < 0 or p < 0 or b < m or e < b or e + p >= t:
                raise ValueError("inconsistent halo descriptor (%d, %d, %d, %d, %d)" % (m, p, b, e, t))
        self.neighbour = [-1] * 27
        for n in range(27):
            if n == 13:
                continue
            es = dir_of(n)  # storage-order direction
            # process dimension I moves with the storage direction of user dimension proc_layout[I]: the reference's
            # nth<layout_transform<reversed data layout, proc layout>, I>(ii, jj, kk) (gcl/halo_exchange.hpp:169,
            # high_level/descriptors.hpp:497-499) -- checked against the reference itself for the two proc layouts
            # that are not their own inverse (tests/test_gcl_reference.py)
            off = [es[layout[proc_layout[i]]] for i in range(3)]
            self.neighbour[n] = grid.proc(*off)

    # common/halo_descriptor.hpp:90-201
    @staticmethod
    def _inside(h, e):
        m, p, b, en, _ = h
        return (en - m + 1 if e == 1 else b), (b + p - 1 if e == -1 else en)

    @staticmethod
    def _outside(h, e):
        m, p, b, en, _ = h
        if e == 0:
            return b, en
        return (en + 1, en + p) if e == 1 else (b - m, b - 1)

    def send_region(self, n):
        """[(lo, hi)] * 3 in storage order (inclusive) of what goes to neighbour n."""
        return [self._inside(self.halos[d], e) for d, e in enumerate(dir_of(n))]

    def recv_region(self, n):
        return [self._outside(self.halos[d], e) for d, e in enumerate(dir_of(n))]

    @staticmethod
    def _count(region):
        c = 1
        for lo, hi in region:
            c *= max(0, hi - lo + 1)
        return c

    def send_count(self, n):
        return 0 if n == 13 or self.neighbour[n] < 0 else self._count(self.send_region(n))

    def recv_count(self, n):
        return 0 if n == 13 or self.neighbour[n] < 0 else self._count(self.recv_region(n))

    def storage_shape(self):
        """numpy shape (slowest first) of a field."""
        return tuple(self.halos[d][4] for d in (2, 1, 0))

    # host codec used by the gloo tests (numpy, dimension 0 fastest == last numpy axis)
    def pack_numpy(self, n, fields):
        (a0, b0), (a1, b1), (a2, b2) = self.send_region(n)
        return np.concatenate([np.ascontiguousarray(f[a2:b2 + 1, a1:b1 + 1, a0:b0 + 1]).ravel() for f in fields])

    def unpack_numpy(self, n, fields, msg):
        (a0, b0), (a1, b1), (a2, b2) = self.recv_region(n)
        shape = (b2 - a2 + 1, b1 - a1 + 1, b0 - a0 + 1)
        cnt = shape[0] * shape[1] * shape[2]
        for i, f in enumerate(fields):
            f[a2:b2 + 1, a1:b1 + 1, a0:b0 + 1] = msg[i * cnt:(i + 1) * cnt].reshape(shape)


def message_tag(n):
    """Tag of the message SENT towards direction n; the receiver posts the matching receive for direction 26 - n
    with the same tag (the reference derives its MPI tag from the direction too, Halo_Exchange_3D.hpp:217-219)."""
    return n


class TorchComm:
    """torch.distributed as the out-of-band channel (blobs) and, for transport='nccl', the message transport."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def all_gather_bytes(self, data: bytes):
        out = [None] * self.size
        self.dist.all_gather_object(out, data, group=self.group)
        return out

    def exchange(self, sends, recvs):
        """sends / recvs: lists of (peer_rank, tag, tensor).  Posts all receives, then all sends, then waits --
        the order of Halo_Exchange_3D::exchange (:792-931).  Messages between the same pair are matched by posting
        order, so both lists must be sorted by tag on the two sides."""
        ops = [self.dist.P2POp(self.dist.irecv, t, peer, group=self.group) for peer, _, t in recvs]
        ops += [self.dist.P2POp(self.dist.isend, t, peer, group=self.group) for peer, _, t in sends]
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def barrier(self):
        self.dist.barrier(group=self.group)


class halo_exchange_dynamic_ut:
    """gcl::halo_exchange_dynamic_ut<DataLayout, ProcLayout, T, Arch> (gcl/halo_exchange.hpp:163-306).

        he = halo_exchange_dynamic_ut(periodicity, grid, dtype, layout=(0,1,2), proc_layout=(0,1,2), comm=...)
        he.add_halo(0, minus, plus, begin, end, total); he.add_halo(1, ...); he.add_halo(2, ...)
        he.setup(max_fields)
        he.pack(f0, f1); he.exchange(); he.unpack(f0, f1)

    Fields are DataStore objects (or anything with raw_ptr()) or raw device pointers (ints) to storage element
    (0,0,0) including the halo -- the `T*` of the reference.
    """

    def __init__(self, periodicity, grid: ProcGrid, dtype, layout=(0, 1, 2), proc_layout=(0, 1, 2), comm=None,
                 transport="p2p", codec=None):
        self.grid = grid
        self.dtype = np.dtype(dtype)
        self.layout, self.proc_layout = tuple(layout), tuple(proc_layout)
        per = [False] * 3
        for d in range(3):
            per[proc_layout[d]] = bool(periodicity[d])  # c.permute<layout2proc_map_abs>() (:202)
        if tuple(per) != grid.periodic:
            raise ValueError("periodicity %s does not match the process grid's %s" % (per, grid.periodic))
        if transport not in ("p2p", "nccl", "host"):
            raise ValueError("transport must be 'p2p', 'nccl' or 'host'")
        self.comm, self.transport, self.codec = comm, transport, codec
        self._halos = [None, None, None]
        self.plan = None
        self._h = None
        self._packed = None

    def comm_grid(self):
        """comm() of the reference (:306): the process grid."""
        return self.grid

    def add_halo(self, dim, minus, plus=None, begin=None, end=None, total=None):
        """add_halo<D>(minus, plus, begin, end, total) or add_halo<D>(halo_descriptor) (:235,:240)."""
        if plus is None:
            minus, plus, begin, end, total = minus
        self._halos[dim] = (minus, plus, begin, end, total)

    def setup(self, max_fields):
        """setup(max_fields) (:216): builds the plan, allocates buffers, connects the neighbours."""
        if any(h is None for h in self._halos):
            raise RuntimeError("setup() called before add_halo() for all three dimensions")
        self.max_fields = int(max_fields)
        self.plan = HaloPlan(self._halos, self.grid, self.layout, self.proc_layout)
        if self.transport == "host":
            return
        self._export()
        if self.transport == "p2p" and self.comm is not None:  # comm=None: several ranks in one process,
            blobs = self.comm.all_gather_bytes(self.blob)      # the caller finishes with connect_local()
            self._connect(blobs)

    def _export(self):
        L = _lib.lib()
        # elements wider than 8 bytes (the reference's own test exchanges array<int, 4>) travel as `f` consecutive
        # 8- (or 4-) byte words: the unit-stride dimension is scaled by f, the byte order of a message is unchanged
        es, f = self.dtype.itemsize, 1
        forced = getattr(self, "_word", None)  # halo_exchange_generic builds its messages from 4-byte words
        if es not in (4, 8) or (forced and forced != es):
            word = forced or (8 if es % 8 == 0 else 4)
            if es % word:
                raise ValueError("element size %d is not a multiple of 4 bytes" % es)
            es, f = word, es // word
        halos = list(self.plan.halos)
        m, p, b, e, t = halos[0]
        halos[0] = (m * f, p * f, b * f, e * f + f - 1, t * f)
        desc = (_lib.HaloDesc * 3)(*[_lib.HaloDesc(*h) for h in halos])
        nbr = (C.c_int * 27)(*self.plan.neighbour)
        h = C.c_void_p()
        _lib.check(L.gtb_halo_create(desc, nbr, self.grid.rank, self.max_fields, es, C.byref(h)))
        self._h = h
        buf = C.create_string_buffer(_lib.HALO_BLOB_BYTES)
        _lib.check(L.gtb_halo_export(self._h, buf))
        self.blob = buf.raw

    def _connect(self, blobs_by_rank):
        keep = [C.create_string_buffer(blobs_by_rank[r], _lib.HALO_BLOB_BYTES) if r >= 0 else None
                for r in self.plan.neighbour]
        arr = (C.c_void_p * 27)(*[C.cast(b, C.c_void_p) if b is not None else None for b in keep])
        _lib.check(_lib.lib().gtb_halo_connect(self._h, arr))

    # ------------------------------------------------------------------ the three phases
    def _ptrs(self, fields):
        if len(fields) == 1 and isinstance(fields[0], (list, tuple)):
            fields = fields[0]  # pack(vector<T*>) overload (:250)
        if len(fields) > self.max_fields:
            raise ValueError("%d fields passed, setup() was told at most %d" % (len(fields), self.max_fields))
        ptrs = [f.raw_ptr() if hasattr(f, "raw_ptr") else int(f) for f in fields]
        return (C.c_void_p * max(1, len(ptrs)))(*ptrs), len(ptrs)

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def pack(self, *fields):
        if self.transport == "host":
            f = fields[0] if len(fields) == 1 and isinstance(fields[0], (list, tuple)) else fields
            self._packed = {n: self.codec.pack(self.plan, n, f) for n in range(27) if self.plan.send_count(n)}
            self._nf = len(f)
            return
        arr, n = self._ptrs(fields)
        self._nf = n
        lost = C.c_int()  # a device-side wait of an earlier exchange that gave up (halo.timeout_ms): host memory read
        _lib.check(_lib.lib().gtb_halo_poll_error(self._h, C.byref(lost)))
        if lost.value:
            raise RuntimeError("halo exchange: the message from direction %d never arrived; the halos of that "
                               "exchange are stale" % (lost.value - 1))
        fn = _lib.lib().gtb_halo_pack_send if self.transport == "p2p" else _lib.lib().gtb_halo_pack
        _lib.check(fn(self._h, arr, n, self._stream()))

    def exchange(self):
        """exchange() = start_exchange() + wait() (:284-304)."""
        self.start_exchange()
        self.wait()

    def post_receives(self):
        pass  # receives are implicit: NVLink stores land in the exported arena / batch_isend_irecv posts them

    def do_sends(self):
        self.start_exchange()

    def start_exchange(self):
        if self.transport == "p2p":
            return  # pack() already pushed the messages and raised the neighbours' flags
        sends, recvs = [], []
        if self.transport == "nccl":
            import torch
            torch.cuda.current_stream().synchronize()  # the reference syncs the device before MPI too (:486)
        for n in range(27):
            peer = self.plan.neighbour[n]
            if n == 13 or peer < 0:
                continue
            if peer == self.grid.rank:  # periodic dimension of extent 1: the message stays on this rank
                if self.plan.recv_count(n):
                    self._recv_tensor(n).copy_(self._send_tensor(26 - n))
                continue
            if self.plan.send_count(n):
                sends.append((peer, message_tag(n), self._send_tensor(n)))
            if self.plan.recv_count(n):
                recvs.append((peer, message_tag(26 - n), self._recv_tensor(n)))
        sends.sort(key=lambda x: x[1])
        recvs.sort(key=lambda x: x[1])
        if sends or recvs:
            self.comm.exchange(sends, recvs)

    def wait(self):
        pass  # p2p: the arrival flags are acquired inside the unpack launch (gtb_halo_wait_unpack)

    def unpack(self, *fields):
        if self.transport == "host":
            f = fields[0] if len(fields) == 1 and isinstance(fields[0], (list, tuple)) else fields
            for n, msg in self._received.items():
                self.codec.unpack(self.plan, n, f, msg)
            return
        arr, n = self._ptrs(fields)
        fn = _lib.lib().gtb_halo_wait_unpack if self.transport == "p2p" else _lib.lib().gtb_halo_unpack
        _lib.check(fn(self._h, arr, n, self._stream()))
        _lib.check(_lib.lib().gtb_halo_next_epoch(self._h))

    def unpacked_flag(self):
        """Device address of the uint64 holding the epoch of the last completed unpack (gtb_halo_unpacked_flag)."""
        return _lib.lib().gtb_halo_unpacked_flag(self._h)

    def epoch(self):
        """Epoch number the next exchange of this object carries (1, 2, ...)."""
        return int(_lib.lib().gtb_halo_epoch(self._h))

    def set_boundary(self, value):
        """distributed_boundaries.hpp:141-200 with value_boundary and proc_grid_predicate: from now on every unpack /
        exchange of this object also writes `value` into the halo regions that face no neighbour, in the same launch.
        None switches it off."""
        if self._h is None:
            raise RuntimeError("set_boundary() needs setup() first (p2p / nccl transports)")
        _lib.check(_lib.lib().gtb_halo_set_boundary(self._h, -1 if value is None else _lib.GTB_BC_VALUE,
                                                    0.0 if value is None else float(value)))

    def bind(self, *fields):
        """Pre-marshals a field list: returns a callable f(stream_handle=None) that runs pack + exchange + unpack for
        these fields with ONE C call (gtb_halo_exchange, two launches) and no per-call Python marshalling -- for time
        loops whose step is a few tens of microseconds.  p2p transport only."""
        if self.transport != "p2p":
            raise ValueError("bind() needs the p2p transport")
        arr, n = self._ptrs(fields)
        fn, h, chk = _lib.lib().gtb_halo_exchange, self._h, _lib.check

        def run(stream_handle=None):
            chk(fn(h, arr, n, stream_handle if stream_handle is not None else self._stream()))
        run.keepalive = (arr, fields)
        return run

    def check(self):
        """0 if every wait so far completed; 1 + direction of a message that never arrived otherwise."""
        code = C.c_int()
        _lib.check(_lib.lib().gtb_halo_error(self._h, C.byref(code)))
        return code.value

    # ------------------------------------------------------------------ staging buffers as tensors
    def _send_tensor(self, n):
        if self.transport == "host":
            import torch
            return torch.from_numpy(np.ascontiguousarray(self._packed[n]))
        nbytes = _lib.lib().gtb_halo_send_bytes(self._h, n, self._nf)
        return _device_tensor(_lib.lib().gtb_halo_send_buffer(self._h, n), nbytes)

    def _recv_tensor(self, n):
        if self.transport == "host":
            import torch
            if not hasattr(self, "_received"):
                self._received = {}
            t = torch.empty(self.plan.recv_count(n) * self._nf,
                            dtype=torch.float64 if self.dtype.itemsize == 8 else torch.float32)
            self._received[n] = t.numpy()
            return t
        nbytes = _lib.lib().gtb_halo_recv_bytes(self._h, n, self._nf)
        return _device_tensor(_lib.lib().gtb_halo_recv_buffer(self._h, n), nbytes)

    def close(self):
        if self._h is not None:
            _lib.lib().gtb_halo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CudaBuffer:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _device_tensor(ptr, nbytes):
    import torch
    return torch.as_tensor(_CudaBuffer(ptr, nbytes), device="cuda")


class NumpyCodec:
    """Host pack/unpack for transport='host' (tests of the choreography on CPU with gloo)."""

    @staticmethod
    def pack(plan, n, fields):
        return plan.pack_numpy(n, fields)

    @staticmethod
    def unpack(plan, n, fields, msg):
        plan.unpack_numpy(n, fields, msg)


class field_on_the_fly:
    """gcl::field_on_the_fly<T, layout, traits>(ptr, halos) (gcl/high_level/field_on_the_fly.hpp:27-95): a field together
    with ITS OWN three halo descriptors (minus, plus, begin, end, total) in increasing-stride order."""

    def __init__(self, field, halos, dtype=None):
        self.field = field
        self.halos = tuple(tuple(int(x) for x in h) for h in halos)
        if len(self.halos) != 3 or any(len(h) != 5 for h in self.halos):
            raise ValueError("field_on_the_fly needs three (minus, plus, begin, end, total) descriptors")
        self.dtype = np.dtype(dtype if dtype is not None else getattr(field, "dtype", np.float64))

    def signature(self):
        return (self.halos, self.dtype.str)


class halo_exchange_generic:
    """gcl::halo_exchange_generic<ProcLayout, Arch> (gcl/halo_exchange.hpp:334-513): every field brings its own halo
    descriptors, so fields of different sizes, halo widths and element types travel in one pack / exchange / unpack.

        hg = halo_exchange_generic(periodicity, grid, comm=TorchComm())
        hg.setup(max_fields, halo_example, typesize)          # field_on_the_fly (or three descriptors), bytes
        hg.pack(field_on_the_fly(a, halos_a), field_on_the_fly(b, halos_b)); hg.exchange(); hg.unpack(...)

    Like the reference (descriptor_generic_manual.hpp:370-796) all fields of a pack() are concatenated into ONE message
    per neighbour: one launch packs every field for every neighbour and stores it in the neighbours' receive buffers,
    one launch waits for the arrival flags and unpacks (gtb_halo_generic_pack_send / _wait_unpack).  The buffers are
    sized by setup(): max_fields fields of the example's halo regions with `typesize`-byte elements
    (hndlr_generic::setup, gcl/halo_exchange.hpp:389-393).  Without an example the enclosing descriptor of the fields
    of the first prepare() / pack() is used -- a collective call like setup() itself."""

    def __init__(self, periodicity, grid: ProcGrid, comm=None, transport="p2p", proc_layout=(0, 1, 2), codec=None,
                 layout=(0, 1, 2)):
        self.periodicity, self.grid, self.comm, self.transport = tuple(periodicity), grid, comm, transport
        self.proc_layout, self.codec, self.layout = tuple(proc_layout), codec, tuple(layout)
        self.max_fields = None
        self._example = None
        self._he = None     # the device object (p2p) / None (host transport: one plan per field)
        self.pending = []   # [exchange object] until connect_local_generic connected it (in-process ranks)
        self._last = []

    def setup(self, max_fields, halo_example=None, typesize=8):
        self.max_fields = int(max_fields)
        if halo_example is not None:
            halos = halo_example.halos if isinstance(halo_example, field_on_the_fly) else halo_example
            self._example = (tuple(tuple(int(x) for x in h) for h in halos), int(typesize))

    @staticmethod
    def _words(fotf):
        es = fotf.dtype.itemsize
        if es % 4:
            raise ValueError("element size %d is not a multiple of 4 bytes" % es)
        return es // 4

    def _plan(self, fotf):
        return HaloPlan(fotf.halos, self.grid, self.layout, self.proc_layout)

    def _create(self, fields):
        if self.max_fields is None:
            raise RuntimeError("pack() called before setup()")
        if self._example is None:  # enclosing descriptor (the maximum per dimension, test_halo_exchange_3D.cpp:224-236)
            enc, typesize = [], max(f.dtype.itemsize for f in fields)
            for d in range(3):
                m = max(f.halos[d][0] for f in fields)
                p = max(f.halos[d][1] for f in fields)
                n = max(f.halos[d][3] - f.halos[d][2] + 1 for f in fields)
                enc.append((m, p, m, m + n - 1, m + n + p))
            self._example = (tuple(enc), typesize)
        halos, typesize = self._example
        he = halo_exchange_dynamic_ut(self.periodicity, self.grid, np.dtype("V%d" % typesize) if typesize not in (4, 8)
                                      else (np.float32 if typesize == 4 else np.float64), layout=self.layout,
                                      proc_layout=self.proc_layout, comm=self.comm, transport="p2p")
        for d in range(3):
            he.add_halo(d, *halos[d])
        he._word = 4  # generic messages are built from 4-byte words
        he.setup(self.max_fields)
        if self.comm is None:
            self.pending.append(he)
        self._he = he

    def prepare(self, *fields):
        """Creates the device object (collective).  Only needed when several ranks live in one process (comm=None):
        call it on every rank, then connect_local_generic([...])."""
        if self.transport != "host" and self._he is None:
            self._create(self._fields(fields))

    @staticmethod
    def _fields(fields):
        if len(fields) == 1 and isinstance(fields[0], (list, tuple)):
            fields = fields[0]  # pack(std::vector<field_on_the_fly>) overload (:421-424)
        return list(fields)

    def _marshal(self, fields):
        arr = (_lib.HaloField * max(1, len(fields)))()
        for a, f in zip(arr, fields):
            plan = self._plan(f)
            w = self._words(f)
            a.ptr = f.field.raw_ptr() if hasattr(f.field, "raw_ptr") else int(f.field)
            for d in range(3):
                m, p, b, e, t = plan.halos[d]
                k = w if d == 0 else 1
                a.desc[d] = _lib.HaloDesc(m * k, p * k, b * k, e * k + k - 1, t * k)
        return arr

    def pack(self, *fields):
        fields = self._fields(fields)
        if self.transport == "host":
            self._last = [(self._plan(f), f) for f in fields]
            self._packed = {}
            for i, (plan, f) in enumerate(self._last):
                for n in range(27):
                    if plan.send_count(n):
                        self._packed[(n, i)] = self.codec.pack(plan, n, [f.field])
            return
        if self._he is None:
            self._create(fields)
        arr = self._marshal(fields)
        _lib.check(_lib.lib().gtb_halo_generic_pack_send(self._he._h, arr, len(fields), self._he._stream()))

    def exchange(self):
        self.start_exchange()
        self.wait()

    def start_exchange(self):
        if self.transport != "host":
            return  # the messages left with pack()
        import torch
        sends, recvs, self._received = [], [], {}
        for i, (plan, f) in enumerate(self._last):
            for n in range(27):
                peer = plan.neighbour[n]
                if n == 13 or peer < 0:
                    continue
                if plan.recv_count(n):
                    t = torch.empty(plan.recv_count(n), dtype=torch.from_numpy(np.zeros(1, f.dtype)).dtype)
                    self._received[(n, i)] = t
                    if peer == self.grid.rank:
                        t.copy_(torch.from_numpy(np.ascontiguousarray(self._packed[(26 - n, i)])))
                    else:
                        recvs.append((peer, (26 - n) * 1000 + i, t))
                if plan.send_count(n) and peer != self.grid.rank:
                    sends.append((peer, n * 1000 + i, torch.from_numpy(np.ascontiguousarray(self._packed[(n, i)]))))
        sends.sort(key=lambda x: x[1])
        recvs.sort(key=lambda x: x[1])
        if sends or recvs:
            self.comm.exchange(sends, recvs)

    def wait(self):
        pass

    def unpack(self, *fields):
        fields = self._fields(fields)
        if self.transport == "host":
            for i, f in enumerate(fields):
                plan = self._plan(f)
                for n in range(27):
                    if plan.recv_count(n):
                        self.codec.unpack(plan, n, [f.field], self._received[(n, i)].numpy())
            return
        arr = self._marshal(fields)
        _lib.check(_lib.lib().gtb_halo_generic_wait_unpack(self._he._h, arr, len(fields), self._he._stream()))
        _lib.check(_lib.lib().gtb_halo_next_epoch(self._he._h))

    def check(self):
        return self._he.check() if self._he is not None else 0

    def close(self):
        if self._he is not None:
            self._he.close()
            self._he = None


def connect_local_generic(exchangers):
    """In-process ranks of halo_exchange_generic objects: connects the device objects created by prepare()."""
    if any(len(hg.pending) != 1 for hg in exchangers):
        raise RuntimeError("every rank must have called prepare() exactly once")
    connect_local([hg.pending[0] for hg in exchangers])
    for hg in exchangers:
        hg.pending = []


def connect_local(exchangers):
    """Several ranks inside ONE process (tests on a single GPU): export all, then connect all."""
    for he in exchangers:
        if any(h is None for h in he._halos):
            raise RuntimeError("add_halo() missing")
    blobs = {he.grid.rank: he.blob for he in exchangers}
    size = exchangers[0].grid.size
    table = [blobs.get(r, b"") for r in range(size)]
    for he in exchangers:
        he._connect(table)
