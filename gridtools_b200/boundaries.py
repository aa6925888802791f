"""Host-side mirror of the reference's boundary-condition module (boundaries/boundary.hpp:57-72) for the predefined
conditions zero_boundary / value_boundary<T> (zero.hpp, value.hpp:27-66) and copy_boundary (copy.hpp:26-46).

    from gridtools_b200 import boundaries as bd
    bd.boundary(halos, bd.value_boundary(3.5)).apply(a, b)            # every field = 3.5 on all 26 outside regions
    bd.boundary(halos, bd.copy_boundary(), predicate).apply(dst, src) # dst = src where predicate(direction) is true

`halos` are the three (minus, plus, begin, end, total) halo descriptors in increasing-stride order, `predicate` a
callable taking the direction (ei, ej, ek) with entries in {-1, 0, 1} (default_predicate: always true;
`proc_grid_predicate(grid)` = grid_predicate.hpp: true where the process grid has no neighbour).  One kernel launch
per apply() for all directions and fields (csrc/halo.cu, gtb_boundary_apply).
"""
import ctypes as C

from . import _lib


class value_boundary:
    def __init__(self, value=0.0):
        self.kind, self.value = _lib.GTB_BC_VALUE, float(value)


class zero_boundary(value_boundary):
    def __init__(self):
        super().__init__(0.0)


class copy_boundary:
    kind, value = _lib.GTB_BC_COPY, 0.0


def default_predicate(direction):
    return True


class proc_grid_predicate:
    """boundaries/grid_predicate.hpp:20-34: apply the condition only where the rank has no neighbour."""

    def __init__(self, grid):
        self.grid = grid

    def __call__(self, direction):
        return self.grid.proc(*direction) < 0


def direction_mask(predicate):
    return [0 if (e0, e1, e2) == (0, 0, 0) else int(bool(predicate((e0, e1, e2))))
            for e2 in (-1, 0, 1) for e1 in (-1, 0, 1) for e0 in (-1, 0, 1)]


class boundary:
    def __init__(self, halos, condition, predicate=default_predicate):
        self.desc = (_lib.HaloDesc * 3)(*[_lib.HaloDesc(*h) for h in halos])
        self.condition = condition
        self.mask = (C.c_int * 27)(*direction_mask(predicate))

    def apply(self, *fields, stream=None):
        """Fields are DataStores (or raw device pointers to storage element (0,0,0), halo included) with the layout the
        descriptors' total lengths describe; copy_boundary takes the source last."""
        if not fields:
            return
        dt = {f.dtype.itemsize for f in fields if hasattr(f, "dtype")}
        if len(dt) > 1:
            raise TypeError("boundary.apply: all fields must have the same element type")
        es = dt.pop() if dt else 8
        n = len(fields)
        src_const = self.condition.kind == _lib.GTB_BC_COPY
        ptrs = [(f.raw_ptr(const=src_const and i == n - 1) if hasattr(f, "raw_ptr") else int(f))
                for i, f in enumerate(fields)]
        arr = (C.c_void_p * n)(*ptrs)
        if stream is None:
            import torch
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().gtb_boundary_apply(self.desc, self.mask, self.condition.kind, self.condition.value, arr, n,
                                                 es, stream))
