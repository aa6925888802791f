"""Host logic of the halo exchange on CPU: process grid and plan against the oracle's restatement of
proc_grids_3D.hpp / halo_descriptor.hpp, and the pack -> exchange -> unpack choreography with world_size 2 and 4
over gloo (the pattern of regression/gcl/test_halo_exchange_3D.cpp:66-123: every cell carries its global
coordinates; after the exchange every halo cell holds the neighbour's value or stays -1 at non-periodic borders)."""
import ctypes as C
import itertools
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridtools_b200.gcl import (HaloPlan, NumpyCodec, ProcGrid, TorchComm, dir_of, field_on_the_fly,
                                halo_exchange_dynamic_ut, halo_exchange_generic)


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 4, 1), (3, 2, 2)])
@pytest.mark.parametrize("periodic", [(0, 0, 0), (1, 0, 1), (1, 1, 1)])
def test_proc_grid_matches_oracle(oracle, dims, periodic):
    lib = oracle.lib()
    cd, cp = (C.c_int * 3)(*dims), (C.c_int * 3)(*periodic)
    size = dims[0] * dims[1] * dims[2]
    for rank in range(size):
        g = ProcGrid(dims, periodic, rank)
        assert g.proc(0, 0, 0) == rank
        for d in itertools.product((-1, 0, 1), repeat=3):
            assert g.proc(*d) == lib.gto_proc_neighbour(cd, cp, *g.coords, *d), (rank, d)


def test_dims_create():
    assert ProcGrid.dims_create(1) == (1, 1, 1)
    assert ProcGrid.dims_create(2) == (1, 2, 1)
    assert ProcGrid.dims_create(4) == (2, 2, 1)
    assert ProcGrid.dims_create(8) == (2, 4, 1)
    assert ProcGrid.dims_create(8, 3) == (2, 2, 2)


HALOS = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (0, 1, 0, 4, 6)]  # minus != plus, k has only a plus halo


def test_plan_counts_match_oracle(oracle):
    lib = oracle.lib()
    plan = HaloPlan(HALOS, ProcGrid((1, 1, 1), (1, 1, 1), 0))
    h = oracle.halos3(HALOS)
    for n in range(27):
        if n == 13:
            continue
        e = dir_of(n)
        assert plan.send_count(n) == lib.gto_halo_send_count(h, *e)
        assert plan.recv_count(n) == lib.gto_halo_recv_count(h, *e)


def test_plan_rejects_bad_descriptor():
    with pytest.raises(ValueError):
        HaloPlan([(2, 2, 1, 9, 14)] * 3, ProcGrid((1, 1, 1), (0, 0, 0), 0))  # begin < minus
    with pytest.raises(ValueError):
        HaloPlan(HALOS, ProcGrid((1, 1, 1), (0, 0, 0), 0), layout=(0, 0, 1))


def stamp(plan, grid, field_id):
    """Global-coordinate stamp on the interior, -1 in the halo (test_halo_exchange_3D.cpp:66-78)."""
    shape = plan.storage_shape()
    a = -np.ones(shape)
    (m0, p0, b0, e0, t0), (m1, p1, b1, e1, t1), (m2, p2, b2, e2, t2) = plan.halos
    n0, n1, n2 = e0 - b0 + 1, e1 - b1 + 1, e2 - b2 + 1
    k, j, i = np.meshgrid(np.arange(n2), np.arange(n1), np.arange(n0), indexing="ij")
    gi, gj, gk = i + n0 * grid.coords[0], j + n1 * grid.coords[1], k + n2 * grid.coords[2]
    a[b2:e2 + 1, b1:e1 + 1, b0:e0 + 1] = field_id * 1e6 + gi * 1e4 + gj * 1e2 + gk
    return a


def run_oracle_exchange(oracle, dims, periodic, n_fields):
    size = dims[0] * dims[1] * dims[2]
    fields = []
    for r in range(size):
        g = ProcGrid(dims, periodic, r)
        plan = HaloPlan(HALOS, g)
        fields.append([stamp(plan, g, f) for f in range(n_fields)])
    oracle.halo_exchange_all(HALOS, dims, periodic, fields, 8)
    return fields


def _worker(rank, size, port, dims, periodic, n_fields, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        grid = ProcGrid(dims, periodic, rank)
        he = halo_exchange_dynamic_ut(periodic, grid, np.float64, comm=TorchComm(), transport="host",
                                      codec=NumpyCodec)
        for d in range(3):
            he.add_halo(d, *HALOS[d])
        he.setup(n_fields)
        fields = [stamp(he.plan, grid, f) for f in range(n_fields)]
        he.pack(fields)
        he.exchange()
        he.unpack(fields)
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.stack(fields))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("dims,periodic", [((2, 1, 1), (0, 0, 0)), ((1, 2, 1), (1, 1, 0)), ((2, 2, 1), (1, 0, 0)),
                                           ((2, 2, 1), (1, 1, 1))])
def test_gloo_exchange_matches_oracle(oracle, tmp_path, dims, periodic):
    size = dims[0] * dims[1] * dims[2]
    n_fields = 3
    expect = run_oracle_exchange(oracle, dims, periodic, n_fields)
    mp.spawn(_worker, args=(size, _free_port(), dims, periodic, n_fields, str(tmp_path)), nprocs=size, join=True)
    for r in range(size):
        got = np.load(tmp_path / ("rank%d.npy" % r))
        assert np.array_equal(got, np.stack(expect[r])), "rank %d differs from the oracle" % r
    if not any(periodic):
        assert (np.stack(expect[0]) == -1).any()  # non-periodic borders keep their -1 (test_halo_exchange_3D.cpp:106-123)


# ------------------------------------------------------------------------------------------- halo_exchange_generic
HALOS_B = [(1, 1, 1, 7, 9), (2, 2, 2, 9, 12), (0, 0, 0, 2, 3)]  # a second field shape with its own halos


def _stamp_h(halos, grid, field_id):
    return stamp(HaloPlan(halos, grid), grid, field_id)


def _generic_worker(rank, size, port, dims, periodic, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        grid = ProcGrid(dims, periodic, rank)
        hg = halo_exchange_generic(periodic, grid, comm=TorchComm(), transport="host", codec=NumpyCodec)
        hg.setup(4)
        a = [_stamp_h(HALOS, grid, f) for f in range(2)]
        b = [_stamp_h(HALOS_B, grid, 7)]
        fields = [field_on_the_fly(a[0], HALOS), field_on_the_fly(b[0], HALOS_B), field_on_the_fly(a[1], HALOS)]
        hg.pack(*fields)
        hg.exchange()
        hg.unpack(*fields)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), a=np.stack(a), b=np.stack(b))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims,periodic", [((2, 1, 1), (1, 0, 0)), ((2, 2, 1), (1, 1, 0))])
def test_gloo_generic_exchange_matches_oracle(oracle, tmp_path, dims, periodic):
    """Fields with different sizes and halos in one pack / exchange / unpack: each group must equal the oracle's
    exchange of that group alone."""
    size = dims[0] * dims[1] * dims[2]
    expect_a, expect_b = [], []
    for r in range(size):
        g = ProcGrid(dims, periodic, r)
        expect_a.append([_stamp_h(HALOS, g, f) for f in range(2)])
        expect_b.append([_stamp_h(HALOS_B, g, 7)])
    oracle.halo_exchange_all(HALOS, dims, periodic, expect_a, 8)
    oracle.halo_exchange_all(HALOS_B, dims, periodic, expect_b, 8)
    mp.spawn(_generic_worker, args=(size, _free_port(), dims, periodic, str(tmp_path)), nprocs=size, join=True)
    for r in range(size):
        got = np.load(tmp_path / ("rank%d.npz" % r))
        assert np.array_equal(got["a"], np.stack(expect_a[r])) and np.array_equal(got["b"], np.stack(expect_b[r])), r
