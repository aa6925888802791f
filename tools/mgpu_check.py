"""Multi-GPU halo exchange check, one process per GPU (torchrun).  Every interior cell carries its global coordinates
(regression/gcl/test_halo_exchange_3D.cpp:66-78); after pack -> exchange -> unpack every halo cell must hold the value
of the neighbouring rank's cell, or -1 at a non-periodic border (:106-123).  Runs both transports.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from gridtools_b200 import _lib, gcl, storage  # noqa: E402


def stamp_global(gi, gj, gk, fid):
    return fid * 1e7 + gi * 1e4 + gj * 10.0 + gk


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().gtb_init(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok_all = True
    ni, nj, nk, H = 48, 40, 6, 2
    for periodic in ((False, False, False), (True, True, False)):
        dims = gcl.ProcGrid.dims_create(world)
        grid = gcl.ProcGrid(dims, periodic, rank)
        for transport in ("p2p", "nccl"):
            n_fields = 3
            he = gcl.halo_exchange_dynamic_ut(periodic, grid, np.float64, comm=gcl.TorchComm(), transport=transport)
            fields = [storage.builder.type(np.float64).dimensions(ni + 2 * H, nj + 2 * H, nk).halos(H, H, 0).build()
                      for _ in range(n_fields)]
            p0 = fields[0].padded_lengths[0]
            he.add_halo(0, H, H, H, H + ni - 1, p0)
            he.add_halo(1, H, H, H, H + nj - 1, nj + 2 * H)
            he.add_halo(2, 0, 0, 0, nk - 1, nk)
            he.setup(n_fields)
            k, j, i = np.meshgrid(np.arange(nk), np.arange(nj), np.arange(ni), indexing="ij")
            for epoch in range(3):
                for f, ds in enumerate(fields):
                    v = ds.host_view()
                    v[...] = -1
                    v[:, H:H + nj, H:H + ni] = stamp_global(i + ni * grid.coords[0], j + nj * grid.coords[1], k,
                                                            f + 10 * epoch)
                he.pack(fields)
                he.exchange()
                he.unpack(fields)
            torch.cuda.synchronize()
            if transport == "p2p":
                assert he.check() == 0, "a halo wait timed out"
            # expected: global function with periodic wrap, -1 outside a non-periodic domain
            Gi, Gj = ni * dims[0], nj * dims[1]
            kk, jj, ii = np.meshgrid(np.arange(nk), np.arange(-H, nj + H), np.arange(-H, ni + H), indexing="ij")
            gi, gj = ii + ni * grid.coords[0], jj + nj * grid.coords[1]
            inside = np.ones(gi.shape, bool)
            if periodic[0]:
                gi = gi % Gi
            else:
                inside &= (gi >= 0) & (gi < Gi)
            if periodic[1]:
                gj = gj % Gj
            else:
                inside &= (gj >= 0) & (gj < Gj)
            for f, ds in enumerate(fields):
                want = np.where(inside, stamp_global(gi, gj, kk, f + 20), -1.0)
                got = ds.to_numpy()
                if not np.array_equal(got, want):
                    ok_all = False
                    bad = np.argwhere(got != want)
                    print("rank %d transport %s periodic %s field %d: %d mismatches, first %s got %r want %r" % (
                        rank, transport, periodic, f, len(bad), bad[0], got[tuple(bad[0])], want[tuple(bad[0])]),
                        flush=True)
            he.close()
            dist.barrier()
    flag = torch.tensor([0 if ok_all else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print("MGPU HALO CHECK", "PASSED" if flag.item() == 0 else "FAILED", "on", world, "GPUs, grid", dims, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
