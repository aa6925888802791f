#!/usr/bin/env python
"""COSMO-like dycore chain (BASELINE.json configs[4]): horizontal_diffusion -> vertical_advection_dycore ->
advection_pdbott_prepare_tracers (11 tracers) per step on an IJ-decomposed domain, halo exchanges of the two fields
with an IJ extent (`in`, H = 2; `wcon`, H = 3) overlapped with the stencils that do not touch them.

    python bench_chain.py [--ni 1024 --nj 1024 --nk 80] [--scaling weak|strong] [--steps K --warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_chain.py ...

One step on every rank (compute stream | communication stream, high priority):
    hori_diff(in, coeff -> out)            | -
    vert_adv(utens_stage, ..., wcon, ...)  | exchange(in)    (hori_diff of this step is done with it)
    prepare_tracers(11 x (rho * in))       | exchange(wcon)  (vert_adv of this step is done with it)
The whole loop is one recorded gtb_seq.  Prints one JSON line (rank 0): whole-job Mpts/s (grid points advanced
through the chain per second), the algorithmic bytes of a step and the achieved HBM bandwidth against the measured
peak.  Fields are the reference repositories' analytic fields evaluated on the device.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

N_TRACERS = 11  # advection_pdbott_prepare_tracers.cpp:44


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ni", type=int, default=1024)
    ap.add_argument("--nj", type=int, default=1024)
    ap.add_argument("--nk", type=int, default=80)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-overlap", action="store_true", help="exchanges on the compute stream (baseline)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from gridtools_b200 import _lib, gcl, stencil, storage

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().gtb_init(local))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = gcl.ProcGrid.dims_create(world)
    grid = gcl.ProcGrid(dims, (False, False, False), rank)
    ni, nj, nk = args.ni, args.nj, args.nk
    gi, gj = (ni, nj) if args.scaling == "strong" else (ni * dims[0], nj * dims[1])
    if args.scaling == "strong":
        ni, nj = ni // dims[0], nj // dims[1]
    i_off, j_off = grid.coords[0] * ni, grid.coords[1] * nj
    hd, _ = bench.device_fields(torch, storage, "hori_diff", ni, nj, nk, np.float64, i_off, j_off, gi, gj)
    va, dtr = bench.device_fields(torch, storage, "vert_adv", ni, nj, nk, np.float64, i_off, j_off, gi, gj)

    def plain(value):
        ds = storage.builder.type(np.float64).dimensions(ni, nj, nk).halos(0, 0, 0).build()
        ds.target_tensor().fill_(value)
        return ds
    tr_in = [plain(1.0 + 0.1 * t) for t in range(N_TRACERS)]
    tr_out = [plain(0.0) for _ in range(N_TRACERS)]
    rho = plain(1.1)
    torch.cuda.synchronize()

    def halo(field, H):
        he = gcl.halo_exchange_dynamic_ut((False, False, False), grid, np.float64, comm=gcl.TorchComm(), transport="p2p")
        p0, d1, d2 = field.padded_lengths
        he.add_halo(0, H, H, H, H + ni - 1, p0)
        he.add_halo(1, H, H, H, H + nj - 1, d1)
        he.add_halo(2, 0, 0, 0, nk - 1, d2)
        he.setup(1)
        return he
    he_in = halo(hd[0], 2) if world > 1 else None
    he_w = halo(va[2], 3) if world > 1 else None

    comp = torch.cuda.current_stream()
    comm = torch.cuda.Stream(priority=-1)
    comp_h = C.c_void_p(comp.cuda_stream)
    comm_h = comp_h if args.no_overlap else C.c_void_p(comm.cuda_stream)
    if world > 1 and not args.no_overlap:
        _lib.set_option("reserve_sms", bench.RESERVE_SMS["vert_adv"])
    seq = stencil.Sequence()
    total = args.warmup + args.steps
    ops = []
    # events: 0 in-halo ready, 1 wcon-halo ready, 2 hori_diff done, 3 vert_adv done
    for s in range(total):
        first = len(seq)
        if world > 1:
            if s == 0:
                seq.halo_exchange(he_in, [hd[0]], comm_h)
                seq.record(0, comm_h)
                seq.halo_exchange(he_w, [va[2]], comm_h)
                seq.record(1, comm_h)
            seq.wait(comp_h, 0)
        seq.horizontal_diffusion(*hd, stream=comp_h)
        seq.record(2, comp_h)
        if world > 1:
            seq.wait(comm_h, 2)
            seq.halo_exchange(he_in, [hd[0]], comm_h)  # for step s + 1, beside vert_adv
            seq.record(0, comm_h)
            seq.wait(comp_h, 1)
        seq.vertical_advection_dycore(*va, dtr, stream=comp_h)
        seq.record(3, comp_h)
        if world > 1:
            seq.wait(comm_h, 3)
            seq.halo_exchange(he_w, [va[2]], comm_h)  # for step s + 1, beside prepare_tracers
            seq.record(1, comm_h)
        seq.prepare_tracers(tr_out, tr_in, rho, stream=comp_h)
        ops.append((first, len(seq) - first))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seq.run(0, ops[args.warmup][0])
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    seq.run(ops[args.warmup][0], len(seq) - ops[args.warmup][0])
    e1.record()
    comm.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        for he in (he_in, he_w):
            if he.check() != 0:
                raise SystemExit("bench_chain.py: a halo wait timed out")
    ms_step = ms / args.steps
    pts = ni * nj * nk
    algo = (24 + 48 + 8 * (2 * N_TRACERS + 1)) * pts  # hd 24, va 48, tracers (2*11+1)*8 B per point
    peak, src = bench.measured_peak()
    line = {"metric": "Mpts/s dycore chain (hori_diff + vert_adv + prepare_tracers x%d) %dx%dx%d fp64" % (
                N_TRACERS, gi if args.scaling == "strong" else ni, gj if args.scaling == "strong" else nj, nk),
            "value": world * pts / (ms_step * 1e-3) / 1e6, "unit": "Mpts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "dtype": "f64", "data": "synthetic (reference repositories' analytic fields, evaluated on the device)",
            "config": {"workload": "chain %dx%dx%d per GPU, %dx%d process grid" % (ni, nj, nk, dims[0], dims[1]),
                       "overlap": "none (exchanges on the compute stream)" if args.no_overlap or world == 1 else
                       "exchange(in) beside vert_adv, exchange(wcon) beside prepare_tracers, %d SMs reserved" %
                       bench.RESERVE_SMS["vert_adv"]},
            "roofline": {"bound": "hbm", "achieved": algo / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (ms_step * 1e-3) / 1e9 / peak, "peak_source": src,
                         "algorithmic_bytes_per_step": algo},
            "gpu_launches": int(_lib.launch_count() - launches0)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
