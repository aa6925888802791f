"""Cost of the exchange/stencil choreography on ONE GPU: the rank is its own periodic neighbour (8 directions), so the
full pack -> peer-store -> flag -> wait -> unpack path runs without a second GPU.  Schemes:
  A  stencil only
  B  exchange(s+1) on a high-priority stream, events both ways (bench.py at N > 1)
  C  exchange(s) and stencil(s) on one stream (no overlap)
  D  pack_send(s+1) on the high-priority stream without any dependency, wait_unpack(s) on the compute stream
  G  scheme B captured into a CUDA graph per rotation (device-side epochs needed: not available -> skipped)
"""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from gridtools_b200 import _lib, gcl, stencil, storage

torch.cuda.set_device(0)
_lib.check(_lib.lib().gtb_init(0))
name = sys.argv[1] if len(sys.argv) > 1 else "hori_diff"
NI = NJ = 256
NK = 80
H = bench.HALO[name]
n_sets = 3 if name == "hori_diff" else 2
sets = []
for s in range(n_sets):
    if name == "vert_adv":
        arrs, dtr = bench.repo_vert_adv(NI, NJ, NK)
        sets.append([storage.from_numpy(a, (H, H, 0)) for a in arrs])
    else:
        inp, coeff = bench.repo_hori_diff(NI, NJ, NK)
        sets.append([storage.from_numpy(inp, (H, H, 0)), storage.from_numpy(coeff, (H, H, 0)),
                     storage.from_numpy(np.zeros_like(inp), (H, H, 0))])
for st in sets:
    for f in st:
        f.const_target_tensor()
per = (True, True, False) if "--j-only" not in sys.argv else (False, True, False)
grid = gcl.ProcGrid((1, 1, 1), per, 0)
he = gcl.halo_exchange_dynamic_ut(per, grid, np.float64, comm=None, transport="p2p")
f0 = sets[0][0]
p0, d1, d2 = f0.padded_lengths
he.add_halo(0, H, H, H, H + NI - 1, p0)
he.add_halo(1, H, H, H, H + NJ - 1, d1)
he.add_halo(2, 0, 0, 0, NK - 1, d2)
he.setup(1)
he._connect([he.blob])
xi = 2 if name == "vert_adv" else 0
L = _lib.lib()
comp = torch.cuda.current_stream()
comm = torch.cuda.Stream(priority=-1)
comp_h, comm_h = C.c_void_p(comp.cuda_stream), C.c_void_p(comm.cuda_stream)
if name == "vert_adv":
    plans = [stencil.plan("vertical_advection_dycore", *st, dtr_stage=dtr) for st in sets]
else:
    plans = [stencil.plan("horizontal_diffusion", *st) for st in sets]
exch = [he.bind(st[xi]) for st in sets]
ptrs = [he._ptrs([st[xi]]) for st in sets]


def pack_send(s, stream):
    arr, n = ptrs[s % n_sets]
    _lib.check(L.gtb_halo_pack_send(he._h, arr, n, stream))


def wait_unpack(s, stream):
    arr, n = ptrs[s % n_sets]
    _lib.check(L.gtb_halo_wait_unpack(he._h, arr, n, stream))
    _lib.check(L.gtb_halo_next_epoch(he._h))


STEPS = 300


def run(scheme):
    ev_x = [torch.cuda.Event() for _ in range(STEPS + 40)]
    ev_c = [torch.cuda.Event() for _ in range(STEPS + 40)]

    def step(s):
        if scheme == "A":
            plans[s % n_sets](comp_h)
        elif scheme == "B":
            def issue(t):
                if t - n_sets >= 0:
                    comm.wait_event(ev_c[t - n_sets])
                exch[t % n_sets](comm_h)
                ev_x[t].record(comm)
            if s == 0:
                issue(0)
            issue(s + 1)
            comp.wait_event(ev_x[s])
            plans[s % n_sets](comp_h)
            ev_c[s].record(comp)
        elif scheme == "C":
            exch[s % n_sets](comp_h)
            plans[s % n_sets](comp_h)
        elif scheme == "D":
            # epochs: pack_send(s) and wait_unpack(s) must use the same epoch -> keep them in one stream order here by
            # issuing pack_send(s) right before wait_unpack(s) but on the other stream (no overlap of pack with the
            # previous stencil is lost: the pack is issued before the stencil of step s-1 finishes on the device)
            pack_send(s, comm_h)
            ev_x[s].record(comm)
            comp.wait_event(ev_x[s])
            wait_unpack(s, comp_h)
            plans[s % n_sets](comp_h)
    if scheme == "G":  # no stream events: the stencil waits on the device for the unpacked epoch, the unpack for the
        # stencil-done counter (gtb_stencil_gate / gtb_halo_gate)
        seq = stencil.Sequence()
        done = torch.zeros(1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        flag, e0 = he.unpacked_flag(), he.epoch()
        ops = []

        def add_xg(t):
            if t - n_sets >= 0:
                seq.halo_gate(he, done.data_ptr(), t - n_sets + 1)
            seq.halo_exchange(he, [sets[t % n_sets][xi]], comm_h)
        for t in range(20 + STEPS):
            first = len(seq)
            if t == 0:
                add_xg(0)
            add_xg(t + 1)
            seq.stencil_gate(flag, e0 + t, done.data_ptr())
            if name == "vert_adv":
                seq.vertical_advection_dycore(*sets[t % n_sets], dtr, stream=comp_h)
            else:
                seq.horizontal_diffusion(*sets[t % n_sets], stream=comp_h)
            ops.append((first, len(seq) - first))
        seq.run(0, ops[20][0])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        seq.run(ops[20][0], len(seq) - ops[20][0])
        b.record()
        torch.cuda.synchronize()
        assert he.check() == 0
        assert int(done.item()) == 20 + STEPS, int(done.item())
        return a.elapsed_time(b) / STEPS * 1e3
    if scheme == "S":  # scheme B recorded as a gtb_seq and issued with one native call
        seq = stencil.Sequence()
        M = n_sets + 2
        ops = []

        def add_x(t):
            if t - n_sets >= 0:
                seq.wait(comm_h, M + (t - n_sets) % M)
            seq.halo_exchange(he, [sets[t % n_sets][xi]], comm_h)
            seq.record(t % M, comm_h)
        for t in range(20 + STEPS):
            first = len(seq)
            if t == 0:
                add_x(0)
            add_x(t + 1)
            seq.wait(comp_h, t % M)
            if name == "vert_adv":
                seq.vertical_advection_dycore(*sets[t % n_sets], dtr, stream=comp_h)
            else:
                seq.horizontal_diffusion(*sets[t % n_sets], stream=comp_h)
            seq.record(M + t % M, comp_h)
            ops.append((first, len(seq) - first))
        seq.run(0, ops[20][0])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        seq.run(ops[20][0], len(seq) - ops[20][0])
        b.record()
        torch.cuda.synchronize()
        assert he.check() == 0
        return a.elapsed_time(b) / STEPS * 1e3
    for s in range(20):
        step(s)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(20, 20 + STEPS):
        step(s)
    b.record()
    torch.cuda.synchronize()
    assert he.check() == 0
    return a.elapsed_time(b) / STEPS * 1e3


for rs in (4,):
    _lib.set_option("reserve_sms", rs)
    for scheme in ("A", "S", "G", "S", "G"):
        print("%s reserve_sms=%d scheme %s: %.2f us per step" % (name, rs, scheme, run(scheme)), flush=True)
