#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--stencil vert_adv|hori_diff]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): vertical_advection_dycore 256x256x80 fp64 per GPU (BASELINE.json configs[1]), the
reference's own perftest (tests/regression/vertical_advection_dycore.cpp:128-152, `perftests 256 256 80`).
A step = one application of the stencil to one set of synthetic fields (the reference repository's analytic
fields); the 5 input fields of a set are 228 MB > L2 (126 MB) and the sets are rotated, so no step finds its inputs
in L2.  At N > 1 the global domain is decomposed in IJ (weak scaling: 256x256x80 per GPU) and every step first
exchanges the halo of wcon with the neighbours (gcl pack -> NVLink stores -> unpack), which is the exchange step a
real dycore has between two stencils.

One JSON line on stdout (rank 0).  `value` = whole-job Mpts/s with inputs resident in HBM; `e2e` = the same metric
through the public API with HOST buffers (pinned host -> device copies of the step's five input fields and the
device -> host copy of the result inside the timed region); `roofline` = algorithmic bytes (48 B/point, SURVEY.md
section 8d) / CUDA-event kernel time against the measured HBM peak; `cpu_baseline` = the reference's own cpu_ifirst /
cpu_kfirst backends (oracle/_ref/libgtref.so, built from the unmodified reference headers) timed on this host.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NI = NJ = 256
NK = 80
ALGO_BYTES = {"vert_adv": 48, "hori_diff": 24}  # per interior point, fp64
P100_MPTS = {"vert_adv": 5117.0, "hori_diff": 10300.0}  # reference stencil::gpu on P100
HALO = {"vert_adv": 3, "hori_diff": 2}
HALO_FUSED = 0   # N > 1: gtb_halo_exchange as one launch (pack, signal, wait, unpack)
RESERVE_SMS = {"vert_adv": 6, "hori_diff": 8}  # N > 1: SMs the persistent stencil grids leave to the concurrent exchange kernels
# (profiles/r02_exchange_variants.txt: an SM pushes only ~10 GB/s over NVLink, the exchange must end inside a step)
EXCHANGES_IN_FLIGHT = 1  # N > 1: pattern objects (and streams) that take the per-step exchanges in turn
HALO_DMA = 0     # N > 1: 1 = the NVLink leg of the exchange on the copy engines (option halo.dma; measured slower:
                 # eight in-stream peer copies cost ~6.5 us each, profiles/r02_exchange_timeline.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stencil", default="vert_adv", choices=["vert_adv", "hori_diff"])
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / secondary stencil")
    ap.add_argument("--ni", type=int, default=256, help="interior size in i (per GPU for weak, global for strong scaling)")
    ap.add_argument("--nj", type=int, default=256)
    ap.add_argument("--nk", type=int, default=80)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    a = ap.parse_args()
    global NI, NJ, NK
    NI, NJ, NK = a.ni, a.nj, a.nk
    return a


# --------------------------------------------------------------------------------------- synthetic fields
def repo_vert_adv(ni, nj, nk, i_off=0, j_off=0, gi=None, gj=None):
    """vertical_advection_repository.hpp:70-93 on the (nk, nj+6, ni+6) box; global indices for multi-GPU like
    copy_stencil_parallel.cpp:102-107."""
    d0, d1 = (gi or ni) + 6, (gj or nj) + 6
    k, j, i = np.meshgrid(np.arange(nk), np.arange(nj + 6) + j_off, np.arange(ni + 6) + i_off, indexing="ij",
                          sparse=True)
    x, y, z = i / d0, j / d1, k / nk
    pi = np.pi
    t = x + y + 0 * z
    u_stage = 7.0 + np.cos(pi * t) + np.sin(2 * pi * t) + 0 * z
    u_pos = u_stage.copy()
    wcon = 2e-4 * (-1.07 + (2. + np.cos(pi * (x + z)) + np.cos(pi * y)) / 2.)
    utens = 3e-6 * (-1.0235 + (2. + np.cos(pi * (x + y)) + np.cos(pi * y * z)) / 2.)
    utens_stage = 7.0 + 1.25 * (2. + np.cos(pi * (x + y)) + np.sin(2 * pi * (x + y))) + 0.1 * k
    shape = (nk, nj + 6, ni + 6)
    return [np.ascontiguousarray(np.broadcast_to(a, shape)) for a in (utens_stage, u_stage, wcon, u_pos, utens)], 3. / 20.


def repo_hori_diff(ni, nj, nk, i_off=0, j_off=0, gi=None, gj=None):
    """horizontal_diffusion_repository.hpp:32-42 on the (nk, nj+4, ni+4) box."""
    d0, d1 = (gi or ni) + 4, (gj or nj) + 4
    k, j, i = np.meshgrid(np.arange(nk), np.arange(nj + 4) + j_off, np.arange(ni + 4) + i_off, indexing="ij",
                          sparse=True)
    x, y = i / d0, j / d1
    inp = 5. + 8 * (2. + np.cos(np.pi * (x + 1.5 * y)) + np.sin(2 * np.pi * (x + 1.5 * y))) / 4. + 0. * k
    shape = (nk, nj + 4, ni + 4)
    return np.ascontiguousarray(np.broadcast_to(inp, shape)), np.full(shape, 0.025)


def device_fields(torch, storage, name, ni, nj, nk, dtype, i_off=0, j_off=0, gi=None, gj=None):
    """The same analytic fields evaluated on the device, straight into the stores' target tensors: a 4096x4096x80
    field is 10.7 GB, which neither a numpy temporary nor a pinned host mirror should have to hold."""
    H = HALO[name]
    d0, d1 = (gi or ni) + 2 * H, (gj or nj) + 2 * H
    tdt = torch.float64
    i = (torch.arange(ni + 2 * H, device="cuda", dtype=tdt) + i_off).view(1, 1, -1)
    j = (torch.arange(nj + 2 * H, device="cuda", dtype=tdt) + j_off).view(1, -1, 1)
    k = torch.arange(nk, device="cuda", dtype=tdt).view(-1, 1, 1)
    x, y, z = i / d0, j / d1, k / nk
    pi = np.pi

    def store(expr):
        ds = storage.builder.type(dtype).dimensions(ni + 2 * H, nj + 2 * H, nk).halos(H, H, 0).build()
        t = ds.target_tensor()
        t[:, :, :ni + 2 * H] = (expr + 0 * (x + y + z)).to(t.dtype)
        return ds

    if name == "hori_diff":
        inp = 5. + 8 * (2. + torch.cos(pi * (x + 1.5 * y)) + torch.sin(2 * pi * (x + 1.5 * y))) / 4.
        return [store(inp), store(torch.full((1, 1, 1), 0.025, device="cuda", dtype=tdt)),
                store(torch.zeros((1, 1, 1), device="cuda", dtype=tdt))], None
    t = x + y
    u = 7.0 + torch.cos(pi * t) + torch.sin(2 * pi * t)
    wcon = 2e-4 * (-1.07 + (2. + torch.cos(pi * (x + z)) + torch.cos(pi * y)) / 2.)
    utens = 3e-6 * (-1.0235 + (2. + torch.cos(pi * (x + y)) + torch.cos(pi * y * z)) / 2.)
    us = 7.0 + 1.25 * (2. + torch.cos(pi * (x + y)) + torch.sin(2 * pi * (x + y))) + 0.1 * k
    return [store(us), store(u), store(wcon), store(u), store(utens)], 3. / 20.


# --------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period = period
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------- reference arm (CPU)
def time_reference(stencil, steps, warmup):
    """The UNMODIFIED reference (its headers compiled into oracle/_ref/libgtref.so by oracle/Makefile) running its own
    cpu_ifirst / cpu_kfirst OpenMP backends on this host's cores.  Returns (best_backend, seconds per step list)."""
    from oracle import pyoracle as o
    if not o.have_ref():
        raise RuntimeError("oracle/_ref/libgtref.so missing (built in the build container by __graft_entry__.build)")
    try:  # torchrun exports OMP_NUM_THREADS=1: give the reference all the host cores it can use
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count() or 1)
    except OSError:
        pass
    best = None
    for backend in ("cpu_ifirst", "cpu_kfirst"):
        if stencil == "vert_adv":
            arrs, dtr = repo_vert_adv(NI, NJ, NK)
            res = np.zeros_like(arrs[0])
            o.ref_run(o.VERT_ADV, backend, arrs, [res], NI, NJ, NK, scalar=dtr, nrep=warmup, flush=True)
            times = o.ref_run(o.VERT_ADV, backend, arrs, [res], NI, NJ, NK, scalar=dtr, nrep=steps, flush=True)
        else:
            inp, coeff = repo_hori_diff(NI, NJ, NK)
            res = np.zeros_like(inp)
            o.ref_run(o.HORI_DIFF, backend, [inp, coeff], [res], NI, NJ, NK, nrep=warmup, flush=True)
            times = o.ref_run(o.HORI_DIFF, backend, [inp, coeff], [res], NI, NJ, NK, nrep=steps, flush=True)
        if best is None or statistics.median(times) < statistics.median(best[1]):
            best = (backend, times, res)
    return best[0], best[1], o.ref_num_threads(), best[2]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 50)  # bounded sample: each step is one full 256x256x80 application (~5-15 ms)
    backend, times, cores, _ = time_reference(args.stencil, steps, min(args.warmup, 5))
    sec = statistics.median(times)
    mpts = NI * NJ * NK / sec / 1e6
    sample = "%d timed applications of %s %dx%dx%d fp64 on stencil::%s, cache flush before each, median" % (
        steps, args.stencil, NI, NJ, NK, backend)
    line = {
        "impl": "reference", "metric": "Mpts/s %s %dx%dx%d fp64" % (args.stencil, NI, NJ, NK), "value": mpts,
        "unit": "Mpts/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 5),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (reference repository's analytic fields)",
        "config": {"workload": "%s %dx%dx%d fp64, reference CPU backend %s" % (args.stencil, NI, NJ, NK, backend)},
        "cpu_baseline": {"value": mpts, "unit": "Mpts/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": mpts, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- B200 arm
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for key in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
                if key in d:
                    return float(d[key]), "measured (MEASURED_PEAKS.json %s)" % key
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def traffic_from_profile(stencil):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(stencil)
        except Exception:
            return None
    return None


def b200_arm(args):
    import torch
    import torch.distributed as dist
    from gridtools_b200 import _lib, gcl, stencil, storage

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().gtb_init(local))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout; stdout carries the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- fields: two rotating sets per rank
    global NI, NJ
    dims = gcl.ProcGrid.dims_create(world)
    # GTB_PERIODIC=1 (diagnosis): periodic process grid, every rank has all 8 IJ neighbours (itself where a dimension
    # has one rank) -- the exchange load of an interior rank of a large grid on as few as 2 GPUs
    periodic = (True, True, False) if os.environ.get("GTB_PERIODIC") == "1" else (False, False, False)
    grid = gcl.ProcGrid(dims, periodic, rank)
    name = args.stencil
    H = HALO[name]
    np_dtype = np.float64 if args.dtype == "f64" else np.float32
    itemsize = np.dtype(np_dtype).itemsize
    global_ni, global_nj = (NI, NJ) if args.scaling == "strong" else (NI * dims[0], NJ * dims[1])
    if args.scaling == "strong":  # the global domain is divided among the ranks
        if NI % dims[0] or NJ % dims[1]:
            raise SystemExit("bench.py: --scaling strong needs sizes divisible by the %dx%d process grid" % dims[:2])
        NI, NJ = NI // dims[0], NJ // dims[1]
    i_off, j_off = grid.coords[0] * NI, grid.coords[1] * NJ
    n_fields = 5 if name == "vert_adv" else 3
    set_bytes = n_fields * (NI + 2 * H + 16) * (NJ + 2 * H) * NK * itemsize
    # enough rotating sets that no step finds its inputs in the 126 MB L2; one set when a set alone is far larger
    n_sets = 1 if set_bytes > (1 << 30) else 4  # >= 4 independent field sets
    n_sets = int(os.environ.get("GTB_NSETS", n_sets))
    on_device = NI * NJ * NK > 32 * 2**20
    sets = []
    dtr = None
    for s in range(n_sets):
        if on_device:
            st, dtr_ = device_fields(torch, storage, name, NI, NJ, NK, np_dtype, i_off, j_off, global_ni, global_nj)
            dtr = dtr_ if dtr_ is not None else dtr
            sets.append(st)
        elif name == "vert_adv":
            arrs, dtr = repo_vert_adv(NI, NJ, NK, i_off, j_off, global_ni, global_nj)
            sets.append([storage.from_numpy(a.astype(np_dtype), (H, H, 0)) for a in arrs])
        else:
            inp, coeff = repo_hori_diff(NI, NJ, NK, i_off, j_off, global_ni, global_nj)
            sets.append([storage.from_numpy(inp.astype(np_dtype), (H, H, 0)),
                         storage.from_numpy(coeff.astype(np_dtype), (H, H, 0)),
                         storage.from_numpy(np.zeros_like(inp, dtype=np_dtype), (H, H, 0))])
    for st in sets:
        for f in st:
            f.const_target_tensor()  # upload now

    # ---- halo exchange object (N > 1): the field with an IJ extent (wcon resp. in), one per set
    he = None
    if world > 1:
        he = gcl.halo_exchange_dynamic_ut(periodic, grid, np_dtype, comm=gcl.TorchComm(), transport="p2p")
        f0 = sets[0][0]
        p0, d1, d2 = f0.padded_lengths
        he.add_halo(0, H, H, H, H + NI - 1, p0)
        he.add_halo(1, H, H, H, H + NJ - 1, d1)
        he.add_halo(2, 0, 0, 0, NK - 1, d2)
        he.setup(1)
    # Exchanges in flight (N > 1): an exchange is a latency chain (pack, NVLink, the neighbours' flags, unpack) of about
    # the length of a hori_diff step, and the exchanges of ONE pattern object are ordered.  PIPE pattern objects (own
    # receive arenas and flags) on PIPE streams take the steps in turn, so the pack of step s+2 does not wait for the
    # flags of step s+1 -- what an application does that exchanges several fields through several patterns.
    pipe = int(os.environ.get("GTB_PIPE", EXCHANGES_IN_FLIGHT)) if he is not None else 1
    if os.environ.get("GTB_TIMELINE") == "1" or os.environ.get("GTB_GATES", "0") == "1":
        pipe = 1
    hes = [he]
    for _ in range(1, pipe):
        h2 = gcl.halo_exchange_dynamic_ut(periodic, grid, np_dtype, comm=gcl.TorchComm(), transport="p2p")
        h2.add_halo(0, H, H, H, H + NI - 1, p0)
        h2.add_halo(1, H, H, H, H + NJ - 1, d1)
        h2.add_halo(2, 0, 0, 0, NK - 1, d2)
        h2.setup(1)
        hes.append(h2)

    exch_index = 2 if name == "vert_adv" else 0

    # Pre-marshalled calls (stencil.plan / halo_exchange.bind): a step is 27-64 us of device time, per-call ctypes
    # marshalling would bound it from the host.
    comp = torch.cuda.current_stream()
    comm = torch.cuda.Stream(priority=-1) if he is not None else None
    comp_h = C.c_void_p(comp.cuda_stream)
    comm_h = C.c_void_p(comm.cuda_stream) if comm is not None else None
    comms = [comm] + [torch.cuda.Stream(priority=-1) for _ in range(1, pipe)]
    comm_hs = [C.c_void_p(c.cuda_stream) if c is not None else None for c in comms]
    if name == "vert_adv":
        stencil_plans = [stencil.plan("vertical_advection_dycore", *st, dtr_stage=dtr) for st in sets]
    else:
        stencil_plans = [stencil.plan("horizontal_diffusion", *st) for st in sets]
    exch_plans = [he.bind(st[exch_index]) for st in sets] if he is not None else None

    def run_stencil(s):
        stencil_plans[s % n_sets](comp_h)

    # N > 1: the halo exchange of step s+1 (comm stream, high priority) overlaps the stencil of step s (compute
    # stream).  exchange(s+1) touches field set (s+1) % n_sets, last read by stencil(s+1-n_sets): event dependency.
    # The whole loop is recorded once as a gtb_seq (include/gtb200.h) and issued natively: three launches and four
    # stream/event operations per 27-60 us step are more than per-call ctypes marshalling leaves room for.
    # Timed region at N > 1: the ranks leave a host barrier hundreds of microseconds apart, and a rank that starts early
    # waits for its neighbours' first halos, so timing from the first launch after the barrier charges that start skew
    # to the steps (round 1: 9-12 us per step at --steps 20).  The K timed steps are therefore issued in the SAME
    # native call as LEAD un-timed lead-in steps; the start mark is recorded in the compute stream after the last
    # lead-in stencil, when the exchange has put the ranks in lock-step, the end mark after the last timed stencil.
    LEAD = 3 if he is not None else 0
    n_warm = max(args.warmup, 3)
    total_steps = n_warm + LEAD + args.steps
    seq, step_ops = None, []
    gated = False
    first_timed_op = [0]

    def build_sequence(use_gates):
        """The loop of total_steps steps as one gtb_seq.  use_gates: the two orderings between the streams (stencil
        after unpack, unpack after the stencil that last read the halos) are waits ON THE DEVICE (gtb_stencil_gate /
        gtb_halo_gate) instead of stream events (GTB_GATES=1)."""
        sq, ops = stencil.Sequence(), []
        M = n_sets + 2  # event slots: exchange done = s % M, stencil done = M + s % M
        done = torch.zeros(1, dtype=torch.int64, device="cuda") if use_gates else None
        flag, e0 = (he.unpacked_flag(), he.epoch()) if use_gates else (None, 0)
        torch.cuda.synchronize()

        nodeps = os.environ.get("GTB_NODEPS") == "1"  # diagnosis ONLY: no ordering between the streams (invalid run)

        def add_exchange(t):
            if t - n_sets >= 0 and not nodeps:
                if use_gates:
                    sq.halo_gate(he, done.data_ptr(), t - n_sets + 1)
                else:
                    sq.wait(comm_hs[t % pipe], M + (t - n_sets) % M)
            sq.halo_exchange(hes[t % pipe], [sets[t % n_sets][exch_index]], comm_hs[t % pipe])
            if not use_gates and not nodeps:
                sq.record(t % M, comm_hs[t % pipe])

        for s in range(total_steps):
            first = len(sq)
            if s == n_warm + LEAD:
                sq.mark(0, comp_h)
                first_timed_op[0] = first
            if s == 0:
                add_exchange(0)
            if s + 1 < total_steps:
                add_exchange(s + 1)
            if use_gates:
                sq.stencil_gate(flag, e0 + s, done.data_ptr())
            elif not nodeps:
                sq.wait(comp_h, s % M)
            if timeline is not None and s < timeline.shape[0]:
                sq.stamp(timeline[s, 0:1].data_ptr(), comp_h)
            if name == "vert_adv":
                sq.vertical_advection_dycore(*sets[s % n_sets], dtr, stream=comp_h)
            else:
                sq.horizontal_diffusion(*sets[s % n_sets], stream=comp_h)
            if timeline is not None and s < timeline.shape[0]:
                sq.stamp(timeline[s, 1:2].data_ptr(), comp_h)
            if not use_gates and not nodeps:
                sq.record(M + s % M, comp_h)
            ops.append((first, len(sq) - first))
        sq.mark(1, comp_h)
        ops[-1] = (ops[-1][0], ops[-1][1] + 1)
        sq.keep = done
        return sq, ops

    # GTB_TIMELINE=1 (diagnosis, distorts the timing slightly): %globaltimer stamps around every stencil launch and
    # inside the transfer kernels; rank 0 prints the steps of the timed region to stderr
    timeline = halo_trace = None
    reserve_sms, halo_dma = 0, 0
    if he is not None and os.environ.get("GTB_TIMELINE") == "1":
        timeline = torch.zeros((total_steps, 2), dtype=torch.int64, device="cuda")
        halo_trace = torch.zeros((256, 8), dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().gtb_halo_set_trace(he._h, C.c_void_p(halo_trace.data_ptr())))
        epoch0 = he.epoch()
    if he is not None:
        reserve_sms = int(os.environ.get("GTB_RESERVE_SMS", RESERVE_SMS[name]))
        halo_dma = int(os.environ.get("GTB_HALO_DMA", HALO_DMA))
        _lib.set_option("reserve_sms", reserve_sms)  # left to the exchange
        _lib.set_option("halo.dma", halo_dma)
        _lib.set_option("halo.fused", int(os.environ.get("GTB_HALO_FUSED", HALO_FUSED)))
        if "GTB_PDL" in os.environ:
            _lib.set_option("pdl", int(os.environ["GTB_PDL"]))
        if "GTB_HALO_MAX_BLOCKS" in os.environ:
            _lib.set_option("halo.max_blocks", int(os.environ["GTB_HALO_MAX_BLOCKS"]))
        gated = name == "vert_adv" and itemsize == 8 and os.environ.get("GTB_GATES", "0") == "1"  # opt-in, see docstring
        seq, step_ops = build_sequence(gated)

    def step(s):
        if seq is not None:
            seq.run(*step_ops[s])
        else:
            run_stencil(s)

    # An NVML query stalls the GPU it reads for tens of microseconds: eight ranks sampling every 5 ms cost 5 us of a
    # 64 us step (profiles/r01_mgpu_ab_8.txt).  One GPU is sampled (rank 0's), immediately when the timed region starts
    # and then every 5 ms (10 ms with several ranks); GTB_NO_SAMPLER=1 switches it off to measure what is left.
    sampler = ClockSampler(local, 0.005 if world == 1 else 0.010)
    if os.environ.get("GTB_NO_SAMPLER") == "1" or rank != 0:
        sampler.nv = None
    for s in range(n_warm):
        step(s)
    barrier()
    if he is not None and any(h.check() != 0 for h in hes):
        raise SystemExit("bench.py: a halo wait timed out during warm-up")
    if gated:  # a device-side wait that gave up means the choreography is broken on this box: use stream events
        t = torch.tensor([_lib.gate_timeouts()], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if int(t.item()) > 0:
            gated = False
            torch.cuda.synchronize()
            seq, step_ops = build_sequence(False)
            # the halo object's epochs went on during the gated warm-up: nothing to reset, events carry no numbers
            for s in range(n_warm):
                step(s)
            barrier()
    sampler.start()
    launches0 = _lib.launch_count()
    if seq is None:
        # the K timed steps are bracketed by ONE pair of events: an event recorded between two launches keeps the
        # second from being staged behind the first and costs ~2.4 us per step (profiles/README.md); the per-step
        # distribution (median / p90 / max) comes from a second pass that pays for its events
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for s in range(args.steps):
            step(n_warm + s)
        ev1.record()
        barrier()
        total_ms = ev0.elapsed_time(ev1)
        launches = _lib.launch_count() - launches0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(min(args.steps, 100) + 1)]
        ev[0].record()
        for s in range(len(ev) - 1):
            step(n_warm + s)
            ev[s + 1].record()
        torch.cuda.synchronize()
        per_step = [ev[s].elapsed_time(ev[s + 1]) for s in range(len(ev) - 1)]
    else:  # one native call issues the K timed steps; the compute stream's events bracket them
        # LEAD lead-in steps; a rendezvous of ALL ranks on the device (the exchanges couple only neighbours, and with
        # n_sets rotating field sets a rank may run n_sets - 1 steps ahead of its neighbours: a rank that is ahead when it
        # passes its start mark would be charged that head start, 8 GPUs: +10 us per step at --steps 20); then the start
        # mark, the K timed steps and the end mark.  Both native calls are issued before the device has worked off the
        # first, so the compute stream never runs dry.
        lead_ops = sum(c for _, c in step_ops[n_warm:n_warm + LEAD])
        seq.run(step_ops[n_warm][0], lead_ops)
        rendezvous = torch.zeros(1, device="cuda")
        dist.all_reduce(rendezvous)  # enqueued on the compute stream: every rank's timed region starts at the same time
        seq.run(step_ops[n_warm][0] + lead_ops, sum(c for _, c in step_ops[n_warm + LEAD:]))
        barrier()
        total_ms = seq.elapsed_ms(0, 1)
        launches = (_lib.launch_count() - launches0) * args.steps // (args.steps + LEAD)
        per_step = [total_ms / args.steps]
        if any(h.check() != 0 for h in hes):
            raise SystemExit("bench.py: a halo wait timed out in the timed region")
    if timeline is not None and rank == 0:
        tl, ht = timeline.cpu().numpy(), halo_trace.cpu().numpy()
        s0 = n_warm + LEAD + 2
        t0 = int(tl[s0, 0])
        sys.stderr.write("timeline (us, relative to the start of stencil %d); exchange e feeds stencil e\n" % s0)
        sys.stderr.write("%5s %9s %9s | %9s %9s %9s | %9s %9s %9s\n" % ("step", "st.start", "st.end", "pk.start", "pk.end",
                                                                        "signal", "up.start", "up.flags", "up.end"))
        for st in range(s0, min(s0 + 8, total_steps)):
            row = ht[(epoch0 + st) % 256]
            rel = lambda v: (int(v) - t0) / 1e3  # noqa: E731
            sys.stderr.write("%5d %9.1f %9.1f | %9.1f %9.1f %9.1f | %9.1f %9.1f %9.1f\n" % (
                st, rel(tl[st, 0]), rel(tl[st, 1]), rel(row[0]), rel(row[1]), rel(row[5]) if row[5] else float("nan"),
                rel(row[2]), rel(row[3]) if row[3] else float("nan"), rel(row[4])))
    rank_ms = [total_ms / args.steps]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x.item()) / args.steps for x in allt]
        total_ms = max(float(x.item()) for x in allt)
    ms_per_step = total_ms / args.steps
    pts = NI * NJ * NK
    value = n_gpus * pts / (ms_per_step * 1e-3) / 1e6
    default_size = (NI, NJ, NK) == (256, 256, 80) and args.scaling == "weak"

    # ---- roofline of the dominant kernel (the stencil kernel: the only launch of a step at N = 1)
    # launch_ms: average duration of one stencil launch inside the timed region.  At N = 1 a step IS one launch, so
    # this is the CUDA-event time of the K back-to-back steps / K; at N > 1 (exchange kernels in the step) the stencil
    # launches are bracketed with their own events after the timed region.  isolated_ms (events around every single
    # launch, includes ~3 us of launch latency that back-to-back launches hide) is reported beside it.
    kern_ev = []
    for s in range(min(args.steps, 100)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_stencil(s)
        b.record()
        kern_ev.append((a, b))
    torch.cuda.synchronize()
    isolated_ms = statistics.mean(a.elapsed_time(b) for a, b in kern_ev)
    launch_ms = ms_per_step if world == 1 else isolated_ms
    clocks = sampler.stop()
    kernel_name = _lib.last_kernel().split(" ")[0]  # the stencil launch is the last thing this thread launched
    peak, peak_src = measured_peak()
    algo_bytes = ALGO_BYTES[name] * itemsize // 8
    achieved = algo_bytes * pts / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic_from_profile(name), "peak_source": peak_src,
                "kernel": kernel_name, "launch_ms": launch_ms,
                "launch_ms_source": "timed region / steps (one launch per step)" if world == 1 else
                "events around each stencil launch", "isolated_launch_ms": isolated_ms,
                "algorithmic_bytes_per_launch": algo_bytes * pts,
                "frac_of_nominal_8TBs": achieved / 8000.0}

    line = {
        "metric": "Mpts/s %s %dx%dx%d %s" % (name, global_ni if args.scaling == "strong" else NI,
                                              global_nj if args.scaling == "strong" else NJ, NK,
                                              "fp64" if itemsize == 8 else "fp32"), "value": value, "unit": "Mpts/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": value / n_gpus / P100_MPTS[name] if n_gpus == 1 and default_size and itemsize == 8 else None,
        "dtype": args.dtype, "data": "synthetic (reference repository's analytic fields%s)" % (
            ", evaluated on the device" if on_device else ""),
        "config": {"workload": "%s %dx%dx%d %s per GPU%s" % (
            name, NI, NJ, NK, "fp64" if itemsize == 8 else "fp32",
            " (BASELINE.json configs[1] family)" if default_size else
            " (%s scaling of a %dx%dx%d global domain)" % (args.scaling, global_ni, global_nj, NK)),
                   "decomposition": ("%dx%dx1 IJ process grid, halo exchange of %s every step over NVLink: local pack kernel, "
                                     "%s, device-side arrival flags, local unpack kernel; on a high-priority stream beside "
                                     "the previous step's stencil, ordered by %s; %d SMs reserved for the pack / unpack "
                                     "kernels; %d pattern object(s) take the steps in turn; loop issued as one recorded gtb_seq" % (
                                         dims[0], dims[1], "wcon" if name == "vert_adv" else "in",
                                         "copy-engine (DMA) transfers into the neighbours' receive buffers" if halo_dma else
                                         "peer stores from the pack kernel", "device-side gates" if gated else "stream events",
                                         reserve_sms, pipe)) if world > 1 else "single GPU",
                   "l2": "inputs of one step (%d MB) exceed L2 and %d field set(s) are rotated" % (
                       sum(f.nbytes_host for f in sets[0][:5 if name == "vert_adv" else 2]) // 2**20, n_sets),
                   "vs_baseline_ref": "reference stencil::gpu on P100, BASELINE.md section 1"},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        "rank_ms_per_step": rank_ms,
        "timing": ("marks in the compute stream of every rank: %d lead-in steps, an all-reduce on the compute stream (all "
                   "ranks start the timed region together), start mark, K steps, end mark; max over ranks" % LEAD)
        if seq is not None else
        "one CUDA event pair around the K back-to-back steps",
        "step_ms_median": statistics.median(per_step), "step_ms_p90": sorted(per_step)[int(0.9 * len(per_step))],
        "step_ms_max": max(per_step), "step_ms_note": "second pass with an event after every step (+~2.4 us per step)"
        if seq is None else "timed region / steps",
    }

    extras = not args.no_extras and default_size and itemsize == 8 and not on_device
    if extras:  # every rank moves its own sub-domain over its own PCIe link; max over ranks
        e = e2e(torch, stencil, storage, name, sets, dtr if name == "vert_adv" else None, args, n_gpus)
        if world > 1:
            t = torch.tensor([e["ms_per_step"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e["ms_per_step"] = float(t.item())
            e["value"] = n_gpus * pts / (e["ms_per_step"] * 1e-3) / 1e6
        line["e2e"] = e
    if rank == 0 and extras and world == 1:
        other = "hori_diff" if name == "vert_adv" else "vert_adv"
        # both halves of the metric in the kept line: the other stencil, same hygiene (>= 4 rotating sets, K launches
        # back to back between one pair of events)
        roofline[other] = secondary(torch, stencil, storage, other, peak, args.steps, max(args.warmup, 3))
        roofline[name] = {k: roofline[k] for k in ("achieved", "frac", "traffic", "kernel", "launch_ms")}
        line["parity_checked"], line["max_rel_err"], line["parity"] = False, None, {}
        for st_name in (name, other):
            try:
                backend, times, cores, want = time_reference(st_name, 20 if st_name == name else 3, 2)
                if st_name == name:
                    sec = statistics.median(times)
                    line["cpu_baseline"] = {
                        "value": pts / sec / 1e6, "unit": "Mpts/s", "cores": cores, "kind": "reference",
                        "sample": "20 timed applications of %s %dx%dx%d fp64 on the reference's stencil::%s (best of "
                                  "cpu_ifirst/cpu_kfirst), cache flush before each, median" % (name, NI, NJ, NK, backend)}
                # correctness inside the benchmarked configuration: one fresh application of the benchmarked call at the
                # benchmarked size against the reference's own CPU backend (oracle/_ref/libgtref.so as the checker)
                line["parity"][st_name] = parity(torch, stencil, storage, st_name, want)
            except Exception as e:  # the checker library is not part of the product
                if st_name == name:
                    line["cpu_baseline"] = {"value": None, "unit": "Mpts/s", "cores": 0, "kind": "reference",
                                            "sample": "unavailable: %s" % e}
                line["parity"][st_name] = {"checked": False, "why": str(e)}
        errs = [v.get("max_rel_err") for v in line["parity"].values() if v.get("checked")]
        if len(errs) == 2:
            line["parity_checked"], line["max_rel_err"] = True, max(errs)
            line["parity_tolerance"] = 1e-12
            if max(errs) > 1e-12:
                raise SystemExit("bench.py: result differs from the reference by %g (> 1e-12)" % max(errs))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def e2e(torch, stencil, storage, name, sets, dtr, args, n_gpus):
    """Same metric through the public API with host buffers: every step copies the step's input fields from pinned
    host memory to the device, runs the stencil and copies the result back to the host.  The three phases of
    consecutive steps are software-pipelined over the rotating device sets the way a user of the storage API would
    do it (upload of step n+1 and download of step n-1 run beside the stencil of step n: PCIe is full duplex), so a
    step costs max(upload, stencil, download) instead of their sum; nothing is skipped or cached."""
    steps = max(4, min(args.steps, 20))
    n_in = 5 if name == "vert_adv" else 2
    oi = 0 if name == "vert_adv" else 2
    H = HALO[name]
    # what a step moves: the points the stencil reads of every input field (compute domain; wcon one column further in
    # i; hori_diff's `in` with its halo of 2) and the compute domain of the result -- not the alignment padding and
    # unread halo rows of the storages (228 -> 210 MB per vert_adv step)
    interior = ((H, H, 0), (H + NI, H + NJ, NK))
    if name == "vert_adv":
        in_boxes = [interior, interior, ((H, H, 0), (H + NI + 1, H + NJ, NK)), interior, interior]
    else:
        in_boxes = [((0, 0, 0), (NI + 2 * H, NJ + 2 * H, NK)), interior]
    box_bytes = lambda b: (b[1][0] - b[0][0]) * (b[1][1] - b[0][1]) * (b[1][2] - b[0][2]) * 8  # noqa: E731
    h2d = sum(box_bytes(b) for b in in_boxes)
    d2h = box_bytes(interior)
    for st in sets:
        for f in st:
            f.const_host_view()
    comp = torch.cuda.current_stream()
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    n = len(sets)
    total = steps + 3
    ev_up = [torch.cuda.Event() for _ in range(total + 1)]
    ev_run = [torch.cuda.Event() for _ in range(total + 1)]
    ev_down = [torch.cuda.Event() for _ in range(total + 1)]

    def upload(s):  # inputs of step s; the set was last used by step s - n (its stencil read them, its download read the result)
        with torch.cuda.stream(up):
            if s - n >= 0:
                up.wait_event(ev_run[s - n])
                up.wait_event(ev_down[s - n])  # in-place result (vert_adv) is an input buffer too
            for f, b in zip(sets[s % n][:n_in], in_boxes):
                f.update_target_box_async(*b)
            ev_up[s].record(up)

    def run(s):
        comp.wait_event(ev_up[s])
        if s - n >= 0:
            comp.wait_event(ev_down[s - n])  # the result buffer of this set has been read back
        if name == "vert_adv":
            stencil.vertical_advection_dycore(*sets[s % n], dtr)
        else:
            stencil.horizontal_diffusion(*sets[s % n])
        ev_run[s].record(comp)

    def download(s):
        with torch.cuda.stream(down):
            down.wait_event(ev_run[s])
            sets[s % n][oi].update_host_box_async(*interior)
            ev_down[s].record(down)

    def one(s):
        if s == 0:
            upload(0)
        if s + 1 < total:
            upload(s + 1)
        run(s)
        download(s)

    for s in range(3):
        one(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(3, total):
        one(s)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    ms = wall * 1e3
    return {"value": n_gpus * NI * NJ * NK / (ms * 1e-3) / 1e6, "unit": "Mpts/s",
            "h2d_bytes_per_step": int(h2d) * n_gpus, "d2h_bytes_per_step": int(d2h) * n_gpus, "ms_per_step": ms,
            "steps": steps, "pipelined": "upload(n+1) | stencil(n) | download(n-1) on three streams",
            "timed_by": "host wall clock around %d steps, device synchronised on both sides" % steps}


def secondary(torch, stencil, storage, name, peak, steps, warmup):
    """The other stencil of the metric, kernel-only, same hygiene as the headline: 4 rotating field sets (> L2 each
    round), W warm-up launches, K launches back to back between ONE pair of events."""
    H = HALO[name]
    sets = []
    if name == "hori_diff":
        for _ in range(4):
            inp, coeff = repo_hori_diff(NI, NJ, NK)
            sets.append([storage.from_numpy(inp, (H, H, 0)), storage.from_numpy(coeff, (H, H, 0)),
                         storage.from_numpy(np.zeros_like(inp), (H, H, 0))])
        plans = [stencil.plan("horizontal_diffusion", *st) for st in sets]
    else:
        for _ in range(4):
            arrs, dtr = repo_vert_adv(NI, NJ, NK)
            sets.append([storage.from_numpy(a, (H, H, 0)) for a in arrs])
        plans = [stencil.plan("vertical_advection_dycore", *st, dtr_stage=0.15) for st in sets]
    for st in sets:
        for f in st:
            f.const_target_tensor()
    h = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for s in range(warmup):
        plans[s % 4](h)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(steps):
        plans[s % 4](h)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    from gridtools_b200 import _lib
    kernel_name = _lib.last_kernel().split(" ")[0]
    pts = NI * NJ * NK
    gbs = ALGO_BYTES[name] * pts / (ms * 1e-3) / 1e9
    return {"metric": "Mpts/s %s %dx%dx%d fp64" % (name, NI, NJ, NK), "value": pts / (ms * 1e-3) / 1e6,
            "unit": "Mpts/s", "launch_ms": ms, "achieved": gbs, "peak": peak, "frac": gbs / peak,
            "traffic": traffic_from_profile(name), "kernel": kernel_name,
            "steps": steps, "sets": 4}


def parity(torch, stencil, storage, name, want):
    """One application of the benchmarked call on fresh fields of the benchmarked size, compared with the reference's
    CPU backend result `want` (relative error as tests/include/verifier.hpp:26-52 defines it)."""
    H = HALO[name]
    if name == "vert_adv":
        arrs, dtr = repo_vert_adv(NI, NJ, NK)
        st = [storage.from_numpy(a, (H, H, 0)) for a in arrs]
        stencil.vertical_advection_dycore(*st, dtr)
        out = st[0]
    else:
        inp, coeff = repo_hori_diff(NI, NJ, NK)
        st = [storage.from_numpy(inp, (H, H, 0)), storage.from_numpy(coeff, (H, H, 0)),
              storage.from_numpy(np.zeros_like(inp), (H, H, 0))]
        stencil.horizontal_diffusion(*st)
        out = st[2]
    torch.cuda.synchronize()
    got = out.to_numpy()[:, H:-H, H:-H]
    ref = want[:, H:-H, H:-H]
    den = np.maximum(np.maximum(np.abs(got), np.abs(ref)), 1e-300)
    err = float(np.max(np.abs(got - ref) / den))
    return {"checked": True, "max_rel_err": err, "points": int(got.size), "against": "reference cpu backend (libgtref)"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        b200_arm(a)
