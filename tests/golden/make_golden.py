"""Generates the golden fixtures of tests/golden/ from the REFERENCE itself.

Runs only in the build container (needs /root/reference through oracle/_ref/libgtref.so, built by oracle/Makefile):
inputs come from the reference's analytic repositories (tests/regression/horizontal_diffusion_repository.hpp:32-79,
vertical_advection_repository.hpp:70-151), outputs from the reference's own cpu_ifirst backend run on them; the
repositories' analytic answers are stored next to them.  The fixtures are what pins oracle/gt_oracle.c and the CUDA
kernels on machines that have no reference tree (the GPU box).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyoracle as o  # noqa: E402


def hori_diff(ni, nj, nk, name):
    inp, coeff, repo_out = o.repo_hori_diff(ni, nj, nk)
    outs = {}
    for be in ("cpu_ifirst", "cpu_kfirst", "naive"):
        res = np.zeros_like(inp)
        o.ref_run(o.HORI_DIFF, be, [inp, coeff], [res], ni, nj, nk)
        outs[be] = res
    assert all(np.array_equal(outs["cpu_ifirst"], v) for v in outs.values()), "reference backends disagree"
    inp32, coeff32 = inp.astype(np.float32), coeff.astype(np.float32)
    res32 = np.zeros_like(inp32)
    o.ref_run(o.HORI_DIFF, "cpu_ifirst", [inp32, coeff32], [res32], ni, nj, nk)
    np.savez_compressed(os.path.join(HERE, name), ni=ni, nj=nj, nk=nk, halo=2, inp=inp, coeff=coeff,
                        out_ref=outs["cpu_ifirst"], out_repo=repo_out, out_ref_f32=res32)


def simple_hori_diff(ni, nj, nk, name):
    inp, coeff, cro, cru, repo_out = o.repo_simple_hori_diff(ni, nj, nk)
    outs = {}
    for be in ("cpu_ifirst", "cpu_kfirst", "naive"):
        res = np.zeros_like(inp)
        o.ref_run(o.SIMPLE_HORI_DIFF, be, [inp, coeff, cro, cru], [res], ni, nj, nk)
        outs[be] = res
    assert all(np.array_equal(outs["cpu_ifirst"], v) for v in outs.values()), "reference backends disagree"
    f32 = [a.astype(np.float32) for a in (inp, coeff, cro, cru)]
    res32 = np.zeros_like(f32[0])
    o.ref_run(o.SIMPLE_HORI_DIFF, "cpu_ifirst", f32, [res32], ni, nj, nk)
    np.savez_compressed(os.path.join(HERE, name), ni=ni, nj=nj, nk=nk, halo=2, inp=inp, coeff=coeff, crlato=cro,
                        crlatu=cru, out_ref=outs["cpu_ifirst"], out_repo=repo_out, out_ref_f32=res32)


def vert_adv(ni, nj, nk, name):
    arrs, repo_out, dtr = o.repo_vert_adv(ni, nj, nk)
    outs = {}
    for be in ("cpu_ifirst", "cpu_kfirst", "naive"):
        res = np.zeros_like(arrs[0])
        o.ref_run(o.VERT_ADV, be, arrs, [res], ni, nj, nk, scalar=dtr)
        outs[be] = res
    assert all(np.array_equal(outs["cpu_ifirst"], v) for v in outs.values()), "reference backends disagree"
    arrs32 = [a.astype(np.float32) for a in arrs]
    res32 = np.zeros_like(arrs32[0])
    o.ref_run(o.VERT_ADV, "cpu_ifirst", arrs32, [res32], ni, nj, nk, scalar=dtr)
    np.savez_compressed(os.path.join(HERE, name), ni=ni, nj=nj, nk=nk, halo=3, dtr_stage=dtr,
                        utens_stage=arrs[0], u_stage=arrs[1], wcon=arrs[2], u_pos=arrs[3], utens=arrs[4],
                        out_ref=outs["cpu_ifirst"], out_repo=repo_out, out_ref_f32=res32)


def tridiagonal(ni, nj, nk, name):
    # tridiagonal.cpp:83-97: inf = -1, diag = 3, sup = 1, rhs = 4 at k=0, 3 inside, 2 at the last level => out == 1
    shape = (nk, nj, ni)
    inf, diag, sup = -np.ones(shape), 3 * np.ones(shape), np.ones(shape)
    rhs = 3 * np.ones(shape)
    rhs[0], rhs[-1] = 4, 2
    out, sup2, rhs2 = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    o.ref_run(o.TRIDIAGONAL, "cpu_ifirst", [inf, diag, sup, rhs], [out, sup2, rhs2], ni, nj, nk)
    np.savez_compressed(os.path.join(HERE, name), ni=ni, nj=nj, nk=nk, inf=inf, diag=diag, sup=sup, rhs=rhs,
                        out_ref=out, sup_ref=sup2, rhs_ref=rhs2)


def boundary(name):
    """boundary<value_boundary / copy_boundary, gcl::cpu, predicate>::apply of the reference on random boxes."""
    h = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (1, 1, 1, 4, 6)]
    shape = (6, 10, 14)
    rng = np.random.default_rng(42)
    mask = [int(x) for x in rng.integers(0, 2, 27)]
    mask[13] = 0
    value = 3.25
    out = dict(halos=np.array(h), mask=np.array(mask), value=value)
    for case, kind, nf, m in (("value_all", 0, 2, None), ("value_masked", 0, 3, mask), ("copy_masked", 1, 3, mask)):
        f = [rng.standard_normal(shape) for _ in range(nf)]
        r = [a.copy() for a in f]
        o.ref_boundary(h, m, kind, value, r)
        out[case + "_in"], out[case + "_ref"] = np.array(f), np.array(r)
    np.savez_compressed(os.path.join(HERE, name), **out)


def halo(name, proc_dims, periodic, layout=(2, 1, 0), n_fields=2):
    """One pack / exchange / unpack of the reference's gcl::halo_exchange_dynamic_ut<layout, <0,1,2>, double, cpu>
    (run with threads as ranks through oracle/mpi_shim/mpi.h) on random boxes; halos in USER dimension order."""
    h = [(2, 3, 2, 9, 14), (1, 2, 1, 6, 10), (0, 1, 0, 4, 6)]
    order = np.argsort(layout)
    shape = tuple(h[d][4] for d in order)
    n = proc_dims[0] * proc_dims[1] * proc_dims[2]
    rng = np.random.default_rng(7)
    start = rng.standard_normal((n, n_fields) + shape)
    res = [[start[r, f].copy() for f in range(n_fields)] for r in range(n)]
    o.ref_gcl_exchange(h, proc_dims, periodic, res, layout=layout)
    np.savez_compressed(os.path.join(HERE, name), halos=np.array(h), proc_dims=np.array(proc_dims),
                        periodic=np.array(periodic), layout=np.array(layout), n_fields=n_fields, start=start,
                        result=np.array(res))


if __name__ == "__main__":
    o.build(ref=True)
    halo("halo_2x2x1_p101.npz", (2, 2, 1), (1, 0, 1))
    halo("halo_2x4x1_p000.npz", (2, 4, 1), (0, 0, 0))
    halo("halo_1x2x2_p010_l021.npz", (1, 2, 2), (0, 1, 0), layout=(0, 2, 1))
    hori_diff(12, 33, 6, "hori_diff_12x33x6.npz")     # test_environment sizes 12x33 (k shortened)
    hori_diff(70, 19, 3, "hori_diff_70x19x3.npz")     # crosses a 64-wide tile boundary
    vert_adv(13, 7, 61, "vert_adv_13x7x61.npz")       # 61 levels like the 12x33x61 environment
    vert_adv(35, 5, 9, "vert_adv_35x5x9.npz")         # crosses a warp boundary, short column
    tridiagonal(12, 33, 6, "tridiagonal_12x33x6.npz") # tridiagonal.cpp sizes
    simple_hori_diff(70, 19, 3, "simple_hori_diff_70x19x3.npz")
    simple_hori_diff(12, 33, 6, "simple_hori_diff_12x33x6.npz")
    boundary("boundary_14x10x6.npz")
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
