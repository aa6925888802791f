"""Pins the oracle (oracle/gt_oracle.c): against the committed golden fixtures (made by the reference itself,
tests/golden/make_golden.py), against the reference's analytic repositories, and -- where /root/reference exists
(build container) -- against the reference's CPU backends run live through oracle/_ref/libgtref.so."""
import numpy as np
import pytest

# tolerance of the reference's own verifier (tests/include/verifier.hpp:26-52): 1e-14 double, 1e-6 float; the
# north star allows 1e-12 / 1e-5
TOL64, TOL32 = 1e-12, 1e-5


def rel_err(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    return float(np.max(np.where(d < 1e-300, 0.0, d / np.maximum(s, 1e-300))))


def verifier_ok(e, a, p):
    """tests/include/verifier.hpp: |e-a| < p or < p*max(|e|,|a|)."""
    d = np.abs(e - a)
    return bool(np.all((d < p) | (d < p * np.maximum(np.abs(e), np.abs(a)))))


@pytest.mark.parametrize("name", ["hori_diff_12x33x6.npz", "hori_diff_70x19x3.npz"])
def test_hori_diff_golden(oracle, golden, name):
    g = golden(name)
    H = int(g["halo"])
    out = oracle.hori_diff(g["inp"], g["coeff"])
    inner = (slice(None), slice(H, -H), slice(H, -H))
    assert verifier_ok(g["out_ref"][inner], out[inner], 1e-14)
    assert verifier_ok(g["out_repo"][inner], out[inner], 1e-14)
    assert rel_err(g["out_ref"][inner], out[inner]) < TOL64
    out32 = oracle.hori_diff(g["inp"].astype(np.float32), g["coeff"].astype(np.float32))
    assert rel_err(g["out_ref_f32"][inner].astype(np.float64), out32[inner].astype(np.float64)) < TOL32


@pytest.mark.parametrize("name", ["simple_hori_diff_12x33x6.npz", "simple_hori_diff_70x19x3.npz"])
def test_simple_hori_diff_golden(oracle, golden, name):
    g = golden(name)
    H = int(g["halo"])
    out = oracle.simple_hori_diff(g["inp"], g["coeff"], g["crlato"], g["crlatu"])
    inner = (slice(None), slice(H, -H), slice(H, -H))
    assert verifier_ok(g["out_ref"][inner], out[inner], 1e-14)      # the reference's cpu_ifirst run
    assert verifier_ok(g["out_repo"][inner], out[inner], 1e-14)     # horizontal_diffusion_repository.hpp:68-79
    f32 = [g[k].astype(np.float32) for k in ("inp", "coeff", "crlato", "crlatu")]
    out32 = oracle.simple_hori_diff(*f32)
    assert rel_err(g["out_ref_f32"][inner].astype(np.float64), out32[inner].astype(np.float64)) < TOL32


@pytest.mark.parametrize("name", ["vert_adv_13x7x61.npz", "vert_adv_35x5x9.npz"])
def test_vert_adv_golden(oracle, golden, name):
    g = golden(name)
    H = int(g["halo"])
    args = [g[k] for k in ("utens_stage", "u_stage", "wcon", "u_pos", "utens")]
    out = oracle.vert_adv(*args, float(g["dtr_stage"]))
    inner = (slice(None), slice(H, -H), slice(H, -H))
    assert rel_err(g["out_ref"][inner], out[inner]) < TOL64
    assert rel_err(g["out_repo"][inner], out[inner]) < TOL64
    assert verifier_ok(g["out_repo"][inner], out[inner], 1e-13)
    out32 = oracle.vert_adv(*[a.astype(np.float32) for a in args], float(g["dtr_stage"]))
    assert rel_err(g["out_ref_f32"][inner].astype(np.float64), out32[inner].astype(np.float64)) < 1e-4


def test_tridiagonal_golden(oracle, golden):
    g = golden("tridiagonal_12x33x6.npz")
    out, sup, rhs = oracle.tridiagonal(g["inf"], g["diag"], g["sup"], g["rhs"])
    assert np.allclose(out, 1.0, rtol=0, atol=1e-14)  # known answer, tridiagonal.cpp:76-98
    assert rel_err(g["out_ref"], out) < TOL64
    assert rel_err(g["sup_ref"], sup) < TOL64 and rel_err(g["rhs_ref"], rhs) < TOL64


def test_copy_and_tracers(oracle):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 7, 9))
    assert np.array_equal(oracle.copy(a), a)
    a32 = a.astype(np.float32)
    assert np.array_equal(oracle.copy(a32), a32)
    rho = rng.standard_normal((3, 4, 6))
    ins = [np.full((3, 4, 6), 1.1 * i) for i in range(11)]  # advection_pdbott_prepare_tracers.cpp:55-56
    outs = oracle.prepare_tracers(ins, rho)
    for i, o_ in enumerate(outs):
        assert np.array_equal(o_, rho * (1.1 * i))


# ----------------------------------------------------------------------------------------------- live reference
needs_ref = pytest.mark.skipif(not __import__("os").path.exists("/root/reference/include/gridtools"),
                               reason="reference tree only exists in the build container")


@needs_ref
@pytest.mark.parametrize("backend", ["cpu_ifirst", "cpu_kfirst", "naive"])
@pytest.mark.parametrize("size", [(12, 33, 5), (23, 11, 4), (64, 16, 2)])
def test_hori_diff_vs_reference(oracle, backend, size):
    ni, nj, nk = size
    rng = np.random.default_rng(ni * 1000 + nj)
    inp = rng.standard_normal((nk, nj + 4, ni + 4))
    coeff = rng.uniform(0.0, 0.05, inp.shape)
    res = np.zeros_like(inp)
    oracle.ref_run(oracle.HORI_DIFF, backend, [inp, coeff], [res], ni, nj, nk)
    out = oracle.hori_diff(inp, coeff)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    assert verifier_ok(res[inner], out[inner], 1e-13)


@needs_ref
@pytest.mark.parametrize("backend", ["cpu_ifirst", "cpu_kfirst", "naive"])
@pytest.mark.parametrize("size", [(12, 33, 5), (70, 9, 2)])
def test_simple_hori_diff_vs_reference(oracle, backend, size):
    ni, nj, nk = size
    rng = np.random.default_rng(ni * 77 + nj)
    inp = rng.standard_normal((nk, nj + 4, ni + 4))
    coeff = rng.uniform(0.0, 0.05, inp.shape)
    cro, cru = rng.uniform(0.5, 1.5, nj + 4), rng.uniform(0.5, 1.5, nj + 4)
    res = np.zeros_like(inp)
    oracle.ref_run(oracle.SIMPLE_HORI_DIFF, backend, [inp, coeff, cro, cru], [res], ni, nj, nk)
    out = oracle.simple_hori_diff(inp, coeff, cro, cru)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    assert verifier_ok(res[inner], out[inner], 1e-13)


@needs_ref
@pytest.mark.parametrize("backend", ["cpu_ifirst", "cpu_kfirst"])
@pytest.mark.parametrize("size", [(12, 33, 61), (23, 11, 43), (5, 3, 2)])
def test_vert_adv_vs_reference(oracle, backend, size):
    ni, nj, nk = size
    arrs, repo_out, dtr = oracle.repo_vert_adv(ni, nj, nk)
    res = np.zeros_like(arrs[0])
    oracle.ref_run(oracle.VERT_ADV, backend, arrs, [res], ni, nj, nk, scalar=dtr)
    out = oracle.vert_adv(*arrs, dtr)
    inner = (slice(None), slice(3, -3), slice(3, -3))
    assert rel_err(res[inner], out[inner]) < TOL64
    assert rel_err(repo_out[inner], out[inner]) < TOL64


@needs_ref
def test_tridiagonal_and_copy_vs_reference(oracle):
    rng = np.random.default_rng(3)
    shape = (7, 6, 10)
    inf, sup = rng.uniform(-1, 0, shape), rng.uniform(0, 1, shape)
    diag, rhs = rng.uniform(3, 4, shape), rng.standard_normal(shape)
    out, sup2, rhs2 = np.zeros(shape), np.zeros(shape), np.zeros(shape)
    oracle.ref_run(oracle.TRIDIAGONAL, "cpu_ifirst", [inf, diag, sup, rhs], [out, sup2, rhs2], 10, 6, 7)
    o_out, o_sup, o_rhs = oracle.tridiagonal(inf, diag, sup, rhs)
    assert rel_err(out, o_out) < TOL64 and rel_err(sup2, o_sup) < TOL64 and rel_err(rhs2, o_rhs) < TOL64
    res = np.zeros(shape)
    oracle.ref_run(oracle.COPY, "cpu_kfirst", [rhs], [res], 10, 6, 7)
    assert np.array_equal(res, rhs) and np.array_equal(oracle.copy(rhs), rhs)
