"""Parity of the CUDA kernels (called through the C ABI) with the oracle and the committed golden fixtures.

Bar: bit-exact for copy; for the arithmetic stencils the kernels evaluate the functor expressions operation by
operation without FMA contraction, so they are also BIT-EXACT against oracle/gt_oracle.c (-ffp-contract=off), and
within 1e-12 (fp64) / 1e-5 (fp32) relative of the reference's cpu_ifirst output stored in tests/golden/ (the
reference build contracts FMAs)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64, TOL32 = 1e-12, 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b)
    s = np.maximum(np.abs(a), np.abs(b))
    return float(np.max(np.where(d == 0, 0.0, d / np.maximum(s, 1e-300))))


@pytest.fixture(scope="module")
def gt():
    import torch
    from gridtools_b200 import _lib, stencil, storage
    _lib.check(_lib.lib().gtb_init(0))
    torch.cuda.set_device(0)

    class NS:
        pass
    ns = NS()
    ns.lib, ns.stencil, ns.storage, ns.torch = _lib, stencil, storage, torch
    return ns


@pytest.fixture(autouse=True)
def reset_options(gt):
    yield
    for k in ("hd.variant", "hd.stages", "hd.ctas_per_sm", "va.variant", "va.threads", "va.unroll", "va.scratch", "va.stagger",
              "va.ctas_per_sm", "va.save_upos", "va.stages", "va.bldg"):
        gt.lib.set_option(k, 0)
    gt.lib.set_option("va.hints", 1)
    gt.lib.set_option("copy.vec", 1)


def run_hd(gt, inp, coeff, alignment=128, out_init=-7.0):
    H = 2
    si = gt.storage.from_numpy(inp, (H, H, 0), alignment)
    sc = gt.storage.from_numpy(coeff, (H, H, 0), alignment)
    so = gt.storage.from_numpy(np.full_like(inp, out_init), (H, H, 0), alignment)
    gt.stencil.horizontal_diffusion(si, sc, so)
    gt.torch.cuda.synchronize()
    return so.to_numpy()


def run_va(gt, arrs, dtr, alignment=128):
    H = 3
    st = [gt.storage.from_numpy(a, (H, H, 0), alignment) for a in arrs]
    gt.stencil.vertical_advection_dycore(*st, dtr)
    gt.torch.cuda.synchronize()
    return st[0].to_numpy(), [s.to_numpy() for s in st[1:]]


# --------------------------------------------------------------------------------------------------- copy
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,halo,alignment", [((7, 9, 16), 0, 128), ((5, 11, 13), 0, 1), ((80, 36, 68), 2, 128),
                                                  ((3, 1, 1), 0, 1)])
def test_copy_bit_exact(gt, oracle, dtype, shape, halo, alignment):
    rng = np.random.default_rng(1)
    a = rng.standard_normal(shape).astype(dtype)
    a.view(np.uint32 if dtype == np.float32 else np.uint64)[0, 0, 0] |= 1  # odd payload bits survive a bit copy
    src = gt.storage.from_numpy(a, (halo, halo, 0), alignment)
    dst = gt.storage.from_numpy(np.zeros_like(a), (halo, halo, 0), alignment)
    gt.stencil.copy(src, dst)
    gt.torch.cuda.synchronize()
    expect = oracle.copy(a, halo)
    assert np.array_equal(dst.to_numpy().view(np.uint8), expect.view(np.uint8))


def test_copy_scalar_path(gt, oracle):
    gt.lib.set_option("copy.vec", 0)
    a = np.random.default_rng(2).standard_normal((4, 6, 32))
    src, dst = gt.storage.from_numpy(a, (0, 0, 0)), gt.storage.from_numpy(np.zeros_like(a), (0, 0, 0))
    gt.stencil.copy(src, dst)
    assert np.array_equal(dst.to_numpy(), a)


# ------------------------------------------------------------------------------------- horizontal diffusion
@pytest.mark.parametrize("variant", [1, 2, 4])
@pytest.mark.parametrize("name", ["hori_diff_12x33x6.npz", "hori_diff_70x19x3.npz"])
def test_hori_diff_golden(gt, oracle, golden, name, variant):
    g = golden(name)
    gt.lib.set_option("hd.variant", variant)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    out = run_hd(gt, g["inp"], g["coeff"])
    assert np.array_equal(out[inner], oracle.hori_diff(g["inp"], g["coeff"])[inner])      # bit exact vs oracle
    assert rel_err(out[inner], g["out_ref"][inner]) < TOL64                                # vs reference cpu_ifirst
    assert rel_err(out[inner], g["out_repo"][inner]) < TOL64                               # vs analytic repository
    out32 = run_hd(gt, g["inp"].astype(np.float32), g["coeff"].astype(np.float32))
    assert rel_err(out32[inner], g["out_ref_f32"][inner]) < TOL32
    halo_mask = np.ones(out.shape, bool)
    halo_mask[inner] = False
    assert np.all(out[halo_mask] == -7.0), "the kernel wrote outside the compute domain"


@pytest.mark.parametrize("variant", [0, 1, 2, 4])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("size,alignment", [((1, 1, 1), 128), ((5, 3, 2), 1), ((64, 16, 3), 128), ((65, 17, 2), 128),
                                            ((129, 47, 5), 1), ((200, 40, 7), 128), ((23, 11, 43), 128)])
def test_hori_diff_random_bit_exact(gt, oracle, variant, dtype, size, alignment):
    ni, nj, nk = size
    rng = np.random.default_rng(ni * 131 + nj * 7 + nk)
    inp = rng.standard_normal((nk, nj + 4, ni + 4)).astype(dtype)
    coeff = rng.uniform(0, 0.05, inp.shape).astype(dtype)
    gt.lib.set_option("hd.variant", variant)
    try:
        out = run_hd(gt, inp, coeff, alignment)
    except gt.lib.GtbError as e:
        # an explicitly requested TMA variant refuses layouts TMA cannot address (variant 0 falls back to cp.async)
        assert variant in (2, 4) and alignment == 1 and e.status == gt.lib.GTB_ERR_LAYOUT
        pytest.skip("layout is not TMA addressable")
    inner = (slice(None), slice(2, -2), slice(2, -2))
    assert np.array_equal(out[inner], oracle.hori_diff(inp, coeff)[inner])


def test_hori_diff_tma_refuses_bad_layout(gt):
    inp = np.zeros((2, 9, 9))  # 9 doubles per row: stride not a multiple of 16 bytes
    gt.lib.set_option("hd.variant", 2)
    with pytest.raises(gt.lib.GtbError) as e:
        run_hd(gt, inp, inp.copy(), alignment=1)
    assert e.value.status == gt.lib.GTB_ERR_LAYOUT


@pytest.mark.parametrize("stages,ctas", [(2, 1), (3, 2), (5, 1)])
def test_hori_diff_pipeline_depths(gt, oracle, stages, ctas):
    rng = np.random.default_rng(5)
    inp = rng.standard_normal((9, 36, 132))
    coeff = rng.uniform(0, 0.05, inp.shape)
    ref = oracle.hori_diff(inp, coeff)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    for variant in (1, 2, 4):
        gt.lib.set_option("hd.variant", variant)
        gt.lib.set_option("hd.stages", stages)
        gt.lib.set_option("hd.ctas_per_sm", ctas)
        assert np.array_equal(run_hd(gt, inp, coeff)[inner], ref[inner])


def test_hori_diff_full_size(gt, oracle):
    """BASELINE.json's 256x256x80 fp64 (and configs[0]'s 128x128x80) against the oracle and the repository."""
    for n in (128, 256):
        ni = nj = n
        nk = 80
        d0, d1 = ni + 4, nj + 4
        k, j, i = np.meshgrid(np.arange(nk), np.arange(d1), np.arange(d0), indexing="ij", sparse=True)
        x, y = i / d0, j / d1  # horizontal_diffusion_repository.hpp:32-42
        inp = 5. + 8 * (2. + np.cos(np.pi * (x + 1.5 * y)) + np.sin(2 * np.pi * (x + 1.5 * y))) / 4. + 0. * k
        coeff = np.full_like(inp, 0.025)
        inner = (slice(None), slice(2, -2), slice(2, -2))
        want = oracle.hori_diff(inp, coeff)[inner]
        for variant in (0, 4):  # default; two pipelines in one CTA per SM
            gt.lib.set_option("hd.variant", variant)
            out = run_hd(gt, inp, coeff)
            assert np.array_equal(out[inner], want), variant


def test_last_kernel_reports_the_automatic_choice(gt):
    """gtb_last_kernel: a short launch takes the one-CTA two-pipeline kernel (programmatic dependent launch), an explicit
    hd.variant 2 or SMs reserved for an exchange the two-CTA kernel; vertical advection fp64 the paired-warp kernel."""
    rng = np.random.default_rng(3)
    inp = rng.standard_normal((8, 36, 132))
    coeff = rng.uniform(0, 0.05, inp.shape)
    run_hd(gt, inp, coeff)
    assert gt.lib.last_kernel().startswith("hd_tma2_kernel")
    gt.lib.set_option("hd.variant", 2)
    run_hd(gt, inp, coeff)
    assert gt.lib.last_kernel().startswith("hd_tma_kernel")
    gt.lib.set_option("hd.variant", 0)
    gt.lib.set_option("reserve_sms", 8)
    try:
        run_hd(gt, inp, coeff)
        assert gt.lib.last_kernel().startswith("hd_tma_kernel")
    finally:
        gt.lib.set_option("reserve_sms", 0)
    shape = (20, 11, 70)
    arrs = [rng.uniform(5, 9, shape), rng.uniform(5, 9, shape), rng.uniform(-3e-4, 3e-4, shape),
            rng.uniform(5, 9, shape), rng.uniform(-1e-5, 1e-5, shape)]
    run_va(gt, arrs, 0.15)
    assert gt.lib.last_kernel().startswith("va_pair_kernel")


def test_hori_diff_linearity(gt):
    """Size-independent property at 512x512x80 fp32 (configs[2]): scaling `in` by a power of two scales `out`
    exactly (the flux limiter only looks at signs)."""
    ni = nj = 512
    nk = 80
    rng = np.random.default_rng(9)
    inp = rng.standard_normal((nk, nj + 4, ni + 4)).astype(np.float32)
    coeff = rng.uniform(0, 0.05, inp.shape).astype(np.float32)
    a = run_hd(gt, inp, coeff)
    b = run_hd(gt, inp * np.float32(4), coeff)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    assert np.array_equal(a[inner] * np.float32(4), b[inner])


# ------------------------------------------------------------------------------------- simple horizontal diffusion
def run_shd(gt, inp, coeff, cro, cru, alignment=128):
    H = 2
    si = gt.storage.from_numpy(inp, (H, H, 0), alignment)
    sc = gt.storage.from_numpy(coeff, (H, H, 0), alignment)
    so = gt.storage.from_numpy(np.full_like(inp, -7.0), (H, H, 0), alignment)
    d2, d1, d0 = inp.shape
    jb = gt.storage.builder.type(inp.dtype).dimensions(d0, d1, d2).halos(H, H, 0).selector(0, 1, 0)
    so_, su_ = jb.build(), jb.build()
    so_.host_view()[0, :, 0] = cro
    su_.host_view()[0, :, 0] = cru
    gt.stencil.simple_hori_diff(sc, si, so, so_, su_)
    gt.torch.cuda.synchronize()
    return so.to_numpy()


@pytest.mark.parametrize("name", ["simple_hori_diff_12x33x6.npz", "simple_hori_diff_70x19x3.npz"])
def test_simple_hori_diff_golden(gt, oracle, golden, name):
    g = golden(name)
    out = run_shd(gt, g["inp"], g["coeff"], g["crlato"], g["crlatu"])
    inner = (slice(None), slice(2, -2), slice(2, -2))
    assert np.array_equal(out[inner], oracle.simple_hori_diff(g["inp"], g["coeff"], g["crlato"], g["crlatu"])[inner])
    assert rel_err(out[inner], g["out_ref"][inner]) < TOL64 and rel_err(out[inner], g["out_repo"][inner]) < TOL64
    halo_mask = np.ones(out.shape, bool)
    halo_mask[inner] = False
    assert np.all(out[halo_mask] == -7.0), "halo of out was modified"
    f32 = [g[k].astype(np.float32) for k in ("inp", "coeff", "crlato", "crlatu")]
    out32 = run_shd(gt, *f32)
    assert np.array_equal(out32[inner], oracle.simple_hori_diff(*f32)[inner])
    assert rel_err(out32[inner], g["out_ref_f32"][inner]) < TOL32


@pytest.mark.parametrize("size", [(1, 1, 1), (64, 8, 2), (65, 9, 3), (130, 17, 4), (256, 256, 5), (33, 70, 2)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_simple_hori_diff_random_bit_exact(gt, oracle, size, dtype):
    ni, nj, nk = size
    rng = np.random.default_rng(ni + 31 * nj)
    inp = rng.standard_normal((nk, nj + 4, ni + 4)).astype(dtype)
    coeff = rng.uniform(0, 0.05, inp.shape).astype(dtype)
    cro, cru = rng.uniform(0.5, 1.5, nj + 4).astype(dtype), rng.uniform(0.5, 1.5, nj + 4).astype(dtype)
    want = oracle.simple_hori_diff(inp, coeff, cro, cru)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    for variant in (0, 1):  # 0: TMA-staged register-tile kernel; 1: plain shared-memory tile kernel (any layout)
        gt.lib.set_option("hd.variant", variant)
        out = run_shd(gt, inp, coeff, cro, cru)
        assert np.array_equal(out[inner], want[inner]), variant
    out = run_shd(gt, inp, coeff, cro, cru, alignment=1)  # not TMA-addressable: falls back by itself
    assert np.array_equal(out[inner], want[inner])


# ------------------------------------------------------------------------------------- vertical advection
VA_CONFIGS = [dict(), dict(variant=7), dict(variant=7, ctas_per_sm=-1), dict(variant=7, ctas_per_sm=3, stages=2, unroll=2),
              dict(variant=7, bldg=1), dict(variant=7, bldg=2), dict(variant=7, bldg=1, ctas_per_sm=-1),
              dict(variant=7, bldg=1, ctas_per_sm=3, stages=2), dict(variant=7, bldg=1, ctas_per_sm=7, stages=3),
              dict(variant=7, ctas_per_sm=7, stages=3, unroll=4), dict(variant=3), dict(variant=3, ctas_per_sm=-2, unroll=8),
              dict(variant=3, ctas_per_sm=-1), dict(variant=3, stages=3, ctas_per_sm=-3), dict(variant=3, stages=6),
              dict(variant=3, ctas_per_sm=-2, save_upos=2), dict(variant=3, threads=32, ctas_per_sm=-1, save_upos=2),
              dict(variant=3, threads=64, ctas_per_sm=-2), dict(variant=3, threads=96, save_upos=2),
              dict(variant=3, unroll=8, save_upos=2, ctas_per_sm=-1),
              dict(variant=1), dict(variant=1, threads=32, unroll=1), dict(variant=1, threads=128, unroll=2),
              dict(variant=1, threads=64, unroll=8, save_upos=2), dict(variant=1, scratch=2, threads=32, unroll=4),
              dict(variant=1, scratch=2, threads=64, unroll=2, hints=0, save_upos=2), dict(variant=1, hints=0, unroll=4),
              dict(variant=1, ctas_per_sm=-3, threads=32, unroll=2), dict(variant=1, ctas_per_sm=-2, threads=32),
              dict(variant=1, scratch=2, threads=32, ctas_per_sm=-1), dict(variant=1, hints=0, unroll=8)]


def set_va(gt, cfg):
    for k, v in cfg.items():
        gt.lib.set_option("va." + k, v)


@pytest.mark.parametrize("cfg", VA_CONFIGS)
@pytest.mark.parametrize("name", ["vert_adv_13x7x61.npz", "vert_adv_35x5x9.npz"])
def test_vert_adv_golden(gt, oracle, golden, name, cfg):
    g = golden(name)
    set_va(gt, cfg)
    arrs = [g[k] for k in ("utens_stage", "u_stage", "wcon", "u_pos", "utens")]
    dtr = float(g["dtr_stage"])
    out, others = run_va(gt, arrs, dtr)
    inner = (slice(None), slice(3, -3), slice(3, -3))
    assert np.array_equal(out[inner], oracle.vert_adv(*arrs, dtr)[inner])       # bit exact vs oracle
    assert rel_err(out[inner], g["out_ref"][inner]) < TOL64                      # vs reference cpu_ifirst
    assert rel_err(out[inner], g["out_repo"][inner]) < TOL64                     # vs analytic Thomas solve
    halo_mask = np.ones(out.shape, bool)
    halo_mask[inner] = False
    assert np.array_equal(out[halo_mask], arrs[0][halo_mask]), "halo of utens_stage was modified"
    for a, b in zip(others, arrs[1:]):
        assert np.array_equal(a, b), "an input field was modified"
    arrs32 = [a.astype(np.float32) for a in arrs]
    out32, _ = run_va(gt, arrs32, dtr)
    assert np.array_equal(out32[inner], oracle.vert_adv(*arrs32, dtr)[inner])
    assert rel_err(out32[inner], g["out_ref_f32"][inner]) < 1e-4  # fp32 Thomas: cancellation in dtr*(x - u_pos)


@pytest.mark.parametrize("size,alignment", [((1, 1, 2), 1), ((33, 2, 3), 1), ((70, 9, 80), 128), ((12, 33, 61), 128),
                                            ((40, 3, 2), 128), ((37, 4, 5), 128), ((64, 2, 9), 128), ((5, 5, 4), 128),
                                            ((33, 3, 3), 128), ((96, 2, 8), 128)])
def test_vert_adv_random_bit_exact(gt, oracle, size, alignment):
    ni, nj, nk = size
    rng = np.random.default_rng(ni + 10 * nj + 100 * nk)
    shape = (nk, nj + 6, ni + 6)
    arrs = [rng.uniform(5, 9, shape), rng.uniform(5, 9, shape), rng.uniform(-3e-4, 3e-4, shape),
            rng.uniform(5, 9, shape), rng.uniform(-1e-5, 1e-5, shape)]
    for cfg in (dict(), dict(variant=7), dict(variant=7, ctas_per_sm=-1, stages=2), dict(variant=3, ctas_per_sm=-3),
                dict(variant=7, bldg=1), dict(variant=7, bldg=1, ctas_per_sm=-1, stages=2), dict(variant=7, bldg=2),
                dict(variant=3, ctas_per_sm=-1, unroll=8), dict(variant=3, ctas_per_sm=-2, threads=32, save_upos=2),
                dict(variant=3, threads=64, save_upos=2), dict(variant=1), dict(variant=1, scratch=2, threads=32),
                dict(variant=1, ctas_per_sm=-2, threads=32)):
        for k in ("variant", "threads", "unroll", "scratch", "ctas_per_sm", "save_upos", "stages", "stagger", "bldg"):
            gt.lib.set_option("va." + k, 0)
        set_va(gt, cfg)
        try:
            out, _ = run_va(gt, arrs, 0.15, alignment)
        except gt.lib.GtbError as e:
            # an explicitly requested TMA variant refuses layouts TMA cannot address (auto falls back)
            assert cfg.get("variant") in (3, 7) and alignment == 1 and e.status == gt.lib.GTB_ERR_LAYOUT
            continue
        inner = (slice(None), slice(3, -3), slice(3, -3))
        assert np.array_equal(out[inner], oracle.vert_adv(*arrs, 0.15)[inner]), cfg


def test_vert_adv_full_size(gt, oracle):
    """BASELINE.json's headline config: 256x256x80 fp64."""
    ni = nj = 256
    nk = 80
    rng = np.random.default_rng(11)
    shape = (nk, nj + 6, ni + 6)
    arrs = [rng.uniform(5, 9, shape), rng.uniform(5, 9, shape), rng.uniform(-3e-4, 3e-4, shape),
            rng.uniform(5, 9, shape), rng.uniform(-1e-5, 1e-5, shape)]
    inner = (slice(None), slice(3, -3), slice(3, -3))
    want = oracle.vert_adv(*arrs, 0.15)[inner]
    for bldg in (0, 1, 2):  # u_pos of the backward sweep: default, register loads, TMA ring
        gt.lib.set_option("va.bldg", bldg)
        out, _ = run_va(gt, arrs, 0.15)
        assert np.array_equal(out[inner], want), bldg


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nk", [49, 50, 51, 78, 79, 97, 200, 700])
def test_vert_adv_tall_columns(gt, oracle, nk, dtype):
    """Columns taller than the TMEM window: the shared-memory slab grows with nk, and beyond what an SM holds the
    paired-warp kernel hands over to the L2-slab kernel (variant 3)."""
    ni, nj = 45, 3
    rng = np.random.default_rng(nk)
    shape = (nk, nj + 6, ni + 6)
    arrs = [rng.uniform(5, 9, shape).astype(dtype), rng.uniform(5, 9, shape).astype(dtype),
            rng.uniform(-3e-4, 3e-4, shape).astype(dtype), rng.uniform(5, 9, shape).astype(dtype),
            rng.uniform(-1e-5, 1e-5, shape).astype(dtype)]
    want = oracle.vert_adv(*arrs, 0.15)
    inner = (slice(None), slice(3, -3), slice(3, -3))
    for cfg in (dict(), dict(variant=7), dict(variant=7, ctas_per_sm=-1), dict(variant=7, bldg=1), dict(variant=7, bldg=1, ctas_per_sm=-1),
                dict(variant=7, bldg=2), dict(variant=3), dict(variant=3, threads=32, stages=2, ctas_per_sm=-2)):
        for k in ("variant", "unroll", "stages", "ctas_per_sm", "bldg"):
            gt.lib.set_option("va." + k, 0)
        set_va(gt, cfg)
        try:
            out, _ = run_va(gt, arrs, 0.15)
        except gt.lib.GtbError as e:
            raise AssertionError("unexpected error %s for %r" % (e, cfg))
            continue
        assert np.array_equal(out[inner], want[inner]), cfg


def test_vert_adv_rejects_single_level(gt):
    a = np.zeros((1, 8, 8))
    with pytest.raises(gt.lib.GtbError):
        run_va(gt, [a] * 5, 0.15)


# ------------------------------------------------------------------------------------- tridiagonal / tracers
def test_tridiagonal_known_answer(gt, oracle, golden):
    g = golden("tridiagonal_12x33x6.npz")
    st = [gt.storage.from_numpy(g[k], (0, 0, 0)) for k in ("inf", "diag", "sup", "rhs")]
    out = gt.storage.from_numpy(np.zeros_like(g["inf"]), (0, 0, 0))
    gt.stencil.tridiagonal(*st, out)
    gt.torch.cuda.synchronize()
    assert np.allclose(out.to_numpy(), 1.0, rtol=0, atol=1e-14)          # tridiagonal.cpp:97
    o_out, o_sup, o_rhs = oracle.tridiagonal(g["inf"], g["diag"], g["sup"], g["rhs"])
    assert np.array_equal(out.to_numpy(), o_out)
    assert np.array_equal(st[2].to_numpy(), o_sup) and np.array_equal(st[3].to_numpy(), o_rhs)
    assert rel_err(out.to_numpy(), g["out_ref"]) < TOL64 and rel_err(st[3].to_numpy(), g["rhs_ref"]) < TOL64


@pytest.mark.parametrize("size", [(23, 11, 6), (40, 3, 1), (64, 8, 37)])
def test_tridiagonal_random(gt, oracle, size):
    ni, nj, nk = size
    rng = np.random.default_rng(ni)
    shape = (nk, nj, ni)
    inf, sup = rng.uniform(-1, 0, shape), rng.uniform(0, 1, shape)
    diag, rhs = rng.uniform(3, 4, shape), rng.standard_normal(shape)
    st = [gt.storage.from_numpy(a, (0, 0, 0), 1) for a in (inf, diag, sup, rhs)]
    out = gt.storage.from_numpy(np.zeros(shape), (0, 0, 0), 1)
    gt.stencil.tridiagonal(*st, out)
    o_out, o_sup, o_rhs = oracle.tridiagonal(inf, diag, sup, rhs)
    assert np.array_equal(out.to_numpy(), o_out)
    assert np.array_equal(st[2].to_numpy(), o_sup) and np.array_equal(st[3].to_numpy(), o_rhs)


@pytest.mark.parametrize("n_tracers,shape,alignment", [(11, (6, 10, 24), 128), (2, (3, 5, 7), 1), (17, (2, 4, 8), 128),
                                                       (0, (2, 2, 2), 128)])
def test_prepare_tracers(gt, oracle, n_tracers, shape, alignment):
    rng = np.random.default_rng(n_tracers)
    rho = rng.standard_normal(shape)
    ins = [np.full(shape, 1.1 * i) + rng.standard_normal(shape) for i in range(n_tracers)]
    s_rho = gt.storage.from_numpy(rho, (0, 0, 0), alignment)
    s_in = [gt.storage.from_numpy(a, (0, 0, 0), alignment) for a in ins]
    s_out = [gt.storage.from_numpy(np.zeros(shape), (0, 0, 0), alignment) for _ in ins]
    gt.stencil.prepare_tracers(s_out, s_in, s_rho)
    gt.torch.cuda.synchronize()
    for got, want in zip(s_out, oracle.prepare_tracers(ins, rho)):
        assert np.array_equal(got.to_numpy(), want)


def test_planned_calls_match_direct_calls(gt, oracle):
    """stencil.plan(): the pre-marshalled call enqueues the same kernel as the direct call."""
    rng = np.random.default_rng(5)
    inp = rng.standard_normal((6, 24, 70))
    coeff = rng.uniform(0, 0.05, inp.shape)
    st = [gt.storage.from_numpy(inp, (2, 2, 0)), gt.storage.from_numpy(coeff, (2, 2, 0)),
          gt.storage.from_numpy(np.zeros_like(inp), (2, 2, 0))]
    f = gt.stencil.plan("horizontal_diffusion", *st)
    f()
    gt.torch.cuda.synchronize()
    inner = (slice(None), slice(2, -2), slice(2, -2))
    assert np.array_equal(st[2].to_numpy()[inner], oracle.hori_diff(inp, coeff)[inner])
    shape = (12, 5 + 6, 40 + 6)
    arrs = [rng.uniform(5, 9, shape), rng.uniform(5, 9, shape), rng.uniform(-3e-4, 3e-4, shape),
            rng.uniform(5, 9, shape), rng.uniform(-1e-5, 1e-5, shape)]
    sv = [gt.storage.from_numpy(a, (3, 3, 0)) for a in arrs]
    g = gt.stencil.plan("vertical_advection_dycore", *sv, dtr_stage=0.15)
    g()
    gt.torch.cuda.synchronize()
    inner = (slice(None), slice(3, -3), slice(3, -3))
    assert np.array_equal(sv[0].to_numpy()[inner], oracle.vert_adv(*arrs, 0.15)[inner])
    with pytest.raises(ValueError):
        gt.stencil.plan("no_such_spec")


def test_launches_are_counted(gt):
    before = gt.lib.launch_count()
    a = np.zeros((2, 8, 16))
    src, dst = gt.storage.from_numpy(a, (0, 0, 0)), gt.storage.from_numpy(a, (0, 0, 0))
    gt.stencil.copy(src, dst)
    assert gt.lib.launch_count() == before + 1


# ------------------------------------------------------------------------------------- recorded sequences (gtb_seq)
def test_sequence_replays_a_two_stream_loop(gt, oracle):
    """A recorded loop (stencil on one stream, ordered against a second stream with events) gives the same fields as
    the direct calls, slice by slice and in one go."""
    torch = gt.torch
    rng = np.random.default_rng(5)
    ni, nj, nk = 70, 19, 5
    inp = rng.standard_normal((nk, nj + 4, ni + 4))
    coeff = rng.uniform(0, 0.05, inp.shape)
    want1 = oracle.hori_diff(inp, coeff)
    inner = (slice(None), slice(2, -2), slice(2, -2))
    a = gt.storage.from_numpy(inp, (2, 2, 0))
    c = gt.storage.from_numpy(coeff, (2, 2, 0))
    b = gt.storage.from_numpy(np.zeros_like(inp), (2, 2, 0))
    d = gt.storage.from_numpy(np.zeros_like(inp), (2, 2, 0))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h1, h2 = C.c_void_p(s1.cuda_stream), C.c_void_p(s2.cuda_stream)
    torch.cuda.synchronize()
    seq = gt.stencil.Sequence()
    seq.horizontal_diffusion(a, c, b, stream=h1)      # b = hd(a)
    seq.record(0, h1)
    seq.wait(h2, 0)
    seq.horizontal_diffusion(b, c, d, stream=h2)      # d = hd(b) on the other stream, after the event
    seq.record(1, h2)
    assert len(seq) == 5
    seq.run(0, 2)
    seq.run(2, 3)
    torch.cuda.synchronize()
    got_b = b.to_numpy()
    assert np.array_equal(got_b[inner], want1[inner])
    want2 = oracle.hori_diff(got_b, coeff)
    assert np.array_equal(d.to_numpy()[inner], want2[inner])
    d.host_view()[...] = 0
    b.host_view()[...] = 0
    b.const_target_tensor(), d.const_target_tensor()
    torch.cuda.synchronize()
    seq.run()
    torch.cuda.synchronize()
    b._host_stale = d._host_stale = True
    assert np.array_equal(b.to_numpy()[inner], want1[inner]) and np.array_equal(d.to_numpy()[inner], want2[inner])
    with pytest.raises(gt.lib.GtbError):
        seq.run(3, 9)
    seq.close()


def test_sequence_vertical_advection(gt, oracle):
    rng = np.random.default_rng(6)
    ni, nj, nk = 40, 6, 17
    shape = (nk, nj + 6, ni + 6)
    arrs = [rng.uniform(5, 9, shape), rng.uniform(5, 9, shape), rng.uniform(-3e-4, 3e-4, shape),
            rng.uniform(5, 9, shape), rng.uniform(-1e-5, 1e-5, shape)]
    for dtype in (np.float64, np.float32):
        xs = [x.astype(dtype) for x in arrs]
        st = [gt.storage.from_numpy(x, (3, 3, 0)) for x in xs]
        seq = gt.stencil.Sequence()
        seq.vertical_advection_dycore(*st, 0.15, stream=None)
        seq.run()
        gt.torch.cuda.synchronize()
        st[0]._host_stale = True
        inner = (slice(None), slice(3, -3), slice(3, -3))
        assert np.array_equal(st[0].to_numpy()[inner], oracle.vert_adv(*xs, 0.15)[inner])


def test_vert_adv_concurrent_streams(gt, oracle):
    """Two launches of the (ticket-scheduled) vertical advection kernel that run at the same time on different streams
    must not share their strip counters."""
    torch = gt.torch
    rng = np.random.default_rng(8)
    ni, nj, nk = 96, 40, 80
    shape = (nk, nj + 6, ni + 6)
    problems = []
    for _ in range(4):
        arrs = [rng.uniform(5, 9, shape), rng.uniform(5, 9, shape), rng.uniform(-3e-4, 3e-4, shape),
                rng.uniform(5, 9, shape), rng.uniform(-1e-5, 1e-5, shape)]
        st = [gt.storage.from_numpy(a, (3, 3, 0)) for a in arrs]
        for f in st:
            f.const_target_tensor()
        problems.append((arrs, st, torch.cuda.Stream()))
    torch.cuda.synchronize()
    for rep in range(3):
        for arrs, st, stream in problems:
            with torch.cuda.stream(stream):
                gt.stencil.vertical_advection_dycore(*st, 0.15)
    torch.cuda.synchronize()
    inner = (slice(None), slice(3, -3), slice(3, -3))
    for arrs, st, _ in problems:
        want = arrs[0]
        for rep in range(3):
            want = oracle.vert_adv(want, *arrs[1:], 0.15)
        st[0]._host_stale = True
        assert np.array_equal(st[0].to_numpy()[inner], want[inner])


def test_box_copies_between_host_mirror_and_device(gt):
    """gtb_copy_box_async through DataStore.update_{target,host}_box_async: only the box moves, in both directions."""
    rng = np.random.default_rng(3)
    for dtype in (np.float64, np.float32):
        box = rng.standard_normal((5, 13, 37)).astype(dtype)
        ds = gt.storage.from_numpy(np.zeros_like(box), (3, 3, 0))
        ds.const_target_tensor()                      # device = zeros
        ds.host_view()[...] = box                     # host = data, device stale
        lo, hi = (3, 2, 1), (30, 11, 4)
        n = ds.update_target_box_async(lo, hi)
        gt.torch.cuda.synchronize()
        assert n == 27 * 9 * 3 * np.dtype(dtype).itemsize
        dev = ds.const_target_tensor().cpu().numpy()[:, :, :37]
        want = np.zeros_like(box)
        want[1:4, 2:11, 3:30] = box[1:4, 2:11, 3:30]
        assert np.array_equal(dev, want)
        ds.target_tensor().mul_(2)                    # device = 2 * box inside the box
        ds._host_np[...] = -1
        ds.update_host_box_async((4, 3, 2), (20, 9, 3))
        gt.torch.cuda.synchronize()
        got = ds._host_np[:, :, :37]
        want = -np.ones_like(box)
        want[2:3, 3:9, 4:20] = 2 * box[2:3, 3:9, 4:20]
        assert np.array_equal(got, want)


def test_removed_variants_are_refused(gt):
    """va.variant 2 / 4 / 5 / 6 and hd.variant 3 (measured-slower experiments of round 1) no longer exist."""
    rng = np.random.default_rng(2)
    shape = (6, 8 + 6, 40 + 6)
    arrs = [rng.uniform(5, 9, shape) for _ in range(5)]
    for v in (2, 4, 5, 6):
        gt.lib.set_option("va.variant", v)
        with pytest.raises(gt.lib.GtbError) as e:
            run_va(gt, arrs, 0.15)
        assert e.value.status == gt.lib.GTB_ERR_ARG
    gt.lib.set_option("va.variant", 0)
    gt.lib.set_option("hd.variant", 3)
    inp = rng.standard_normal((4, 12, 40))
    with pytest.raises(gt.lib.GtbError) as e:
        si = gt.storage.from_numpy(inp, (2, 2, 0))
        gt.stencil.horizontal_diffusion(si, gt.storage.from_numpy(inp.copy(), (2, 2, 0)), gt.storage.from_numpy(np.zeros_like(inp), (2, 2, 0)))
    assert e.value.status == gt.lib.GTB_ERR_ARG
    gt.lib.set_option("hd.variant", 0)
