"""The reference's OWN test sources, compiled unchanged where they lie under /root/reference/tests (tests/cpp/Makefile,
target reference_sources) with two stand-ins for what this image lacks: tests/cpp/gtest_shim (googletest) and
oracle/mpi_shim (MPI, threads as ranks), plus the one block a maintainer adds to tests/include/{stencil,gcl}_select.hpp
(tests/cpp/select/).  The binaries are built in the build container and travel to the GPU box.

  gcl_reference_cpu    tests/regression/gcl/test_halo_exchange_3D.cpp on the reference's gcl::cpu: proves the stand-ins
  gcl_reference_b200   the same source with gcl_arch_t = gridtools::gcl::b200 (only the arch tag differs)
  regression_b200      tests/regression/*.cpp (19 sources) + tests/src/regression_main.cpp with
                       stencil_backend_t = gridtools::stencil::b200<>, no registration lines: the generic fused path
                       (incl. boundary_conditions.cpp with gcl_arch_t = gridtools::gcl::b200)
  regression_emulated  the same 18 stencil sources with stencil_backend_t = emulated::backend<> (g++ only): the very bodies
                       of the generic path -- register tiles, shared-memory tiles, register windows, chained sweeps -- run
                       by emulated CTAs on the host, no GPU
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def run(name, *args, timeout=900):
    exe = os.path.join(BUILD, name)
    if not os.path.exists(exe):
        pytest.skip("tests/_build/%s not built (make -C tests/cpp reference_sources in the build container)" % name)
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    tail = r.stdout[-6000:] + r.stderr[-2000:]
    return r.returncode, r.stdout, tail


def test_reference_gcl_test_passes_on_its_own_cpu_arch_through_the_shims():
    """halo_exchange_3D_all + halo_exchange_3D_generic (6 layouts x vector/variadic x 2^3 periodicities, 6 parameter
    sets) on gcl::cpu with 2 thread-ranks: the unmodified reference passes its own test on gtest_shim + mpi_shim."""
    rc, out, tail = run("gcl_reference_cpu", 2)
    assert rc == 0 and "ALL PASSED" in out, tail


def test_reference_regression_sources_pass_on_emulated_ctas():
    """tests/regression/*.cpp, unchanged, through the generic path's per-thread bodies on emulated CTAs (cpu_ifirst
    stores): float and double, both inlined domain sizes, verified by the reference's own verifier.  No GPU."""
    rc, out, tail = run("regression_emulated", timeout=600)
    assert rc == 0 and "[  FAILED  ]" not in out, tail
    assert "tests ran" in out and int(out.split("[==========] ")[-1].split(" tests ran")[0]) >= 50, tail


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [2, 4])
def test_reference_gcl_test_passes_on_the_b200_arch(ranks):
    """The same unmodified source with gcl_arch_t = gridtools::gcl::b200: reference class templates, reference ctor
    (periodicity, MPI_Comm), device storages; rank threads share the visible GPUs."""
    rc, out, tail = run("gcl_reference_b200", ranks)
    assert rc == 0 and "ALL PASSED" in out, tail


@pytest.mark.gpu
def test_reference_regression_sources_pass_on_the_b200_backend():
    """tests/regression/*.cpp through stencil::b200<> at the harness' inlined domain sizes (12x33x61, 23x11x43),
    float and double, verified by the reference's own verifier against its own analytic repositories."""
    rc, out, tail = run("regression_b200")
    assert rc == 0 and "[  FAILED  ]" not in out, tail
    assert "tests ran" in out and int(out.split("[==========] ")[-1].split(" tests ran")[0]) >= 40, tail


@pytest.mark.gpu
def test_reference_perftests_mode_on_the_b200_backend():
    """`perftests 256 256 80 3` (tests/src/regression_main.cpp:27-60): the command-line sized cases, timed by the
    harness' own timer_cuda, results verified."""
    rc, out, tail = run("regression_b200", 256, 256, 80, 3)
    assert rc == 0 and "[  FAILED  ]" not in out, tail
    assert '"outputs"' in out and "horizontal_diffusion" in out and "vertical_advection_dycore" in out, tail


@pytest.mark.gpu
def test_reference_regression_sources_bind_to_the_named_kernels():
    """regression_b200_named = the same unchanged sources with tests/cpp/register_reference_specs.hpp force-included
    (the GTB200_REGISTER_SPEC lines): every registered spec passes the shape check of b200_shapes.hpp and runs on its
    hand-written kernel, results verified by the reference's verifier."""
    exe = os.path.join(BUILD, "regression_b200_named")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/regression_b200_named not built")
    env = dict(os.environ, GTB200_TRACE_DISPATCH="1")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=env)
    tail = r.stdout[-5000:] + r.stderr[-3000:]
    assert r.returncode == 0 and "[  FAILED  ]" not in r.stdout, tail
    named = [l for l in r.stderr.splitlines() if "-> named kernel" in l]
    generic = [l for l in r.stderr.splitlines() if "-> generic path" in l]
    # copy, hori_diff, hori_diff_fused, simple_hori_diff, vert_adv (float + double each), tridiagonal (double),
    # prepare_tracers chunks of 2 and 1 (double; float has no kernel: generic)
    kinds = {l.split("kernel id ")[1].split(",")[0] for l in named}
    assert kinds >= {"1", "2", "3", "4", "5", "6", "7"}, (sorted(kinds), generic, tail)


@pytest.mark.gpu
@pytest.mark.parametrize("ranks", [1, 4])
def test_reference_boundary_unit_tests_pass_on_the_b200_arch(ranks):
    """tests/unit_tests/boundaries/test_boundary_conditions.cpp and test_distributed_boundaries.cpp, unchanged, with
    gcl_arch_t = gridtools::gcl::b200: user boundary functors (gtb200/boundaries/b200.hpp) and the reference's own
    distributed_boundaries<comm_traits<storage, gcl::b200, timer>> over the b200 halo exchange."""
    rc, out, tail = run("boundaries_reference_b200", ranks)
    assert rc == 0 and "ALL PASSED" in out, tail
