"""The C++ side of the drop-in boundary: include/gtb200/stencil/b200.hpp behind GridTools' own stencil::run.

* no GPU (build container only, needs the reference headers): a translation unit that runs a REGISTERED spec through
  stencil::b200 compiles with plain g++ -- the named path is pure host code over the C ABI;
* GPU: tests/_build/b200_regression (built in the build container by tests/cpp/Makefile from tests/cpp/*.cu against
  the unmodified reference headers) runs named and generic specs on the device and compares every one with the
  reference's cpu_ifirst backend in the same process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/include"
BIN = os.path.join(ROOT, "tests", "_build", "b200_regression")

HOST_TU = r"""
#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include <gridtools/storage/sid.hpp>
#include <gtb200/stencil/b200.hpp>
#include "functors.hpp"
GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff, user::lap_f<0>, user::flx_f<0>, user::fly_f<0>, user::out_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff_fused, user::fused_out_f<0>);
namespace gt = gridtools; namespace st = gridtools::stencil;
void g() {  // horizontal_diffusion_fused.cpp:86-95: (out, in, coeff)
    auto mk = [] { return gt::storage::builder<gt::storage::cpu_ifirst>.type<double>().dimensions(20, 20, 4).halos(2, 2, 0).build(); };
    auto h = gt::halo_descriptor(2, 2, 2, 17, 20);
    st::run_single_stage(user::fused_out_f<0>(), st::b200<>(), st::make_grid(h, h, st::axis<1>(4)), mk(), mk(), mk());
}
void f() {
    auto mk = [] { return gt::storage::builder<gt::storage::cpu_ifirst>.type<double>().dimensions(20, 20, 4).halos(2, 2, 0).build(); };
    auto h = gt::halo_descriptor(2, 2, 2, 17, 20);
    auto grid = st::make_grid(h, h, st::axis<1>(4));
    st::run(user::hori_diff_spec<double, 0>(), st::b200<>(), grid, mk(), mk(), mk());
}
"""


@pytest.mark.skipif(not os.path.exists(REF), reason="reference headers only exist in the build container")
def test_named_path_is_plain_host_code(tmp_path):
    src = tmp_path / "tu.cpp"
    src.write_text(HOST_TU)
    cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I" + REF, "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "tests", "cpp"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
def test_b200_tag_through_gridtools_frontend():
    if not os.path.exists(BIN):
        pytest.skip("tests/_build/b200_regression not built (make -C tests/cpp in the build container)")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout
    assert r.stdout.count(" ok ") >= 17


GENERIC_BIN = os.path.join(ROOT, "tests", "_build", "b200_generic")


@pytest.mark.gpu
def test_generic_paths_of_the_tag_on_the_device():
    """tests/_build/b200_generic: every case of tests/cpp/cases.hpp (17 specs) through st::b200<> (fused generic path),
    the stage-by-stage fallback and a second block geometry, on storage::gpu stores against cpu_ifirst."""
    if not os.path.exists(GENERIC_BIN):
        pytest.skip("tests/_build/b200_generic not built (make -C tests/cpp in the build container)")
    r = subprocess.run([GENERIC_BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout
    assert r.stdout.count(" ok ") >= 5 * 17


SELECT_TU = r"""
#include <cstring>
#include <gridtools/stencil/cartesian.hpp>
#include <gtb200/stencil/b200_select.hpp>
namespace st = gridtools::stencil;
using tag_t = st::b200<>;
using staged_t = st::b200<gtb200::default_stream, gtb200::stage_by_stage>;
static_assert(std::is_same<decltype(backend_storage_traits(tag_t())), gridtools::storage::gpu>::value, "");
static_assert(std::is_same<decltype(backend_timer_impl(staged_t())), gridtools::timer_cuda>::value, "");
static_assert(!decltype(backend_supports_icosahedral(tag_t()))::value, "");
static_assert(decltype(backend_supports_vertical_stencils(tag_t()))::value, "");
int main() { return std::strcmp(backend_name(tag_t()), "b200"); }
"""


@pytest.mark.skipif(not os.path.exists(REF), reason="reference headers only exist in the build container")
def test_harness_traits_of_the_tag(tmp_path):
    """What tests/include/stencil_select.hpp:129-231 asks of a backend tag (storage traits, timer, name, capability
    probes) is found by ADL on stencil::b200 (include/gtb200/stencil/b200_select.hpp)."""
    src = tmp_path / "sel.cpp"
    src.write_text(SELECT_TU)
    exe = tmp_path / "sel"
    cmd = ["/usr/bin/g++", "-std=c++17", "-I" + REF, "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           str(src), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([str(exe)]).returncode == 0


SHAPE_TU = r"""
#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/stencil/global_parameter.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/cpu_ifirst.hpp>
#include <gridtools/storage/sid.hpp>
#include <gtb200/stencil/b200.hpp>
#include "functors.hpp"
GTB200_REGISTER_SPEC(gtb200::kernel::copy, user::copy_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::hori_diff, user::lap_f<0>, user::flx_f<0>, user::fly_f<0>, user::out_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::simple_hori_diff, user::wlap_f<0>, user::divflux_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::vert_adv, user::va_forward_f<0>, user::va_backward_f<0>);
GTB200_REGISTER_SPEC(gtb200::kernel::tridiagonal, user::td_forward_f<0>, user::td_backward_f<0>);
namespace gt = gridtools; namespace st = gridtools::stencil; using gtb200::kernel;
// a backend tag that only asks: which named kernel would stencil::b200 bind this spec to?
template <kernel Expected>
struct probe {
    template <class Spec, class Grid, class DataStores>
    friend void gridtools_backend_entry_point(probe, Spec, Grid const &, DataStores) {
        static_assert(st::b200_backend::named_kernel_of<Spec, Grid, DataStores>::value == Expected, "binding");
    }
};
template <class T>
auto mk() { return gt::storage::builder<gt::storage::cpu_ifirst>.template type<T>().dimensions(20, 20, 8).halos(3, 3, 0).build(); }
void checks() {
    auto h = gt::halo_descriptor(3, 3, 3, 16, 20);
    auto grid = st::make_grid(h, h, st::axis<1>(8));
    // the reference shapes bind ...
    st::run(user::hori_diff_spec<double, 0>(), probe<kernel::hori_diff>(), grid, mk<double>(), mk<double>(), mk<double>());
    st::run(user::hori_diff_spec<float, 0>(), probe<kernel::hori_diff>(), grid, mk<float>(), mk<float>(), mk<float>());
    st::run_single_stage(user::copy_f<0>(), probe<kernel::copy>(), grid, mk<double>(), mk<double>());
    st::run(user::tridiagonal_spec<0>(), probe<kernel::tridiagonal>(), grid, mk<double>(), mk<double>(), mk<double>(), mk<double>(), mk<double>());
    auto vgrid = st::make_grid(h, h, user::va_axis_t(8));
    st::run(user::vert_adv_spec<double, 0>(), probe<kernel::vert_adv>(), vgrid, mk<double>(), mk<double>(), mk<double>(),
        mk<double>(), mk<double>(), st::global_parameter(0.15));
    // ... the Thomas solve in float does not (there is only gtb_tridiagonal_f64) ...
    st::run(user::tridiagonal_spec<0>(), probe<kernel::none>(), grid, mk<float>(), mk<float>(), mk<float>(), mk<float>(), mk<float>());
    // ... nor do the same functors with another wiring: in and coeff swapped in the last stage,
    st::run([](auto in, auto coeff, auto out) {
            GT_DECLARE_TMP(double, lap, flx, fly);
            return st::execute_parallel().ij_cached(lap, flx, fly).stage(user::lap_f<0>(), lap, in)
                .stage(user::flx_f<0>(), flx, in, lap).stage(user::fly_f<0>(), fly, in, lap)
                .stage(user::out_f<0>(), out, coeff, flx, fly, in); },
        probe<kernel::none>(), grid, mk<double>(), mk<double>(), mk<double>());
    // another order of the run() arguments,
    st::run([](auto out, auto in, auto coeff) {
            GT_DECLARE_TMP(double, lap, flx, fly);
            return st::execute_parallel().ij_cached(lap, flx, fly).stage(user::lap_f<0>(), lap, in)
                .stage(user::flx_f<0>(), flx, in, lap).stage(user::fly_f<0>(), fly, in, lap)
                .stage(user::out_f<0>(), out, in, flx, fly, coeff); },
        probe<kernel::none>(), grid, mk<double>(), mk<double>(), mk<double>());
    // temporaries that are not ij-cached,
    st::run([](auto in, auto coeff, auto out) {
            GT_DECLARE_TMP(double, lap, flx, fly);
            return st::execute_parallel().stage(user::lap_f<0>(), lap, in)
                .stage(user::flx_f<0>(), flx, in, lap).stage(user::fly_f<0>(), fly, in, lap)
                .stage(user::out_f<0>(), out, in, flx, fly, coeff); },
        probe<kernel::none>(), grid, mk<double>(), mk<double>(), mk<double>());
    // or a copy functor used twice.
    st::run([](auto a, auto b) { return st::execute_parallel().stage(user::copy_f<0>(), a, b).stage(user::copy_f<0>(), b, a); },
        probe<kernel::none>(), grid, mk<double>(), mk<double>());
}
"""


@pytest.mark.skipif(not os.path.exists(REF), reason="reference headers only exist in the build container")
def test_named_binding_needs_the_reference_shape(tmp_path):
    """GTB200_REGISTER_SPEC alone does not bind: the spec must be type-identical (temporaries renumbered) to the
    reference spec the kernel implements, built by the reference's own frontend (b200_shapes.hpp).  Same functors with
    swapped wiring, another run() argument order, without the caches, or an element type the kernel does not exist
    for fall through to the generic path."""
    src = tmp_path / "shape.cpp"
    src.write_text(SHAPE_TU)
    cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I" + REF, "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "tests", "cpp"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


FUSABLE_TU = r"""
#include <cstdio>
#include <gtb200/stencil/b200_fused.hpp>   // first: the header must be self-sufficient
#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/cpu_kfirst.hpp>
#include <gridtools/storage/sid.hpp>
#include "cases.hpp"
namespace st = gridtools::stencil;
namespace gt = gridtools;
struct probe {
    static bool &fusable() { static bool v; return v; }
    template <class Spec, class Grid, class DataStores>
    friend void gridtools_backend_entry_point(probe, Spec, Grid const &, DataStores) {
        fusable() = st::b200_backend::fused::fusable<Spec>::value;
    }
};
int main() {
    auto h = gt::halo_descriptor(0, 0, 0, 3, 4);
    auto grid = st::make_grid(h, h, cases::kc_axis_t(4));
    auto mk = [] { return gt::storage::builder<gt::storage::cpu_kfirst>.type<double>().dimensions(4, 4, 4).build(); };
    st::run([](auto in, auto out) { GT_DECLARE_TMP(double, acc);
        return st::execute_parallel().k_cached(acc).stage(cases::carry_f(), in, acc).stage(cases::emit_f(), acc, out); },
        probe(), grid, mk(), mk());
    bool parallel_with_k_cache = probe::fusable();
    st::run([](auto in, auto out) { GT_DECLARE_TMP(double, acc);
        return st::execute_forward().k_cached(acc).stage(cases::carry_f(), in, acc).stage(cases::emit_f(), acc, out); },
        probe(), grid, mk(), mk());
    bool sweep_with_k_cache = probe::fusable();
    std::printf("%d %d\n", parallel_with_k_cache, sweep_with_k_cache);
    return !(!parallel_with_k_cache && sweep_with_k_cache);
}
"""


@pytest.mark.skipif(not os.path.exists(REF), reason="reference headers only exist in the build container")
def test_fusable_predicate_and_header_self_sufficiency(tmp_path):
    """b200_fused.hpp compiles on its own as host code; the one shape the fused path declines (k caches inside a
    parallel multi-stage) is reported by fused::fusable, a sweep with a k cache is taken."""
    src = tmp_path / "fusable.cpp"
    src.write_text(FUSABLE_TU)
    exe = tmp_path / "fusable"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O0", "-I" + REF, "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "tests", "cpp"), str(src), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([str(exe)]).returncode == 0


EMU_BIN = os.path.join(ROOT, "tests", "_build", "fused_emulation")


def test_fused_generic_path_on_emulated_ctas():
    """tests/_build/fused_emulation (g++ only): the per-thread body of the fused generic path -- thread-to-point mapping,
    shared-memory tiles, register windows with fill / flush, k blocking -- run by emulated CTAs (one OpenMP team per
    CTA, `omp barrier` for `__syncthreads`) and compared with the reference's cpu_ifirst backend on 17 specs, eight
    block geometries / domain sizes / options; also one launch per multi-stage."""
    if not os.path.exists(EMU_BIN):
        pytest.skip("tests/_build/fused_emulation not built (make -C tests/cpp in the build container)")
    env = dict(os.environ, OMP_WAIT_POLICY="passive")
    r = subprocess.run([EMU_BIN], capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout and r.stdout.count(" ok ") >= 169


# ------------------------------------------------------------------------------------------------ gcl (C++ class)
GCL_BIN = os.path.join(ROOT, "tests", "_build", "gcl_regression")

GCL_TU = r"""
#include <gtb200/gcl/halo_exchange.hpp>
namespace gcl = gtb200::gcl;
int f(double *a, double *b) {
    gcl::proc_grid grid(gcl::proc_grid::dims_create(8), {false, true, false}, 3);
    gcl::halo_exchange_dynamic_ut<gcl::layout_map<2, 1, 0>, gcl::layout_map<0, 1, 2>, double> he(
        {false, true, false}, grid, gcl::file_channel("/tmp", "x", 3, 8));
    he.add_halo<0>(2, 2, 2, 65, 80);
    he.add_halo<1>(gcl::halo_descriptor{2, 2, 2, 33, 36});
    he.add_halo<2>(0, 0, 0, 79, 80);
    he.setup(2);
    he.pack(a, b); he.exchange(); he.unpack(a, b);
    he.post_receives(); he.do_sends(); he.start_exchange(); he.wait();
    int i, j, k; he.comm().coords(i, j, k);
    return he.comm().proc(0, 1, 0) + i;
}
"""


def test_gcl_header_is_plain_host_code(tmp_path):
    """include/gtb200/gcl/halo_exchange.hpp needs nothing but the C ABI header and the standard library."""
    src = tmp_path / "tu.cpp"
    src.write_text(GCL_TU)
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_gcl_proc_grid_matches_python_host():
    """The C++ proc_grid (coords, proc with periodicity, dims_create) against the Python mirror that the oracle
    pins (tests/test_gcl_host.py), through a tiny g++-built probe."""
    import itertools
    import tempfile
    from gridtools_b200 import gcl
    prog = r"""
#include <cstdio>
#include <gtb200/gcl/halo_exchange.hpp>
int main() {
    namespace gcl = gtb200::gcl;
    for (int n : {1, 2, 4, 6, 8, 12}) { auto d = gcl::proc_grid::dims_create(n); std::printf("D %d %d %d %d\n", n, d[0], d[1], d[2]); }
    int dims[2][3] = {{2, 4, 1}, {3, 2, 2}};
    bool pers[3][3] = {{false, false, false}, {true, false, true}, {true, true, true}};
    for (auto &dm : dims) for (auto &pr : pers) {
        int size = dm[0] * dm[1] * dm[2];
        for (int r = 0; r < size; ++r) {
            gcl::proc_grid g({dm[0], dm[1], dm[2]}, {pr[0], pr[1], pr[2]}, r);
            for (int i = -1; i <= 1; ++i) for (int j = -1; j <= 1; ++j) for (int k = -1; k <= 1; ++k)
                std::printf("P %d %d %d %d %d %d %d %d %d %d %d\n", dm[0], dm[1], dm[2], (int)pr[0], (int)pr[1], (int)pr[2], r, i, j, k, g.proc(i, j, k));
        }
    }
}
"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.cpp")
        open(src, "w").write(prog)
        exe = os.path.join(d, "p")
        r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-o", exe],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    n_checked = 0
    for line in out:
        t = line.split()
        v = [int(x) for x in t[1:]]
        if t[0] == "D":
            assert tuple(v[1:]) == gcl.ProcGrid.dims_create(v[0]), line
        else:
            g = gcl.ProcGrid(v[0:3], v[3:6], v[6])
            assert g.proc(*v[7:10]) == v[10], line
        n_checked += 1
    assert n_checked > 1000


@pytest.mark.gpu
def test_gcl_class_exchanges_like_the_reference_test():
    """tests/_build/gcl_regression: the C++ class with ranks as threads on one GPU, coordinate-stamp check of
    test_halo_exchange_3D.cpp:66-123 (expectation computed from global coordinates only)."""
    if not os.path.exists(GCL_BIN):
        pytest.skip("tests/_build/gcl_regression not built (make -C tests/cpp)")
    r = subprocess.run([GCL_BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout and r.stdout.count(" ok ") >= 7


# ------------------------------------------------------------------------------------------------ boundaries (C++ class)
BC_TU = r"""
#include <gtb200/boundaries/boundary.hpp>
namespace bd = gtb200::boundaries;
void f(double *a, double *b, float *c) {
    std::array<gtb_halo_desc, 3> h = {{{2, 2, 2, 33, 36}, {2, 2, 2, 17, 20}, {0, 0, 0, 4, 5}}};
    bd::make_boundary(h, bd::value_boundary<double>(3.5)).apply(a, b);
    bd::make_boundary(h, bd::zero_boundary<float>()).apply(c);
    bd::make_boundary(h, bd::copy_boundary(), [](int ei, int, int) { return ei < 0; }).apply_on(nullptr, a, b);
}
"""


def test_boundary_header_is_plain_host_code(tmp_path):
    src = tmp_path / "bc.cpp"
    src.write_text(BC_TU)
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
