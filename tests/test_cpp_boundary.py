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
namespace gt = gridtools; namespace st = gridtools::stencil;
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
