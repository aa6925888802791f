"""patches/gridtools-host-mirror-through-traits.patch: the reference-side change that lets a storage traits type own
the HOST mirror of a data_store (storage/data_store.hpp:89,109,114,178,187 allocate it with std::make_unique).  With it
`gridtools::storage::b200` hands out page-locked memory (storage_allocate_host -> gtb_host_malloc) and
gtb_staged_upload / gtb_staged_download copy from / to the mirror directly.  Here: the patch applies to a copy of the
reference headers, every traits type without storage_allocate_host keeps its plain array, storage::b200 gets the pinned
one, and the storage test program still compiles against the patched tree.  (No GPU: compile-time checks only.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/include"
PATCH = os.path.join(ROOT, "patches", "gridtools-host-mirror-through-traits.patch")

TU = r"""
#include <memory>
#include <type_traits>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/gpu.hpp>
#include <gtb200/storage/b200.hpp>
namespace gs = gridtools::storage;
// a traits type that says nothing about the host mirror keeps the plain array
static_assert(std::is_same<gs::traits::host_ptr_type<gs::gpu, double>, std::unique_ptr<double[]>>::value, "");
// storage::b200 provides the page-locked one
static_assert(std::is_same<gs::traits::host_ptr_type<gs::b200, double>,
                  std::unique_ptr<double[], gs::b200_impl_::host_free>>::value, "");
// and a data_store built with it compiles (mutable and const element types)
void instantiate() {
    auto a = gs::builder<gs::b200>.type<double>().dimensions(8, 4, 2).build();
    auto b = gs::builder<gs::b200>.type<double const>().dimensions(8, 4, 2).initializer([](int, int, int) { return 1.; }).build();
    (void)a->host_view();
    (void)b->const_host_view();
}
int main() { return 0; }
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference headers (build container)")
def test_host_mirror_patch_applies_and_b200_traits_pin_the_mirror(tmp_path):
    tree = tmp_path / "gridtools"
    shutil.copytree(REF, tree / "include")
    r = subprocess.run(["patch", "-p1", "-i", PATCH], cwd=tree, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = tmp_path / "tu.cpp"
    src.write_text(TU)
    inc = ["-I", str(tree / "include"), "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include"]
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only"] + inc + [str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    # the storage test program of this repository against the patched tree
    prog = os.path.join(ROOT, "tests", "cpp", "storage_b200.cpp")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-fopenmp"] + inc + [prog], capture_output=True,
        text=True)
    assert r.returncode == 0, r.stderr[-4000:]
