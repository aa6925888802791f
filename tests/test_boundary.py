"""The drop-in boundary without a GPU: the C-ABI library loads, exports every symbol include/gtb200.h declares,
validates its arguments like the reference's asserts would, fails loudly (no fallback) when there is no device,
and the product never touches oracle/."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from gridtools_b200 import _lib, build
    build.build_library()
    return _lib


def test_every_declared_symbol_is_exported(L):
    header = open(os.path.join(ROOT, "include", "gtb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gtb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    handle = L.lib()
    for name in sorted(declared):
        assert hasattr(handle, name), "libgtb200.so does not export %s" % name
    assert declared == set(L.exported_symbols()), declared ^ set(L.exported_symbols())


def test_version_and_options(L):
    assert L.lib().gtb_version() == 100
    L.set_option("hd.variant", 1)
    assert L.get_option("hd.variant") == 1
    L.set_option("hd.variant", 0)
    with pytest.raises(L.GtbError) as e:
        L.set_option("no.such.option", 1)
    assert e.value.status == L.GTB_ERR_ARG and "no.such.option" in str(e.value)


def test_argument_validation_needs_no_device(L):
    h = L.lib()
    f = L.Field(0x1000, 1, 8, 64)
    null = L.Field(None, 1, 8, 64)
    assert h.gtb_hori_diff_f64(C.byref(null), C.byref(f), C.byref(f), 4, 4, 4, None) == L.GTB_ERR_ARG
    assert h.gtb_hori_diff_f64(C.byref(f), C.byref(f), C.byref(f), -1, 4, 4, None) == L.GTB_ERR_ARG
    strided = L.Field(0x1000, 2, 8, 64)
    g = L.Field(0x2000, 1, 8, 64)
    assert h.gtb_hori_diff_f64(C.byref(strided), C.byref(f), C.byref(g), 4, 4, 4, None) == L.GTB_ERR_LAYOUT
    assert "stride_i" in L.last_error()
    assert h.gtb_hori_diff_f64(C.byref(f), C.byref(g), C.byref(f), 4, 4, 4, None) == L.GTB_ERR_ARG  # out aliases in
    assert h.gtb_copy(C.byref(f), C.byref(g), 4, 4, 4, 3, None) == L.GTB_ERR_ARG
    assert h.gtb_vert_adv_f64(*[C.byref(f)] * 5, 0.15, 4, 4, 1, None) == L.GTB_ERR_ARG  # nk < 2
    assert h.gtb_tridiagonal_f64(*[C.byref(f)] * 4, C.byref(null), 4, 4, 4, None) == L.GTB_ERR_ARG
    assert h.gtb_halo_create(None, None, 0, 1, 8, None) == L.GTB_ERR_ARG
    assert h.gtb_halo_send_bytes(None, 0, 1) == 0
    assert h.gtb_seq_create(None) == L.GTB_ERR_ARG
    assert h.gtb_seq_size(None) == 0 and h.gtb_seq_destroy(None) == L.GTB_OK
    assert h.gtb_seq_run(None, 0, 0) == L.GTB_ERR_ARG
    assert h.gtb_seq_add_record(None, 0, None) == L.GTB_ERR_ARG


def test_no_cpu_fallback_without_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = L.lib()
    assert h.gtb_device_count() == 0
    f, g = L.Field(0x1000, 1, 8, 64), L.Field(0x2000, 1, 8, 64)
    st = h.gtb_copy(C.byref(f), C.byref(g), 4, 4, 4, 8, None)
    assert st == L.GTB_ERR_CUDA and "CUDA" in L.last_error()
    assert h.gtb_init(0) == L.GTB_ERR_CUDA
    seq = C.c_void_p()
    assert h.gtb_seq_create(C.byref(seq)) == L.GTB_ERR_CUDA  # a sequence only ever replays device work


def test_product_never_uses_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "gridtools_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(base, fn)).read()
                # imports, includes, dlopens -- a comment that names oracle/gt_oracle.c as the checker is fine
                if re.search(r"(from|import)\s+oracle|#include.*gt_oracle|libgtoracle|libgtref|pyoracle", text):
                    bad.append(fn)
    for fn in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", fn)
        if os.path.isfile(p) and re.search(r"#include.*gt_oracle|libgtoracle", open(p).read()):
            bad.append(fn)
    assert not bad, "product files reference the oracle: %s" % bad


def test_missing_library_is_a_loud_import_error(monkeypatch):
    from gridtools_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgtb200.so")
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.lib()
