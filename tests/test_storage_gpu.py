"""storage::b200 traits (include/gtb200/storage/b200.hpp) behind the reference's storage::builder / data_store, in a
plain host program (tests/cpp/storage_b200.cpp, g++ only) that also runs a registered spec through stencil::b200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "_build", "storage_b200")


@pytest.mark.gpu
def test_storage_b200_traits_round_trip_and_transfer_rate():
    if not os.path.exists(BIN):
        pytest.skip("tests/_build/storage_b200 not built (make -C tests/cpp in the build container)")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_storage_b200_program_is_plain_host_code():
    """Built by g++ (tests/cpp/Makefile): the traits and a registered spec need no nvcc.  Without a device the program
    says so and exits with 2 (there is no CPU fallback)."""
    if not os.path.exists(BIN):
        pytest.skip("tests/_build/storage_b200 not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no device" in r.stdout
