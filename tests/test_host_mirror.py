"""The C++ host mirror (openfhe-development_b200/host/dcrtpoly_b200.h) over the C ABI: the reference's core unit
tests for this path re-expressed in C++ (tests/cpp/host_mirror_test.cpp).  CPU tier: it compiles, links against the
in-tree library and fails loudly without a device.  GPU tier: it runs and passes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "openfhe-development_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "_host_mirror_test")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(PKG, "host"), src, "-L" + PKG, "-lofhe_b200",
                           "-Wl,-rpath," + PKG, "-o", EXE])
    return EXE


def test_host_mirror_compiles_links_and_has_no_cpu_fallback():
    exe = _build()
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_host_mirror_runs_reference_unit_tests():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
