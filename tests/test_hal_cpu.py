"""CPU tier for the B200 lattice backend (openfhe-development_b200/hal): the reference relinked against it exists, carries the
device entry points as undefined symbols of libofhe_b200.so (so the hot methods cannot be served by anything else), and
FAILS LOUDLY without a CUDA device -- there is no CPU fallback behind the HAL class."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAL = os.path.join(ROOT, "oracle", "_ref", "hal")


def _need(name):
    p = os.path.join(HAL, name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (oracle/Makefile.hal needs /root/reference)")
    return p


def test_relinked_reference_binds_the_c_abi():
    core, pke = _need("libOPENFHEcore_b200.so"), _need("libOPENFHEpke_b200.so")
    und = subprocess.run(["nm", "-D", "--undefined-only", core], capture_output=True, text=True, check=True).stdout
    for sym in ("ofhe_ntt_fwd", "ofhe_ntt_inv", "ofhe_ew_binop", "ofhe_bconv_apply", "ofhe_automorphism",
                "ofhe_switch_modulus", "ofhe_stream_create"):
        assert f" U {sym}" in und, sym
    und = subprocess.run(["nm", "-D", "--undefined-only", pke], capture_output=True, text=True, check=True).stdout
    for sym in ("ofhe_hks_modup", "ofhe_hks_inner_product", "ofhe_hks_approx_mod_down", "ofhe_hks_keyswitch_core",
                "ofhe_evk_create"):
        assert f" U {sym}" in und, sym
    # the rebound KeySwitchHYBRID members are the strong definitions of this repo's translation unit
    dem = subprocess.run(["nm", "-DC", "--defined-only", pke], capture_output=True, text=True, check=True).stdout
    assert "lbcrypto::KeySwitchHYBRID::KeySwitchCore(" in dem
    assert "lbcrypto::b200::DCRTPolyB200<" in dem or "DCRTPolyB200" in subprocess.run(
        ["nm", "-DC", core], capture_output=True, text=True, check=True).stdout


def test_no_cpu_fallback_behind_the_hal_class():
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present: covered by tests/test_gpu_hal.py")
    exe = _need("core_tests_b200")
    r = subprocess.run([exe, "--gtest_filter=UTNTT.*"], capture_output=True, text=True, timeout=120)
    out = r.stdout + r.stderr
    assert r.returncode != 0, "the reference's NTT tests passed without a GPU: something computed on the CPU"
    assert "no CUDA device" in out or "CUDA" in out, out[-2000:]
