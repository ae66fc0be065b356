"""The reference's OWN unit tests on the device path (SURVEY.md section 8 b/c).

oracle/Makefile.hal compiles the unmodified reference with this repo's lattice backend selected
(openfhe-development_b200/hal: lbcrypto::DCRTPoly = b200::DCRTPolyB200, KeySwitchHYBRID's four hot members rebound
to the fused device entry points) and builds the reference's gtest suites on top of it.  These tests run those
binaries on the GPU box and require (i) every selected test to pass and (ii) the backend's own tally of device work
(printed at exit when OFHE_B200_HAL_STATS is set) to show that the kernels, not the host mirror, did the arithmetic.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAL = os.path.join(ROOT, "oracle", "_ref", "hal")

pytestmark = pytest.mark.gpu


def run_suite(binary, gtest_filter, limit, need):
    exe = os.path.join(HAL, binary)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (oracle/Makefile.hal needs the reference sources; it travels prebuilt)")
    env = dict(os.environ, OFHE_B200_HAL_STATS="1", OMP_NUM_THREADS="8")
    p = subprocess.run([exe, f"--gtest_filter={gtest_filter}"], capture_output=True, text=True, timeout=limit, env=env)
    out = p.stdout + p.stderr
    tail = out[-3000:]
    assert p.returncode == 0, tail
    m = re.search(r"(\d+) tests? ran", out) or re.search(r"\[  PASSED  \] (\d+) test", out)
    assert "FAILED" not in out, tail
    s = re.search(r"\[ofhe_b200 hal\] device work: (.*)", out)
    assert s, "the backend printed no device-work tally:\n" + tail
    tally = {k: int(v) for k, v in (kv.split("=") for kv in s.group(1).split())}
    for key in need:
        assert tally[key] > 0, f"{key} == 0: the device path did not run ({tally})"
    return (int(m.group(1)) if m else None), tally


def test_core_ntt_transform_dcrtpoly():
    """UTNTT.* (UnitTestNTT.cpp:53-119), UTTransform.*, UTDCRTPoly.* (UnitTestDCRTElements.cpp): exact tests"""
    run_suite("core_tests_b200", "UTNTT.*:UTTransform.*:UTDCRTPoly.*", 600, ["limb_ntt", "ew_limbs"])


def test_bfvrns_crt_operations():
    """UTBFVRNS_CRT.* (UnitTestBFVrnsCRTOperations.cpp:195-694): exact vs BigInteger CRT interpolation"""
    run_suite("pke_tests_b200", "UTBFVRNS_CRT.*", 600, ["bconv", "limb_ntt"])


def test_general_evalmult():
    """UTGENERAL_EVALMULT (UnitTestEvalMult.cpp:656)"""
    run_suite("pke_tests_b200", "*UTGENERAL_EVALMULT*", 900, ["limb_ntt", "ew_limbs", "modup", "moddown"])


def test_ckksrns():
    """UTCKKSRNS.* (UnitTestCKKSrns.cpp): decrypt-within-eps on the device path"""
    run_suite("pke_tests_b200", "*UTCKKSRNS.*", 1500, ["limb_ntt", "ew_limbs", "modup", "moddown", "rescale", "automorph_limbs"])


def test_bgvrns():
    """UTBGVRNS.* (UnitTestBGVrns.cpp): BGV ModReduce and the t-scaled ModDown on the device path"""
    run_suite("pke_tests_b200", "*UTBGVRNS.*", 1200, ["limb_ntt", "rescale", "moddown"])


def test_bfvrns():
    """UTBFVRNS.* (UnitTestBFVrns.cpp)"""
    run_suite("pke_tests_b200", "*UTBFVRNS.*", 1200, ["limb_ntt", "bconv"])


def test_ckksrns_bootstrap():
    """UTCKKSRNS_BOOT (UnitTestBootstrap.cpp): EvalBootstrap = ModRaise + CoeffsToSlots / SlotsToCoeffs BSGS linear
    transforms over hoisted rotations + EvalChebyshev, all through the backend (BASELINE config 4's code path)"""
    run_suite("pke_tests_b200", "*UTCKKSRNS_BOOT*", 1800, ["limb_ntt", "modup", "moddown", "automorph_limbs"])


def test_ckksrns_serialization():
    """UTCKKSRNS_SER (UnitTestCKKSrnsSerialize.cpp): contexts, keys and ciphertexts written and read back through the
    backend class's save/load (same three fields, same order as dcrtpoly.h:337-351), then used on the device path"""
    run_suite("pke_tests_b200", "*UTCKKSRNS_SER*", 900, ["limb_ntt"])


def test_general_she():
    """UTGENERAL_SHE (UnitTestSHE.cpp): EvalAdd/Sub/Mult/Square/Sum/Merge/AtIndex over BFV (all four multiplication
    techniques), BGV and CKKS"""
    run_suite("pke_tests_b200", "*UTGENERAL_SHE*", 900, ["limb_ntt", "ew_limbs", "bconv"])


def test_ckksrns_automorphism():
    """UTCKKSRNS_AUTOMORPHISM (UnitTestCKKSrnsAutomorphism.cpp)"""
    run_suite("pke_tests_b200", "*UTCKKSRNS_AUTOMORPHISM*", 600, ["limb_ntt", "automorph_limbs", "moddown"])


def test_bgvrns_bfvrns_serialization():
    """UTBGVRNS_SER / UTBFVRNS_SER (UnitTestBGVrnsSerialize.cpp, UnitTestBFVrnsSerialize.cpp)"""
    run_suite("pke_tests_b200", "*UTBGVRNS_SER*:UTBFVRNS_SER*", 900, ["limb_ntt"])


def test_bootstrap_driver_config4_parameters_small_ring():
    """tools/hal_bootstrap.cpp (BASELINE config 4's parameters: simple-ckks-bootstrapping.cpp:63-156) at N = 2^12 on the
    device path: the bootstrapped ciphertext decrypts to the input within the scheme's precision."""
    import json

    exe = os.path.join(HAL, "bootstrap_b200")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built")
    env = dict(os.environ, OFHE_B200_HAL_STATS="1", OMP_NUM_THREADS="8")
    p = subprocess.run([exe, "--logn", "12", "--slots", "8"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, (p.stdout + p.stderr)[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["precision_bits"] > 12 and line["levels_remaining_after"] >= 9, line
    assert "device work" in p.stderr and "moddown=0 " not in p.stderr, p.stderr[-500:]
