"""The reference's OWN test-suite (test/testFFT.cpp, test/testIIR.cpp), compiled unmodified against the
drop-in headers include/sdsp/*.h + libsdsp_b200.so (oracle/Makefile target `reftests`, built in the build
container where /root/reference exists; the binary and its CSV fixtures travel under oracle/_ref/)."""
import os
import subprocess

import pytest

from tests.util import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_tests_dropin")


def test_reference_test_suite_passes_through_the_gpu_dropin():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_tests_dropin not built (needs /root/reference at build time)")
    # the tests open ../../../test_data/impulse_response relative to the working directory
    cwd = os.path.join(ROOT, "oracle", "_ref", "run", "a", "b")
    os.makedirs(cwd, exist_ok=True)
    r = subprocess.run([BIN, "--order", "rand", "--warn", "NoAssertions"], cwd=cwd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 failed" in r.stdout
    assert r.stdout.count("[ ok ]") >= 10  # 5 FFT + 5 IIR test cases


def test_additive_cpp_api_matches_the_reference_shaped_api():
    """tests/cpp/additions_test.cpp: batched / fp32 / real-input / iir_bank entry points of include/sdsp/*.h against the
    single-object calls the reference's tests pin (bit-exact where the arithmetic is the same, within tolerance else)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "additions_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/additions_test not built (make -C oracle additions)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "all additions ok" in r.stdout
