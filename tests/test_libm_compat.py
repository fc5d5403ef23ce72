"""The glibc-compatible expf/logf used by the kernels, checked on the host: dspnet_b200/csrc/libm_compat.h is
compiled as plain C++ (oracle/libm_exhaustive.cc) and compared with the platform libm.  The full 2^32 sweep takes
~20 s on 8 cores and was run for both glibc builds (0 mismatches); CI uses a stride."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("libm") / "libm_exhaustive")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-o", exe,
                           os.path.join(ROOT, "oracle", "libm_exhaustive.cc")])
    return exe


def _host_has_fma():
    with open("/proc/cpuinfo") as f:
        flags = f.read()
    return " fma " in flags and " avx2 " in flags


def test_matches_host_libm(checker):
    """Variant auto-detected the way glibc's ifunc resolver does it (FMA && AVX2)."""
    out = subprocess.run([checker, "1" if _host_has_fma() else "0", "37"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


def test_matches_non_fma_build(checker):
    env = dict(os.environ, GLIBC_TUNABLES="glibc.cpu.hwcaps=-AVX2,-FMA")
    out = subprocess.run([checker, "0", "37"], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stdout
