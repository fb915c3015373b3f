"""Index algebra of the GPU FFT convolution (fft16k.cuh), emulated on the host: the same
pass functions compiled as host code (scalar arithmetic, a loop in place of the CTA) against a
direct convolution in double; also checks that the shared-memory swizzle is conflict free.
No GPU needed (nvcc compiles the host side only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hosttest(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("fft16k") / "fft16k_hosttest")
    r = subprocess.run([nvcc, "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-diag-suppress", "20014", "-o", exe,
                        os.path.join(ROOT, "tools", "fft16k_hosttest.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("args", [["512"], ["1536"], ["4096"], ["8192"], ["0", "2"]])
def test_fft16k_passes_convolve(hosttest, args):
    """Half-tap counts of the CLI block sizes 1024/8192/16384, the plugin's 3072-tap FIR, and the
    two-partition form used for block size 32768."""
    r = subprocess.run([hosttest] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
