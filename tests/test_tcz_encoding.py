"""CPU check of the ORB "Z" operand encoding (easysfm_b200/csrc/tc_layout.cuh): nvcc builds a host-only program that decodes the
FP8 bytes the pack kernel and the query writers emit and verifies q.t + q'.t' == 20480 + 2^15 * hamming + column in exact integers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_z_encoding_is_exact(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "tcz_encoding_check"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", str(exe), os.path.join(ROOT, "tests", "host", "tcz_encoding_check.cu")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "TCZ ENCODING OK" in out.stdout, out.stdout + out.stderr
