"""CPU emulation of the Stockham FFT index arithmetic shared with the CUDA kernels (csrc/fft_core.cuh)."""
import os
import subprocess


def test_fft_core_emulation(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fft_core_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tests", "cpu", "fft_core_test.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
