"""Host emulation of the register-level FFT building blocks (no GPU needed)."""
import os
import subprocess
import sys


def test_fft_core_on_cpu(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fft_core_test")
    src = os.path.join(root, "tests", "cpu", "fft_core_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0 and "ALL OK" in res.stdout


def test_fourstep_1d_path_on_cpu(tmp_path):
    """The index maps / twiddles / scrambled pairing of the fused 1-D kernels (kernels1d.cuh), emulated on the host with
    the same butterflies, against a long-double O(N^2) evaluation of fft(|ifft(X)|)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "fourstep_test")
    src = os.path.join(root, "tests", "cpu", "fourstep_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0 and "ALL OK" in res.stdout


def test_halfplane_3d_transform_on_cpu(tmp_path):
    """The two-pass 3-D transform of kernels3d.cuh (radix-2 along O inside the M-axis kernels, half-plane 2-D transforms,
    scrambled spatial positions), emulated on the host with the same butterflies, against O(N^2) DFTs."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "halfplane3d_test")
    src = os.path.join(root, "tests", "cpu", "halfplane3d_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, src], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0 and "ALL OK" in res.stdout
