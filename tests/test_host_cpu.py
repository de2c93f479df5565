"""CPU: host-side logic -- scenes, replay format, FFT core simulation."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from blah2_b200.scene import make_scene, read_rspduo, write_rspduo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scene_is_seeded_and_int16():
    a = make_scene(5000, 2e6, seed=3)
    b = make_scene(5000, 2e6, seed=3)
    assert np.array_equal(a.x, b.x) and np.array_equal(a.y, b.y)
    for v in (a.x, a.y):
        assert np.all(v.real == np.round(v.real)) and np.all(np.abs(v.real) <= 32767)
        assert np.array_equal(v.astype(np.complex64).astype(np.complex128), v)   # exact in float32


def test_rspduo_roundtrip(tmp_path):
    sc = make_scene(1234, 2e6, seed=1)
    p = str(tmp_path / "t.rspduo")
    write_rspduo(p, sc.x, sc.y)
    assert os.path.getsize(p) == 1234 * 8          # int16 I1 Q1 I2 Q2 (TestAmbiguity.cpp:39-69)
    x, y = read_rspduo(p)
    assert np.array_equal(x, sc.x) and np.array_equal(y, sc.y)
    x2, _ = read_rspduo(p, 100)
    assert x2.shape[0] == 100


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_cta_fft_core_host_simulation(tmp_path):
    """blah2_b200/csrc/fft_core.cuh executed on the CPU (same __host__ __device__ code the
    kernels run), all plan sizes 2^8..2^14, float and double, against a long-double DFT."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_sim")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "fft_sim.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "FFT_SIM OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_cta_fft_dit_core_host_simulation(tmp_path):
    """blah2_b200/csrc/fft_dit.cuh (the fused-butterfly decimation-in-time core of the WienerHopf kernels) executed
    on the CPU: forward against a long-double DFT and forward -> inverse through the register hand-off, M = 512..4096,
    float and double."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_dit_sim")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "fft_dit_sim.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "FFT_DIT_SIM OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_split_toeplitz_solve_host_simulation(tmp_path):
    """blah2_b200/csrc/solve_steps.cuh (the per-step bodies of wh_solve_split_kernel: Schur row, pivot chain, queue
    entry, Levinson row) executed on the CPU in the kernel's shared-memory layout and block / boundary-warp
    structure: weights against a dense long-double Cholesky of the matrix of WienerHopf.cpp:85-97 (1 ... 448 taps)
    and the 'not positive definite' verdict for a first bad minor at the initial check, early, mid-way, at the last
    pivots (WienerHopf.cpp:111-117)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "solve_split_sim")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "solve_split_sim.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "SOLVE_SPLIT_SIM OK" in r.stdout, r.stdout + r.stderr
