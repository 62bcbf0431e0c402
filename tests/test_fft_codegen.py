"""The generated register FFTs (csrc/fft_gen.cuh) compiled with g++ and checked against numpy.
The text is the same one nvcc compiles: lm_f2.cuh maps the packed ops to plain C on the host."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "audio_classification_icbhi_b200", "csrc")

HARNESS = r"""
#include "fft_gen.cuh"
extern "C" void fft32(float* re, float* im){ float a[32], b[32]; for(int i=0;i<32;i++){a[i]=re[i];b[i]=im[i];} lm_fft32(a,b); for(int i=0;i<32;i++){re[i]=a[i];im[i]=b[i];} }
extern "C" void fft16(float* re, float* im){ float a[16], b[16]; for(int i=0;i<16;i++){a[i]=re[i];b[i]=im[i];} lm_fft16(a,b); for(int i=0;i<16;i++){re[i]=a[i];im[i]=b[i];} }
extern "C" void fft32_aos(float* re, float* im){ lm_f2 z[32]; for(int i=0;i<32;i++) z[i]=lm_pack(re[i],im[i]); lm_fft32_aos(z); for(int i=0;i<32;i++){re[i]=lm_lo(z[i]);im[i]=lm_hi(z[i]);} }
extern "C" void fft32_aos_from2(float* re, float* im){ lm_f2 z[32]; for(int r=0;r<16;r++){ lm_f2 a=lm_pack(re[r],im[r]), b=lm_pack(re[r+16],im[r+16]); z[r]=lm_add2(a,b); z[r+16]=lm_sub2(a,b);} lm_fft32_aos_from2(z); for(int i=0;i<32;i++){re[i]=lm_lo(z[i]);im[i]=lm_hi(z[i]);} }
extern "C" void fft32_soa(float* re, float* im){ lm_f2 pr[16], pi[16]; for(int m=0;m<16;m++){ pr[m]=lm_pack(re[2*m],re[2*m+1]); pi[m]=lm_pack(im[2*m],im[2*m+1]); } float a[32], b[32]; lm_fft32_soa(pr,pi,a,b); for(int i=0;i<32;i++){re[i]=a[i];im[i]=b[i];} }
"""


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("fft")
    src = d / "h.cpp"
    src.write_text(HARNESS)
    so = d / "h.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I", CSRC, str(src), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def test_generated_header_is_current():
    out = subprocess.run([sys.executable, os.path.join(CSRC, "gen_fft.py")], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(CSRC, "fft_gen.cuh")).read(), "run gen_fft.py > fft_gen.cuh"


@pytest.mark.parametrize("name,n", [("fft16", 16), ("fft32", 32), ("fft32_aos", 32), ("fft32_aos_from2", 32),
                                    ("fft32_soa", 32)])
def test_fft_matches_numpy(harness, name, n):
    f = getattr(harness, name)
    p = ctypes.POINTER(ctypes.c_float)
    rs = np.random.RandomState(0)
    worst = 0.0
    for trial in range(20):
        x = rs.standard_normal(n) + 1j * rs.standard_normal(n)
        if trial == 0:
            x = np.zeros(n, dtype=complex); x[1] = 1.0          # impulse: exercises every twiddle
        re, im = x.real.astype(np.float32), x.imag.astype(np.float32)
        ref = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64))
        f(re.ctypes.data_as(p), im.ctypes.data_as(p))
        worst = max(worst, np.abs((re + 1j * im) - ref).max() / np.abs(ref).max())
    assert worst < 5e-7


def test_packed_and_scalar_flavours_agree_bitwise_enough(harness):
    """Same butterfly network, same Linzer-Feig constants: flavours differ by a few ulp at most."""
    p = ctypes.POINTER(ctypes.c_float)
    rs = np.random.RandomState(3)
    x = rs.standard_normal(32) + 1j * rs.standard_normal(32)
    outs = []
    for name in ("fft32", "fft32_aos", "fft32_soa", "fft32_aos_from2"):
        re, im = x.real.astype(np.float32), x.imag.astype(np.float32)
        getattr(harness, name)(re.ctypes.data_as(p), im.ctypes.data_as(p))
        outs.append(re + 1j * im)
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])
    np.testing.assert_array_equal(outs[0], outs[3])
