"""The kernel's branch-free sincos/asin (csrc/mcs_math.cuh) compiled for the host: every operation is an IEEE
add/mul/fma/div/sqrt in a fixed order, so the host build reproduces the device bit for bit; here it is held to
<= 2 ulp of libm over the argument ranges the kernel produces."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dm") / "libdm.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", out,
                    os.path.join(HERE, "device_math_host.cpp")], check=True)
    return C.CDLL(out)


def ulp_err(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_sincos_bf(mlib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-5 * np.pi, 5 * np.pi, 400_000), rng.uniform(-1e-3, 1e-3, 50_000),
                        np.arange(-10, 11) * (np.pi / 2), np.arange(-10, 11) * (np.pi / 2) + 1e-9, [0.0, -0.0, 1e-300]])
    s, c = np.zeros_like(x), np.zeros_like(x)
    P = C.POINTER(C.c_double)
    mlib.t_sincos(x.ctypes.data_as(P), s.ctypes.data_as(P), c.ctypes.data_as(P), C.c_long(len(x)))
    ws, wc = np.sin(x), np.cos(x)
    big = np.abs(ws) > 1e-12   # near zeros of sin/cos compare absolutely (libm's exact reduction vs 2-term Cody-Waite)
    assert ulp_err(s[big], ws[big]).max() <= 2.0
    assert np.abs(s[~big] - ws[~big]).max() < 2e-16
    big = np.abs(wc) > 1e-12
    assert ulp_err(c[big], wc[big]).max() <= 2.0
    assert np.abs(c[~big] - wc[~big]).max() < 2e-16
    assert np.abs(s * s + c * c - 1).max() < 5e-16


def test_asin_bf(mlib):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-1, 1, 400_000), rng.uniform(-1e-6, 1e-6, 20_000), 1 - 10.0 ** rng.uniform(-16, -1, 20_000),
                        [0.0, 0.5, -0.5, 1.0, -1.0, np.nextafter(1.0, 0), -np.nextafter(1.0, 0), np.nextafter(0.5, 1)]])
    y = np.zeros_like(x)
    P = C.POINTER(C.c_double)
    mlib.t_asin(x.ctypes.data_as(P), y.ctypes.data_as(P), C.c_long(len(x)))
    w = np.arcsin(x)
    nz = w != 0
    assert ulp_err(y[nz], w[nz]).max() <= 2.0
    assert np.all(np.sign(y) == np.sign(x)) and y[x == 0].tolist() == [0.0]
