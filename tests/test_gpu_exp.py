"""GPU: the table-based exponential used inside the forward-model hooks (csrc/vb_exp.cuh) against exact
references: at most 1 ULP (what CUDA documents for its own exp), over the whole range the kernels use it."""
import ctypes as C
from decimal import Decimal, getcontext

import numpy as np
import pytest

from fabber_core_b200 import device

pytestmark = pytest.mark.gpu


def test_exp_fast_within_one_ulp():
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(200000) * 40, rng.uniform(-707.9, 707.9, 100000), rng.uniform(-1e-3, 1e-3, 20000),
                        np.array([0.0, -0.0, 1e-300, -1e-300, -707.99, 707.99, np.log(2) / 128, -np.log(2) / 128])])
    L = device.lib()
    L.fabber_cuda_exp_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    dx = device.DeviceArray.from_host(x)
    df, dr = device.DeviceArray(x.shape, np.float64), device.DeviceArray(x.shape, np.float64)
    device.check(L.fabber_cuda_exp_probe(dx.ptr, df.ptr, dr.ptr, x.size, None), "exp probe")
    fast, ref = df.to_host(), dr.to_host()
    ulp = np.spacing(np.abs(ref))
    d = np.abs(fast - ref) / ulp
    # against CUDA's exp (itself <= 1 ULP): never more than 2 ULP apart, and identical most of the time
    assert d.max() <= 2.0
    assert np.mean(d == 0) > 0.8
    # against exact references on a sample: <= 1 ULP
    getcontext().prec = 50
    idx = rng.choice(x.size, 3000, replace=False)
    worst = 0.0
    for i in idx:
        exact = Decimal(float(x[i])).exp()
        err = abs(Decimal(float(fast[i])) - exact) / Decimal(float(np.spacing(float(exact))))
        worst = max(worst, float(err))
    assert worst <= 1.0, worst


def test_exp_fast_returns_zero_for_decayed_arguments():
    """x < -708 (a decay rate that has run away): 0 where exp() gives a denormal or 0 - less than 3e-308 of the
    amplitude it multiplies - so such voxels stay on the table pass instead of dragging their warp on to libm"""
    x = np.concatenate([np.linspace(-708.5, -2000.0, 500), np.array([-1e6, -1e300, -np.inf])])
    L = device.lib()
    L.fabber_cuda_exp_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    dx = device.DeviceArray.from_host(x)
    df, dr = device.DeviceArray(x.shape, np.float64), device.DeviceArray(x.shape, np.float64)
    device.check(L.fabber_cuda_exp_probe(dx.ptr, df.ptr, dr.ptr, x.size, None), "exp probe")
    fast, ref = df.to_host(), dr.to_host()
    assert np.all(fast == 0.0)
    assert np.all(ref < 3e-308)
