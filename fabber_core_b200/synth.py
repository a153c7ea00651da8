"""Synthetic volumes for the configurations in BASELINE.json.

All generators return a float32 tensor in the C-API voxel layout [T][N] (row t = volume t, voxel
index x fastest) on the requested torch device, filled row by row so that a 256^3 x 96 volume never
needs more than one extra [N] temporary. The seeds are the ones SURVEY.md names.
"""
import math

import numpy as np
import torch


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def poly_volume(n_voxels, n_times=64, degree=3, seed=1002, device="cpu"):
    """C2: y = sum_n c_n i^n (i = 1..T) + N(0, 1); c0~U(50,150), c1~U(-2,2), c2~U(-.05,.05), c3~U(-5e-4,5e-4)."""
    g = _gen(device, seed)
    lo = [50.0, -2.0, -0.05, -5e-4] + [0.0] * 8
    hi = [150.0, 2.0, 0.05, 5e-4] + [0.0] * 8
    coef = [lo[n] + (hi[n] - lo[n]) * torch.rand(n_voxels, generator=g, device=device, dtype=torch.float64)
            for n in range(degree + 1)]
    y = torch.empty((n_times, n_voxels), dtype=torch.float32, device=device)
    for t in range(n_times):
        i = float(t + 1)
        acc = torch.zeros(n_voxels, dtype=torch.float64, device=device)
        for n in range(degree + 1):
            acc += coef[n] * (i ** n)
        acc += torch.randn(n_voxels, generator=g, device=device, dtype=torch.float64)
        y[t] = acc.to(torch.float32)
    return y


def biexp_volume(n_voxels, n_times=96, dt=0.02, noise=0.02, seed=1003, device="cpu", smooth_shape=None,
                 voxel_offset=0):
    """C3 (and C5 with smooth_shape=(nx,ny,nz)): y = amp1 exp(-r1 t) + 0.5 exp(-6 t) + N(0, noise^2).

    Truth as in the reference's examples/test_biexp.py:17-22: amp1~U(0.5,1), r1~U(0.8,1), amp2=0.5, r2=6.
    With smooth_shape the truth fields are the smooth ones of SURVEY.md 8d (C5)."""
    g = _gen(device, seed)
    if smooth_shape is None:
        amp1 = 0.5 + 0.5 * torch.rand(n_voxels, generator=g, device=device, dtype=torch.float64)
        r1 = 0.8 + 0.2 * torch.rand(n_voxels, generator=g, device=device, dtype=torch.float64)
    else:
        nx, ny, nz = smooth_shape
        assert nx * ny * nz >= n_voxels + voxel_offset
        # voxel_offset: this block starts at that voxel of the (nx, ny, nz) volume - z-slabs of a larger volume
        idx = torch.arange(n_voxels, device=device) + voxel_offset
        x = (idx % nx).to(torch.float64)
        yy = ((idx // nx) % ny).to(torch.float64)
        z = (idx // (nx * ny)).to(torch.float64)
        amp1 = 0.75 + 0.25 * torch.cos(2 * math.pi * x / 64) * torch.cos(2 * math.pi * yy / 64)
        r1 = 0.9 + 0.1 * torch.sin(2 * math.pi * z / 64)
    y = torch.empty((n_times, n_voxels), dtype=torch.float32, device=device)
    for t in range(n_times):
        tt = t * dt
        acc = amp1 * torch.exp(-r1 * tt) + 0.5 * math.exp(-6.0 * tt)
        acc += noise * torch.randn(n_voxels, generator=g, device=device, dtype=torch.float64)
        y[t] = acc.to(torch.float32)
    return y


def ar_design(n_times=200):
    """C4 design: columns 1, t/T, sin(2 pi t/50), cos(2 pi t/50)."""
    t = np.arange(n_times, dtype=np.float64)
    return np.stack([np.ones(n_times), t / n_times, np.sin(2 * np.pi * t / 50), np.cos(2 * np.pi * t / 50)], axis=1)


def dual_echo_design(n_pairs=100):
    """Two-echo design, samples interleaved TE1 TE2 TE1 .. (2 n_pairs rows): a baseline, a slow BOLD-like
    regressor that doubles at the second echo, and a tag/control alternation that is weaker there."""
    t = np.arange(n_pairs, dtype=np.float64)
    slow = np.sin(2 * np.pi * t / 20)
    tag = np.where(t % 2 == 0, 1.0, -1.0)
    d = np.zeros((2 * n_pairs, 3))
    d[0::2] = np.stack([np.ones(n_pairs), slow, tag], axis=1)
    d[1::2] = np.stack([0.7 * np.ones(n_pairs), 2 * slow, 0.3 * tag], axis=1)
    return d


def dual_echo_volume(n_voxels, n_pairs=100, rho1=0.3, rho2=0.2, cross=0.4, seed=1006, device="cpu"):
    """y = design beta + noise; each echo has AR(1) noise of its own and the second echo also carries `cross`
    times the first echo's noise of the same time point (what ar1-cross-terms models)."""
    g = _gen(device, seed)
    design = torch.as_tensor(dual_echo_design(n_pairs), device=device)
    beta = 50.0 * torch.randn((3, n_voxels), generator=g, device=device, dtype=torch.float64)
    y = torch.empty((2 * n_pairs, n_voxels), dtype=torch.float32, device=device)
    e1 = torch.zeros(n_voxels, dtype=torch.float64, device=device)
    e2 = torch.zeros(n_voxels, dtype=torch.float64, device=device)
    for t in range(n_pairs):
        e1 = rho1 * e1 + torch.randn(n_voxels, generator=g, device=device, dtype=torch.float64)
        e2 = rho2 * e2 + cross * e1 + 2.0 * torch.randn(n_voxels, generator=g, device=device, dtype=torch.float64)
        y[2 * t] = (design[2 * t] @ beta + e1).to(torch.float32)
        y[2 * t + 1] = (design[2 * t + 1] @ beta + e2).to(torch.float32)
    return y


def linear_ar_volume(n_voxels, n_times=200, rho=0.3, seed=1004, device="cpu"):
    """C4: y = design beta + AR(1) noise (rho, unit innovations); beta ~ N(0, 100^2)^4."""
    g = _gen(device, seed)
    design = torch.as_tensor(ar_design(n_times), device=device)
    beta = 100.0 * torch.randn((4, n_voxels), generator=g, device=device, dtype=torch.float64)
    y = torch.empty((n_times, n_voxels), dtype=torch.float32, device=device)
    e = torch.zeros(n_voxels, dtype=torch.float64, device=device)
    for t in range(n_times):
        e = rho * e + torch.randn(n_voxels, generator=g, device=device, dtype=torch.float64)
        y[t] = (design[t] @ beta + e).to(torch.float32)
    return y
