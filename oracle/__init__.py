"""TEST INFRASTRUCTURE ONLY - Python driver for the CPU oracle (oracle/vb_oracle.cc).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product (fabber_core_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from fabber_core_b200 import cuda_abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib(variant=""):
    """variant "" = the oracle proper; "fma" / "ulp" = noise-floor probes (same source with FMA
    contraction / with the model's exp() perturbed by <= 1 ULP, see oracle/Makefile); "ld" = the same
    source in 80-bit extended precision, the higher-precision "truth" of tests/parity.py."""
    if variant not in _LIBS:
        name = {"fma": "libvb_oracle_fma.so", "ulp": "libvb_oracle_ulp.so",
                "ld": "libvb_oracle_ld.so"}.get(variant, "libvb_oracle.so")
        path = os.path.join(_HERE, "_build", name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.vb_oracle_gammaln.restype = C.c_double
        L.vb_oracle_gammaln.argtypes = [C.c_double]
        L.vb_oracle_digamma.restype = C.c_double
        L.vb_oracle_digamma.argtypes = [C.c_double]
        _LIBS[variant] = L
    return _LIBS[variant]


def _ptr(a):
    return a.ctypes.data if a is not None else None


def ar2_noise_scale(noise, n_alphas):
    """Per-field scale for comparing the two-echo AR noise block (b1 c1 b2 c2, alpha means, packed alpha
    precisions): an alpha that is truly ~0 (no cross coupling in the data) or an off-diagonal precision that
    cancels has no meaningful RELATIVE error; the posterior std of that alpha, and sqrt(P_rr P_cc), are the
    natural units. Zero rows mean plain relative error (the Gammas)."""
    nA = n_alphas
    scale = np.zeros_like(noise)
    tri = lambda i, j: i * (i + 1) // 2 + j
    n = noise.shape[1]
    prec = np.zeros((n, nA, nA))
    for r in range(nA):
        for c in range(r + 1):
            prec[:, r, c] = prec[:, c, r] = noise[4 + nA + tri(r, c)]
    with np.errstate(all="ignore"):
        try:
            cov = np.linalg.inv(prec)
        except np.linalg.LinAlgError:
            cov = np.full_like(prec, np.nan)
    for i in range(nA):
        scale[4 + i] = np.sqrt(np.abs(cov[:, i, i]))
    for r in range(nA):
        for c in range(r + 1):
            scale[4 + nA + tri(r, c)] = np.sqrt(np.abs(prec[:, r, r] * prec[:, c, c]))
    return np.nan_to_num(scale, nan=0.0, posinf=0.0)


def run(spec, data, spatial=False, image_priors=None, coords=None, init_mean=None, init_cov=None,
        init_noise=None, variant="", lock_centre=None):
    """Run the oracle. data: float32 [T][N]. Returns dict of numpy arrays (see fabber_cuda.h layouts)."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    T, N = data.shape
    assert T == spec.n_times
    prob = spec.prob
    prob.n_voxels = N
    P, NN = spec.P, spec.NN
    out = {
        "mean": np.zeros((P, N)),
        "cov": np.zeros((spec.ncov, N)),
        "noise": np.zeros((NN, N)),
        "free_energy": np.zeros(N),
        "iterations": np.zeros(N, dtype=np.int32),
        "status": np.zeros(N, dtype=np.int32),
    }
    buf = abi.VbBuffers()
    keep = [data]
    buf.data = data.ctypes.data
    if image_priors:
        for k, img in image_priors.items():
            arr = np.ascontiguousarray(img, dtype=np.float64)
            keep.append(arr)
            buf.image_prior[k] = arr.ctypes.data
    for name, arr in (("init_mean", init_mean), ("init_cov", init_cov), ("init_noise", init_noise),
                      ("lock_centre", lock_centre)):
        if arr is not None:
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            keep.append(arr)
            setattr(buf, name, arr.ctypes.data)
    if coords is not None:
        coords = np.ascontiguousarray(coords, dtype=np.int32)
        keep.append(coords)
        buf.coords = coords.ctypes.data
    if prob.f_history_len > 0:
        out["f_history"] = np.zeros((prob.f_history_len, N))
        buf.f_history = out["f_history"].ctypes.data
    if spatial:
        out["spatial_ak"] = np.zeros((prob.max_iterations + 1, P))
        buf.spatial_ak = out["spatial_ak"].ctypes.data
    buf.mean = out["mean"].ctypes.data
    buf.cov = out["cov"].ctypes.data
    buf.noise = out["noise"].ctypes.data
    buf.free_energy = out["free_energy"].ctypes.data
    buf.iterations = out["iterations"].ctypes.data
    buf.status = out["status"].ctypes.data
    fn = lib(variant).vb_oracle_spatial if spatial else lib(variant).vb_oracle_voxelwise
    fn.restype = C.c_int
    rc = fn(C.byref(prob), C.byref(buf))
    out["rc"] = rc
    out["n_times"] = T
    if spec.prob.noise_type == abi.NOISE_AR1 and spec.prob.n_phis == 2:
        out["noise_scale"] = ar2_noise_scale(out["noise"], spec.n_alphas)
    return out


def model_fit(spec, mean):
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    P, N = mean.shape
    spec.prob.n_voxels = N
    fit = np.zeros((spec.n_times, N))
    lib().vb_oracle_model_fit(C.byref(spec.prob), C.c_void_p(mean.ctypes.data), C.c_void_p(fit.ctypes.data))
    return fit


def neighbours(coords, spatial_dims=3):
    """coords int [3][N] -> (list of 1-based neighbour id lists, list of second-neighbour lists)"""
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    N = coords.shape[1]
    off1 = np.zeros(N + 1, dtype=np.int32)
    off2 = np.zeros(N + 1, dtype=np.int32)
    ids1 = np.zeros(max(1, 6 * N), dtype=np.int32)
    ids2 = np.zeros(max(1, 36 * N), dtype=np.int32)
    rc = lib().vb_oracle_neighbours(C.c_void_p(coords.ctypes.data), N, spatial_dims,
                                    C.c_void_p(off1.ctypes.data), C.c_void_p(ids1.ctypes.data), ids1.size,
                                    C.c_void_p(off2.ctypes.data), C.c_void_p(ids2.ctypes.data), ids2.size)
    if rc != 0:
        raise RuntimeError("neighbour calculation failed: %d" % rc)
    n1 = [list(ids1[off1[v]:off1[v + 1]]) for v in range(N)]
    n2 = [list(ids2[off2[v]:off2[v + 1]]) for v in range(N)]
    return n1, n2


def convergence_trace(name, F, max_its=10, fchange=0.01, max_trials=10):
    F = np.ascontiguousarray(F, dtype=np.float64)
    n = F.size
    t = np.zeros(n, dtype=np.int32)
    s = np.zeros(n, dtype=np.int32)
    r = np.zeros(n, dtype=np.int32)
    a = np.zeros(n, dtype=np.float32)
    lib().vb_oracle_convergence_trace(abi.CONV_BY_NAME[name], max_its, C.c_double(fchange), max_trials,
                                      C.c_void_p(F.ctypes.data), n, C.c_void_p(t.ctypes.data),
                                      C.c_void_p(s.ctypes.data), C.c_void_p(r.ctypes.data),
                                      C.c_void_p(a.ctypes.data))
    return t.astype(bool), s.astype(bool), r.astype(bool), a
