"""Shared helpers for the parity tests: compare a CUDA run with the oracle on the same inputs.

Tolerances follow BASELINE.json north_star: posterior means and variances, noise parameters and free
energy within 1e-6 relative in FP64; iteration counts and status masks identical.
"""
import numpy as np

RTOL = 1e-6


def tri(i, j):
    return i * (i + 1) // 2 + j if i >= j else j * (j + 1) // 2 + i


def rel_err(a, b, scale=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    den = np.maximum(den, 1e-300)
    return np.abs(a - b) / den


def compare(gpu, ref, P, rtol=RTOL, check_f=True, label=""):
    """Returns a dict of max relative errors; raises AssertionError with a readable message."""
    ok = (ref["status"] == 0)
    assert np.array_equal(gpu["status"], ref["status"]), "%s status masks differ: gpu %s ref %s" % (
        label, np.unique(gpu["status"], return_counts=True), np.unique(ref["status"], return_counts=True))
    assert np.array_equal(gpu["iterations"], ref["iterations"]), "%s iteration counts differ at %d voxels" % (
        label, np.count_nonzero(gpu["iterations"] != ref["iterations"]))
    errs = {}
    std = np.sqrt(np.abs(np.stack([ref["cov"][tri(i, i)] for i in range(P)])))
    # means: relative to max(|mean|, posterior std) - a mean that is zero within its own uncertainty
    # has no meaningful relative error
    errs["mean"] = float(np.max(rel_err(gpu["mean"], ref["mean"], scale=std)[:, ok], initial=0.0))
    var_g = np.stack([gpu["cov"][tri(i, i)] for i in range(P)])
    var_r = np.stack([ref["cov"][tri(i, i)] for i in range(P)])
    errs["var"] = float(np.max(rel_err(var_g, var_r)[:, ok], initial=0.0))
    # off-diagonal covariances relative to the geometric mean of the variances
    worst = 0.0
    for i in range(P):
        for j in range(i):
            sc = np.sqrt(np.abs(var_r[i] * var_r[j]))
            worst = max(worst, float(np.max(rel_err(gpu["cov"][tri(i, j)], ref["cov"][tri(i, j)], scale=sc)[ok],
                                            initial=0.0)))
    errs["cov_offdiag"] = worst
    errs["noise"] = float(np.max(rel_err(gpu["noise"], ref["noise"])[:, ok], initial=0.0))
    if check_f:
        errs["F"] = float(np.max(rel_err(gpu["free_energy"], ref["free_energy"])[ok], initial=0.0))
    bad = {k: v for k, v in errs.items() if not (v <= rtol)}
    assert not bad, "%s parity outside %g: %s (all: %s)" % (label, rtol, bad, errs)
    return errs
