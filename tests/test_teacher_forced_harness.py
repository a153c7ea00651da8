"""CPU: the teacher-forced, single-iteration harness of tests/parity.py run oracle against oracle (the FMA
build stands in for the device): it must accept an equally valid build of the same algorithm at every step of
BASELINE config 5 ('MMMM' bi-exponential) although the two builds' full trajectories drift apart by O(1), and
it must REJECT a wrong implementation (a perturbed one-iteration result), and compare() must refuse a
multi-iteration comparison whose floor-based tolerance exceeds the cap."""
import numpy as np
import pytest

import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import synth
from parity import CAP, compare, teacher_forced

C3 = dict(num_exps=2, dt=0.02, param_overrides={"r2": {"mean": 6.0}})


def c5_case():
    nx, ny, nz = 8, 6, 4
    y = synth.biexp_volume(nx * ny * nz, 96, 0.02, 0.02, seed=1005, smooth_shape=(nx, ny, nz)).numpy()
    idx = np.arange(nx * ny * nz)
    coords = np.ascontiguousarray(np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32))

    def mk(its, **extra):
        sp = abi.ProblemSpec("exp", 96, prior_types=list("MMMM"), need_f=True, max_iterations=its, allow_bad_voxels=True,
                             **dict(C3, **extra))
        sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
        return sp

    return mk, y, coords


def fma_as_device(spec, data, **kw):
    return oracle.run(spec, data, variant="fma", **kw)


def test_harness_accepts_an_equally_valid_build_at_every_step():
    mk, y, coords = c5_case()
    reps = teacher_forced(mk, y, 4, (0, 1, 3, 6, 9), "C5 MMMM (cpu harness)", spatial=True, variants=("ulp",),
                          device_run=fma_as_device, coords=coords)
    assert all(max(v for k, v in r["tolerance"].items()) <= CAP for r in reps)


def test_harness_rejects_a_wrong_iteration():
    mk, y, coords = c5_case()

    def wrong(spec, data, **kw):
        out = oracle.run(spec, data, **kw)
        out["mean"] = out["mean"] * (1 + 3e-5)   # a 3e-5 relative error: far below what the old floor rule let through
        return out

    with pytest.raises(AssertionError, match="parity outside tolerance|further from the extended-precision truth"):
        teacher_forced(mk, y, 4, (3,), "C5 MMMM wrong", spatial=True, variants=("ulp",), device_run=wrong, coords=coords)


def test_uncapped_full_trajectory_comparison_is_refused():
    mk, y, coords = c5_case()
    ref = oracle.run(mk(8), y, spatial=True, coords=coords)
    fma = oracle.run(mk(8), y, spatial=True, coords=coords, variant="fma")
    ulp = oracle.run(mk(8), y, spatial=True, coords=coords, variant="ulp")
    with pytest.raises(AssertionError, match="UNINFORMATIVE"):
        compare(fma, ref, 4, [ulp], label="C5 MMMM full trajectory")


def test_truth_criterion_on_ill_conditioned_c2():
    """C2's cubic normal equations: the tolerance is clamped at the cap, and the distances to the
    extended-precision truth are reported side by side"""
    y = synth.poly_volume(1500, 64, 3, seed=1002).numpy()
    mk = lambda: abi.ProblemSpec("poly", 64, degree=3, need_f=True)
    ref, fma, truth = oracle.run(mk(), y), oracle.run(mk(), y, variant="fma"), oracle.run(mk(), y, variant="ld")
    rep = compare(fma, ref, 4, [fma], truth=truth, label="C2 (cpu harness)")
    assert max(rep["tolerance"].values()) <= CAP
    assert set(rep["truth"]["oracle_vs_truth"]) == set(rep["truth"]["gpu_vs_truth"])
