"""GPU: randomised configurations of the voxelwise paths - model x noise model (white with patterns / masked
samples, AR(1), two-echo AR(1) with every cross-term setting) x all five convergence detectors x N / ARD / image
prior mixes, and NLLS - through the CUDA kernels (C ABI) and through the oracle, under the same capped parity rule as
the hand-picked tests (tests/parity.py). The oracle side of the same draw is pinned on the reference's own code in
tests/test_reference_fuzz.py. Seeds are fixed."""
import numpy as np
import pytest

import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device
from parity import compare

pytestmark = pytest.mark.gpu

N = 96


def draw_model(rng, T):
    kind = rng.choice(["poly", "exp1", "linear"])
    if kind == "poly":
        deg = int(rng.integers(0, 4))
        i = np.arange(1, T + 1)[:, None]
        y = sum(rng.normal(0, 1, N) * (i / T) ** k * 5 for k in range(deg + 1)) + 10 + rng.normal(0, 1, (T, N))
        return "poly", deg + 1, y, dict(degree=deg)
    if kind == "exp1":
        t = np.arange(T) * 0.05
        y = rng.uniform(5, 10, N) * np.exp(-rng.uniform(0.5, 2, N) * t[:, None]) + rng.normal(0, 0.05, (T, N))
        return "exp", 2, y, dict(num_exps=1, dt=0.05)
    P = int(rng.integers(1, 7))
    design = rng.normal(0, 1, (T, P))
    y = design @ rng.normal(0, 5, (P, N)) + rng.normal(0, 1, (T, N))
    return "linear", P, y, dict(design=design)


def both(name, T, spec, y, image_priors=None):
    mk = lambda: abi.ProblemSpec(name, T, **spec)
    variants = ("fma", "ulp") if name == "exp" else ("fma",)
    ref = oracle.run(mk(), y, image_priors=image_priors)
    probes = [oracle.run(mk(), y, variant=v, image_priors=image_priors) for v in variants]
    truth = oracle.run(mk(), y, variant="ld", image_priors=image_priors)
    gpu = device.run(mk(), y, image_priors=image_priors)
    return gpu, ref, probes, truth


@pytest.mark.parametrize("seed", range(8))
def test_vb_random_configurations(seed):
    rng = np.random.default_rng(4000 + seed)
    for _ in range(6):
        T = 2 * int(rng.integers(10, 40))
        name, P, y, spec = draw_model(rng, T)
        y = y.astype(np.float32)
        conv = str(rng.choice(["maxits", "pointzeroone", "freduce", "trialmode", "lm"]))
        spec.update(convergence=conv, max_iterations=int(rng.integers(2, 9)), need_f=True, allow_bad_voxels=True)
        noise = rng.choice(["white", "white", "ar", "ar2"])
        if noise == "white":
            if rng.random() < 0.4:
                spec["noise_pattern"] = str(rng.choice(["12", "112", "21", "123"]))
            if rng.random() < 0.4:
                spec["masked_timepoints"] = tuple(sorted(set(int(x) for x in rng.integers(1, T + 1, 3))))
        elif noise == "ar":
            spec["noise"] = "ar"
        else:
            spec.update(noise="ar", num_echoes=2, ar_cross_terms=str(rng.choice(["none", "same", "dual"])))
        types = [str(rng.choice(["N", "N", "A", "I"])) for _ in range(P)]
        spec["prior_types"] = types
        # positive images: the exp model's parameters are log-transformed, and an image prior also seeds the initial
        # posterior (fwdmodel.cc:292-299) - a negative value there is a set-up failure that ends the reference's run
        images = {k: np.abs(rng.normal(0, 1, N)) + 0.5 for k, ty in enumerate(types) if ty == "I"}
        gpu, ref, probes, truth = both(name, T, spec, y, images or None)
        label = "fuzz %d: %s P%d T%d %s %s %s" % (seed, name, P, T, noise, conv, "".join(types))
        compare(gpu, ref, P, probes, truth=truth, label=label, max_ambiguous=0.25)


@pytest.mark.parametrize("seed", range(3))
def test_nlls_random_configurations(seed):
    rng = np.random.default_rng(5000 + seed)
    for _ in range(5):
        T = int(rng.integers(16, 60))
        name, P, y, spec = draw_model(rng, T)
        y = y.astype(np.float32)
        spec.update(method="nlls", nlls_lm=bool(rng.random() < 0.5))
        if rng.random() < 0.5:
            spec["masked_timepoints"] = tuple(sorted(set(int(x) for x in rng.integers(1, T + 1, 2))))
        gpu, ref, probes, truth = both(name, T, spec, y)
        label = "fuzz nlls %d: %s P%d T%d lm=%s" % (seed, name, P, T, spec["nlls_lm"])
        compare(gpu, ref, P, probes, truth=truth, check_f=False, label=label, max_ambiguous=0.25)


@pytest.mark.parametrize("seed", range(4))
def test_spatial_random_configurations(seed):
    """method=spatialvb on odd little volumes: random prior-type strings over M / m / P / p / N / A, holes in the
    mask, one to three spatial dimensions, update-on-first-iteration, a speed limit (the draw of
    tests/test_reference_fuzz.py, where the oracle side is pinned on the reference's own code)"""
    rng = np.random.default_rng(6000 + seed)
    for _ in range(6):
        nx, ny, nz = int(rng.integers(2, 9)), int(rng.integers(2, 7)), int(rng.integers(1, 6))
        n = nx * ny * nz
        T = int(rng.integers(12, 30))
        if rng.random() < 0.5:
            deg = int(rng.integers(0, 3))
            P = deg + 1
            i = np.arange(1, T + 1)[:, None]
            y = sum(rng.normal(0, 1, n) * (i / T) ** k * 5 for k in range(P)) + 10 + rng.normal(0, 1, (T, n))
            name, spec = "poly", dict(degree=deg)
        else:
            P = int(rng.integers(1, 4))
            design = rng.normal(0, 1, (T, P))
            y = design @ rng.normal(0, 5, (P, n)) + rng.normal(0, 1, (T, n))
            name, spec = "linear", dict(design=design)
        y = y.astype(np.float32)
        mask = rng.random((nx, ny, nz)) > 0.15
        if mask.sum() < 2:
            mask[:] = True
        sel = mask.reshape(-1, order="F")
        idx = np.arange(n)[sel]
        coords = np.ascontiguousarray(np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32))
        ys = np.ascontiguousarray(y[:, sel])
        types = [str(rng.choice(list("MmPpNA"))) for _ in range(P)]
        spec.update(prior_types=types, spatial_dims=int(rng.integers(1, 4)), max_iterations=int(rng.integers(2, 7)),
                    need_f=True, update_first_iter=bool(rng.random() < 0.5),
                    spatial_speed=float(rng.choice([-1.0, -1.0, 2.0, 10.0])))

        def mk():
            sp = abi.ProblemSpec(name, T, **spec)
            sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
            return sp

        ref = oracle.run(mk(), ys, spatial=True, coords=coords)
        if ref["rc"] != 0:
            continue   # a draw the reference itself cannot run (tests/test_reference_fuzz.py checks that agreement)
        probes = [oracle.run(mk(), ys, spatial=True, coords=coords, variant="fma")]
        truth = oracle.run(mk(), ys, spatial=True, coords=coords, variant="ld")
        gpu = device.run(mk(), ys, spatial=True, coords=coords)
        label = "fuzz spatial %d: %s P%d %dx%dx%d %s dims %d" % (seed, name, P, nx, ny, nz, "".join(types), spec["spatial_dims"])
        compare(gpu, ref, P, probes, truth=truth, label=label, max_ambiguous=0.25)
        assert np.max(np.abs(gpu["spatial_ak"] - ref["spatial_ak"]) / np.maximum(np.abs(ref["spatial_ak"]), 1e-300)) < 1e-5, label
