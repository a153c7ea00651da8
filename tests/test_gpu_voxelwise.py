"""GPU parity: the CUDA voxelwise VB path (through the C ABI) against the CPU oracle and the
reference's golden outputs. Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest
import torch

import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device, synth
from parity import compare, tri

pytestmark = pytest.mark.gpu


# biexp with the reference's default priors starts both exponentials at the same rate: the fit then
# relies on rounding noise to break the symmetry and is chaotic in the reference itself (two CPU builds
# of the oracle disagree by O(1)). A prior mean on r2 (--PSP_byname1=r2
# --PSP_byname1_mean=6) makes the problem well posed; that is the C3 configuration used throughout.
C3 = dict(model="exp", num_exps=2, dt=0.02, param_overrides={"r2": {"mean": 6.0}})


class Probes(list):
    """the oracle's noise-floor runs, plus .truth: the extended-precision run compare() falls back on where
    the floor-based tolerance would exceed the cap (tests/parity.py)"""
    truth = None


def both(spec_kwargs, data, floor=True, **run_kwargs):
    """-> (gpu, oracle, [oracle noise-floor probes]) on the same inputs"""
    spec_kwargs = dict(spec_kwargs)
    model = spec_kwargs.pop("model")
    n_times = data.shape[0]
    ref = oracle.run(abi.ProblemSpec(model, n_times, **spec_kwargs), data, **run_kwargs)
    probes = None
    if floor:
        variants = ("fma", "ulp") if model == "exp" else ("fma",)
        probes = Probes(oracle.run(abi.ProblemSpec(model, n_times, **spec_kwargs), data, variant=vr, **run_kwargs)
                        for vr in variants)
        probes.truth = oracle.run(abi.ProblemSpec(model, n_times, **spec_kwargs), data, variant="ld", **run_kwargs)
    gpu = device.run(abi.ProblemSpec(model, n_times, **spec_kwargs), data, **run_kwargs)
    return gpu, ref, probes


def test_library_reports_device():
    assert device.lib().fabber_cuda_device_count() >= 1


def test_c1_linear_golden(golden):
    gpu, ref, fma = both(dict(model="linear", design=golden["design"]), golden["data"])
    compare(gpu, ref, 4, fma, truth=fma.truth, check_f=False, label="C1 linear")
    for i in range(4):
        g = golden["linear_vb/mean_Parameter_%d" % (i + 1)][0]
        assert np.max(np.abs(gpu["mean"][i] - g)) < 1e-3  # test/test_commandline.cc:10 ALLOWED_DELTA
        assert np.max(np.abs(gpu["mean"][i] - g) / np.abs(g)) < 5e-6
        z = gpu["mean"][i] / np.sqrt(gpu["cov"][tri(i, i)])
        gz = golden["linear_vb/zstat_Parameter_%d" % (i + 1)][0]
        assert np.max(np.abs(z - gz) / np.abs(gz)) < 5e-6


def test_c1_poly_golden(golden):
    gpu, ref, fma = both(dict(model="poly", degree=2), golden["data"])
    compare(gpu, ref, 3, fma, truth=fma.truth, check_f=False, label="C1 poly")
    for i in range(3):
        g = golden["poly/mean_c%d" % i][0]
        assert np.max(np.abs(gpu["mean"][i] - g) / np.abs(g)) < 5e-6
        gs = golden["poly/std_c%d" % i][0]
        assert np.max(np.abs(np.sqrt(gpu["cov"][tri(i, i)]) - gs) / gs) < 5e-6


@pytest.mark.parametrize("conv", ["maxits", "pointzeroone", "freduce", "trialmode", "lm"])
def test_c2_poly_synthetic(conv):
    y = synth.poly_volume(3000, 64, 3, seed=1002).numpy()
    gpu, ref, fma = both(dict(model="poly", degree=3, convergence=conv, need_f=True), y)
    compare(gpu, ref, 4, fma, truth=fma.truth, label="C2 poly %s" % conv)


@pytest.mark.parametrize("conv", ["maxits", "pointzeroone", "freduce", "trialmode", "lm"])
def test_c3_biexp_synthetic(conv):
    y = synth.biexp_volume(3000, 96, 0.02, 0.02, seed=1003).numpy()
    gpu, ref, fma = both(dict(C3, convergence=conv, need_f=True, allow_bad_voxels=True), y)
    compare(gpu, ref, 4, fma, truth=fma.truth, label="C3 biexp %s" % conv)


def test_c3_biexp_noisy_stress_masks_and_counts():
    """noise 0.1 (the reference example's level, examples/test_biexp.py) under LM: status masks and iteration
    counts must agree wherever the two CPU builds agree with each other. The posteriors of this trajectory are
    not compared end to end (the reference's own FP64 run is > 1e-5 from exact arithmetic, so no tolerance
    under the cap exists) - they are pinned iteration by iteration in tests/test_gpu_teacher_forced.py, as are
    the reference's default (symmetric) biexp priors."""
    y = synth.biexp_volume(2000, 96, 0.02, 0.1, seed=7).numpy()
    gpu, ref, fma = both(dict(C3, convergence="lm", need_f=True, allow_bad_voxels=True), y)
    stable = np.ones(y.shape[1], dtype=bool)
    for pr in fma:
        stable &= (pr["status"] == ref["status"]) & (pr["iterations"] == ref["iterations"])
    assert np.count_nonzero(~stable) <= 0.05 * stable.size
    assert np.array_equal(gpu["status"][stable], ref["status"][stable])
    assert np.array_equal(gpu["iterations"][stable], ref["iterations"][stable])


@pytest.mark.parametrize("degree", [0, 1, 4, 5])
def test_poly_other_sizes(degree):
    # T: short enough that the reference's own FP64 arithmetic resolves the normal equations of a quartic /
    # quintic in i = 1..T (at T = 40 its degree-4 result is 9e-6 from exact arithmetic, measured with the "ld" build)
    y = synth.poly_volume(500, {4: 28, 5: 24}.get(degree, 40), min(degree, 3), seed=11).numpy()
    gpu, ref, fma = both(dict(model="poly", degree=degree, need_f=True), y)
    compare(gpu, ref, degree + 1, fma, truth=fma.truth, label="poly degree %d" % degree)


def test_constant_data_recovers_value():
    """test/test_inference.cc:108-160: constant data -> mean == VAL to float precision."""
    y = np.full((10, 7), 7.32, dtype=np.float32)
    gpu, ref, fma = both(dict(model="poly", degree=0), y)
    assert np.allclose(gpu["mean"][0], np.float32(7.32), rtol=1e-6)
    compare(gpu, ref, 1, fma, truth=fma.truth, check_f=False, label="constant")


def test_noise_pattern_and_masked_timepoints():
    y = synth.poly_volume(800, 64, 2, seed=5).numpy()
    gpu, ref, fma = both(dict(model="poly", degree=2, noise_pattern="12", masked_timepoints=(3, 10, 64),
                              need_f=True, convergence="pointzeroone"), y)
    compare(gpu, ref, 3, fma, truth=fma.truth, label="pattern+mask")


def test_masked_timepoints_single_phi():
    y = synth.poly_volume(800, 64, 2, seed=6).numpy()
    y[4] = 1e4  # corrupt a sample, then mask it (test/test_inference.cc:485-560)
    gpu, ref, fma = both(dict(model="poly", degree=2, masked_timepoints=(5,), need_f=True), y)
    compare(gpu, ref, 3, fma, truth=fma.truth, label="mask")


def test_ard_and_image_priors():
    rng = np.random.default_rng(3)
    design = rng.standard_normal((50, 3))
    beta = rng.standard_normal((3, 600)) * np.array([[10.0], [0.0], [5.0]])
    y = (design @ beta + rng.standard_normal((50, 600))).astype(np.float32)
    img = beta[2] + 0.1 * rng.standard_normal(600)
    kw = dict(model="linear", design=design, prior_types=["N", "A", "I"], need_f=True,
              param_overrides={"Parameter_3": {"prec": 4.0}}, convergence="trialmode")
    gpu, ref, fma = both(kw, y, image_priors={2: img})
    compare(gpu, ref, 3, fma, truth=fma.truth, label="ARD+image")


def test_noise_options():
    y = synth.poly_volume(500, 64, 1, seed=8).numpy()
    gpu, ref, fma = both(dict(model="poly", degree=1, prior_noise_stddev=2.0, need_f=True), y)
    compare(gpu, ref, 2, fma, truth=fma.truth, label="prior-noise-stddev")
    gpu, ref, fma = both(dict(model="poly", degree=1, locked_noise_stdev=1.5, need_f=True), y)
    compare(gpu, ref, 2, fma, truth=fma.truth, label="locked-noise-stdev")


def test_restart_from_mvn():
    """continue-from-mvn (inference_vb.cc:181-216): second run starts from the first run's posterior."""
    y = synth.biexp_volume(500, 96, 0.02, 0.02, seed=9).numpy()
    kw = dict(C3, max_iterations=3, need_f=True)
    first = oracle.run(abi.ProblemSpec("exp", 96, **{k: v for k, v in kw.items() if k != "model"}), y)
    gpu, ref, fma = both(kw, y, init_mean=first["mean"], init_cov=first["cov"], init_noise=first["noise"])
    compare(gpu, ref, 4, fma, truth=fma.truth, label="restart")


def test_empty_volume_ok():
    """test/test_inference.cc:57-73: zero voxels is not an error."""
    spec = abi.ProblemSpec("poly", 10, degree=1)
    out = device.run(spec, np.zeros((10, 0), dtype=np.float32))
    assert out["rc"] == 0 and out["mean"].shape == (2, 0)


def test_bad_voxel_halts_by_default():
    y = synth.biexp_volume(64, 96, 0.02, 0.02, seed=10).numpy()
    y[:, 5] = np.inf
    gpu, ref, _ = both(dict(C3), y, floor=False)
    assert ref["rc"] == abi.ERR_BAD_VOXEL and gpu["rc"] == abi.ERR_BAD_VOXEL
    assert gpu["status"][5] == ref["status"][5] != 0


def test_invalid_arguments_are_rejected():
    spec = abi.ProblemSpec("poly", 10, degree=1, prior_types=["M", "N"])
    with pytest.raises(device.CudaError):
        device.run(spec, np.zeros((10, 4), dtype=np.float32))
    spec = abi.ProblemSpec("poly", 10, degree=1, max_iterations=0)
    with pytest.raises(device.CudaError):
        device.run(spec, np.zeros((10, 4), dtype=np.float32))


def test_full_size_properties_c2():
    """Size-independent properties at BASELINE's C2 size (128^3 x 64): the fit of voxel i depends on
    voxel i only, so a strided sample re-run as its own small volume must reproduce the big run bit for
    bit, and that sample must match the oracle."""
    n = 128 ** 3
    y = synth.poly_volume(n, 64, 3, seed=1002, device="cuda")
    spec = abi.ProblemSpec("poly", 64, degree=3)
    run = device.VbRun(spec, n)
    run.set_data_device(y.data_ptr())
    assert run.launch(torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    big = run.results()
    run.close()
    assert np.all(big["status"] == 0) and np.all(big["iterations"] == 10)
    pick = np.arange(0, n, 4099)
    ys = y[:, torch.as_tensor(pick, device="cuda")].cpu().numpy()
    small = device.run(abi.ProblemSpec("poly", 64, degree=3), ys)
    for k in ("mean", "cov", "noise"):
        assert np.array_equal(small[k], big[k][:, pick]), k
    ref = oracle.run(abi.ProblemSpec("poly", 64, degree=3), ys)
    fma = oracle.run(abi.ProblemSpec("poly", 64, degree=3), ys, variant="fma")
    compare(small, ref, 4, fma, check_f=False, label="C2 full-size sample")


# ---- AR(1) noise (noisemodel_ar.cc, num-echoes=1, ar1-cross-terms=none) ---------------------------
@pytest.mark.parametrize("conv", ["maxits", "pointzeroone", "trialmode", "lm"])
def test_c4_linear_ar1(conv):
    y = synth.linear_ar_volume(1500, 200, 0.3, seed=1004).numpy()
    gpu, ref, probes = both(dict(model="linear", design=synth.ar_design(200), noise="ar", convergence=conv,
                                 need_f=True), y)
    compare(gpu, ref, 4, probes, label="C4 linear AR1 %s" % conv)
    # the AR coefficient is recovered (truth 0.3)
    assert abs(np.median(gpu["noise"][2]) - 0.3) < 0.05


def test_poly_ar1_fit():
    """test/test_vb.cc:617-694: polynomial data + noise, AR noise model, coefficients within 0.2."""
    rng = np.random.default_rng(12)
    T, N = 50, 64
    i = np.arange(1, T + 1, dtype=np.float64)[:, None]
    y = (2.0 + 0.5 * i + 0.01 * i * i + 0.05 * rng.standard_normal((T, N))).astype(np.float32)
    gpu, ref, probes = both(dict(model="poly", degree=2, noise="ar", need_f=True), y)
    compare(gpu, ref, 3, probes, label="poly AR1")
    assert np.all(np.abs(gpu["mean"][0] - 2.0) < 0.2) and np.all(np.abs(gpu["mean"][1] - 0.5) < 0.2)


def test_ar1_with_masked_timepoints_is_rejected():
    """test/test_inference.cc:564-633: AR + masked time points must fail."""
    spec = abi.ProblemSpec("poly", 20, degree=1, noise="ar", masked_timepoints=(3,))
    with pytest.raises(device.CudaError):
        device.run(spec, np.ones((20, 4), dtype=np.float32))


def test_biexp_ar1():
    y = synth.biexp_volume(600, 96, 0.02, 0.02, seed=21).numpy()
    gpu, ref, probes = both(dict(C3, noise="ar", need_f=True, convergence="pointzeroone"), y)
    compare(gpu, ref, 4, probes, label="biexp AR1")


def _full_size_sample_check(spec_factory, y, P, label, stride):
    """Fit of voxel i depends on voxel i only: a strided sample re-run as its own small volume must
    reproduce the big run bit for bit, and that sample must match the oracle."""
    n = y.shape[1]
    run = device.VbRun(spec_factory(), n)
    run.set_data_device(y.data_ptr())
    assert run.launch(torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    big = run.results()
    run.close()
    assert np.all(big["status"] == 0), np.unique(big["status"], return_counts=True)
    pick = np.arange(0, n, stride)
    ys = y[:, torch.as_tensor(pick, device="cuda")].cpu().numpy()
    small = device.run(spec_factory(), ys)
    for k in ("mean", "cov", "noise", "iterations", "free_energy"):
        assert np.array_equal(small[k], big[k][..., pick]), k
    ref = oracle.run(spec_factory(), ys)
    probes = [oracle.run(spec_factory(), ys, variant=vr) for vr in ("fma", "ulp")]
    compare(small, ref, P, probes, label=label)
    return big


def test_full_size_properties_c3():
    """BASELINE configs[2] at full size: biexp, LM, 256^3 x 96 (6.4 GB of series resident in HBM)."""
    n = 256 ** 3
    y = synth.biexp_volume(n, 96, 0.02, 0.02, seed=1003, device="cuda")
    mk = lambda: abi.ProblemSpec("exp", 96, num_exps=2, dt=0.02, convergence="lm", need_f=True,
                                 param_overrides={"r2": {"mean": 6.0}})
    big = _full_size_sample_check(mk, y, 4, "C3 full-size sample", 16411)
    assert 1 <= big["iterations"].min() and big["iterations"].max() <= 140
    # truth recovered on average: amp1 ~ U(0.5, 1), r1 ~ U(0.8, 1)
    assert abs(np.mean(np.exp(big["mean"][0])) - 0.75) < 0.02 and abs(np.mean(np.exp(big["mean"][1])) - 0.9) < 0.03


def test_full_size_properties_c4():
    """BASELINE configs[3]: linear model, AR(1) noise, 256^3 x 200 over 8 GPUs = 2.1 M voxels per GPU."""
    n = 256 ** 3 // 8
    y = synth.linear_ar_volume(n, 200, 0.3, seed=1004, device="cuda")
    mk = lambda: abi.ProblemSpec("linear", 200, design=synth.ar_design(200), noise="ar", need_f=True)
    big = _full_size_sample_check(mk, y, 4, "C4 per-GPU-size sample", 2053)
    assert abs(np.median(big["noise"][2]) - 0.3) < 0.02   # AR coefficient recovered


@pytest.mark.parametrize("noise", ["white", "ar"])
def test_voxel_range_launches_tile_the_volume(noise):
    """fabber_cuda_vb_voxelwise_range: launching disjoint voxel ranges (uneven, not multiples of the CTA
    size, one of them empty) gives bit for bit what one launch over everything gives - what the host relies
    on when it starts each uploaded block of voxels on its own."""
    n, T = 1000, 40
    y = synth.poly_volume(n, T, 2, seed=77).numpy()
    kw = dict(degree=2, noise=noise, need_f=True, max_iterations=6)
    whole = device.run(abi.ProblemSpec("poly", T, **kw), y)
    run = device.VbRun(abi.ProblemSpec("poly", T, **kw), n)
    try:
        run.set_data(y)
        for a, b in ((300, 301), (0, 130), (130, 300), (301, 301), (301, 1000)):
            assert run.launch_range(a, b) == abi.OK
        run.sync()
        parts = run.results()
        for k in ("mean", "cov", "noise", "free_energy", "iterations", "status"):
            assert np.array_equal(parts[k], whole[k]), k
        assert run.launch_range(-1, 10) == abi.ERR_INVALID and run.launch_range(5, 1001) == abi.ERR_INVALID
        assert run.launch_range(10, 5) == abi.ERR_INVALID
    finally:
        run.close()


# ---- opt-in basis-row Jacobian (FABBER_B200_BASIS_JACOBIAN=1, recentre_loop in vb_voxelwise.cuh) ------------
@pytest.mark.parametrize("case", ["c2_poly", "c2_poly_lm", "c4_linear_ar1", "c1_linear", "linear_spatial"])
def test_basis_jacobian_option_stays_inside_the_parity_rule(case, golden, monkeypatch):
    """Models that are linear in their model-space parameters may form J as basis row x transform
    difference quotient instead of 2P+1 evaluations. Opt-in; held to the same parity rule as the default
    path, and it must actually be a different code path (results differ in the last bits)."""
    if case == "c2_poly" or case == "c2_poly_lm":
        y = synth.poly_volume(3000, 64, 3, seed=1002).numpy()
        kw = dict(model="poly", degree=3, need_f=True, convergence="lm" if case.endswith("lm") else "maxits")
        P, run_kw = 4, {}
    elif case == "c4_linear_ar1":
        y = synth.linear_ar_volume(1500, 200, 0.3, seed=1004).numpy()
        kw = dict(model="linear", design=synth.ar_design(200), noise="ar", need_f=True)
        P, run_kw = 4, {}
    elif case == "c1_linear":
        y = golden["data"]
        kw = dict(model="linear", design=golden["design"])
        P, run_kw = 4, {}
    else:
        nx, ny, nz = 10, 9, 5
        y = synth.poly_volume(nx * ny * nz, 40, 2, seed=35).numpy()
        idx = np.arange(nx * ny * nz)
        coords = np.ascontiguousarray(np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32))
        kw = dict(model="poly", degree=2, prior_types=list("pMm"), need_f=True, max_iterations=5)
        P, run_kw = 3, dict(spatial=True, coords=coords)
    spec_kw = dict(kw)
    model = spec_kw.pop("model")

    def mk():
        sp = abi.ProblemSpec(model, y.shape[0], **spec_kw)
        if "spatial" in run_kw:
            sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
        return sp

    ref = oracle.run(mk(), y, **run_kw)
    probes = [oracle.run(mk(), y, variant="fma", **run_kw)]
    default = device.run(mk(), y, **run_kw)
    monkeypatch.setenv("FABBER_B200_BASIS_JACOBIAN", "1")
    basis = device.run(mk(), y, **run_kw)
    monkeypatch.delenv("FABBER_B200_BASIS_JACOBIAN")
    compare(basis, ref, P, probes, check_f="need_f" in kw, label="basis jacobian %s" % case)
    assert np.mean(basis["iterations"] != default["iterations"]) < 0.01  # a voxel may sit on a detector threshold
    assert not np.array_equal(basis["mean"], default["mean"])


@pytest.mark.parametrize("noise", ["white", "ar"])
def test_linear_design_too_long_for_shared_memory(noise):
    """3000 samples x 4 columns = 96 KB of design: more than the kernels stage in shared memory next to their parked
    state, so the rows are read from global memory (one broadcast load per warp) - same arithmetic, same result"""
    rng = np.random.default_rng(31)
    T, N = 3000, 96
    t = np.arange(T, dtype=np.float64)
    design = np.stack([np.ones(T), t / T, np.sin(2 * np.pi * t / 50), np.cos(2 * np.pi * t / 300)], axis=1)
    beta = rng.standard_normal((4, N)) * 50
    y = (design @ beta + rng.standard_normal((T, N))).astype(np.float32)
    gpu, ref, probes = both(dict(model="linear", design=design, noise=noise, need_f=True, max_iterations=4), y)
    compare(gpu, ref, 4, probes, truth=probes.truth, label="linear, 3000-sample design, %s" % noise)


def test_six_parameter_models():
    """the widest hooks compiled (P = 6): 21-element packed covariance, 168-register kernels with the state parked"""
    rng = np.random.default_rng(32)
    T, N = 120, 256
    design = rng.standard_normal((T, 6))
    y = (design @ (rng.standard_normal((6, N)) * 10) + rng.standard_normal((T, N))).astype(np.float32)
    for noise in ("white", "ar"):
        gpu, ref, probes = both(dict(model="linear", design=design, noise=noise, need_f=True, convergence="pointzeroone"), y)
        compare(gpu, ref, 6, probes, truth=probes.truth, label="linear P6 %s" % noise)
