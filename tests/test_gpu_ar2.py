"""GPU parity: AR(1) noise on two interleaved echoes (noisemodel_ar.cc with num-echoes=2, every ar1-cross-terms
setting) - the CUDA kernel of csrc/vb_voxelwise_ar2.cuh through the C ABI against the CPU oracle, whose two-echo
restatement is pinned on the reference's own code in tests/test_reference_build.py."""
import numpy as np
import pytest

import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device, synth
from parity import compare, teacher_forced

pytestmark = pytest.mark.gpu

CROSS = ["none", "same", "dual"]


def both(spec_kwargs, data, variants=("fma",)):
    kw = dict(spec_kwargs)
    model = kw.pop("model")
    T = data.shape[0]
    ref = oracle.run(abi.ProblemSpec(model, T, **kw), data)
    probes = [oracle.run(abi.ProblemSpec(model, T, **kw), data, variant=v) for v in variants]
    truth = oracle.run(abi.ProblemSpec(model, T, **kw), data, variant="ld")
    gpu = device.run(abi.ProblemSpec(model, T, **kw), data)
    return gpu, ref, probes, truth


@pytest.mark.parametrize("cross", CROSS)
@pytest.mark.parametrize("conv", ["maxits", "pointzeroone", "trialmode"])
def test_linear_two_echoes(cross, conv):
    y = synth.dual_echo_volume(300, 100, seed=2001).numpy()
    gpu, ref, probes, truth = both(dict(model="linear", design=synth.dual_echo_design(100), noise="ar", num_echoes=2,
                                        ar_cross_terms=cross, convergence=conv, need_f=True), y)
    compare(gpu, ref, 3, probes, truth=truth, label="two-echo linear AR1 %s %s" % (cross, conv))
    nA = {"none": 2, "same": 3, "dual": 4}[cross]
    assert gpu["noise"].shape[0] == abi.ar2_noise_fields(nA)
    # each echo's AR coefficient is recovered (truth 0.3 / 0.2; the cross coupling leaks into alpha2 without
    # cross terms), and with "dual" the second echo's cross-term alpha finds the 0.4 of the first echo's noise
    assert abs(np.median(gpu["noise"][4]) - 0.3) < 0.08
    if cross == "dual":
        assert abs(np.median(gpu["noise"][4 + 3]) - 0.4) < 0.15


@pytest.mark.parametrize("cross", CROSS)
def test_linear_two_echoes_one_iteration_from_every_state(cross):
    y = synth.dual_echo_volume(250, 100, seed=2002).numpy()
    design = synth.dual_echo_design(100)
    mk = lambda its, **extra: abi.ProblemSpec("linear", 200, design=design, noise="ar", num_echoes=2,
                                              ar_cross_terms=cross, need_f=True, max_iterations=its, **extra)
    teacher_forced(mk, y, 3, (0, 1, 4, 9), "two-echo linear AR1 %s" % cross)


@pytest.mark.parametrize("cross", ["none", "dual"])
def test_nonlinear_model_two_echoes(cross):
    """the numerical-Jacobian (and table-exponential) pass under the two-echo noise model: a bi-exponential read
    as 48 echo pairs. ARD on one amplitude exercises the prior's free-energy term on this path too."""
    y = synth.biexp_volume(300, 96, 0.02, 0.02, seed=2003).numpy()
    gpu, ref, probes, truth = both(dict(model="exp", num_exps=2, dt=0.02, param_overrides={"r2": {"mean": 6.0}},
                                        noise="ar", num_echoes=2, ar_cross_terms=cross, need_f=True,
                                        convergence="pointzeroone", prior_types=["N", "N", "A", "N"]), y,
                                   variants=("fma", "ulp"))
    compare(gpu, ref, 4, probes, truth=truth, label="two-echo biexp AR1 %s" % cross)


def test_poly_two_echoes_freduce_and_lm():
    rng = np.random.default_rng(2004)
    T, N = 60, 128
    i = np.arange(1, T + 1, dtype=np.float64)[:, None]
    y = (2.0 + 0.5 * i + 0.01 * i * i + 0.5 * rng.standard_normal((T, N))).astype(np.float32)
    for conv in ("freduce", "lm"):
        gpu, ref, probes, truth = both(dict(model="poly", degree=2, noise="ar", num_echoes=2, ar_cross_terms="same",
                                            convergence=conv, need_f=True), y)
        compare(gpu, ref, 3, probes, truth=truth, label="two-echo poly AR1 same %s" % conv)


def test_two_echoes_need_an_even_series_and_cross_terms_need_two_echoes():
    with pytest.raises(device.CudaError):
        device.run(abi.ProblemSpec("poly", 21, degree=1, noise="ar", num_echoes=2), np.ones((21, 4), dtype=np.float32))
    spec = abi.ProblemSpec("poly", 20, degree=1, noise="ar")
    spec.prob.ar_cross_terms = 2
    with pytest.raises(device.CudaError):
        device.run(spec, np.ones((20, 4), dtype=np.float32))
