"""GPU: the drop-in claim, literally. The same option strings and the same float32 volumes go through the
SAME C API of two libraries - libfabbercore_b200.so (this repo, CUDA) and oracle/_ref/libfabbercore_ref.so
(the reference's own sources on the test-only NEWMAT stand-in, CPU) - with the same ctypes wrapper, and
every output volume the reference produces must come back the same (float32 outputs; tolerance 2e-5
relative to the output's scale, the reference algorithm's own noise floor documented in DESIGN.md)."""
import numpy as np
import pytest

import refbuild
from fabber_core_b200 import fabber as fab
from fabber_core_b200 import synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refbuild.available(), reason="oracle/_ref not built")]

SAVE = {"save-mean": True, "save-std": True, "save-zstat": True, "save-var": True, "save-noise-mean": True,
        "save-noise-std": True, "save-mvn": True, "save-free-energy": True, "save-model-fit": True,
        "save-residuals": True}


def both(opts, data, mask=None):
    o = dict(opts)
    o.update(SAVE)
    ours = fab.Fabber().run_with_data(o, data, mask=mask)
    ref = refbuild.ReferenceFabber().run_with_data(o, data, mask=mask)
    assert sorted(ours.data.keys()) == sorted(ref.data.keys())
    return ours, ref


def assert_same(ours, ref, rtol=2e-5, skip=()):
    for k in ref.data:
        if k in skip:
            continue
        a, b = ours.data[k].astype(np.float64), ref.data[k].astype(np.float64)
        assert a.shape == b.shape, k
        if k.startswith("zstat_") or k.startswith("mean_"):
            # a mean is only determined to within its own posterior std: compare on that scale
            name = k.split("_", 1)[1]
            scale = np.maximum(np.abs(b), ref.data["std_" + name] if k.startswith("mean_") else 1.0)
        elif k == "finalMVN":
            scale = np.maximum(np.abs(b), np.max(np.abs(b), axis=(0, 1, 2), keepdims=True) * 1e-3)
        elif k in ("modelfit", "residuals"):
            scale = np.maximum(np.abs(ref.data["modelfit"].astype(np.float64)), 1e-3)
        else:
            scale = np.abs(b)
        err = np.max(np.abs(a - b) / np.maximum(scale, 1e-30))
        assert err < rtol, "%s differs: %g" % (k, err)


def test_dropin_poly_white_lm():
    nx, ny, nz = 6, 5, 4
    y = synth.poly_volume(nx * ny * nz, 40, 2, seed=71).numpy()
    ours, ref = both({"model": "poly", "degree": 2, "noise": "white", "method": "vb", "convergence": "lm"},
                     {"data": refbuild.volume(y, (nx, ny, nz))})
    assert_same(ours, ref)


def test_dropin_biexp_trialmode_with_mask():
    nx, ny, nz = 6, 5, 4
    y = synth.biexp_volume(nx * ny * nz, 96, 0.02, 0.02, seed=72).numpy()
    mask = np.ones((nx, ny, nz), dtype=np.int32)
    mask[1, :, 2] = 0
    mask[5, 4, :] = 0
    ours, ref = both({"model": "exp", "num-exps": 2, "dt": 0.02, "noise": "white", "method": "vb",
                      "convergence": "trialmode", "PSP_byname1": "r2", "PSP_byname1_mean": 6.0},
                     {"data": refbuild.volume(y, (nx, ny, nz))}, mask=mask)
    assert_same(ours, ref)
    assert np.all(ours.data["mean_amp1"][mask == 0] == 0)


def test_dropin_linear_ar1(tmp_path):
    nx, ny, nz = 5, 4, 3
    y = synth.linear_ar_volume(nx * ny * nz, 120, 0.3, seed=73).numpy()
    basis = str(tmp_path / "ar.mat")
    np.savetxt(basis, synth.ar_design(120), fmt="%.17g")
    ours, ref = both({"model": "linear", "basis": basis, "noise": "ar", "method": "vb", "convergence": "pointzeroone"},
                     {"data": refbuild.volume(y, (nx, ny, nz))})
    assert_same(ours, ref)


@pytest.mark.parametrize("cross", ["none", "same", "dual"])
def test_dropin_linear_ar1_two_echoes(cross, tmp_path):
    """num-echoes=2: every output volume, including noise_means / noise_stdevs (TWO volumes: alpha1 and alpha2,
    because Ar1cNoiseModel::NumParams() returns nPhis) and the finalMVN with its (alphas, phi1, phi2) noise block"""
    nx, ny, nz = 5, 4, 3
    y = synth.dual_echo_volume(nx * ny * nz, 60, seed=76).numpy()
    basis = str(tmp_path / "de.mat")
    np.savetxt(basis, synth.dual_echo_design(60), fmt="%.17g")
    ours, ref = both({"model": "linear", "basis": basis, "noise": "ar", "num-echoes": 2, "ar1-cross-terms": cross,
                      "method": "vb", "convergence": "pointzeroone"}, {"data": refbuild.volume(y, (nx, ny, nz))})
    assert_same(ours, ref)
    nA = {"none": 2, "same": 3, "dual": 4}[cross]
    n_all = 3 + nA + 2
    assert ours.data["noise_means"].shape[-1] == 2
    assert ours.data["finalMVN"].shape[-1] == n_all * (n_all + 1) // 2 + n_all + 1


@pytest.mark.parametrize("types", ["M+", "P+", "MA"])
def test_dropin_spatialvb(types):
    nx, ny, nz = 6, 5, 4
    y = synth.poly_volume(nx * ny * nz, 30, 1, seed=74).numpy()
    ours, ref = both({"model": "poly", "degree": 1, "noise": "white", "method": "spatialvb",
                      "param-spatial-priors": types, "max-iterations": 5},
                     {"data": refbuild.volume(y, (nx, ny, nz))})
    assert_same(ours, ref)


def test_dropin_noise_pattern_masked_timepoints_image_prior():
    nx, ny, nz = 5, 4, 3
    n = nx * ny * nz
    y = synth.poly_volume(n, 40, 2, seed=75).numpy()
    img = np.linspace(-1e-3, 1e-3, n).astype(np.float32)
    ours, ref = both({"model": "poly", "degree": 2, "noise": "white", "method": "vb", "noise-pattern": "12",
                      "mt1": 4, "mt2": 40, "param-spatial-priors": "NAN", "PSP_byname1": "c2",
                      "PSP_byname1_type": "I", "PSP_byname1_image": "img", "PSP_byname1_prec": 1e6},
                     {"data": refbuild.volume(y, (nx, ny, nz)), "img": refbuild.volume(img[None], (nx, ny, nz))[..., 0]})
    assert_same(ours, ref)


def test_dropin_locked_linear_from_mvn():
    """locked-linear-from-mvn through the C API of both libraries: the MVN volume is one more named data item."""
    from test_reference_build import locked_mvn

    nx, ny, nz = 5, 4, 3
    n = nx * ny * nz
    y = synth.biexp_volume(n, 96, 0.02, 0.02, seed=76, smooth_shape=(nx, ny, nz)).numpy()
    rng = np.random.default_rng(7)
    centres = (np.array([[1.0], [1.2], [0.8], [5.0]]) * (1 + 0.05 * rng.standard_normal((4, n)))).astype(np.float32)
    opts = {"model": "exp", "num-exps": 2, "dt": 0.02, "noise": "white", "method": "spatialvb",
            "param-spatial-priors": "MMMM", "max-iterations": 4, "PSP_byname1": "r2", "PSP_byname1_mean": 6.0,
            "locked-linear-from-mvn": "lockmvn"}
    data = {"data": refbuild.volume(y, (nx, ny, nz)), "lockmvn": refbuild.volume(locked_mvn(4, 1, centres), (nx, ny, nz))}
    ours, ref = both(opts, data)
    assert_same(ours, ref)
    unlocked, _ = both({k: v for k, v in opts.items() if k != "locked-linear-from-mvn"}, {"data": data["data"]})
    assert np.max(np.abs(unlocked.data["mean_amp1"] - ours.data["mean_amp1"])) > 1e-4   # the option is live


def test_dropin_noise_initial_prior_and_posterior_files(tmp_path):
    """noise-initial-prior / noise-initial-posterior (inference_vb.cc:132-142,204-205): MVN matrix files whose
    means and variances set each phi's Gamma; two phis (noise-pattern=12), both libraries read the same files."""
    def mvn_file(name, means, variances):
        n = len(means)
        m = np.zeros((n + 1, n + 1))
        m[np.arange(n), np.arange(n)] = variances
        m[:n, n] = m[n, :n] = means
        m[n, n] = 1.0
        path = str(tmp_path / name)
        np.savetxt(path, m, fmt="%.17g")
        return path

    nx, ny, nz = 5, 4, 3
    y = synth.poly_volume(nx * ny * nz, 40, 2, seed=78).numpy()
    opts = {"model": "poly", "degree": 2, "noise": "white", "method": "vb", "noise-pattern": "12",
            "noise-initial-prior": mvn_file("prior.mat", [2.0, 0.5], [4e6, 1e5]),
            "noise-initial-posterior": mvn_file("post.mat", [1e-3, 5e-4], [2e-8, 5e-9])}
    data = {"data": refbuild.volume(y, (nx, ny, nz))}
    ours, ref = both(opts, data)
    assert_same(ours, ref)
    plain, _ = both({k: v for k, v in opts.items() if not k.startswith("noise-initial")}, data)
    assert np.max(np.abs(plain.data["noise_means"] - ours.data["noise_means"])
                  / np.abs(plain.data["noise_means"])) > 1e-4   # the options are live
    # a malformed file is refused with the reference's message
    bad = str(tmp_path / "bad.mat")
    np.savetxt(bad, np.array([[1.0, 0.5, 1.0], [0.0, 1.0, 1.0], [1.0, 1.0, 1.0]]))
    for lib in (fab.Fabber(), refbuild.ReferenceFabber()):
        with pytest.raises(Exception, match="MVNs must be symmetric"):
            lib.run_with_data(dict(opts, **{"noise-initial-prior": bad}), data)
