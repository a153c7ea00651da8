"""CPU: option translation in the host library (Vb::DoCalculations up to the first CUDA call). Invalid run
options are refused with the option's name and the reason before any device work; a valid run then fails
LOUDLY at the first CUDA call because there is no CPU inference path. Reference behaviour restated from
inference_vb.cc:31-76, noisemodel_white.cc:166-215, noisemodel_ar.cc:318-349, convergence.cc, setup.cc."""
import numpy as np
import pytest

from fabber_core_b200 import fabber as fab

BASE = {"model": "poly", "degree": 1, "noise": "white", "method": "vb"}
DATA = (np.random.default_rng(0).standard_normal((3, 2, 2, 12)) + 10).astype(np.float32)


def run(extra):
    opts = dict(BASE)
    opts.update(extra)
    return fab.Fabber().run_with_data(opts, {"data": DATA})


@pytest.mark.parametrize("extra, fragments", [
    ({"convergence": "nosuch"}, ("convergence", "Unrecognized convergence detector")),
    ({"noise-pattern": "1?2"}, ("noise-pattern=?", "Invalid character")),
    ({"mt1": "13"}, ("mt", "beyond the end of the data")),                      # 12 time points
    ({"noise": "ar", "mt1": "2"}, ("AR noise model does not support masked time points",)),
    ({"noise": "ar", "num-echoes": "3"}, ("num-echoes", "Must be 1 or 2")),
    ({"noise": "ar", "num-echoes": "2", "ar1-cross-terms": "both"}, ("ar1-cross-terms", "Must be dual, same or none")),
    ({"noise": "ar", "ar1-cross-terms": "dual"}, ("ar1-cross-terms", "ar1-cross-terms=none with num-echoes=1")),
    ({"noise": "pink"}, ("noise=pink", "Unrecognized noise type")),
    ({"prior-noise-stddev": "-2"}, ("prior-noise-stddev", "Must be > 0")),
    ({"model": "nosuch"}, ("model", "Unrecognized forward model")),
    ({"method": "mcmc"}, ("method", "Unrecognized inference method")),
    ({"method": "nlls", "fwd-inital-posterior": "/nonexistent/file.mat"}, ("Could not read matrix file",)),
    ({"degree": "-1"}, ("degree", "Minimum 0")),
    ({"max-iterations": "0"}, ("max_iterations=0", "Must be positive")),          # sic: convergence.cc:39
    ({"convergence": "lm", "max-iterations": "0"}, ("max-iterations=0", "Must be positive")),
    ({"convergence": "trialmode", "max-trials": "0"}, ("max-trials=0", "Must be positive")),
    ({"degree": "abc"}, ("degree=abc", "Failed to convert to required type")),
    ({"locked-linear-from-mvn": "nosuch"}, ("Voxel data not found: nosuch",)),
    ({"method": "spatialvb", "param-spatial-priors": "M+", "spatial-dims": "4"}, ("spatial-dims=4", "Maximum 3")),
])
def test_invalid_options_are_refused_before_any_device_work(extra, fragments):
    with pytest.raises(fab.FabberException) as e:
        run(extra)
    for frag in fragments:
        assert frag in str(e.value), str(e.value)
    assert "cuda" not in str(e.value).lower()


def test_noise_initial_distribution_files_are_validated(tmp_path):
    """inference_vb.cc:132-142 / dist_mvn.cc:287-309 / noisemodel_white.cc:70-79"""
    def mat(name, rows):
        p = str(tmp_path / name)
        np.savetxt(p, np.array(rows, dtype=np.float64), fmt="%.17g")
        return p

    two = {"noise-pattern": "12"}
    with pytest.raises(fab.FabberException, match="MVNs must be symmetric"):
        run(dict(two, **{"noise-initial-prior": mat("asym.mat", [[1, 0.5, 1], [0, 1, 1], [1, 1, 1]])}))
    with pytest.raises(fab.FabberException, match="MVNs must be symmetric"):        # corner element must be 1
        run(dict(two, **{"noise-initial-posterior": mat("corner.mat", [[1, 0, 1], [0, 1, 1], [1, 1, 2]])}))
    with pytest.raises(fab.FabberException, match="Phis should have zero covariance"):
        run(dict(two, **{"noise-initial-prior": mat("corr.mat", [[4e6, 1, 2], [1, 1e5, 0.5], [2, 0.5, 1]])}))
    with pytest.raises(fab.FabberException, match="fewer rows"):
        run(dict(two, **{"noise-initial-prior": mat("short.mat", [[4e6, 2], [2, 1]])}))
    with pytest.raises(fab.FabberException, match="Could not read matrix file"):
        run(dict(two, **{"noise-initial-prior": str(tmp_path / "nofile.mat")}))
    with pytest.raises(fab.FabberException, match="not supported with noise=ar"):
        run({"noise": "ar", "noise-initial-prior": mat("ok.mat", [[4e6, 2], [2, 1]])})


def test_a_valid_run_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU: the run would succeed")
    with pytest.raises(fab.FabberException) as e:
        run({})
    assert "cuda" in str(e.value).lower()   # the CUDA error itself: nothing falls back to the CPU
