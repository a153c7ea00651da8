"""GPU: models from a plug-in library (fabber_core_b200/examples -> libfabber_models_example.so).
 - "sine": a model the core library has never seen, kernels compiled in the plug-in, checked against the test
   oracle's restatement of it (same parity rule as every other model);
 - "exp" from the plug-in against the built-in exp: the same device struct compiled twice must give the
   same bits, through the C API, for voxelwise, AR(1) and spatial runs;
 - through the command line tool with --loadmodels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import niftiutil
import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device, synth
from fabber_core_b200 import fabber as fab
from parity import compare

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "fabber_core_b200", "libfabber_models_example.so")
CLI = os.path.join(ROOT, "fabber_core_b200", "fabber_b200")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(PLUGIN), reason="example plug-in not built")]


def sine_volume(n, T, dt, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(T) * dt
    a = 1.2 + 0.1 * rng.standard_normal(n)
    b = 1.05 + 0.02 * rng.standard_normal(n)
    c = 0.1 + 0.05 * rng.standard_normal(n)
    d = 0.3 + 0.1 * rng.standard_normal(n)
    y = a[None] * np.sin(b[None] * (t[:, None] - c[None])) + d[None] + 0.02 * rng.standard_normal((T, n))
    return np.ascontiguousarray(y, dtype=np.float32)


def sine_launchers():
    device.lib()                                                        # libfabber_cuda.so first
    C.CDLL(os.path.join(ROOT, "fabber_core_b200", "libfabbercore_b200.so"), mode=C.RTLD_GLOBAL)
    lib = C.CDLL(PLUGIN)
    lib.fabber_example_sine_launchers.restype = C.c_void_p
    return lib.fabber_example_sine_launchers()


@pytest.mark.parametrize("kw", [dict(), dict(convergence="lm", need_f=True), dict(noise="ar"),
                                dict(convergence="trialmode", need_f=True, max_iterations=20)])
def test_sine_plugin_model_against_the_oracle(kw):
    T, n, dt = 48, 700, 0.25
    y = sine_volume(n, T, dt, seed=51)
    base = dict(dt=dt, max_iterations=10, need_f=True)
    base.update(kw)
    ref = oracle.run(abi.ProblemSpec("sine", T, **base), y)
    probes = [oracle.run(abi.ProblemSpec("sine", T, **base), y, variant="fma")]
    gpu = device.run(abi.ProblemSpec("sine", T, plugin_launchers=sine_launchers(), **base), y)
    compare(gpu, ref, 4, probes, label="plug-in sine %s" % kw)
    assert np.abs(np.median(gpu["mean"][0]) - 1.2) < 0.05 and np.abs(np.median(gpu["mean"][1]) - 1.05) < 0.02


def test_sine_plugin_spatial_against_the_oracle():
    nx, ny, nz, T, dt = 6, 5, 4, 48, 0.25
    y = sine_volume(nx * ny * nz, T, dt, seed=52)
    idx = np.arange(nx * ny * nz)
    coords = np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32)

    def mk(launchers=None):
        sp = abi.ProblemSpec("sine", T, dt=dt, max_iterations=5, need_f=True, prior_types=list("MNMN"),
                             plugin_launchers=launchers)
        sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
        return sp

    ref = oracle.run(mk(), y, spatial=True, coords=coords)
    probes = [oracle.run(mk(), y, spatial=True, coords=coords, variant="fma")]
    gpu = device.run(mk(sine_launchers()), y, spatial=True, coords=coords)
    compare(gpu, ref, 4, probes, label="plug-in sine spatial MNMN")


def test_plugin_exp_is_bit_identical_to_the_builtin_exp():
    nx, ny, nz, T = 6, 5, 4, 96
    y = synth.biexp_volume(nx * ny * nz, T, 0.02, 0.02, seed=53, smooth_shape=(nx, ny, nz)).numpy()
    vol = niftiutil.series_to_volume(y, (nx, ny, nz))
    save = {"save-mean": True, "save-std": True, "save-mvn": True, "save-free-energy": True, "save-model-fit": True}
    cases = ({"method": "vb", "noise": "white", "convergence": "lm"},
             {"method": "vb", "noise": "ar"},
             {"method": "spatialvb", "noise": "white", "param-spatial-priors": "MMMM", "max-iterations": 4})

    def options(extra):
        opts = {"model": "exp", "num-exps": 2, "dt": 0.02, "PSP_byname1": "r2", "PSP_byname1_mean": 6.0}
        opts.update(extra)
        opts.update(save)
        return opts

    # the model registry is per process, as the reference's FwdModelFactory is: once the library is loaded its
    # "exp" replaces the built-in one - so every built-in run comes first
    builtin = [fab.Fabber().run_with_data(options(extra), {"data": vol}) for extra in cases]
    for extra, ref in zip(cases, builtin):
        assert "Forward Model version b200" in ref.log
        plugged = fab.Fabber().run_with_data(dict(options(extra), loadmodels=PLUGIN), {"data": vol})
        assert "Forward Model version example plug-in 1.0" in plugged.log
        assert sorted(plugged.data) == sorted(ref.data)
        for k in ref.data:
            assert np.array_equal(plugged.data[k], ref.data[k]), (extra, k)


def test_sine_plugin_through_the_command_line(tmp_path):
    nx, ny, nz, T, dt = 5, 4, 3, 48, 0.25
    y = sine_volume(nx * ny * nz, T, dt, seed=54)
    niftiutil.write(str(tmp_path / "data.nii.gz"), niftiutil.series_to_volume(y, (nx, ny, nz)))
    p = subprocess.run([CLI, "--loadmodels=" + PLUGIN, "--model=sine", "--dt=0.25", "--method=vb", "--noise=white",
                        "--data=data", "--output=out", "--save-model-fit"], cwd=str(tmp_path), capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert (tmp_path / "out" / "paramnames.txt").read_text().split() == ["a", "b", "c", "d"]
    log = (tmp_path / "out" / "logfile").read_text()
    assert "Loading model sine" in log and "WARNING" not in log
    a, _ = niftiutil.read(str(tmp_path / "out" / "mean_a.nii.gz"))
    fit, _ = niftiutil.read(str(tmp_path / "out" / "modelfit.nii.gz"))
    ref = oracle.run(abi.ProblemSpec("sine", T, dt=dt, max_iterations=10), y)
    assert np.max(np.abs(niftiutil.volume_to_series(a)[0] - ref["mean"][0]) / np.abs(ref["mean"][0])) < 1e-5
    assert np.max(np.abs(niftiutil.volume_to_series(fit) - y)) < 0.2       # the fit follows the data
