"""GPU: end-to-end through the reference's public C API (libfabbercore_b200.so, include/fabber_capi.h),
driven exactly as py/fabber.py drives the reference: options -> extent -> data -> dorun -> get_data."""
import os

import numpy as np
import pytest

import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import fabber as fab
from fabber_core_b200 import synth
from parity import tri

pytestmark = pytest.mark.gpu


def volume(series, shape):
    """[T][N] (x fastest) -> [x, y, z, t]"""
    nx, ny, nz = shape
    return np.ascontiguousarray(series.T.reshape(nz, ny, nx, -1).transpose(2, 1, 0, 3))


def flat(img):
    """[x, y, z(, k)] -> [k][N] in voxel order"""
    if img.ndim == 3:
        img = img[..., None]
    return np.stack([img[..., k].reshape(-1, order="F") for k in range(img.shape[3])])


def design_file(tmp_path, design):
    p = tmp_path / "design.mat"
    p.write_text("\n".join(" ".join("%.17g" % x for x in row) for row in design) + "\n")
    return str(p)


def test_c1_regression_case_through_capi(golden, tmp_path):
    """BASELINE configs[0]: --model=linear --basis=... --noise=white --method=vb on test_data_small,
    compared with test/outdata_linear_vb exactly as test/test_commandline.cc:108-139 does (1e-3 absolute)
    and at float32 precision."""
    f = fab.Fabber()
    opts = {"model": "linear", "basis": design_file(tmp_path, golden["design"]), "noise": "white", "method": "vb",
            "save-mean": True, "save-zstat": True, "save-std": True, "save-mvn": True, "save-noise-mean": True,
            "save-noise-std": True, "save-model-fit": True, "save-residuals": True}
    run = f.run_with_data(opts, {"data": volume(golden["data"], (3, 3, 2))})
    assert "Parameter_1" in " ".join(run.data.keys())
    for i in range(1, 5):
        m = flat(run.data["mean_Parameter_%d" % i])[0]
        g = golden["linear_vb/mean_Parameter_%d" % i][0]
        assert np.max(np.abs(m - g)) < 1e-3
        assert np.max(np.abs(m - g) / np.abs(g)) < 5e-6
        z = flat(run.data["zstat_Parameter_%d" % i])[0]
        gz = golden["linear_vb/zstat_Parameter_%d" % i][0]
        assert np.max(np.abs(z - gz) / np.abs(gz)) < 5e-6
    mvn = flat(run.data["finalMVN"])
    g = golden["linear_vb/finalMVN"]
    assert mvn.shape == g.shape == (21, 18)
    scale = np.maximum(np.abs(g), np.abs(g).max(axis=0, keepdims=True) * 1e-7)
    assert np.max(np.abs(mvn - g) / np.maximum(scale, 1e-30)) < 2e-5
    assert np.all(mvn[20] == 1.0)
    nm = flat(run.data["noise_means"])[0]
    assert np.max(np.abs(nm - g[19]) / g[19]) < 5e-6
    # modelfit + residuals = data
    fit, res = flat(run.data["modelfit"]), flat(run.data["residuals"])
    assert np.max(np.abs(fit + res - golden["data"])) < 1e-2
    assert "Vb::" in run.log and "Duration" in run.log


def test_poly_golden_through_capi(golden):
    f = fab.Fabber()
    run = f.run_with_data({"model": "poly", "degree": 2, "noise": "white", "method": "vb", "save-mean": True,
                           "save-std": True, "save-noise-mean": True, "save-noise-std": True},
                          {"data": volume(golden["data"], (3, 3, 2))})
    for i in range(3):
        for kind in ("mean", "std"):
            v = flat(run.data["%s_c%d" % (kind, i)])[0]
            g = golden["poly/%s_c%d" % (kind, i)][0]
            assert np.max(np.abs(v - g) / np.abs(g)) < 5e-6
    assert np.max(np.abs(flat(run.data["noise_means"])[0] / golden["poly/noise_means"][0] - 1)) < 5e-6
    assert np.max(np.abs(flat(run.data["noise_stdevs"])[0] / golden["poly/noise_stdevs"][0] - 1)) < 5e-6


@pytest.mark.parametrize("method", ["vb", "spatialvb"])
def test_constant_data_and_mask(method):
    """test/test_inference.cc:108-160 (param. over vb, spatialvb): constant data -> mean_c0 == VAL;
    voxels outside the mask come back as zeros (rundata_array.cc:68-98)."""
    nx, ny, nz, nt, val = 5, 4, 3, 10, 7.32
    data = np.full((nx, ny, nz, nt), val, dtype=np.float32)
    mask = np.ones((nx, ny, nz), dtype=np.int32)
    mask[0, :, :] = 0
    mask[2, 1, 1] = 0
    f = fab.Fabber()
    seen = []
    run = f.run_with_data({"model": "poly", "degree": 0, "noise": "white", "method": method, "save-mean": True},
                          {"data": data}, mask=mask, progress_cb=lambda v, n: seen.append((v, n)))
    mean = run.data["mean_c0"]
    assert mean.shape == (nx, ny, nz)
    assert np.all(mean[mask == 0] == 0)
    assert np.allclose(mean[mask != 0], np.float32(val), rtol=1e-6)
    assert seen and seen[-1][0] == seen[-1][1] == int(mask.sum())


def test_capi_matches_inner_abi_on_biexp_lm():
    """Same problem through the option strings and through the inner ABI: identical numbers."""
    nx, ny, nz = 8, 6, 5
    n = nx * ny * nz
    y = synth.biexp_volume(n, 96, 0.02, 0.02, seed=41).numpy()
    f = fab.Fabber()
    opts = {"model": "exp", "num-exps": 2, "dt": 0.02, "noise": "white", "method": "vb", "convergence": "lm",
            "PSP_byname1": "r2", "PSP_byname1_mean": 6.0, "save-mvn": True, "save-free-energy": True,
            "save-mean": True, "save-var": True}
    run = f.run_with_data(opts, {"data": volume(y, (nx, ny, nz))})
    ref = oracle.run(abi.ProblemSpec("exp", 96, num_exps=2, dt=0.02, convergence="lm", need_f=True,
                                     param_overrides={"r2": {"mean": 6.0}}), y)
    mvn = flat(run.data["finalMVN"]).astype(np.float64)
    # float32 outputs of a float64 computation: compare at float32 precision
    for i in range(4):
        assert np.allclose(mvn[15 + i], ref["mean"][i], rtol=2e-5, atol=1e-6)
        assert np.allclose(mvn[tri(i, i)], ref["cov"][tri(i, i)], rtol=2e-5)
        # model-space outputs: log transform (transforms.h:135-156)
        assert np.allclose(flat(run.data["mean_" + ["amp1", "r1", "amp2", "r2"][i]])[0], np.exp(ref["mean"][i]), rtol=2e-5)
        assert np.allclose(flat(run.data["var_" + ["amp1", "r1", "amp2", "r2"][i]])[0], np.exp(ref["cov"][tri(i, i)]), rtol=2e-5)
    assert np.allclose(flat(run.data["freeEnergy"])[0], ref["free_energy"], rtol=2e-5)


def test_restart_from_mvn_and_output_only():
    """test/test_vb.cc:305-498: continue-from-mvn restarts from a saved finalMVN; output-only re-emits it."""
    nx, ny, nz = 6, 5, 4
    y = synth.poly_volume(nx * ny * nz, 30, 2, seed=42).numpy()
    data = volume(y, (nx, ny, nz))
    f = fab.Fabber()
    base = {"model": "poly", "degree": 2, "noise": "white", "method": "vb", "save-mvn": True, "save-mean": True}
    first = f.run_with_data(dict(base, **{"max-iterations": 3}), {"data": data})
    second = f.run_with_data(dict(base, **{"max-iterations": 7, "continue-from-mvn": "mvn_in"}),
                             {"data": data, "mvn_in": first.data["finalMVN"]})
    full = f.run_with_data(dict(base, **{"max-iterations": 10}), {"data": data})
    # 3 + 7 iterations from the float32 checkpoint land where 10 iterations do (to checkpoint precision)
    assert np.allclose(second.data["mean_c0"], full.data["mean_c0"], rtol=1e-4, atol=1e-4)
    echo = f.run_with_data(dict(base, **{"output-only": True, "continue-from-mvn": "mvn_in"}),
                           {"data": data, "mvn_in": first.data["finalMVN"]})
    assert np.allclose(echo.data["finalMVN"][..., :6], first.data["finalMVN"][..., :6], rtol=1e-6)
    assert np.array_equal(echo.data["mean_c0"], first.data["mean_c0"])


def test_ar_noise_and_error_paths():
    nx, ny, nz = 6, 5, 3
    y = synth.linear_ar_volume(nx * ny * nz, 200, 0.3, seed=43).numpy()
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        basis = os.path.join(d, "ar.mat")
        np.savetxt(basis, synth.ar_design(200), fmt="%.17g")
        f = fab.Fabber()
        run = f.run_with_data({"model": "linear", "basis": basis, "noise": "ar", "method": "vb", "save-mvn": True,
                               "save-noise-mean": True}, {"data": volume(y, (nx, ny, nz))})
        nm = flat(run.data["noise_means"])
        # quirk kept: one volume (alpha1) because Ar1cNoiseModel::NumParams() returns nPhis (noisemodel_ar.cc:362)
        assert nm.shape[0] == 1 and abs(np.median(nm[0]) - 0.3) < 0.1
        assert flat(run.data["finalMVN"]).shape[0] == 7 * 8 // 2 + 7 + 1
        # AR + masked time points must fail (test/test_inference.cc:564-633)
        with pytest.raises(fab.FabberException):
            f.run_with_data({"model": "linear", "basis": basis, "noise": "ar", "method": "vb", "mt1": 3},
                            {"data": volume(y, (nx, ny, nz))})
        # unknown method / missing noise option
        with pytest.raises(fab.FabberException):
            f.run_with_data({"model": "linear", "basis": basis, "noise": "white", "method": "mcmc"},
                            {"data": volume(y, (nx, ny, nz))})
        with pytest.raises(fab.FabberException) as e:
            f.run_with_data({"model": "linear", "basis": basis, "method": "vb"}, {"data": volume(y, (nx, ny, nz))})
        assert "noise" in str(e.value)


def test_two_echo_ar_noise_restart_and_odd_series(tmp_path):
    """num-echoes=2 through the C API: a run continued from its own finalMVN (Ar1cParams::InputFromMVN order: the
    alphas, then the phis) for one more iteration lands where a straight run of one more iteration lands (up to
    the float32 MVN in between), and an odd series is refused."""
    nx, ny, nz = 5, 4, 2
    y = synth.dual_echo_volume(nx * ny * nz, 50, seed=45).numpy()
    basis = str(tmp_path / "de.mat")
    np.savetxt(basis, synth.dual_echo_design(50), fmt="%.17g")
    opts = {"model": "linear", "basis": basis, "noise": "ar", "num-echoes": 2, "ar1-cross-terms": "dual", "method": "vb",
            "save-mvn": True, "save-noise-mean": True, "save-noise-std": True, "save-mean": True}
    f = fab.Fabber()
    first = f.run_with_data(opts, {"data": volume(y, (nx, ny, nz))})
    mvn = first.data["finalMVN"]
    assert mvn.shape[-1] == 9 * 10 // 2 + 9 + 1 and flat(first.data["noise_means"]).shape[0] == 2
    again = dict(opts)
    again.update({"continue-from-mvn": "mvn", "max-iterations": 1})
    second = f.run_with_data(again, {"data": volume(y, (nx, ny, nz)), "mvn": mvn})
    straight = dict(opts)
    straight["max-iterations"] = 11
    third = f.run_with_data(straight, {"data": volume(y, (nx, ny, nz))})
    a, b = flat(second.data["finalMVN"]), flat(third.data["finalMVN"])
    assert np.max(np.abs(flat(mvn) - b)) > 0   # the extra iteration did move something
    scale = np.maximum(np.abs(b), np.max(np.abs(b), axis=1, keepdims=True) * 1e-3)
    assert np.max(np.abs(a - b) / np.maximum(scale, 1e-30)) < 5e-3
    with pytest.raises(fab.FabberException) as e:
        f.run_with_data(opts, {"data": volume(y[:-1], (nx, ny, nz))})
    assert "num-echoes" in str(e.value)


def test_bad_voxel_policy_through_capi():
    nx, ny, nz = 4, 4, 2
    y = synth.biexp_volume(nx * ny * nz, 96, 0.02, 0.02, seed=44).numpy()
    y[:, 5] = np.inf
    f = fab.Fabber()
    opts = {"model": "exp", "num-exps": 2, "dt": 0.02, "noise": "white", "method": "vb", "PSP_byname1": "r2",
            "PSP_byname1_mean": 6.0, "save-mean": True}
    with pytest.raises(fab.FabberException) as e:
        f.run_with_data(opts, {"data": volume(y, (nx, ny, nz))})
    assert e.value.errcode == fab.FABBER_ERR_FATAL and "Non-finite" in str(e.value)


def test_blockwise_upload_with_a_mask_equals_one_block(monkeypatch):
    """The main series is staged and uploaded one block of voxels at a time by a pool of host threads, and
    voxelwise VB is launched block by block (rundata.cc SetVoxelDataArray, Vb::DoCalculations). With a mask
    (gathered, not contiguous, staging) and two blocks the outputs must be bit for bit those of one block."""
    n = 72
    g = np.arange(n) - (n - 1) / 2.0
    r2 = g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2
    mask = (r2 < (0.45 * n) ** 2).astype(np.int32)
    n_in = int(mask.sum())
    assert n_in >= 2 * 65536                     # enough voxels for two blocks
    rng = np.random.default_rng(91)
    t = np.arange(1, 9, dtype=np.float32)
    data = (rng.uniform(50, 150, (n, n, n, 1)) + rng.uniform(-2, 2, (n, n, n, 1)) * t
            + rng.standard_normal((n, n, n, 8))).astype(np.float32)
    opts = {"model": "poly", "degree": 1, "noise": "white", "method": "vb", "save-mean": True, "save-std": True,
            "save-noise-mean": True}
    monkeypatch.setenv("FABBER_B200_UPLOAD_BLOCK_MB", "1")      # 8 x n_in x 4 B = 4.5 MB -> two blocks
    two = fab.Fabber().run_with_data(opts, {"data": data}, mask=mask)
    monkeypatch.setenv("FABBER_B200_UPLOAD_BLOCK_MB", "4096")   # one block
    one = fab.Fabber().run_with_data(opts, {"data": data}, mask=mask)
    for k in one.data:
        assert np.array_equal(one.data[k], two.data[k]), k
    assert np.all(two.data["mean_c0"][mask == 0] == 0)
    # and it is the right answer: a straight-line fit of a few voxels, done independently
    for (x, y, z) in ((36, 36, 36), (20, 40, 50), (36, 10, 36)):
        assert mask[x, y, z]
        c1, c0 = np.polyfit(t.astype(np.float64), data[x, y, z].astype(np.float64), 1)
        assert abs(two.data["mean_c1"][x, y, z] - c1) < 1e-3 * max(1.0, abs(c1))
        assert abs(two.data["mean_c0"][x, y, z] - c0) < 1e-3 * abs(c0)


def _run_ordered(opts, items, late_opts=None):
    """options -> extent -> data items IN THE GIVEN ORDER -> (optional late options) -> dorun -> mean / std / F"""
    f = fab.Fabber()
    f._set_options(opts)
    shape = items[0][1].shape
    n = shape[0] * shape[1] * shape[2]
    f._trycall(f.clib.fabber_set_extent, f.handle, shape[0], shape[1], shape[2], np.ones(n, dtype=np.int32), f.errbuf)
    for key, item in items:
        size = 1 if item.ndim == 3 else item.shape[3]
        flat = np.ascontiguousarray(np.asarray(item).flatten(order="F"), dtype=np.float32)
        f._trycall(f.clib.fabber_set_data, f.handle, key.encode(), size, flat, f.errbuf)
    if late_opts:
        f._set_options(late_opts)
    f._trycall(f.clib.fabber_dorun, f.handle, len(f.outbuf), f.outbuf, f.errbuf, f.progress_cb_type(0))
    log = f.outbuf.value.decode(errors="replace")
    f._trycall(f.clib.fabber_get_model_params, f.handle, len(f.outbuf), f.outbuf, f.errbuf)
    out = {}
    for key in ["mean_" + p for p in f.outbuf.value.decode().splitlines()] + ["freeEnergy", "finalMVN"]:
        size = f._trycall(f.clib.fabber_get_data_size, f.handle, key.encode(), f.errbuf)
        arr = np.empty(n * size, dtype=np.float32)
        f._trycall(f.clib.fabber_get_data, f.handle, key.encode(), arr, f.errbuf)
        out[key] = arr
    return out, log


def test_speculative_start_is_adopted_only_when_nothing_changed(monkeypatch):
    """fabber_set_data("data") starts voxelwise VB while the series is still being uploaded; fabber_dorun adopts
    that run only if no option or data item was set since - and the outputs are bit-identical to a run started
    from scratch in every case."""
    n, T = 70000, 40   # more than one upload block's worth of CTAs, a few blocks with the knob below
    monkeypatch.setenv("FABBER_B200_UPLOAD_BLOCK_MB", "2")
    y = synth.poly_volume(n, T, 2, seed=90).numpy()
    vol = volume(y, (n, 1, 1))
    img = np.linspace(-0.05, 0.05, n).astype(np.float32).reshape(n, 1, 1)
    base = {"model": "poly", "degree": 2, "noise": "white", "method": "vb", "convergence": "lm", "save-mean": True,
            "save-free-energy": True, "save-mvn": True}
    ADOPT = "Adopting the run started while the data was being set"

    monkeypatch.setenv("FABBER_B200_SPECULATE", "0")
    ref, log = _run_ordered(base, [("data", vol)])
    assert ADOPT not in log
    ref_it, _ = _run_ordered(dict(base, **{"max-iterations": 3}), [("data", vol)])
    img_opts = dict(base, PSP_byname1="c2", PSP_byname1_type="I", PSP_byname1_image="c2img", PSP_byname1_prec=1e4)
    ref_img, _ = _run_ordered(img_opts, [("c2img", img), ("data", vol)])
    monkeypatch.delenv("FABBER_B200_SPECULATE")

    got, log = _run_ordered(base, [("data", vol)])
    assert ADOPT in log
    for k in ref:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    # an option set after the data: the speculative run is thrown away, the new option holds
    got, log = _run_ordered(base, [("data", vol)], late_opts={"max-iterations": 3})
    assert ADOPT not in log
    for k in ref_it:
        assert np.array_equal(got[k], ref_it[k], equal_nan=True), k
    assert not np.array_equal(ref_it["mean_c1"], ref["mean_c1"])
    # the image prior arrives AFTER the main data: no speculation possible (or it is discarded), same result;
    # and when it arrives first the speculative run uses it
    got, log = _run_ordered(img_opts, [("data", vol), ("c2img", img)])
    assert ADOPT not in log
    for k in ref_img:
        assert np.array_equal(got[k], ref_img[k], equal_nan=True), k
    got, log = _run_ordered(img_opts, [("c2img", img), ("data", vol)])
    assert ADOPT in log
    for k in ref_img:
        assert np.array_equal(got[k], ref_img[k], equal_nan=True), k


@pytest.mark.parametrize("method", ["vb", "nlls", "spatialvb"])
def test_no_voxels_is_not_an_error(method):
    """test/test_inference.cc:57-73 (NoVoxels, run for every method): an empty mask runs through and yields empty maps"""
    nx, ny, nz, T = 3, 2, 2, 10
    f = fab.Fabber()
    data = np.ones((nx, ny, nz, T), dtype=np.float32)
    run = f.run_with_data({"model": "poly", "degree": 1, "noise": "white", "method": method, "save-mean": True},
                          {"data": data}, mask=np.zeros((nx, ny, nz), dtype=np.int32))
    assert run.data["mean_c0"].shape == (nx, ny, nz) and not np.any(run.data["mean_c0"])
