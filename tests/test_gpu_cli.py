"""GPU: the command line tool end to end, files in -> files out, the way test/test_commandline.cc drives the
reference's `fabber` (option strings on the command line, NIfTI volumes, logfile, output directory).
Inputs are written and outputs read back with the tests' own numpy NIfTI code (niftiutil.py)."""
import os
import subprocess

import numpy as np
import pytest

import niftiutil
from fabber_core_b200 import fabber as fab
from fabber_core_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "fabber_core_b200", "fabber_b200")

pytestmark = pytest.mark.gpu

ALLOWED_DELTA = 0.001   # test/test_commandline.cc:10


def run(args, cwd, expect=0):
    p = subprocess.run([CLI] + args, cwd=str(cwd), capture_output=True, text=True, timeout=600)
    assert p.returncode == expect, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout, p.stderr


def out_series(path):
    vol, hdr = niftiutil.read(path)
    return niftiutil.volume_to_series(vol), hdr


def write_c1(tmp_path, golden, sform=None):
    vol = niftiutil.series_to_volume(golden["data"], (3, 3, 2)).astype(np.int16)
    niftiutil.write(str(tmp_path / "test_data_small.nii.gz"), vol, pixdim=(3.5, 3.5, 3.5, 1.0), sform=sform)
    (tmp_path / "design.mat").write_text(
        "/NumWaves 4\n/NumPoints 106\n/Matrix\n" + "\n".join(" ".join("%.17g" % x for x in r) for r in golden["design"]) + "\n")


@pytest.mark.parametrize("method", ["vb", "spatialvb", "nlls"])
def test_linear_model_vest_regression_case(tmp_path, golden, method):
    """test_commandline.cc LinearModelVest (BASELINE configs[0]): outputs against test/outdata_linear_<method>"""
    write_c1(tmp_path, golden)
    run(["--output=out.tmp", "--model=linear", "--basis=design.mat", "--data=test_data_small.nii.gz", "--noise=white",
         "--method=" + method], tmp_path)
    log = (tmp_path / "out.tmp" / "logfile").read_text()
    for txt in ("model=linear", "method=" + method, "test_data_small.nii.gz"):
        assert txt in log
    for i in range(1, 5):
        for kind in ("mean", "zstat"):
            got, hdr = out_series(str(tmp_path / "out.tmp" / ("%s_Parameter_%d.nii.gz" % (kind, i))))
            want = golden["linear_%s/%s_Parameter_%d" % (method, kind, i)]
            assert got.shape == want.shape
            assert np.max(np.abs(got - want)) < ALLOWED_DELTA
            assert np.max(np.abs(got - want) / np.abs(want)) < 1e-5
            assert hdr["datatype"] == 16 and hdr["magic"] == b"n+1\0"
    mvn, hdr = out_series(str(tmp_path / "out.tmp" / "finalMVN.nii.gz"))
    n_all = 4 if method == "nlls" else 5                                          # NLLS: no noise parameter
    assert hdr["intent_code"] == 1005 and mvn.shape == (n_all * (n_all + 1) // 2 + n_all + 1, 18)   # NIFTI_INTENT_SYMMATRIX
    want = golden["linear_%s/finalMVN" % method]
    scale = np.maximum(np.abs(want), np.abs(want).max(axis=0, keepdims=True) * 1e-7)
    assert np.max(np.abs(mvn - want) / scale) < (3e-5 if method == "nlls" else 2e-5)
    assert (tmp_path / "out.tmp" / "paramnames.txt").read_text().split() == ["Parameter_%d" % i for i in range(1, 5)]
    assert os.path.realpath(str(tmp_path / "out.tmp_latest")) == os.path.realpath(str(tmp_path / "out.tmp"))


def test_poly_model_outputs_and_header_properties(tmp_path, golden):
    """PolyModel + OutputCopiesPropsNoMask: mean / std against test/outdata_poly, voxel sizes copied from the input"""
    write_c1(tmp_path, golden)
    run(["--model=poly", "--output=out.tmp", "--degree=2", "--method=vb", "--noise=white",
         "--data=test_data_small.nii.gz"], tmp_path)
    for name in ("c0", "c1", "c2"):
        for kind in ("mean", "std"):
            got, hdr = out_series(str(tmp_path / "out.tmp" / ("%s_%s.nii.gz" % (kind, name))))
            want = golden["poly/%s_%s" % (kind, name)]
            assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-3)) < 2e-5
            assert hdr["pixdim"][1:4] == (3.5, 3.5, 3.5)
            assert abs(hdr["cal_max"] - got.max()) <= 1e-6 * abs(got.max())
    # (the golden freeEnergy volume is the constant 9999 - written by an older fabber that left resultFs at its
    #  initial value, inference_vb.cc:165; the current sources compute F, pinned in test_reference_build.py)
    assert np.all(golden["poly/freeEnergy"] == 9999)
    fe, _ = out_series(str(tmp_path / "out.tmp" / "freeEnergy.nii.gz"))
    assert np.all(np.isfinite(fe)) and np.all(fe < 0)
    for k in ("noise_means", "noise_stdevs"):
        got, _ = out_series(str(tmp_path / "out.tmp" / (k + ".nii.gz")))
        want = golden["poly/" + k]
        assert np.max(np.abs(got - want) / np.abs(want)) < 2e-5


def test_mask_empty_mask_unused_option_and_uncompressed_output(tmp_path):
    nx, ny, nz, T = 6, 5, 4, 40
    y = synth.poly_volume(nx * ny * nz, T, 2, seed=31).numpy()
    niftiutil.write(str(tmp_path / "data.nii.gz"), niftiutil.series_to_volume(y, (nx, ny, nz)))
    rng = np.random.default_rng(4)
    mask = (rng.random((nx, ny, nz)) > 0.5) * rng.random((nx, ny, nz))          # float mask, binarised > 0
    niftiutil.write(str(tmp_path / "mask.nii.gz"), mask.astype(np.float32))
    args = ["--model=poly", "--degree=2", "--method=vb", "--noise=white", "--data=data", "--output=out", "--overwrite"]
    run(args + ["--mask=mask"], tmp_path)
    got, _ = out_series(str(tmp_path / "out" / "mean_c1.nii.gz"))
    sel = mask.reshape(-1, order="F") > 0
    assert np.all(got[0][~sel] == 0) and np.all(got[0][sel] != 0)
    # same voxels, same numbers as the C API on arrays
    api = fab.Fabber().run_with_data({"model": "poly", "degree": 2, "method": "vb", "noise": "white", "save-mean": True},
                                     {"data": niftiutil.series_to_volume(y, (nx, ny, nz))},
                                     mask=(mask > 0).astype(np.int32))
    assert np.array_equal(niftiutil.volume_to_series(api.data["mean_c1"][..., None])[0].astype(np.float32),
                          got[0].astype(np.float32))
    log = (tmp_path / "out" / "logfile").read_text()
    assert "WARNING" not in log
    # EmptyMask (test_commandline.cc:204-218): nothing to do is not an error
    niftiutil.write(str(tmp_path / "empty.nii.gz"), np.zeros((nx, ny, nz), dtype=np.int16))
    run(args + ["--mask=empty"], tmp_path)
    # UnusedParams (:341-355)
    run(args + ["--squaffle"], tmp_path)
    log = (tmp_path / "out" / "logfile").read_text()
    assert "WARNING" in log and "Unused option" in log
    # FSLOUTPUTTYPE=NIFTI writes .nii
    env = dict(os.environ, FSLOUTPUTTYPE="NIFTI")
    p = subprocess.run([CLI] + args, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and os.path.exists(str(tmp_path / "out" / "mean_c0.nii"))
    a, _ = niftiutil.read(str(tmp_path / "out" / "mean_c0.nii"))
    assert a.shape == (nx, ny, nz, 1)


def test_neurological_files_are_mirrored_in_memory_like_newimage(tmp_path):
    """A file whose sform has a positive determinant is held x-mirrored (NEWIMAGE's radiological storage),
    which changes the ORDER spatial VB sweeps the voxels in. File-level result for an MRF prior must equal
    the C API run on the mirrored arrays, mirrored back - and differ from the unmirrored run."""
    nx, ny, nz, T = 6, 5, 3, 30
    y = synth.poly_volume(nx * ny * nz, T, 1, seed=32).numpy()
    vol = niftiutil.series_to_volume(y, (nx, ny, nz))
    neuro = [[2, 0, 0, 0], [0, 2, 0, 0], [0, 0, 2, 0]]      # det > 0
    radio = [[-2, 0, 0, 0], [0, 2, 0, 0], [0, 0, 2, 0]]     # det < 0
    opts = {"model": "poly", "degree": 1, "method": "spatialvb", "noise": "white", "param-spatial-priors": "M+",
            "max-iterations": 5, "save-mean": True}
    args = ["--model=poly", "--degree=1", "--method=spatialvb", "--noise=white", "--param-spatial-priors=M+",
            "--max-iterations=5", "--overwrite"]
    niftiutil.write(str(tmp_path / "neuro.nii.gz"), vol, sform=neuro)
    niftiutil.write(str(tmp_path / "radio.nii.gz"), vol, sform=radio)
    run(args + ["--data=neuro", "--output=o_neuro"], tmp_path)
    run(args + ["--data=radio", "--output=o_radio"], tmp_path)
    got_n, hdr_n = niftiutil.read(str(tmp_path / "o_neuro" / "mean_c1.nii.gz"))
    got_r, _ = niftiutil.read(str(tmp_path / "o_radio" / "mean_c1.nii.gz"))
    assert hdr_n["sform_code"] == 1 and hdr_n["srow"][0] == 2.0                 # orientation copied to the output
    api_plain = fab.Fabber().run_with_data(opts, {"data": vol}).data["mean_c1"]
    api_mirror = fab.Fabber().run_with_data(opts, {"data": np.ascontiguousarray(vol[::-1])}).data["mean_c1"][::-1]
    assert np.array_equal(got_r[..., 0].astype(np.float32), api_plain.astype(np.float32))
    assert np.array_equal(got_n[..., 0].astype(np.float32), api_mirror.astype(np.float32))
    assert not np.array_equal(got_n, got_r)


def test_multiple_data_files_and_restart_from_mvn_file(tmp_path):
    """data1 / data2 with data-order (rundata.cc:821-905) and continue-from-mvn / image priors named by file"""
    nx, ny, nz, T = 4, 4, 2, 24
    y = synth.poly_volume(nx * ny * nz, T, 1, seed=33).numpy()
    vol = niftiutil.series_to_volume(y, (nx, ny, nz))
    niftiutil.write(str(tmp_path / "all.nii.gz"), vol)
    niftiutil.write(str(tmp_path / "even.nii.gz"), np.ascontiguousarray(vol[..., 0::2]))
    niftiutil.write(str(tmp_path / "odd.nii.gz"), np.ascontiguousarray(vol[..., 1::2]))
    niftiutil.write(str(tmp_path / "first.nii.gz"), np.ascontiguousarray(vol[..., :10]))
    niftiutil.write(str(tmp_path / "rest.nii.gz"), np.ascontiguousarray(vol[..., 10:]))
    base = ["--model=poly", "--degree=1", "--method=vb", "--noise=white", "--overwrite"]
    run(base + ["--data=all", "--output=o_all"], tmp_path)
    run(base + ["--data1=even", "--data2=odd", "--output=o_il"], tmp_path)                       # interleave is the default
    run(base + ["--data1=first", "--data2=rest", "--data-order=concatenate", "--output=o_cat"], tmp_path)
    ref, _ = niftiutil.read(str(tmp_path / "o_all" / "mean_c1.nii.gz"))
    for d in ("o_il", "o_cat"):
        got, _ = niftiutil.read(str(tmp_path / d / "mean_c1.nii.gz"))
        assert np.array_equal(got, ref), d
    _, err = run(base + ["--data1=even", "--data2=rest", "--output=o_bad"], tmp_path, expect=1)
    assert "same number of time points" in err
    # restart: 2 iterations, then 3 more from the saved MVN == 5 in one go (test_vb.cc restart cases)
    run(base + ["--data=all", "--output=o_2", "--max-iterations=2"], tmp_path)
    run(base + ["--data=all", "--output=o_2_3", "--max-iterations=3", "--continue-from-mvn=o_2/finalMVN"], tmp_path)
    run(base + ["--data=all", "--output=o_5", "--max-iterations=5"], tmp_path)
    a, _ = niftiutil.read(str(tmp_path / "o_2_3" / "mean_c0.nii.gz"))
    b, _ = niftiutil.read(str(tmp_path / "o_5" / "mean_c0.nii.gz"))
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-4       # the MVN file is float32
