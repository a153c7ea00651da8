"""CPU: the command line front end (fabber_core_b200/host/fabber_main.cc, rundata_newimage.cc, nifti_io.cc)
up to the point where the GPU is needed - option handling, usage / list commands, option files, output
directory rules, NIfTI reading (against an independent numpy reader / writer, every datatype, both byte
orders, scaling, masks, orientation) - and that without a CUDA device the run FAILS LOUDLY instead of
falling back to anything. Mirrors test/test_commandline.cc of the reference where no GPU is involved."""
import os
import re
import subprocess

import numpy as np
import pytest

import niftiutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "fabber_core_b200", "fabber_b200")

pytestmark = pytest.mark.skipif(not os.path.exists(CLI), reason="command line tool not built")


def run(args, cwd):
    p = subprocess.run([CLI] + args, cwd=str(cwd), capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout, p.stderr


def no_gpu():
    import torch

    return not torch.cuda.is_available()


def test_help_and_lists(tmp_path):
    rc, out, _ = run(["--help"], tmp_path)
    assert rc == 0 and "Usage" in out and "--method" in out
    rc, out, _ = run([], tmp_path)
    assert rc == 0 and "Usage" in out
    rc, out, _ = run(["--listmodels"], tmp_path)
    assert rc == 0 and {"poly", "linear", "exp"} <= set(out.split())
    rc, out, _ = run(["--listmethods"], tmp_path)
    assert rc == 0 and {"vb", "spatialvb"} <= set(out.split())
    rc, out, _ = run(["--help", "--model=poly"], tmp_path)
    assert rc == 0 and "--degree" in out
    rc, out, _ = run(["--help", "--method=vb"], tmp_path)
    assert rc == 0 and "--max-iterations" in out and "--convergence" in out
    rc, out, _ = run(["--listparams", "--model=poly", "--degree=3"], tmp_path)
    assert rc == 0 and out.split() == ["c0", "c1", "c2", "c3"]
    rc, out, _ = run(["--listparams", "--model=exp", "--num-exps=2", "--dt=0.1"], tmp_path)
    assert rc == 0 and out.split() == ["amp1", "r1", "amp2", "r2"]
    rc, out, _ = run(["--version"], tmp_path)
    assert rc == 0 and out.startswith("Fabber ")


def test_bad_command_lines(tmp_path):
    rc, _, err = run(["model=poly"], tmp_path)
    assert rc == 1 and "doesn't begin with --" in err
    rc, _, err = run(["-f"], tmp_path)
    assert rc == 1 and "No filename specified" in err
    rc, _, err = run(["-f", "missing.txt"], tmp_path)
    assert rc == 1 and "Couldn't read input options file" in err
    rc, _, err = run(["--model=poly", "--model=linear", "--listparams"], tmp_path)
    assert rc == 1 and "Already has a value" in err
    rc, _, err = run(["--listparams", "--model=nosuchmodel"], tmp_path)
    assert rc == 1 and "nosuchmodel" in err
    rc, _, err = run(["--loadmodels=libx.so", "--listmodels"], tmp_path)
    assert rc == 1 and "loadmodels" in err
    # a run without data: the reference reports the missing file
    rc, _, err = run(["--model=poly", "--degree=1", "--method=vb", "--noise=white", "--output=o", "--data=nothere"], tmp_path)
    assert rc == 1 and "nothere" in err


def test_evaluate_model(tmp_path):
    (tmp_path / "p.mat").write_text("2\n0.5\n0.25\n")
    rc, out, _ = run(["--evaluate", "--model=poly", "--degree=2", "--evaluate-params=p.mat", "--evaluate-nt=4"], tmp_path)
    assert rc == 0
    got = np.array([float(x) for x in out.split()])
    t = np.arange(1, 5)
    assert np.allclose(got, 2 + 0.5 * t + 0.25 * t * t, rtol=1e-6)


def logged_mean(log, name):
    m = re.search(re.escape(name) + r" mean value=([-0-9.e+]+)", log)
    assert m, log
    return float(m.group(1))


@pytest.mark.skipif(not no_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
@pytest.mark.parametrize("dtype,endian,scale", [(np.int16, "<", (0.0, 0.0)), (np.float32, ">", (0.0, 0.0)),
                                                (np.uint8, "<", (2.0, -3.0)), (np.float64, "<", (0.0, 0.0)),
                                                (np.int32, ">", (0.5, 1.0)), (np.uint16, "<", (0.0, 0.0))])
def test_reads_nifti_then_fails_loudly_without_a_gpu(tmp_path, dtype, endian, scale):
    """every datatype / byte order / scaling goes through the loader: the log's 'mean value' of the masked
    matrix must equal numpy's; then the run stops with the CUDA error and exit code 1 - no CPU fallback."""
    rng = np.random.default_rng(3)
    nx, ny, nz, nt = 5, 4, 3, 7
    vol = (rng.random((nx, ny, nz, nt)) * 100).astype(dtype)
    mask = (rng.random((nx, ny, nz)) > 0.4).astype(np.int16)
    niftiutil.write(str(tmp_path / "data.nii.gz"), vol, slope=scale[0], inter=scale[1], endian=endian)
    niftiutil.write(str(tmp_path / "mask.nii"), mask)
    rc, out, err = run(["--model=poly", "--degree=1", "--method=vb", "--noise=white", "--output=out", "--data=data",
                        "--mask=mask.nii"], tmp_path)
    assert rc == 1 and "Exception caught in fabber" in err and "cuda" in err.lower()
    log = (tmp_path / "out" / "logfile").read_text()
    assert "x=5, y=4, z=3, vols=7" in log
    v = vol.astype(np.float64) * (scale[0] if scale[0] else 1.0) + (scale[1] if scale[0] else 0.0)
    expect = v[mask > 0].mean()
    assert abs(logged_mean(log, "data") - expect) <= 2e-5 * abs(expect)
    assert sorted(os.listdir(str(tmp_path / "out"))) == ["logfile", "paramnames.txt"]   # nothing was "computed"
    assert (tmp_path / "out" / "paramnames.txt").read_text().split() == ["c0", "c1"]
    assert os.path.islink(str(tmp_path / "out_latest"))


@pytest.mark.skipif(not no_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
def test_output_directory_rules_and_option_files(tmp_path):
    vol = np.arange(2 * 2 * 2 * 5, dtype=np.float32).reshape(2, 2, 2, 5)
    niftiutil.write(str(tmp_path / "d.nii.gz"), vol)
    base = ["--model=poly", "--method=vb", "--noise=white", "--output=out.tmp", "--data=d.nii.gz"]
    run(base + ["--degree=2"], tmp_path)
    run(base + ["--degree=1"], tmp_path)
    assert "degree=2" in (tmp_path / "out.tmp" / "logfile").read_text()
    assert "degree=1" in (tmp_path / "out.tmp+" / "logfile").read_text()        # test_commandline.cc NoOverwrite
    run(base + ["--degree=3", "--overwrite"], tmp_path)
    assert "degree=3" in (tmp_path / "out.tmp" / "logfile").read_text()         # Overwrite
    (tmp_path / "new.txt").write_text("# comment\nmodel=poly\noutput=o1\ndegree=2   # trailing comment\nmethod=vb\n"
                                      "noise=white\noverwrite\n")
    run(["-f", "new.txt", "--data=d.nii.gz"], tmp_path)
    log = (tmp_path / "o1" / "logfile").read_text()
    assert "model=poly" in log and "degree=2\n" in log and "overwrite" in log   # OptFileNewStyle
    (tmp_path / "old.txt").write_text("--model=poly\n--output=o2 --degree=2\n# comment --degree=9\n--method=vb\n"
                                      "--noise=white\n--overwrite")
    run(["-@", "old.txt", "--data=d.nii.gz"], tmp_path)
    assert "degree=2" in (tmp_path / "o2" / "logfile").read_text()              # OptFileOldStyleOldName
    run(["--optfile=old.txt", "--data=d.nii.gz", "--no-compat-output"], tmp_path)
    log = (tmp_path / "o2" / "logfile").read_text()
    assert "degree=2" in log and "save-mean" not in log                         # OptFileOldStyle


@pytest.mark.skipif(not os.path.exists("/root/reference/test/test_data_small.nii.gz"), reason="reference not mounted")
@pytest.mark.skipif(not no_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
def test_reads_the_references_own_fixture(tmp_path, golden):
    """the regression time-series of the reference (int16, 3x3x2x106) loads to the values the golden holds"""
    rc, _, _ = run(["--model=poly", "--degree=2", "--method=vb", "--noise=white", "--output=out",
                    "--data=/root/reference/test/test_data_small.nii.gz"], tmp_path)
    assert rc == 1
    log = (tmp_path / "out" / "logfile").read_text()
    assert "x=3, y=3, z=2, vols=106" in log
    expect = float(golden["data"].mean())
    assert abs(logged_mean(log, "/root/reference/test/test_data_small.nii.gz") - expect) <= 1e-5 * expect
