"""CPU: the two shared libraries load and export every symbol the public headers declare
(include/fabber_cuda.h, include/fabber_capi.h); the ctypes mirrors match the compiled structs; the host
logic that needs no GPU (options, model registry, host model evaluation, error paths) behaves like the
reference. No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import fabber as fab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fabber_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    lib = C.CDLL(abi.library_path())
    names = declared_functions("fabber_cuda.h")
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n


def test_host_library_exports_the_reference_capi():
    lib = C.CDLL(fab.library_path())
    names = declared_functions("fabber_capi.h")
    assert len(names) == 17  # the reference's fabber_capi.h:40-279
    for n in names:
        assert hasattr(lib, n), n


def test_struct_mirrors_match():
    lib = C.CDLL(abi.library_path())
    assert lib.fabber_cuda_sizeof_problem() == C.sizeof(abi.VbProblem)
    assert lib.fabber_cuda_sizeof_buffers() == C.sizeof(abi.VbBuffers)


def test_registry_names_and_options():
    f = fab.Fabber()
    assert f.get_models() == ["exp", "linear", "poly"]       # setup.cc:44-47 + examples/exp_models.cc
    assert f.get_methods() == ["nlls", "spatialvb", "vb"]    # setup.cc:28-33
    opts, desc = f.get_options(model="poly")
    assert [o["name"] for o in opts] == ["degree"] and "polynomial" in desc
    opts, _ = f.get_options(method="vb")
    names = [o["name"] for o in opts]
    for k in ("noise", "convergence", "max-iterations", "param-spatial-priors", "allow-bad-voxels"):
        assert k in names
    opts, _ = f.get_options()
    assert "save-mean" in [o["name"] for o in opts]


def test_model_params_and_priors_string():
    f = fab.Fabber()
    assert f.get_model_params({"model": "poly", "degree": 2}) == ["c0", "c1", "c2"]
    assert f.get_model_params({"model": "exp", "dt": 0.02, "num-exps": 2}) == ["amp1", "r1", "amp2", "r2"]
    with pytest.raises(fab.FabberException) as e:
        f.get_model_params({"model": "poly"})
    assert e.value.errcode == fab.FABBER_ERR_FATAL and "degree" in str(e.value)
    with pytest.raises(fab.FabberException):
        f.get_model_params({"model": "nosuchmodel"})


def test_host_model_evaluate_matches_reference_formulas():
    f = fab.Fabber()
    out = f.model_evaluate({"model": "poly", "degree": 2}, [1.0, 2.0, 0.5], 5)
    i = np.arange(1, 6)
    assert np.allclose(out, 1 + 2 * i + 0.5 * i * i)        # fwdmodel_poly.cc:68-79, i = 1..T
    out = f.model_evaluate({"model": "exp", "dt": 0.1, "num-exps": 1}, [2.0, 3.0], 4)
    assert np.allclose(out, 2 * np.exp(-3 * 0.1 * np.arange(4)), rtol=1e-6)   # fwdmodel_exp.cc:71-81
    with pytest.raises(fab.FabberException):
        f.model_evaluate({"model": "poly", "degree": 2}, [1.0], 5)


def test_error_conventions():
    f = fab.Fabber()
    err = C.create_string_buffer(255)
    assert f.clib.fabber_get_data_size(f.handle, b"nothing", err) == -1       # fabber_capi.cc:172-175
    assert f.clib.fabber_set_opt(None, b"a", b"b", err) == fab.FABBER_ERR_FATAL
    assert b"NULL" in err.value
    small = C.create_string_buffer(3)
    assert f.clib.fabber_get_models(f.handle, 3, small, err) == -1 and b"small" in err.value
    assert f.clib.fabber_load_models(f.handle, b"libfabber_models_x.so", err) == fab.FABBER_ERR_FATAL


def test_matrix_file_readers(tmp_path):
    """tools.cc:27-40: VEST and plain ASCII design files give the same model."""
    d = np.arange(12, dtype=float).reshape(4, 3) + 0.5
    ascii_f = tmp_path / "d_ascii.mat"
    ascii_f.write_text("# comment\n" + "\n".join(" ".join("%.17g" % x for x in row) for row in d) + "\n")
    vest_f = tmp_path / "d_vest.mat"
    vest_f.write_text("/NumWaves 3\n/NumPoints 4\n/PPheights 1 1 1\n\n/Matrix\n"
                      + "\n".join("\t".join("%e" % x for x in row) for row in d) + "\n")
    f = fab.Fabber()
    for path in (ascii_f, vest_f):
        assert f.get_model_params({"model": "linear", "basis": str(path)}) == ["Parameter_1", "Parameter_2", "Parameter_3"]
        out = f.model_evaluate({"model": "linear", "basis": str(path)}, [1.0, -1.0, 2.0], 4)
        assert np.allclose(out, d @ np.array([1.0, -1.0, 2.0]), rtol=1e-6)


@pytest.mark.parametrize("masked", [False, True])
def test_set_data_get_data_round_trip_on_the_host(masked):
    """fabber_set_data stages the caller's volume (cache-bypassing copy for a full mask, gather under a mask)
    and fabber_get_data hands volumes back (zeros outside the mask, rundata_array.cc:68-133): what comes out
    is bit for bit what went in. Odd sizes, so the vector copy's head / tail paths are taken. No GPU needed."""
    nx, ny, nz, nt = 37, 29, 23, 5
    n = nx * ny * nz
    rng = np.random.default_rng(3)
    vol = rng.standard_normal(nt * n).astype(np.float32)
    mask = np.ones(n, dtype=np.int32)
    if masked:
        mask[rng.random(n) < 0.3] = 0
    f = fab.Fabber()
    err = C.create_string_buffer(255)
    assert f.clib.fabber_set_extent(f.handle, nx, ny, nz, mask, err) == 0, err.value
    for key, rows in ((b"data", nt), (b"other", 2)):
        assert f.clib.fabber_set_data(f.handle, key, rows, vol[: rows * n], err) == 0, err.value
        assert f.clib.fabber_get_data_size(f.handle, key, err) == rows
        back = np.full(rows * n, np.nan, dtype=np.float32)
        assert f.clib.fabber_get_data(f.handle, key, back, err) == 0, err.value
        expect = vol[: rows * n].reshape(rows, n) * (mask != 0)
        assert np.array_equal(back.reshape(rows, n), expect.astype(np.float32))
