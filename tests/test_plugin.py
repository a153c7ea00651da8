"""CPU: the model plug-in mechanism (include/fabber_model_plugin.h; --loadmodels / fabber_load_models as in
fwdmodel.cc:63-129) up to where the GPU is needed: the example library exports the reference's three symbols,
registers its models, they list / describe / evaluate through the command line tool, and things that are not
plug-ins of this ABI are refused with the loader's messages."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "fabber_core_b200", "libfabber_models_example.so")
CLI = os.path.join(ROOT, "fabber_core_b200", "fabber_b200")
HOSTLIB = os.path.join(ROOT, "fabber_core_b200", "libfabbercore_b200.so")

pytestmark = pytest.mark.skipif(not os.path.exists(PLUGIN), reason="example plug-in not built")


def cli(args, cwd):
    p = subprocess.run([CLI] + args, cwd=str(cwd), capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout, p.stderr


def test_plugin_exports_the_reference_symbols():
    C.CDLL(HOSTLIB, mode=C.RTLD_GLOBAL)
    lib = C.CDLL(PLUGIN)
    lib.get_model_name.restype = C.c_char_p
    assert lib.fabber_b200_plugin_abi() == 2
    assert lib.get_num_models() == 2
    assert [lib.get_model_name(i) for i in range(2)] == [b"sine", b"exp"]
    assert lib.get_model_name(2) is None
    lib.get_new_instance_func.restype = C.c_void_p
    assert lib.get_new_instance_func(b"sine") and not lib.get_new_instance_func(b"nosuch")
    for sym in ("fabber_example_sine_launchers", "fabber_example_exp_launchers"):
        fn = getattr(lib, sym)
        fn.restype = C.c_void_p
        assert fn()


def test_loadmodels_on_the_command_line(tmp_path):
    rc, out, _ = cli(["--listmodels"], tmp_path)
    assert rc == 0 and "sine" not in out.split()
    rc, out, _ = cli(["--loadmodels=" + PLUGIN, "--listmodels"], tmp_path)
    assert rc == 0 and {"sine", "exp", "poly", "linear"} <= set(out.split())
    rc, out, _ = cli(["--loadmodels=" + PLUGIN, "--help", "--model=sine"], tmp_path)
    assert rc == 0 and "a*sin(b*(t-c))+d" in out and "--dt" in out
    rc, out, _ = cli(["--loadmodels=" + PLUGIN, "--listparams", "--model=sine"], tmp_path)
    assert rc == 0 and out.split() == ["a", "b", "c", "d"]
    # the plug-in's "exp" replaces the built-in registration, as a loaded library does in the reference
    rc, out, _ = cli(["--version", "--model=exp"], tmp_path)
    assert out.strip() == "b200"
    rc, out, _ = cli(["--loadmodels=" + PLUGIN, "--version", "--model=exp"], tmp_path)
    assert out.strip() == "example plug-in 1.0"
    (tmp_path / "p.mat").write_text("2\n0.5\n0.1\n3\n")
    rc, out, _ = cli(["--loadmodels=" + PLUGIN, "--evaluate", "--model=sine", "--dt=0.5", "--evaluate-params=p.mat",
                      "--evaluate-nt=5"], tmp_path)
    t = np.arange(5) * 0.5
    assert rc == 0 and np.allclose([float(x) for x in out.split()], 2 * np.sin(0.5 * (t - 0.1)) + 3, rtol=1e-5)


def test_loader_refuses_what_is_not_a_plugin(tmp_path):
    rc, _, err = cli(["--loadmodels=/nonexistent/lib.so", "--listmodels"], tmp_path)
    assert rc == 1 and "Failed to open library" in err
    rc, _, err = cli(["--loadmodels=" + HOSTLIB, "--listmodels"], tmp_path)       # a library, but no plug-in symbols
    assert rc == 1 and "get_num_models" in err
    from fabber_core_b200 import fabber as fab

    f = fab.Fabber()
    with pytest.raises(fab.FabberException) as e:
        f.load_models("/nonexistent/lib.so")
    assert "Failed to open library" in str(e.value)
    f.load_models(PLUGIN)
    assert "sine" in f.get_models()
