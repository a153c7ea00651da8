"""GPU: the spatial, z-slab and two-echo kernels under the DEBUG build of the CUDA library
(csrc: make checked -> libfabber_cuda_checked.so, -DFAB_BOUNDS_CHECK): every table-driven index (neighbour lists,
hyper-plane order, slab ghost maps, mailbox slots, voxel permutations) is tested inside the kernels and failures are
counted on the device. compute-sanitizer is closed on the GPU pool this is developed on
(profiles/r2m_compute_sanitizer_closed.txt), so this is the memory-safety evidence for those kernels: the same
parity tests, run in a sub-session against the checked library, must pass with ZERO counted failures - and the
counters must be live (a deliberate failure is seen). Races show up as non-determinism: the ordered sweep and the
slab engine are run repeatedly and must reproduce themselves bit for bit."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "fabber_core_b200", "csrc", "libfabber_cuda_checked.so")


@pytest.mark.skipif(not os.path.exists(CHECKED), reason="debug build not made (make -C fabber_core_b200/csrc checked)")
def test_checked_build_counts_no_bad_index(tmp_path):
    report = str(tmp_path / "check.json")
    env = dict(os.environ, FABBER_CUDA_LIB=CHECKED, FABBER_CHECK_REPORT=report)
    sel = ("golden or irregular or struck or dirichlet or dims or mrf_slabs or one_part or uneven_slabs "
           "or over_slabs or poly_two_echoes or even_series")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_spatial.py"), os.path.join(ROOT, "tests", "test_gpu_spatial_multi.py"),
           os.path.join(ROOT, "tests", "test_gpu_ar2.py"), "-k", sel]
    run = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = run.stdout[-3000:] + run.stderr[-2000:]
    assert run.returncode == 0, tail
    rep = json.load(open(report))
    log = os.environ.get("FABBER_CHECK_LOG")
    if log:
        with open(log, "a") as f:
            f.write(json.dumps({"report": rep, "pytest_tail": run.stdout.strip().splitlines()[-1]}) + "\n")
    assert rep["compiled_in"] == 1, rep
    assert rep["failures"] == 0, "index checks failed in the kernels: %s\n%s" % (rep, tail)


@pytest.mark.skipif(not os.path.exists(CHECKED), reason="debug build not made")
def test_the_counters_are_live():
    code = ("import ctypes as C, os\n"
            "L = C.CDLL(os.environ['FABBER_CUDA_LIB'])\n"
            "out = (C.c_ulonglong * 2)()\n"
            "assert L.fabber_cuda_check_selftest() == 0\n"
            "rc = L.fabber_cuda_check_report(out)\n"
            "print(rc, out[0], out[1])\n"
            "rc = L.fabber_cuda_check_report(out)\n"
            "print(rc, out[0], out[1])\n")
    run = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FABBER_CUDA_LIB=CHECKED),
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = run.stdout.strip().splitlines()
    assert lines[-2].split() == ["1", "1", "999"] and lines[-1].split() == ["1", "0", "0"], run.stdout


def test_production_build_has_no_checks_compiled_in():
    L = device.lib()
    out = (C.c_ulonglong * 2)()
    L.fabber_cuda_check_report.argtypes = [C.POINTER(C.c_ulonglong)]
    assert L.fabber_cuda_check_report(out) == 0


def test_ordered_sweep_and_slab_engine_reproduce_themselves_bit_for_bit():
    """a data race in the wavefront sweep, the split barrier or the cross-slab flags would make runs differ"""
    nx, ny, nz, T = 12, 10, 16, 24
    y = synth.poly_volume(nx * ny * nz, T, 1, seed=91).numpy()
    idx = np.arange(nx * ny * nz)
    coords = np.ascontiguousarray(np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32))

    def spec():
        sp = abi.ProblemSpec("poly", T, degree=1, prior_types=list("MM"), max_iterations=6)
        sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
        return sp

    first = device.run(spec(), y, spatial=True, coords=coords)
    for _ in range(4):
        again = device.run(spec(), y, spatial=True, coords=coords)
        for k in ("mean", "cov", "noise", "spatial_ak"):
            assert np.array_equal(first[k], again[k]), k
    slabs = [device.run_spatial_multi(spec(), y, coords, 4) for _ in range(4)]
    for s in slabs[1:]:
        for k in ("mean", "cov", "noise"):
            assert np.array_equal(slabs[0][k], s[k]), k
