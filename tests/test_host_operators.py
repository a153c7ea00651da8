"""CPU: the host side of the reference's operator plug-in interface (fabber_core_b200/host/operators.h:
ConvergenceDetector, NoiseModel, Prior, PriorFactory, MVNDist / GammaDist / RunContext) under the reference's own
unit tests, restated in C++ against the same class names (tests/cpp/operators_test.cc follows
test/test_convergence.cc and test/test_priors.cc). The detectors' state machines are additionally compared, call
by call, with the oracle's (which the device detectors are compared with in the GPU tests)."""
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fabber_core_b200")


def build(tmp_path, src, name):
    exe = str(tmp_path / name)
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, src, "-L" + LIBDIR, "-l:libfabbercore_b200.so",
           "-L" + os.path.join(LIBDIR, "csrc"), "-l:libfabber_cuda.so", "-Wl,-rpath," + LIBDIR,
           "-Wl,-rpath," + os.path.join(LIBDIR, "csrc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_reference_unit_tests_restated(tmp_path):
    exe = build(tmp_path, os.path.join(ROOT, "tests", "cpp", "operators_test.cc"), "operators_test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok (0 failures)")


TRACE_SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "%s/fabber_core_b200/host/operators.h"
using namespace fabber_b200;
int main(int argc, char **argv)
{
    FabberRunData rd;
    rd.Set("max-iterations", argv[2]);
    rd.Set("min-fchange", argv[3]);
    rd.Set("max-fchange", argv[3]);
    rd.Set("max-trials", argv[4]);
    std::unique_ptr<ConvergenceDetector> c(ConvergenceDetector::NewFromName(argv[1]));
    c->Initialize(rd);
    for (int i = 5; i < argc; i++)
    {
        const bool t = c->Test(atof(argv[i]));
        printf("%%d %%d %%d %%.9g\n", (int)t, (int)c->NeedSave(), (int)c->NeedRevert(), (double)c->LMalpha());
    }
    return 0;
}
'''


@pytest.mark.parametrize("name", ["maxits", "pointzeroone", "freduce", "trialmode", "lm"])
def test_detector_sequences_equal_the_oracle(tmp_path, name):
    src = tmp_path / "trace.cc"
    src.write_text(TRACE_SRC % ROOT)
    exe = build(tmp_path, str(src), "trace")
    rng = np.random.default_rng(3)
    for trial in range(20):
        # a random walk of F with drops, plateaus and recoveries; the detectors keep being called after they
        # have answered true, exactly as the trace does
        F = np.cumsum(rng.normal(0.5, 2.0, 30)) * rng.choice([1.0, 0.01])
        F = np.round(F, 6)
        t, s, r, a = oracle.convergence_trace(name, F, max_its=6, fchange=0.05, max_trials=3)
        out = subprocess.run([exe, name, "6", "0.05", "3"] + ["%.17g" % x for x in F], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        rows = [l.split() for l in out.stdout.strip().splitlines()]
        assert len(rows) == len(F)
        for i, row in enumerate(rows):
            assert (int(row[0]), int(row[1]), int(row[2])) == (int(t[i]), int(s[i]), int(r[i])), (name, trial, i)
            assert np.float32(float(row[3])) == a[i], (name, trial, i)
            if t[i]:
                break   # the reference stops calling Test() once it has returned true
