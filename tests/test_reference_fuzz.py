"""CPU: randomised configurations - model x noise model (white with patterns / masked samples, AR(1), two-echo AR(1)
with every cross-term setting) x all five convergence detectors x N / ARD prior mixes, and NLLS with masks and both
damping rules - through the reference's OWN code (oracle/_ref, compiled from its unchanged sources) and through the
restated oracle, compared voxel by voxel in double. The hand-picked cases of tests/test_reference_build.py pin what
was thought of; this looks for what was not (it is how the masked-Jacobian quirk of inference_nlls.cc:172 would have
been found, had it not been found by hand first). Seeds are fixed: the
300 VB + 80 NLLS + 120 spatial configurations are the same every run."""
import numpy as np
import pytest

import oracle
import refbuild
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import fabber as fab
from parity import tri

pytestmark = pytest.mark.skipif(not refbuild.available(), reason="oracle/_ref not built (no /root/reference here)")

SHAPE, N = (3, 2, 2), 12


def rel(a, b, scale):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), scale)))


def draw_model(rng, tmp_path, T):
    kind = rng.choice(["poly", "exp1", "linear"])
    if kind == "poly":
        deg = int(rng.integers(0, 3))
        i = np.arange(1, T + 1)[:, None]
        y = sum(rng.normal(0, 1, N) * (i / T) ** k * 5 for k in range(deg + 1)) + 10 + rng.normal(0, 1, (T, N))
        return "poly", deg + 1, y, {"model": "poly", "degree": deg}, dict(degree=deg)
    if kind == "exp1":
        t = np.arange(T) * 0.05
        y = rng.uniform(5, 10, N) * np.exp(-rng.uniform(0.5, 2, N) * t[:, None]) + rng.normal(0, 0.05, (T, N))
        return "exp", 2, y, {"model": "exp", "num-exps": 1, "dt": 0.05}, dict(num_exps=1, dt=0.05)
    P = int(rng.integers(1, 4))
    design = rng.normal(0, 1, (T, P))
    y = design @ rng.normal(0, 5, (P, N)) + rng.normal(0, 1, (T, N))
    path = str(tmp_path / ("design_%d.mat" % rng.integers(1 << 30)))
    np.savetxt(path, design, fmt="%.17g")
    return "linear", P, y, {"model": "linear", "basis": path}, dict(design=design)


@pytest.mark.parametrize("seed", range(25))
def test_vb_random_configurations(seed, tmp_path):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        T = 2 * int(rng.integers(6, 20))   # even: two-echo runs need it
        name, P, y, opts, spec = draw_model(rng, tmp_path, T)
        y = y.astype(np.float32)
        conv = str(rng.choice(["maxits", "pointzeroone", "freduce", "trialmode", "lm"]))
        its = int(rng.integers(2, 8))
        opts.update({"method": "vb", "convergence": conv, "max-iterations": its, "allow-bad-voxels": True,
                     "save-mvn": True, "save-free-energy": True})
        spec.update(convergence=conv, max_iterations=its, need_f=True, allow_bad_voxels=True)
        noise = rng.choice(["white", "white", "ar", "ar2"])
        n_noise = 1
        if noise == "white":
            opts["noise"] = "white"
            if rng.random() < 0.4:
                pat = str(rng.choice(["12", "112", "21"]))
                opts["noise-pattern"], spec["noise_pattern"], n_noise = pat, pat, 2
            if rng.random() < 0.4:
                mt = sorted(set(int(x) for x in rng.integers(1, T + 1, 2)))
                for k, m in enumerate(mt):
                    opts["mt%d" % (k + 1)] = m
                spec["masked_timepoints"] = tuple(mt)
        elif noise == "ar":
            opts["noise"], spec["noise"], n_noise = "ar", "ar", 3
        else:
            cross = str(rng.choice(["none", "same", "dual"]))
            opts.update({"noise": "ar", "num-echoes": 2, "ar1-cross-terms": cross})
            spec.update(noise="ar", num_echoes=2, ar_cross_terms=cross)
            n_noise = {"none": 2, "same": 3, "dual": 4}[cross] + 2
        types = "".join(rng.choice(["N", "N", "A"]) for _ in range(P))
        opts["param-spatial-priors"], spec["prior_types"] = types, list(types)
        f = refbuild.ReferenceFabber()
        f.run_with_data(opts, {"data": refbuild.volume(y, SHAPE)})
        mvn, F = f.doubles("finalMVN", N), f.doubles("freeEnergy", N)[0]
        f._destroy_handle()   # NOW: fabber_destroy tears the reference's factories down, a late __del__ would do that under the next run
        ref = oracle.run(abi.ProblemSpec(name, T, **spec), y)
        n_all = P + n_noise
        n_cov = n_all * (n_all + 1) // 2
        assert mvn.shape[0] == n_cov + n_all + 1, (opts, mvn.shape)
        ok = ref["status"] == 0
        assert ok.any(), opts
        std = np.sqrt(np.abs(np.stack([ref["cov"][tri(i, i)] for i in range(P)])))
        for i in range(P):
            assert rel(mvn[n_cov + i][ok], ref["mean"][i][ok], std[i][ok]) < 1e-7, ("mean", i, opts)
            assert rel(mvn[tri(i, i)][ok], ref["cov"][tri(i, i)][ok], 1e-300) < 1e-7, ("var", i, opts)
        assert rel(F[ok], ref["free_energy"][ok], 1.0) < 1e-7, ("F", opts)
        if noise == "white" and name != "exp":   # white noise, linear-in-parameter models: last-bit differences at most
            assert all(rel(mvn[n_cov + i][ok], ref["mean"][i][ok], std[i][ok]) < 1e-11 for i in range(P)), opts


@pytest.mark.skipif(not refbuild.nlls_available(), reason="oracle/_ref/libfabbercore_ref_nlls.so not built")
@pytest.mark.parametrize("seed", range(10))
def test_nlls_random_configurations(seed, tmp_path):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(8):
        T = int(rng.integers(12, 40))
        name, P, y, opts, spec = draw_model(rng, tmp_path, T)
        y = y.astype(np.float32)
        lm = bool(rng.random() < 0.5)
        opts.update({"method": "nlls", "save-mvn": True})
        spec.update(method="nlls", nlls_lm=lm)
        if lm:
            opts["lm"] = True
        if rng.random() < 0.5:
            mt = sorted(set(int(x) for x in rng.integers(1, T + 1, 2)))
            for k, m in enumerate(mt):
                opts["mt%d" % (k + 1)] = m
            spec["masked_timepoints"] = tuple(mt)
        f = refbuild.ReferenceFabber(lib=refbuild.REF_NLLS_LIB)
        f.run_with_data(opts, {"data": refbuild.volume(y, SHAPE)})
        mvn = f.doubles("finalMVN", N)
        f._destroy_handle()
        ref = oracle.run(abi.ProblemSpec(name, T, **spec), y)
        n_cov = P * (P + 1) // 2
        assert mvn.shape[0] == n_cov + P + 1 and np.all(ref["status"] == 0), opts
        for i in range(P):
            scale = np.median(np.abs(ref["mean"][i]))
            assert rel(mvn[n_cov + i], ref["mean"][i], scale) < 1e-6, ("mean", i, opts)
            assert rel(mvn[tri(i, i)], ref["cov"][tri(i, i)], 1e-300) < 1e-6, ("var", i, opts)


@pytest.mark.parametrize("seed", range(15))
def test_spatial_random_configurations(seed, tmp_path):
    """method=spatialvb: random prior-type strings over M / m / P / p / N / A, holes in the mask, spatial-dims,
    update-spatial-prior-on-first-iteration, a speed limit - the ordered sweep and the aK updates of the reference's
    own code against the oracle"""
    rng = np.random.default_rng(3000 + seed)
    n_failed = 0
    for _ in range(8):
        nx, ny, nz = int(rng.integers(2, 6)), int(rng.integers(2, 5)), int(rng.integers(1, 4))
        n = nx * ny * nz
        T = int(rng.integers(12, 30))
        kind = rng.choice(["poly", "linear"])
        if kind == "poly":
            deg = int(rng.integers(0, 3))
            P = deg + 1
            i = np.arange(1, T + 1)[:, None]
            y = sum(rng.normal(0, 1, n) * (i / T) ** k * 5 for k in range(P)) + 10 + rng.normal(0, 1, (T, n))
            opts, spec, name = {"model": "poly", "degree": deg}, dict(degree=deg), "poly"
        else:
            P = int(rng.integers(1, 4))
            design = rng.normal(0, 1, (T, P))
            y = design @ rng.normal(0, 5, (P, n)) + rng.normal(0, 1, (T, n))
            path = str(tmp_path / ("sdesign_%d.mat" % rng.integers(1 << 30)))
            np.savetxt(path, design, fmt="%.17g")
            opts, spec, name = {"model": "linear", "basis": path}, dict(design=design), "linear"
        y = y.astype(np.float32)
        mask = (rng.random((nx, ny, nz)) > 0.15).astype(np.int32)
        if mask.sum() < 2:
            mask[:] = 1
        types = "".join(rng.choice(list("MmPpNA")) for _ in range(P))
        dims = int(rng.integers(1, 4))
        its = int(rng.integers(2, 7))
        first = bool(rng.random() < 0.5)
        speed = float(rng.choice([-1.0, -1.0, 2.0, 10.0]))
        opts.update({"noise": "white", "method": "spatialvb", "param-spatial-priors": types, "spatial-dims": dims,
                     "max-iterations": its, "save-mvn": True, "save-free-energy": True, "spatial-speed": speed})
        if first:
            opts["update-spatial-prior-on-first-iteration"] = True
        spec.update(prior_types=list(types), spatial_dims=dims, max_iterations=its, need_f=True,
                    update_first_iter=first, spatial_speed=speed)
        sel = mask.reshape(-1, order="F") != 0
        idx = np.arange(n)[sel]
        coords = np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32)
        sp = abi.ProblemSpec(name, T, **spec)
        sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
        ref = oracle.run(sp, np.ascontiguousarray(y[:, sel]), spatial=True, coords=coords)
        f = refbuild.ReferenceFabber()
        try:
            f.run_with_data(opts, {"data": refbuild.volume(y, (nx, ny, nz))}, mask=mask)
        except fab.FabberException as e:
            # a configuration the reference itself cannot run (e.g. a blown-up aK makes F non-finite): the oracle
            # must stop with a numerical failure too, not sail through
            n_failed += 1
            f._destroy_handle()
            assert ref["rc"] != 0, (str(e), opts)
            continue
        nv = int(sel.sum())
        mvn, F = f.doubles("finalMVN", nv), f.doubles("freeEnergy", nv)[0]
        f._destroy_handle()
        n_all = P + 1
        n_cov = n_all * (n_all + 1) // 2
        assert np.all(ref["status"] == 0), opts
        std = np.sqrt(np.abs(np.stack([ref["cov"][tri(i, i)] for i in range(P)])))
        for i in range(P):
            assert rel(mvn[n_cov + i], ref["mean"][i], std[i]) < 1e-7, ("mean", i, opts, mask.tolist())
            assert rel(mvn[tri(i, i)], ref["cov"][tri(i, i)], 1e-300) < 1e-7, ("var", i, opts)
        assert rel(F, ref["free_energy"], 1.0) < 1e-7, ("F", opts)
    assert n_failed <= 2
