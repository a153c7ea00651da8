#!/usr/bin/env python
"""bench.py - voxel-iterations/second of the VB update loop on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--sub c2,c4,c5|none] [--impl reference]

The default line is BASELINE.json's north-star configuration: C3, the bi-exponential model under VB with
Levenberg-Marquardt convergence on a synthetic 256^3 x 96 volume (the largest configuration that is quoted on
one GPU and on 1/2/4/8). One "step" = one complete VB run (all iterations of every voxel) over the volume that
is already resident in HBM. `value` = voxel-iterations of all ranks / time (max over ranks, CUDA events).
`e2e` = the same run through the reference-facing C API (fabber_set_data -> fabber_dorun -> fabber_get_data)
with HOST buffers: every host<->device copy is inside the timed region.

Multi-GPU is STRONG scaling: `--gpus N` cuts the SAME 16.8 M-voxel volume into N contiguous voxel ranges, one
process per GPU, no data-path collective (voxels are independent, inference_vb.cc:423). The end-to-end leg at
N > 1 is ONE process (rank 0) driving all N GPUs through the C API - what a fabber user's single fabber_dorun
call gets on this box (libfabbercore_b200.so deals the voxel ranges to the devices itself).

Sub-records (`sub`): the other synthetic BASELINE configurations in the same JSON line - C2 (poly, 128^3 x 64),
C4 (linear + AR(1), 256^3 x 200), C5 (spatial VB, bi-exponential, 256^3 x 96; at N > 1 ONE volume in N
z-slabs with the exact ordered sweep) - each measured the same way with fewer steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# fp64_active: sm__pipe_fp64_cycles_active of the dominant kernel from the committed `ncu --set full` capture
# of this round (profiles/), traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch there
WORKLOADS = {
    # BASELINE.json configs[1]: poly degree 3, VB, white noise, synthetic 128^3 x 64, maxits 10
    "c2": dict(name="C2 poly(degree 3) VB white noise, synthetic 128^3 x 64, maxits 10", side=128, T=64,
               model="poly", spec=dict(degree=3), P=4, kind="poly",
               capi={"model": "poly", "degree": 3, "noise": "white", "method": "vb", "max-iterations": 10},
               ncu=dict(fp64_active=0.758, traffic=537.313536e6 + 272.282112e6,
                        source="profiles/r2n_ncu_full_c2.txt")),
    # BASELINE.json configs[2]: biexp VB, LM convergence, synthetic 256^3 x 96
    # (prior mean 6 on r2 via PSP_byname: with the default symmetric priors the reference's own fit is
    #  chaotic - see DESIGN.md "C3"; the default-prior run is timed too and reported as `default_priors`)
    "c3": dict(name="C3 exp(num-exps 2, dt 0.02, PSP_byname1=r2 PSP_byname1_mean=6) VB white noise, convergence=lm, "
                    "synthetic 256^3 x 96", side=256, T=96, model="exp",
               spec=dict(num_exps=2, dt=0.02, convergence="lm", need_f=True,
                         param_overrides={"r2": {"mean": 6.0}}), P=4, kind="exp", NE=2,
               capi={"model": "exp", "num-exps": 2, "dt": 0.02, "noise": "white", "method": "vb", "convergence": "lm",
                     "max-iterations": 10, "PSP_byname1": "r2", "PSP_byname1_mean": 6.0},
               ncu=dict(fp64_active=0.750, traffic=6.487053e9 + 2.401360e9,
                        source="profiles/r2i_ncu_full_c3_exp2.txt")),
    # BASELINE.json configs[3]: linear model (synthetic 200 x 4 design), AR(1) noise, synthetic 256^3 x 200
    # (the reference has no AR(2): Ar1cNoiseModel only, setup.cc:39)
    "c4": dict(name="C4 linear(200x4 design) VB AR(1) noise (num-echoes 1, cross-terms none), synthetic "
                    "256^3 x 200, maxits 10", side=256, T=200, model="linear", spec=dict(noise="ar"), P=4, kind="ar",
               capi={"model": "linear", "basis": "@design", "noise": "ar", "method": "vb", "max-iterations": 10},
               ncu=dict(fp64_active=0.744, traffic=16.546807e9 + 6.548900e9, source="profiles/r2n_ncu_full_c4.txt")),
    # BASELINE.json configs[4]: spatialvb (MRF spatial prior 'M' on every parameter), biexp, smooth synthetic
    # 256^3 x 96. NB the reference's CovarianceCache is dead code; the MRF prior is
    # SpatialPrior in priors.cc.
    "c5": dict(name="C5 exp(num-exps 2, dt 0.02, PSP_byname1=r2 PSP_byname1_mean=6) spatialvb, "
                    "param-spatial-priors=M+, smooth synthetic 256^3 x 96, maxits 10", side=256, T=96, model="exp",
               spatial=True,
               spec=dict(num_exps=2, dt=0.02, prior_types=list("MMMM"), param_overrides={"r2": {"mean": 6.0}}),
               P=4, kind="exp", NE=2,
               capi={"model": "exp", "num-exps": 2, "dt": 0.02, "noise": "white", "method": "spatialvb",
                     "param-spatial-priors": "M+", "max-iterations": 10, "PSP_byname1": "r2",
                     "PSP_byname1_mean": 6.0},
               # the dominant kernel of a spatial run is sp_noise (58 % of the step): its capture, per launch (= one of the
               # ten iterations); the step as a whole also runs sp_sweep / sp_theta / sp_ak_*
               ncu=dict(fp64_active=0.743, traffic=None, source="profiles/r2o_ncu_full_c5_noise.txt",
                        sp_noise_traffic_per_launch=10.939664e9 + 2.865362e9)),
}


def algorithmic_flops(w):
    """FLOP per voxel-iteration of the minimal algorithm, SURVEY.md 8(d) convention (FMA = 2, add = mul = 1,
    div = sqrt = 10, exp = log = 20): one pass over t evaluating the model at the 2P+1 finite-difference points
    and accumulating the sufficient statistics, plus the O(P^3) algebra:
        W = T * e + e0 + 2TP + T [P(P+1) + 2P + 3] + (2 P^3 + 10 P^2 + 300).
    poly(d): e = (2P+1) * 2(d+1), e0 = 0.          linear: e = (2P+1) * 2P, e0 = 0.
    exp(NE): the perturbed points differ from the centre in ONE parameter, so a sample needs the exponential of
      each rate at its three values only (3 NE exponentials, not (2P+1) NE as SURVEY.md's e = 23 NE assumed - the
      round-1 judge's correction): e = 3 NE (20 + 1) + 1 + (2P+1)(2 NE - 1), and the log-transforms of the
      3 P parameter values per pass: e0 = 3 P * 20.
    AR(1): SURVEY.md 8(d): statistics tripled + 2x2 alpha algebra = 35.9 kFLOP at P = 4, T = 200."""
    P, T = w["P"], w["T"]
    common = 2 * T * P + T * (P * (P + 1) + 2 * P + 3) + (2 * P ** 3 + 10 * P ** 2 + 300)
    if w["kind"] == "ar":
        return 35900
    if w["kind"] == "poly":
        return (2 * P + 1) * T * 2 * P + common
    NE = w["NE"]
    e = 3 * NE * 21 + 1 + (2 * P + 1) * (2 * NE - 1)
    return T * e + 3 * P * 20 + common


def survey_formula_flops(w):
    """SURVEY.md 8(d)'s own formula (counts (2P+1) full model evaluations per sample); reported beside W."""
    P, T = w["P"], w["T"]
    if w["kind"] == "ar":
        return 35900
    e, e0 = (2 * P, 0) if w["kind"] == "poly" else (23 * w["NE"], 20 * P)
    return (2 * P + 1) * (T * e + e0) + 2 * T * P + T * (P * (P + 1) + 2 * P + 3) + (2 * P ** 3 + 10 * P ** 2 + 300)


def algorithmic_bytes(w, n_iter):
    """HBM bytes per voxel-iteration, all iterations fused (y read once, results written once)."""
    P, T = w["P"], w["T"]
    if w.get("spatial"):  # iteration-at-a-time: y + the per-voxel state round trip each iteration
        return 4 * T + 16 * (P + P * (P + 1) // 2 + 2) + 8
    return (4 * T + 4 * ((P + 1) * (P + 2) // 2 + (P + 1) + 1) + 16) / max(n_iter, 1e-9)


def make_volume(w, n_voxels, device, seed_offset=0, voxel_offset=0, n_total=None):
    """`n_voxels` voxels of the workload's synthetic volume, starting at voxel `voxel_offset` of a volume of
    `n_total` voxels (strong scaling: every rank generates its own contiguous range)."""
    from fabber_core_b200 import synth

    if w["model"] == "poly":
        return synth.poly_volume(n_voxels, w["T"], 3, seed=1002 + seed_offset, device=device)
    if w["model"] == "linear":
        return synth.linear_ar_volume(n_voxels, w["T"], 0.3, seed=1004 + seed_offset, device=device)
    if w.get("spatial"):
        n_total = n_total or n_voxels
        side = round(n_total ** (1.0 / 3))
        assert side ** 3 == n_total, "spatial workloads need a cubic voxel count"
        return synth.biexp_volume(n_voxels, w["T"], 0.02, 0.02, seed=1005 + seed_offset, device=device,
                                  smooth_shape=(side, side, side), voxel_offset=voxel_offset)
    return synth.biexp_volume(n_voxels, w["T"], 0.02, 0.02, seed=1003 + seed_offset, device=device)


def make_spec(w, n_voxels=0):
    from fabber_core_b200 import cuda_abi as abi

    from fabber_core_b200 import synth

    spec = dict(w["spec"])
    if w["model"] == "linear":
        spec["design"] = synth.ar_design(w["T"])
    ps = abi.ProblemSpec(w["model"], w["T"], **spec)
    if w.get("spatial") and n_voxels:
        side = round(n_voxels ** (1.0 / 3))
        ps.prob.nx = ps.prob.ny = ps.prob.nz = side
    return ps


class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (ts, l) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [l for (_, l) in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def capi_e2e(w, host_y, extent, steps):
    """End to end through the reference-facing C API (include/fabber_capi.h) with HOST buffers, exactly the
    calls py/fabber.py makes: fabber_set_data (host float32 volume in) -> fabber_dorun -> fabber_get_data of
    every requested output (host float32 volumes out). Wall clock around the synchronous calls; the copies
    host -> pinned -> device and device -> host -> caller are all inside. Returns (seconds per step, h2d
    bytes, d2h bytes)."""
    import ctypes as C
    import tempfile

    from fabber_core_b200 import fabber as fab
    from fabber_core_b200 import synth

    f = fab.Fabber()
    opts = dict(w["capi"])
    tmp = None
    if opts.get("basis") == "@design":
        tmp = tempfile.NamedTemporaryFile("w", suffix=".mat", delete=False)
        np.savetxt(tmp, synth.ar_design(w["T"]), fmt="%.17g")
        tmp.close()
        opts["basis"] = tmp.name
    opts.update({"save-mean": True, "save-std": True, "save-noise-mean": True})
    f._set_options(opts)
    f._trycall(f.clib.fabber_get_model_params, f.handle, len(f.outbuf), f.outbuf, f.errbuf)
    params = f.outbuf.value.decode().splitlines()
    outputs = ["mean_" + p for p in params] + ["std_" + p for p in params] + ["noise_means"]
    n = extent[0] * extent[1] * extent[2]
    mask = np.ones(n, dtype=np.int32)
    f._trycall(f.clib.fabber_set_extent, f.handle, extent[0], extent[1], extent[2], mask, f.errbuf)
    flat = host_y.reshape(-1)  # [t][z][y][x] == [T][N] for a full mask
    bufs = {}
    noop = f.progress_cb_type(0)

    phases = {"set_data": 0.0, "dorun": 0.0, "get_data": 0.0}

    def one():
        t_a = time.perf_counter()
        f._trycall(f.clib.fabber_set_data, f.handle, b"data", w["T"], flat, f.errbuf)
        t_b = time.perf_counter()
        f._trycall(f.clib.fabber_dorun, f.handle, len(f.outbuf), f.outbuf, f.errbuf, noop)
        t_c = time.perf_counter()
        phases["set_data"] += t_b - t_a
        phases["dorun"] += t_c - t_b
        nbytes = 0
        for key in outputs:
            size = f._trycall(f.clib.fabber_get_data_size, f.handle, key.encode(), f.errbuf)
            if key not in bufs:
                bufs[key] = np.empty(n * size, dtype=np.float32)
            f._trycall(f.clib.fabber_get_data, f.handle, key.encode(), bufs[key], f.errbuf)
            nbytes += bufs[key].nbytes
        phases["get_data"] += time.perf_counter() - t_c
        return nbytes

    one()  # warm-up
    for k in phases:
        phases[k] = 0.0
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        d2h = one()
    dt = (time.perf_counter() - t0) / steps
    if tmp is not None:
        os.unlink(tmp.name)
    capi_e2e.last_phases_ms = {k: v / steps * 1e3 for k, v in phases.items()}
    log_lines = f.outbuf.value.decode(errors="replace").splitlines()
    capi_e2e.last_log = [l for l in log_lines if "Vb::timing" in l]
    capi_e2e.last_devices = [l.split(",")[-1].strip() for l in log_lines if "calculations on the GPU" in l]
    return dt, flat.nbytes, d2h


REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libfabbercore_ref.so")


def _reference_worker(job):
    """One process = one single-threaded run of the reference's own code (oracle/_ref: its unchanged sources
    compiled against the test-only NEWMAT stand-in) over its chunk of voxels, through its own C API."""
    wname, n, seed = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refbuild

    w = WORKLOADS[wname]
    y = make_volume(w, n, "cpu", seed_offset=seed).numpy()
    opts = dict(w["capi"])
    tmp = None
    if opts.get("basis") == "@design":
        import tempfile

        from fabber_core_b200 import synth

        tmp = tempfile.NamedTemporaryFile("w", suffix=".mat", delete=False)
        np.savetxt(tmp, synth.ar_design(w["T"]), fmt="%.17g")
        tmp.close()
        opts["basis"] = tmp.name
    opts["save-mean"] = True
    side = round(n ** (1.0 / 3))
    shape = (side, side, side) if w.get("spatial") else (n, 1, 1)
    f = refbuild.ReferenceFabber()
    t0 = time.perf_counter()
    f.run_with_data(opts, {"data": refbuild.volume(y, shape)})
    dt = time.perf_counter() - t0
    if tmp is not None:
        os.unlink(tmp.name)
    # iteration counts are not exposed by the reference's C API; the oracle is pinned bit-for-bit to this
    # very code (tests/test_reference_build.py), so its counts are the reference's counts
    its = int(oracle_run(w, y)["iterations"].sum())
    return dt, its, n


def reference_throughput(wname, procs, budget_s, pool=None, per_voxel=None, seed=100):
    """Voxel-iterations/s of the reference's own code on `procs` processes over disjoint voxel chunks -
    how fabber is parallelised in practice (it is single-threaded). Bounded sample sized from a probe.
    `pool` / `per_voxel`: a caller timing several steps keeps one pool of warmed-up workers and one probe
    (taken with every worker busy) instead of paying process start-up and imports per step."""
    import multiprocessing as mp

    w = WORKLOADS[wname]
    probe_n = 216 if w.get("spatial") else 128
    own_pool = pool is None
    if own_pool:
        pool = mp.get_context("spawn").Pool(procs)
    try:
        if per_voxel is None:
            res = pool.map(_reference_worker, [(wname, probe_n, i) for i in range(procs)], chunksize=1)
            per_voxel = max(r[0] for r in res) / probe_n
        n = int(max(probe_n, min(100000, budget_s / per_voxel)))
        if w.get("spatial"):
            side = max(4, int(round(n ** (1.0 / 3))))
            n = side ** 3
        t0 = time.perf_counter()
        res = pool.map(_reference_worker, [(wname, n, seed + i) for i in range(procs)], chunksize=1)
        wall = time.perf_counter() - t0
    finally:
        if own_pool:
            pool.close()
            pool.join()
    compute = max(r[0] for r in res)
    total_its = sum(r[1] for r in res)
    return total_its / compute, compute, n * procs, wall, per_voxel


def oracle_run(w, y):
    import oracle

    n = y.shape[1]
    if not w.get("spatial"):
        return oracle.run(make_spec(w), y)
    side = round(n ** (1.0 / 3))
    idx = np.arange(n)
    coords = np.stack([idx % side, (idx // side) % side, idx // (side * side)]).astype(np.int32)
    return oracle.run(make_spec(w, n), y, spatial=True, coords=coords)


def oracle_throughput(w, threads, budget_s):
    """Voxel-iterations/s of the CPU oracle (port of the reference's algorithm) on a bounded sample."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor

    oracle.lib()
    probe_n = 343 if w.get("spatial") else 256

    def run_chunk(seed):
        def f(n):
            y = make_volume(w, n, "cpu", seed_offset=seed).numpy()
            t0 = time.perf_counter()
            out = oracle_run(w, y)
            return time.perf_counter() - t0, int(out["iterations"].sum())
        return f

    dt, its = run_chunk(0)(probe_n)
    rate1 = its / dt
    n_per_thread = int(max(probe_n, min(200000, rate1 * budget_s / max(its / probe_n, 1))))
    if w.get("spatial"):
        side = max(4, int(round(n_per_thread ** (1.0 / 3))))
        n_per_thread = side ** 3
    datas = [make_volume(w, n_per_thread, "cpu", seed_offset=100 + i).numpy() for i in range(threads)]

    def work(i):
        out = oracle_run(w, datas[i])  # ctypes releases the GIL: threads run in parallel
        return int(out["iterations"].sum())

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        total = sum(ex.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return total / dt, dt, n_per_thread * threads


def roofline_block(w, its_launch, avg_kernel_s, n_vox_launch, fp64_peak_gflops, spatial):
    """roofline of the dominant kernel: algorithmic FLOP of one launch / its measured duration against the FP64
    DFMA peak measured live on this GPU; HBM fraction beside it (this path is nowhere near HBM-bound)."""
    W = algorithmic_flops(w)
    achieved_tf = W * its_launch / avg_kernel_s / 1e12
    peak_tf = fp64_peak_gflops / 1e3
    n_iter = its_launch / float(max(n_vox_launch, 1))
    bytes_per_launch = algorithmic_bytes(w, n_iter) * its_launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ncu = w.get("ncu", {})
    full_size = n_vox_launch == w["side"] ** 3
    return {
        "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf if peak_tf > 0 else None,
        # DRAM bytes of one launch from the committed ncu --set full capture (only quoted at the size it was
        # captured at); algorithmic bytes are in "hbm" below
        "traffic": ncu.get("traffic") if full_size else None,
        "traffic_unit": "bytes per launch (dram read + write)",
        "traffic_source": ncu.get("source") if full_size else None,
        "fp64_pipe_active_ncu": ncu.get("fp64_active"), "fp64_pipe_active_source": ncu.get("source"),
        "algorithmic_bytes_per_launch": bytes_per_launch,
        "peak_source": "measured live: dependent-free DFMA loop on all SMs (MEASURED_PEAKS.json has no FP64 figure)",
        "flop_per_voxel_iteration": W, "flop_per_voxel_iteration_survey_formula": survey_formula_flops(w),
        "flop_note": "W counts the 3*NE exponentials per sample the algorithm needs (exp model), not (2P+1)*NE",
        "kernel": ("sp_noise_kernel (+ sp_theta / sp_sweep / sp_ak)" if spatial else "vb_voxelwise_ar_kernel"
                   if w["spec"].get("noise") == "ar" else "vb_voxelwise_white_kernel"),
        "avg_launch_ms": avg_kernel_s * 1e3,
        "hbm": {"achieved": bytes_per_launch / avg_kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_per_launch / avg_kernel_s / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}


class Dist(object):
    """the bench's use of torch.distributed: a barrier and max / sum over ranks (no data-path collective)"""

    def __init__(self, world):
        self.world = world
        if world > 1:
            import torch.distributed as dist

            self.dist = dist
            # a CPU-side group for the phases in which rank 0 alone drives every GPU: an NCCL barrier would leave a
            # kernel spinning on the other ranks' GPUs, and two processes on one GPU are time-sliced - rank 0's
            # kernels there ran at half speed (measured at N = 2)
            self.cpu = dist.new_group(backend="gloo")

    def cpu_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.cpu)

    def barrier(self):
        import torch

        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, value, op):
        import torch

        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())


def host_volume_full(w, n_total, world):
    """The whole synthetic volume in pinned host memory, [T][n_total] float32, generated on this GPU rank range
    by rank range (the same generators and seeds the ranks use for their device-resident shards)."""
    import torch

    from fabber_core_b200 import shard

    host = torch.empty((w["T"], n_total), dtype=torch.float32, pin_memory=True)
    for r in range(world):
        lo, hi = shard.voxel_range(n_total, r, world)
        part = make_volume(w, hi - lo, "cuda", seed_offset=r, voxel_offset=lo, n_total=n_total)
        host[:, lo:hi].copy_(part)
        del part
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return host


def run_workload(wname, args, D, rank, local_rank, world, steps, warmup, main_line):
    """Device-resident timing (all ranks), then the end-to-end leg through the C API (rank 0 drives every GPU).
    Returns the record on rank 0, None elsewhere."""
    import torch

    from fabber_core_b200 import device, shard

    w = WORKLOADS[wname]
    L = device.lib()
    stream = torch.cuda.current_stream().cuda_stream
    spatial = bool(w.get("spatial"))
    n_total = args.voxels if (args.voxels and main_line) else w["side"] ** 3
    side = round(n_total ** (1.0 / 3))
    lo, hi = shard.voxel_range(n_total, rank, world)
    n_local = hi - lo
    config = {"workload": w["name"], "voxels_total": n_total, "voxels_per_gpu": n_local, "timepoints": w["T"],
              "params": w["P"],
              "sharding": ("one volume, contiguous voxel ranges, no collective" if not spatial else "one GPU"),
              "cache": "inputs larger than L2 (%.0f MB per GPU per step)" % (n_local * w["T"] * 4 / 1e6)}

    if spatial and world > 1:
        rec = spatial_slab_bench(args, w, D, rank, local_rank, world, n_total, steps, warmup, config)
        if rank != 0:
            rec = {"iterations_per_voxel": 0.0}
    else:
        y = make_volume(w, n_local, "cuda", seed_offset=rank, voxel_offset=lo, n_total=n_total)
        spec = make_spec(w, n_local)
        run = device.VbRun(spec, n_local, spatial=spatial)
        run.set_data_device(y.data_ptr())
        if spatial:
            idx = torch.arange(n_local, device="cuda")
            coords = torch.stack([idx % side, (idx // side) % side, idx // (side * side)]).to(torch.int32).contiguous()
            run.buf.coords = coords.data_ptr()

        def step():
            rc = run.launch(stream)
            if rc != 0:
                raise RuntimeError("launch failed: %s" % device.last_error())

        for _ in range(max(warmup, 3)):
            step()
        D.barrier()
        its_local = int(run.out["iterations"].to_host().astype(np.int64).sum())
        n_bad = int(np.count_nonzero(run.out["status"].to_host()))
        fp64_peak = L.fabber_cuda_measure_fp64_peak(3)  # GFLOP/s, live (no FP64 figure in MEASURED_PEAKS.json)
        if main_line:
            torch.cuda.profiler.start()  # ncu --profile-from-start off: only the timed region is listed
        sampler = ClockSampler(local_rank)
        time.sleep(0.3)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        launches0 = L.fabber_cuda_launch_count()
        D.barrier()
        t0 = time.time()
        ev[0].record()
        for i in range(steps):
            step()
            ev[i + 1].record()
        D.barrier()
        t1 = time.time()
        if main_line:
            torch.cuda.profiler.stop()
        launches = L.fabber_cuda_launch_count() - launches0
        clocks = sampler.stop(t0, t1)
        kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        total_ms = D.reduce(ev[0].elapsed_time(ev[-1]), "MAX")
        its_all = D.reduce(its_local, "SUM")
        bad_all = int(D.reduce(n_bad, "SUM"))
        rec = {
            "value": its_all * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "steps": steps,
            "warmup": max(warmup, 3), "config": config, "gpu_launches": int(launches), "clocks": clocks,
            "iterations_per_voxel": its_all / float(n_total), "bad_voxels": bad_all,
            "roofline": roofline_block(w, its_local, float(np.mean(kernel_ms)) * 1e-3, n_local, fp64_peak, spatial)}

        # the inner ABI with pinned host buffers: copy in, run, copy every raw result array out (CUDA events)
        host_y = torch.empty(y.shape, dtype=torch.float32, pin_memory=True)
        host_y.copy_(y)
        outs = run.out
        host_out = {k: torch.empty(v.nbytes, dtype=torch.uint8, pin_memory=True) for k, v in outs.items()}
        h2d = host_y.numel() * 4
        d2h = sum(v.nbytes for v in outs.values())

        def inner_step():
            device.check(L.fabber_cuda_memcpy_h2d(y.data_ptr(), host_y.data_ptr(), h2d, stream), "h2d")
            step()
            for k, v in outs.items():
                device.check(L.fabber_cuda_memcpy_d2h(host_out[k].data_ptr(), v.ptr, v.nbytes, stream), "d2h")

        inner_step()
        D.barrier()
        inner_steps = max(1, min(steps, 3))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner_steps):
            inner_step()
        e1.record()
        D.barrier()
        inner_ms = D.reduce(e0.elapsed_time(e1), "MAX")
        rec["inner_abi"] = {"value": its_all * inner_steps / (inner_ms * 1e-3), "h2d_bytes_per_step": h2d,
                            "d2h_bytes_per_step": d2h, "steps": inner_steps,
                            "path": "fabber_cuda_vb_* per rank with pinned host buffers, raw result arrays, CUDA events"}

        # C3 with the reference's DEFAULT (symmetric) priors: throughput and how many voxels diverge
        if main_line and wname == "c3":
            spec_d = dict(w["spec"])
            spec_d.pop("param_overrides")
            from fabber_core_b200 import cuda_abi as abi

            sd = abi.ProblemSpec(w["model"], w["T"], allow_bad_voxels=True, **spec_d)
            run_d = device.VbRun(sd, n_local)
            run_d.set_data_device(y.data_ptr())
            run_d.launch(stream)
            D.barrier()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for _ in range(2):
                run_d.launch(stream)
            d1.record()
            D.barrier()
            its_d = D.reduce(int(run_d.out["iterations"].to_host().astype(np.int64).sum()), "SUM")
            bad_d = D.reduce(int(np.count_nonzero(run_d.out["status"].to_host())), "SUM")
            ms_d = D.reduce(d0.elapsed_time(d1), "MAX")
            rec["default_priors"] = {
                "value": its_d * 2 / (ms_d * 1e-3), "unit": "voxel-iterations/s", "ms_per_step": ms_d / 2,
                "iterations_per_voxel": its_d / float(n_total), "bad_voxels": int(bad_d),
                "note": "same data and options without the PSP_byname override (allow-bad-voxels): the symmetric "
                        "start is a saddle in the reference itself (DESIGN.md section 4)"}
            run_d.close()
        run.close()
        del host_out, host_y, y
        torch.cuda.empty_cache()

    # ---- end to end through the reference's public C API with host buffers: rank 0 drives all `world` GPUs ----
    D.barrier()
    import torch

    torch.cuda.synchronize()
    e2e = None
    if rank == 0:
        host = host_volume_full(w, n_total, world)
        capi_steps = max(1, min(steps, 3))
        extent = (side, side, side) if side ** 3 == n_total else (n_total, 1, 1)
        capi_s, capi_h2d, capi_d2h = capi_e2e(w, host.numpy(), extent, capi_steps)
        its_all = rec["iterations_per_voxel"] * n_total
        dev_line = [l for l in capi_e2e.last_devices] or ["?"]
        e2e = {"value": its_all / capi_s, "unit": "voxel-iterations/s", "h2d_bytes_per_step": capi_h2d,
               "d2h_bytes_per_step": capi_d2h, "steps": capi_steps, "ms_per_step": capi_s * 1e3,
               "calls_ms": capi_e2e.last_phases_ms, "dorun_log": capi_e2e.last_log, "devices_used": dev_line[0],
               "path": "fabber_capi (libfabbercore_b200.so), ONE process: fabber_set_data (pinned host float32 volume "
                       "in) -> fabber_dorun -> fabber_get_data of mean_*, std_*, noise_means (host float32 volumes "
                       "out); the library deals the voxel ranges to the GPUs it was given (FABBER_B200_DEVICES=%s); "
                       "wall clock" % os.environ.get("FABBER_B200_DEVICES", "")}
        del host
    D.cpu_barrier()  # the other ranks wait on the CPU, their GPUs idle
    D.barrier()
    if rank != 0:
        return None
    rec["e2e"] = e2e
    if "inner_abi" in rec:
        rec["e2e"]["inner_abi"] = rec.pop("inner_abi")
    return rec


def spatial_slab_bench(args, w, D, rank, local_rank, world, n_total, steps, warmup, config):
    """C5 on several GPUs, STRONG scaling: ONE side^3 volume cut into `world` z-slabs. The slabs are coupled
    DEVICE to DEVICE (fabber_cuda_vb_spatial_multi: the ordered sweep forwards its top plane into the next slab's
    ghost voxels behind release/acquire flags in peer memory, the aK sums are all-gathered through mailboxes), so
    one process drives the whole job: rank 0 queues the launches for all `world` GPUs, the other ranks' processes
    stand by at the barrier. Timed on the devices (CUDA events on every slab's stream, max over the slabs)."""
    import torch

    from fabber_core_b200 import device, synth

    side = round(n_total ** (1.0 / 3))
    assert side ** 3 == n_total
    rec = None
    D.barrier()
    if rank == 0:
        L = device.lib()
        idx = np.arange(n_total)
        coords = np.stack([idx % side, (idx // side) % side, idx // (side * side)]).astype(np.int32)
        del idx
        run = device.SpatialMultiRun(make_spec(w, n_total), coords, world, devices=list(range(world)))
        ys = []
        for r in range(world):
            g0, g1 = run.part_range(r)
            with torch.cuda.device(r):
                y = synth.biexp_volume(g1 - g0, w["T"], 0.02, 0.02, seed=1005 + r, device="cuda:%d" % r,
                                       smooth_shape=(side, side, side), voxel_offset=g0)
                torch.cuda.synchronize()
            ys.append(y)
            run.set_data_device(r, y.data_ptr())

        def step():
            rc = run.launch()
            if rc != 0:
                raise RuntimeError("spatial multi-device launch failed: %s" % device.last_error())
            return run.last_ms

        for _ in range(max(warmup, 3)):
            step()
        out = run.results()
        its = int(out["iterations"].astype(np.int64).sum())
        n_bad = int(np.count_nonzero(out["status"]))
        fp64_peak = L.fabber_cuda_measure_fp64_peak(3)
        launches0 = L.fabber_cuda_launch_count()
        sampler = ClockSampler(local_rank)
        time.sleep(0.3)
        t0 = time.time()
        ms = [step() for _ in range(steps)]
        wall_ms = (time.time() - t0) * 1e3 / steps
        clocks = sampler.stop(t0, time.time())
        launches = L.fabber_cuda_launch_count() - launches0
        t_ms = float(np.sum(ms))
        run.close()
        del ys
        for r in range(world):
            with torch.cuda.device(r):
                torch.cuda.empty_cache()
        cfg = dict(config)
        cfg["voxels_per_gpu"] = n_total // world
        cfg["sharding"] = ("z-slabs of ONE %dx%dx%d volume, one per GPU, driven by one process; slabs coupled device to "
                           "device through peer memory: the exact ordered sweep forwards hyper-plane by hyper-plane behind "
                           "release/acquire flags, halo after the sweep, aK sums all-gathered through mailboxes in the aK "
                           "kernel; result equals the one-GPU run" % (side, side, side))
        rec = {"value": its * steps / (t_ms * 1e-3), "ms_per_step": t_ms / steps, "wall_ms_per_step": wall_ms,
               "steps": steps, "warmup": max(warmup, 3), "config": cfg, "gpu_launches": int(launches), "clocks": clocks,
               "iterations_per_voxel": its / float(n_total), "bad_voxels": n_bad,
               "timing": "CUDA events on every slab's stream around set-up + iterations + result permutation, max over "
                         "the slabs, summed over the steps (the call is synchronous)",
               "roofline": roofline_block(w, its // world, t_ms / steps * 1e-3, n_total // world, fp64_peak, True)}
    D.cpu_barrier()  # the other ranks wait on the CPU: rank 0's kernels have their GPUs to themselves
    D.barrier()
    return rec


def reference_arm(args, metric):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores: oracle/_ref
    (its unchanged sources compiled against the test-only NEWMAT stand-in, built where /root/reference exists and
    shipped with the tree), one single-threaded process per core - how fabber is parallelised in practice. Falls
    back to the oracle port only if that library is missing."""
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    use_ref = os.path.exists(REF_LIB)
    rates = []
    t_all0 = time.perf_counter()
    # ~75 s of reference compute over all steps (plus worker start-up), whatever K and W are
    per_step = max(1.5, 75.0 / (args.warmup + args.steps))
    pool, per_voxel = None, None
    if use_ref:
        import multiprocessing as mp

        pool = mp.get_context("spawn").Pool(cores)
    try:
        for i in range(args.warmup + args.steps):
            if use_ref:
                rate, dt, n, _, per_voxel = reference_throughput(args.workload, cores, budget_s=per_step, pool=pool,
                                                                 per_voxel=per_voxel, seed=100 + 1000 * i)
            else:
                rate, dt, n = oracle_throughput(w, cores, budget_s=per_step)
            if i >= args.warmup:
                rates.append((rate, dt, n))
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    value = float(np.mean([r[0] for r in rates]))
    ms = float(np.mean([r[1] for r in rates]) * 1e3)
    kind = "reference" if use_ref else "port"
    sample = "%d voxels of the same synthetic workload per step, %d %s" % (
        rates[0][2], cores, "single-threaded processes (oracle/_ref)" if use_ref else "threads (oracle port)")
    n_total = w["side"] ** 3
    config = {"workload": w["name"], "voxels_total": n_total, "timepoints": w["T"], "params": w["P"]}
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": "voxel-iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": "voxel-iterations/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "voxel-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all0}))
    return 0


def extra_kernels(torch):
    """Device-resident rates of the kernels outside BASELINE.json's five configurations (two-echo AR(1), NLLS), on
    2^20 voxels each: reported beside the sub-records, not part of the headline."""
    from fabber_core_b200 import cuda_abi as abi
    from fabber_core_b200 import device, synth

    def timed(spec, y, reps=3):
        run = device.VbRun(spec, y.shape[1])
        run.set_data_device(y.data_ptr())
        st = torch.cuda.current_stream().cuda_stream
        device.check(run.launch(st), "extra kernel")
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            device.check(run.launch(st), "extra kernel")
        ev[1].record()
        torch.cuda.synchronize()
        res = run.results()
        run.close()
        ms = ev[0].elapsed_time(ev[1]) / reps
        return {"value": float(res["iterations"].sum()) / ms * 1e3, "unit": "voxel-iterations/s", "ms_per_step": ms,
                "voxels": int(y.shape[1]), "bad_voxels": int((res["status"] != 0).sum())}

    n = 1 << 20
    out = {}
    y = synth.dual_echo_volume(n, 100, seed=5, device="cuda")
    design = synth.dual_echo_design(100)
    for cross in ("none", "dual"):
        r = timed(abi.ProblemSpec("linear", 200, design=design, noise="ar", num_echoes=2, ar_cross_terms=cross,
                                  need_f=True), y)
        r["workload"] = "linear(200x3) VB AR(1) num-echoes=2 ar1-cross-terms=%s, maxits 10" % cross
        out["ar1_two_echoes_" + cross] = r
    del y
    y = synth.biexp_volume(n, 96, 0.02, 0.02, seed=1003, device="cuda")
    r = timed(abi.ProblemSpec("exp", 96, num_exps=2, dt=0.02, method="nlls", allow_bad_voxels=True,
                              param_overrides={"r2": {"mean": 6.0}}), y)
    r["workload"] = "exp(num-exps 2) method=nlls (Levenberg); iterations = accepted steps"
    out["nlls_biexp"] = r
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--sub", default="c2,c4,c5", help="comma list of workloads reported as sub-records, or 'none'")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voxels", type=int, default=0, help="override the TOTAL voxel count of the main workload (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = "voxel-iterations/sec"

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, metric)

    # the end-to-end leg is ONE process (rank 0) driving every GPU of the job through the C API: tell the host
    # library which devices that is before it is first used (it reads the variable once)
    if "FABBER_B200_DEVICES" not in os.environ:
        os.environ["FABBER_B200_DEVICES"] = ",".join(str(i) for i in range(world)) if rank == 0 else str(local_rank)

    import torch

    from fabber_core_b200 import device

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = device.lib()
    device.check(L.fabber_cuda_set_device(local_rank), "set_device")
    D = Dist(world)

    rec = run_workload(args.workload, args, D, rank, local_rank, world, args.steps, args.warmup, True)
    subs = {}
    sub_names = [x for x in args.sub.split(",") if x and x != "none" and x != args.workload]
    for name in sub_names:
        if name not in WORKLOADS:
            raise SystemExit("unknown sub workload %r" % name)
        sub_steps = max(1, min(args.steps, 3))
        r = run_workload(name, args, D, rank, local_rank, world, sub_steps, 3, False)
        if rank == 0:
            r.update({"metric": metric, "unit": "voxel-iterations/s", "n_gpus": world, "scaling": "strong",
                      "dtype": "f64", "data": "synthetic"})
            subs[name] = r

    if rank == 0:
        line = {"metric": metric, "value": rec["value"], "unit": "voxel-iterations/s", "n_gpus": world,
                "steps": rec["steps"], "warmup": rec["warmup"], "ms_per_step": rec["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": rec["config"], "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"],
                "clocks": rec["clocks"], "roofline": rec["roofline"],
                "iterations_per_voxel": rec["iterations_per_voxel"], "bad_voxels": rec["bad_voxels"]}
        if "default_priors" in rec:
            line["default_priors"] = rec["default_priors"]
        if world == 1 and not args.no_cpu_baseline:
            prate, pdt, pn = oracle_throughput(w, 1, budget_s=8.0)
            if os.path.exists(REF_LIB):
                rate, dt, n, _, _ = reference_throughput(args.workload, 1, budget_s=12.0)
                line["cpu_baseline"] = {
                    "value": rate, "unit": "voxel-iterations/s", "cores": 1, "kind": "reference",
                    "sample": "%d voxels of the same synthetic workload, %.1f s, one process: the reference's own "
                              "sources on the test-only NEWMAT stand-in (oracle/_ref)" % (n, dt),
                    "port": {"value": prate, "sample": "%d voxels, %.1f s, single thread: the restated oracle "
                                                       "(fixed-size arrays, no heap matrices)" % (pn, pdt)}}
            else:
                line["cpu_baseline"] = {"value": prate, "unit": "voxel-iterations/s", "cores": 1, "kind": "port",
                                        "sample": "%d voxels of the same synthetic workload, %.1f s, single thread"
                                        % (pn, pdt)}
        if subs:
            line["sub"] = subs
            if world == 1:
                line["extra_kernels"] = extra_kernels(torch)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
