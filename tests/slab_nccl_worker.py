"""torchrun worker for test_gpu_spatial_slab.py::test_nccl_slab_run_equals_one_gpu_run: every rank runs its
z-slab through TorchDistComm (NCCL); rank 0 also runs the whole volume on its own GPU and compares."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fabber_core_b200 import cuda_abi as abi  # noqa: E402
from fabber_core_b200 import device, synth  # noqa: E402
from fabber_core_b200.spatial_mgpu import SlabPlan, TorchDistComm, run_slab  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    nx, ny, nz = 12, 10, 16
    n = nx * ny * nz
    y = synth.biexp_volume(n, 96, 0.02, 0.02, seed=91, smooth_shape=(nx, ny, nz)).numpy()
    its = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    mk = lambda: abi.ProblemSpec("exp", 96, num_exps=2, dt=0.02, prior_types=list("MMMM"), max_iterations=its,
                                 need_f=True, param_overrides={"r2": {"mean": 6.0}})
    plan = SlabPlan(nx, ny, nz, rank, world)
    g0, g1 = plan.global_columns()
    spec = mk()
    res = run_slab(spec, np.ascontiguousarray(y[:, g0:g1]), plan, TorchDistComm(rank, world, spec.P))
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: res[k] for k in ("mean", "cov", "noise", "free_energy", "status", "spatial_ak")})
    report = {}
    if rank == 0:
        whole = {k: np.concatenate([g[k] for g in gathered], axis=-1) for k in ("mean", "cov", "noise", "free_energy", "status")}
        spec1 = mk()
        spec1.prob.nx, spec1.prob.ny, spec1.prob.nz = nx, ny, nz
        idx = np.arange(n)
        coords = np.stack([idx % nx, (idx // nx) % ny, idx // (nx * ny)]).astype(np.int32)
        one = device.run(spec1, y, spatial=True, coords=coords)
        P = spec1.P
        std = np.sqrt(np.stack([one["cov"][i * (i + 1) // 2 + i] for i in range(P)]))
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
        report = {
            "world": world, "iterations": its,
            "mean_err_in_std": float(np.max(np.abs(whole["mean"] - one["mean"]) / std)),
            "noise_rel": rel(whole["noise"], one["noise"]),
            "f_rel": rel(whole["free_energy"], one["free_energy"]),
            "ak_rel": rel(gathered[0]["spatial_ak"], one["spatial_ak"]),
            "ak_same_on_all_ranks": all(np.array_equal(g["spatial_ak"], gathered[0]["spatial_ak"]) for g in gathered),
            "bad": int(np.count_nonzero(whole["status"])) + int(np.count_nonzero(one["status"])),
        }
        print("SLAB_NCCL_REPORT " + json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
