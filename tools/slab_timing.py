"""Where the time of a z-slab spatial run goes: the same 256^3 C5 volume over all visible GPUs at 2, 6 and 10
iterations - the slope is the cost of an iteration, the intercept the per-call set-up (neighbour tables, hyper-plane
sort, permutation of the series, links). python tools/slab_timing.py [side]  (needs >= 1 GPU; one process)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device, synth

side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
world = torch.cuda.device_count()
n_total = side ** 3
idx = np.arange(n_total)
coords = np.stack([idx % side, (idx // side) % side, idx // (side * side)]).astype(np.int32)
out = {"gpus": world, "side": side, "runs": {}}
for its in (2, 6, 10):
    spec = abi.ProblemSpec("exp", 96, num_exps=2, dt=0.02, prior_types=list("MMMM"), max_iterations=its,
                           param_overrides={"r2": {"mean": 6.0}})
    spec.prob.nx = spec.prob.ny = spec.prob.nz = side
    run = device.SpatialMultiRun(spec, coords, world, devices=list(range(world)))
    ys = []
    for r in range(world):
        g0, g1 = run.part_range(r)
        with torch.cuda.device(r):
            y = synth.biexp_volume(g1 - g0, 96, 0.02, 0.02, seed=1005 + r, device="cuda:%d" % r,
                                   smooth_shape=(side, side, side), voxel_offset=g0)
            torch.cuda.synchronize()
        ys.append(y)
        run.set_data_device(r, y.data_ptr())
    ms = []
    for k in range(6):
        assert run.launch() == 0, device.last_error()
        ms.append(run.last_ms)
    out["runs"][its] = {"ms": ms[2:], "mean_ms": float(np.mean(ms[2:]))}
    run.close()
    del ys
a, b = out["runs"][2]["mean_ms"], out["runs"][10]["mean_ms"]
out["ms_per_iteration"] = (b - a) / 8
out["setup_ms"] = a - 2 * out["ms_per_iteration"]
print(json.dumps(out))
