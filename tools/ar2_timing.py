"""Device time of the two-echo AR(1) kernel next to the one-echo kernel on the same series (linear model, 4 / 3
columns, 200 samples, 2^21 voxels, 10 iterations): python tools/ar2_timing.py  (needs a GPU)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import device, synth


def timed(spec, y, reps=3):
    run = device.VbRun(spec, y.shape[1])
    run.set_data_device(y.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    assert run.launch(st) == 0
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        assert run.launch(st) == 0
    ev[1].record()
    torch.cuda.synchronize()
    res = run.results()
    run.close()
    assert np.all(res["status"] == 0)
    return ev[0].elapsed_time(ev[1]) / reps, int(res["iterations"].sum())


n = 1 << 21
y = synth.dual_echo_volume(n, 100, seed=5, device="cuda")
design = synth.dual_echo_design(100)
out = {}
ms, its = timed(abi.ProblemSpec("linear", 200, design=design, noise="ar", need_f=True), y)
out["one_echo"] = {"ms": ms, "voxel_iterations_per_s": its / ms * 1e3}
for cross in ("none", "same", "dual"):
    ms, its = timed(abi.ProblemSpec("linear", 200, design=design, noise="ar", num_echoes=2, ar_cross_terms=cross,
                                    need_f=True), y)
    out["two_echoes_" + cross] = {"ms": ms, "voxel_iterations_per_s": its / ms * 1e3}
print(json.dumps(out))
