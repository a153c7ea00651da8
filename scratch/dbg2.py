import sys; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np, oracle
from fabber_core_b200 import cuda_abi as abi, synth, device
from test_gpu_spatial import grid_coords, C5
np.set_printoptions(linewidth=220, precision=8)
nx, ny, nz = 12, 10, 6
y = synth.biexp_volume(nx*ny*nz, 96, 0.02, 0.02, seed=1005, smooth_shape=(nx,ny,nz)).numpy()
coords = grid_coords(nx, ny, nz)
for types in ("pppp",):
  for mi in (1,2):
    def mk():
        kw = dict(C5); kw.pop("model")
        sp = abi.ProblemSpec("exp", 96, prior_types=list(types), need_f=True, max_iterations=mi, allow_bad_voxels=True, **kw)
        sp.prob.nx, sp.prob.ny, sp.prob.nz = nx, ny, nz
        return sp
    ref = oracle.run(mk(), y, spatial=True, coords=coords)
    gpu = device.run(mk(), y, spatial=True, coords=coords)
    print(types, mi, "status gpu", np.unique(gpu["status"], return_counts=True), "ref", np.unique(ref["status"], return_counts=True))
    for v in (0, 5, 100):
        print(" v", v, "gpu m", gpu["mean"][:, v], "noise", gpu["noise"][:, v], "F", gpu["free_energy"][v])
        print(" v", v, "ref m", ref["mean"][:, v], "noise", ref["noise"][:, v], "F", ref["free_energy"][v])
        print("   gpu var", gpu["cov"][[0,2,5,9], v], "ref var", ref["cov"][[0,2,5,9], v])
    print(" ak gpu", gpu["spatial_ak"], "\n ak ref", ref["spatial_ak"])
