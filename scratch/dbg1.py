import sys; sys.path.insert(0,".")
import numpy as np, oracle
sys.path.insert(0,'.')
from fabber_core_b200 import cuda_abi as abi, synth, device
y = synth.biexp_volume(300, 96, 0.02, 0.02, seed=1003).numpy()
kw = dict(num_exps=2, dt=0.02, convergence="maxits", max_iterations=10, need_f=True, allow_bad_voxels=True, f_history_len=12)
ref = oracle.run(abi.ProblemSpec("exp", 96, **kw), y)
gpu = device.run(abi.ProblemSpec("exp", 96, **kw), y)
print("ref status", np.unique(ref['status'], return_counts=True))
print("gpu status", np.unique(gpu['status'], return_counts=True))
bad = np.where(gpu['status']!=0)[0][:3]
np.set_printoptions(linewidth=200, precision=10)
for b in bad:
    print("voxel", b, "its", gpu['iterations'][b], ref['iterations'][b])
    print(" gpu F", gpu['f_history'][:, b])
    print(" ref F", ref['f_history'][:, b])
    print(" gpu m", gpu['mean'][:, b], "noise", gpu['noise'][:, b])
    print(" ref m", ref['mean'][:, b], "noise", ref['noise'][:, b])
    print(" gpu cov", gpu['cov'][:, b])
    print(" ref cov", ref['cov'][:, b])
for mi in (1,2,3):
    kw['max_iterations']=mi
    ref = oracle.run(abi.ProblemSpec("exp", 96, **kw), y)
    gpu = device.run(abi.ProblemSpec("exp", 96, **kw), y)
    b = bad[0]
    print("maxits", mi, "gpu m", gpu['mean'][:, b], gpu['noise'][:, b], gpu['free_energy'][b], gpu['status'][b])
    print("maxits", mi, "ref m", ref['mean'][:, b], ref['noise'][:, b], ref['free_energy'][b], ref['status'][b])
    print(" gpu cov", gpu['cov'][:, b])
    print(" ref cov", ref['cov'][:, b])
