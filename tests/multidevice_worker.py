"""Worker of tests/test_gpu_multidevice.py: one C-API run in a fresh process (the library reads
FABBER_B200_DEVICES once per process), results saved to an .npz."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fabber_core_b200 import device, synth  # noqa: E402
from fabber_core_b200 import fabber as fab  # noqa: E402


def main():
    out, case, pinned = sys.argv[1], sys.argv[2], sys.argv[3] == "pinned"
    n, T = 6000, 40
    shape = None
    if case == "spatial":
        # spatial VB through the C API: the library cuts the volume into z-slabs, one per device
        shape = (16, 15, 12)
        n = shape[0] * shape[1] * shape[2]
        y = synth.poly_volume(n, T, 2, seed=79).numpy()
        opts = {"model": "poly", "degree": 2, "noise": "white", "method": "spatialvb", "param-spatial-priors": "MMM",
                "max-iterations": 5, "save-mean": True, "save-std": True, "save-mvn": True, "save-noise-mean": True,
                "save-free-energy": True}
        extra = {}
    elif case == "poly_image_lm":
        y = synth.poly_volume(n, T, 2, seed=77).numpy()
        opts = {"model": "poly", "degree": 2, "noise": "white", "method": "vb", "convergence": "lm",
                "PSP_byname1": "c2", "PSP_byname1_type": "I", "PSP_byname1_image": "c2img", "PSP_byname1_prec": 1e4,
                "save-mean": True, "save-std": True, "save-mvn": True, "save-free-energy": True,
                "save-noise-mean": True, "save-model-fit": True, "save-residuals": True,
                "save-free-energy-history": True}
        extra = {"c2img": np.linspace(-0.05, 0.05, n).astype(np.float32)}
    elif case == "ar1":
        y = synth.linear_ar_volume(n, T, 0.3, seed=78).numpy()
        design = os.path.join(os.path.dirname(out), "design.mat")
        np.savetxt(design, synth.ar_design(T), fmt="%.17g")
        opts = {"model": "linear", "basis": design, "noise": "ar", "method": "vb", "save-mean": True, "save-mvn": True,
                "save-noise-mean": True}
        extra = {}
    elif case == "nlls":
        # --method=nlls rides the same voxel-range plumbing (no noise outputs: NLLS has no noise parameters)
        y = synth.poly_volume(n, T, 2, seed=80).numpy()
        opts = {"model": "poly", "degree": 2, "method": "nlls", "save-mean": True, "save-mvn": True, "save-std": True}
        extra = {}
    elif case == "ar2":
        y = synth.dual_echo_volume(n, T // 2, seed=81).numpy()
        design = os.path.join(os.path.dirname(out), "design2.mat")
        np.savetxt(design, synth.dual_echo_design(T // 2), fmt="%.17g")
        opts = {"model": "linear", "basis": design, "noise": "ar", "num-echoes": 2, "ar1-cross-terms": "dual",
                "method": "vb", "save-mean": True, "save-mvn": True, "save-noise-mean": True}
        extra = {}
    else:
        raise SystemExit("unknown case")
    f = fab.Fabber()
    f._set_options(opts)
    mask = np.ones(n, dtype=np.int32)
    ext = shape or (n, 1, 1)
    f._trycall(f.clib.fabber_set_extent, f.handle, ext[0], ext[1], ext[2], mask, f.errbuf)
    for k, v in extra.items():
        f._trycall(f.clib.fabber_set_data, f.handle, k.encode(), 1, np.ascontiguousarray(v), f.errbuf)
    flat = np.ascontiguousarray(y.reshape(-1))
    if pinned:
        # a page-locked caller buffer: the library must DMA from it in place (no staging copy)
        L = device.lib()
        ptr = L.fabber_cuda_host_alloc(flat.nbytes)
        buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(flat.size,))
        buf[:] = flat
        flat = buf
    f._trycall(f.clib.fabber_set_data, f.handle, b"data", T, flat, f.errbuf)
    if pinned:
        flat[:] = -1.0  # the call has returned: the caller may scribble on its buffer
    noop = f.progress_cb_type(0)
    f._trycall(f.clib.fabber_dorun, f.handle, len(f.outbuf), f.outbuf, f.errbuf, noop)
    log = f.outbuf.value.decode(errors="replace")
    res = {}
    f._trycall(f.clib.fabber_get_model_params, f.handle, len(f.outbuf), f.outbuf, f.errbuf)
    params = f.outbuf.value.decode().splitlines()
    keys = ["mean_" + p for p in params] + ["finalMVN"] + ([] if case == "nlls" else ["noise_means"])
    if case == "spatial":
        keys += ["std_" + p for p in params] + ["freeEnergy"]
    if case == "poly_image_lm":
        keys += ["std_" + p for p in params] + ["freeEnergy", "modelfit", "residuals", "freeEnergyHistory", "data"]
    for key in keys:
        size = f._trycall(f.clib.fabber_get_data_size, f.handle, key.encode(), f.errbuf)
        b = np.empty(n * size, dtype=np.float32)
        f._trycall(f.clib.fabber_get_data, f.handle, key.encode(), b, f.errbuf)
        res[key] = b.reshape(size, n)
    devices_line = [l for l in log.splitlines() if "calculations on the GPU" in l]
    res["n_devices"] = np.array([int(devices_line[0].split(",")[-1].split()[0])])
    np.savez(out, **res)


if __name__ == "__main__":
    main()
