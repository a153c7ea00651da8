"""GPU: one process, several GPUs inside the host library (libfabbercore_b200.so). Voxelwise VB deals one
contiguous voxel range to each device (voxels are independent, inference_vb.cc:423-571) - upload, kernels,
SaveResults and download per device - and the result must be BIT-identical to the one-device run. On a
one-GPU box the ranges are dealt to the same GPU twice (FABBER_B200_DEVICES=0,0), which exercises the whole
range / pitch / gather logic; with more GPUs present real devices are used. Also: a page-locked caller buffer
is uploaded by DMA in place, and the call only returns when the caller may reuse it."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def run_worker(tmp_path, name, case, devices, pinned=False, extra_env=None):
    out = str(tmp_path / (name + ".npz"))
    env = dict(os.environ)
    env.update({"FABBER_B200_DEVICES": devices, "FABBER_B200_MIN_VOXELS_PER_DEVICE": "1000"})
    env.pop("LOCAL_RANK", None)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(HERE, "multidevice_worker.py"), out, case,
                        "pinned" if pinned else "pageable"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.load(out)


def device_list(n):
    import torch

    have = torch.cuda.device_count()
    return ",".join(str(i % have) for i in range(n))


@pytest.mark.parametrize("case", ["poly_image_lm", "ar1", "ar2", "nlls"])
def test_ranges_over_devices_equal_the_one_device_run(tmp_path, case):
    one = run_worker(tmp_path, "one", case, "0")
    three = run_worker(tmp_path, "three", case, device_list(3))
    assert int(one["n_devices"][0]) == 1 and int(three["n_devices"][0]) == 3
    for k in one.files:
        if k != "n_devices":
            assert np.array_equal(one[k], three[k], equal_nan=True), k


def test_pinned_caller_buffer_is_uploaded_in_place(tmp_path):
    one = run_worker(tmp_path, "one", "poly_image_lm", "0")
    direct = run_worker(tmp_path, "direct", "poly_image_lm", device_list(2), pinned=True)
    forced = run_worker(tmp_path, "forced", "poly_image_lm", "0", pinned=True, extra_env={"FABBER_B200_DIRECT_UPLOAD": "1"})
    for got in (direct, forced):
        for k in one.files:
            if k != "n_devices":
                assert np.array_equal(one[k], got[k], equal_nan=True), k   # incl. "data": read back from the device


def test_spatial_vb_through_the_capi_is_cut_into_z_slabs(tmp_path):
    """method=spatialvb with several devices: the host library deals z-slabs (own planes + ghost planes) to the
    devices and runs fabber_cuda_vb_spatial_multi; the outputs equal the one-device run (float32 outputs; the
    aK sums differ in summation order only)."""
    env = {"FABBER_B200_MIN_VOXELS_PER_DEVICE": "500"}
    one = run_worker(tmp_path, "one", "spatial", "0", extra_env=env)
    three = run_worker(tmp_path, "three", "spatial", device_list(3), extra_env=env)
    assert int(one["n_devices"][0]) == 1 and int(three["n_devices"][0]) == 3
    for k in one.files:
        if k == "n_devices":
            continue
        a, b = one[k].astype(np.float64), three[k].astype(np.float64)
        scale = np.maximum(np.abs(a), np.abs(a).max(axis=-1, keepdims=True) * 1e-6)
        assert np.max(np.abs(a - b) / np.maximum(scale, 1e-30)) < 2e-6, k
