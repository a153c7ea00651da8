#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own test fixtures and golden outputs.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs (reference repo, read-only):
  test/test_data_small.nii.gz            3x3x2x106 int16 - the regression time-series
  test/test_linear_design.mat            VEST design matrix 106x4
  test/outdata_linear_vb/*.nii.gz        golden outputs of `--model=linear --noise=white --method=vb`
  test/outdata_linear_spatialvb/*.nii.gz same with --method=spatialvb (only 'N' priors)
  test/outdata_linear_nlls/*.nii.gz      same model with --method=nlls
  test/outdata_poly/*.nii.gz             golden outputs of `--model=poly --degree=2`
The goldens live on the 64x64x42 grid of the (missing) test_data.nii.gz; test_data_small is the
crop [30:33, 30:33, 20:22] (0-based x,y,z) of it, so the 18 golden voxels
at those indices pin the 18 voxels of test_data_small.
Reference assertion tolerance: test/test_commandline.cc:10 ALLOWED_DELTA 0.001 absolute.
"""
import gzip
import os
import struct
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

_DT = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 512: np.uint16}


def read_nifti(path):
    """Minimal NIfTI-1 reader -> array indexed [x, y, z, t] (Fortran order on disk)."""
    raw = gzip.open(path, "rb").read()
    hdr = raw[:348]
    endian = "<" if struct.unpack("<i", hdr[:4])[0] == 348 else ">"
    dim = struct.unpack(endian + "8h", hdr[40:56])
    datatype = struct.unpack(endian + "h", hdr[70:72])[0]
    vox_offset = int(struct.unpack(endian + "f", hdr[108:112])[0])
    slope, inter = struct.unpack(endian + "2f", hdr[112:120])
    nd = dim[0]
    shape = tuple(int(d) for d in dim[1 : nd + 1])
    dt = np.dtype(_DT[datatype]).newbyteorder(endian)
    n = int(np.prod(shape))
    arr = np.frombuffer(raw, dtype=dt, count=n, offset=vox_offset).reshape(shape, order="F")
    arr = arr.astype(np.float64)
    if slope not in (0.0,) and np.isfinite(slope):
        arr = arr * slope + inter
    while arr.ndim < 4:
        arr = arr[..., None]
    return arr


def read_vest(path):
    rows = []
    in_matrix = False
    for line in open(path):
        s = line.strip()
        if s.startswith("/Matrix"):
            in_matrix = True
            continue
        if not in_matrix or not s:
            continue
        rows.append([float(x) for x in s.split()])
    return np.array(rows)


def crop(vol):
    """golden 64x64x42xK -> [K][18] in x-fastest voxel order of the 3x3x2 crop"""
    sub = vol[30:33, 30:33, 20:22, :]  # [x,y,z,k]
    k = sub.shape[3]
    # voxel order: x fastest, then y, then z  (rundata_array.cc:45-62)
    return np.stack([sub[:, :, :, i].reshape(-1, order="F") for i in range(k)], axis=0)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not present; goldens are committed, nothing to do")
    t = os.path.join(REF, "test")
    data = read_nifti(os.path.join(t, "test_data_small.nii.gz"))  # [3,3,2,106]
    nx, ny, nz, nt = data.shape
    assert (nx, ny, nz, nt) == (3, 3, 2, 106)
    # C-API layout [t][z][y][x] == [T][N] with x fastest
    series = np.stack([data[:, :, :, i].reshape(-1, order="F") for i in range(nt)], axis=0).astype(np.float32)
    design = read_vest(os.path.join(t, "test_linear_design.mat"))
    assert design.shape == (106, 4)
    out = dict(data=series, design=design, shape=np.array([nx, ny, nz, nt]))

    def grab(dirname, names, prefix):
        for n in names:
            p = os.path.join(t, dirname, n + ".nii.gz")
            if os.path.exists(p):
                out[prefix + n] = crop(read_nifti(p)).astype(np.float32)

    lin_names = ["mean_Parameter_%d" % i for i in range(1, 5)] + ["zstat_Parameter_%d" % i for i in range(1, 5)]
    lin_names += ["std_Parameter_%d" % i for i in range(1, 5)] + ["finalMVN", "noise_means", "noise_stdevs", "freeEnergy"]
    grab("outdata_linear_vb", lin_names, "linear_vb/")
    grab("outdata_linear_spatialvb", lin_names, "linear_spatialvb/")
    grab("outdata_linear_nlls", lin_names, "linear_nlls/")  # --method=nlls (Levenberg, the default)
    poly_names = []
    for c in range(3):
        poly_names += ["mean_c%d" % c, "std_c%d" % c, "zstat_c%d" % c]
    poly_names += ["finalMVN", "noise_means", "noise_stdevs", "freeEnergy"]
    grab("outdata_poly", poly_names, "poly/")
    np.savez_compressed(os.path.join(OUT, "c1_regression.npz"), **out)
    for k in sorted(out):
        print(k, out[k].shape, out[k].dtype)


if __name__ == "__main__":
    main()
