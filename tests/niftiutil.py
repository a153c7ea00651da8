"""Independent NIfTI-1 reader / writer for the tests (numpy + gzip only): the command line tool's own I/O
(fabber_core_b200/host/nifti_io.cc) is checked against this, never against itself."""
import gzip
import struct

import numpy as np

_DT = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16}
_CODE = {np.dtype(v): k for k, v in _DT.items()}


def write(path, arr, pixdim=(1.0, 1.0, 1.0, 1.0), sform=None, slope=0.0, inter=0.0, endian="<", intent=0):
    """arr indexed [x, y, z] or [x, y, z, t]; sform: 3x4 voxel->world rows (None = no orientation info)"""
    arr = np.asarray(arr)
    dt = np.dtype(arr.dtype)
    hdr = bytearray(348)
    e = endian
    struct.pack_into(e + "i", hdr, 0, 348)
    shape = list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into(e + "8h", hdr, 40, arr.ndim, *shape)
    struct.pack_into(e + "h", hdr, 68, intent)
    struct.pack_into(e + "h", hdr, 70, _CODE[dt])
    struct.pack_into(e + "h", hdr, 72, dt.itemsize * 8)
    struct.pack_into(e + "8f", hdr, 76, 1.0, pixdim[0], pixdim[1], pixdim[2], pixdim[3], 1.0, 1.0, 1.0)
    struct.pack_into(e + "f", hdr, 108, 352.0)
    struct.pack_into(e + "2f", hdr, 112, slope, inter)
    if sform is not None:
        struct.pack_into(e + "h", hdr, 254, 1)
        struct.pack_into(e + "12f", hdr, 280, *np.asarray(sform, dtype=np.float64).reshape(-1))
    hdr[344:348] = b"n+1\0"
    body = np.asfortranarray(arr).astype(dt.newbyteorder(e)).tobytes(order="F")
    raw = bytes(hdr) + b"\0\0\0\0" + body
    if path.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(raw)
    else:
        with open(path, "wb") as f:
            f.write(raw)


def read(path):
    """-> (array [x, y, z, t] float64 with scaling applied, header dict)"""
    raw = gzip.open(path, "rb").read() if path.endswith(".gz") else open(path, "rb").read()
    e = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    dim = struct.unpack(e + "8h", raw[40:56])
    datatype = struct.unpack(e + "h", raw[70:72])[0]
    pixdim = struct.unpack(e + "8f", raw[76:108])
    vox_offset = int(struct.unpack(e + "f", raw[108:112])[0])
    slope, inter = struct.unpack(e + "2f", raw[112:120])
    cal_max, cal_min = struct.unpack(e + "2f", raw[124:132])
    shape = tuple(int(d) for d in dim[1:dim[0] + 1])
    dt = np.dtype(_DT[datatype]).newbyteorder(e)
    arr = np.frombuffer(raw, dtype=dt, count=int(np.prod(shape)), offset=vox_offset).reshape(shape, order="F")
    arr = arr.astype(np.float64)
    if slope != 0.0 and np.isfinite(slope):
        arr = arr * slope + inter
    while arr.ndim < 4:
        arr = arr[..., None]
    hdr = {"dim": dim, "datatype": datatype, "pixdim": pixdim, "intent_code": struct.unpack(e + "h", raw[68:70])[0],
           "sform_code": struct.unpack(e + "h", raw[254:256])[0], "srow": struct.unpack(e + "12f", raw[280:328]),
           "cal_max": cal_max, "cal_min": cal_min, "magic": raw[344:348]}
    return arr, hdr


def series_to_volume(series, shape):
    """[T][N] voxel series (x fastest) -> [x, y, z, t]"""
    nx, ny, nz = shape
    t = series.shape[0]
    return np.ascontiguousarray(series.reshape(t, nz, ny, nx).transpose(3, 2, 1, 0))


def volume_to_series(vol):
    """[x, y, z, t] -> [T][N]"""
    nx, ny, nz, t = vol.shape
    return np.ascontiguousarray(vol.transpose(3, 2, 1, 0).reshape(t, nx * ny * nz))
