"""Test helper: the reference's own sources compiled against the test-only NEWMAT shim
(oracle/_ref/libfabbercore_ref.so, built by oracle/Makefile where /root/reference exists), driven through
the reference's own C API with the same ctypes wrapper the product library uses."""
import os

import numpy as np

from fabber_core_b200 import fabber as fab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libfabbercore_ref.so")
# the same sources plus inference_nlls.cc, built against the stand-in for FSL's MISCMATHS::nonlin (oracle/_shim)
REF_NLLS_LIB = os.path.join(ROOT, "oracle", "_ref", "libfabbercore_ref_nlls.so")


def available():
    return os.path.exists(REF_LIB)


def nlls_available():
    return os.path.exists(REF_NLLS_LIB)


class ReferenceFabber(fab.Fabber):
    def __init__(self, lib=None):
        fab.Fabber.__init__(self, lib=lib or REF_LIB)

    def _new_handle(self):
        fab.Fabber._new_handle(self)
        # fabber_destroy tears the model factory down (fabber_capi.cc:279): re-register the example model
        self.clib.fabber_ref_register_exp()

    def doubles(self, name, n_voxels, max_rows=512):
        """The reference's result matrix `name` as float64 [rows][n_voxels] (masked voxels only), through
        the test-only accessor in oracle/_shim/ref_extra.cc - the public C API narrows to float32."""
        import ctypes as C

        buf = np.zeros(max_rows * n_voxels)
        fn = self.clib.fabber_ref_get_data_double
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
        rows = fn(self.handle, name.encode(), buf.ctypes.data, buf.size)
        if rows < 0:
            raise RuntimeError("no such data in the reference run: %s" % name)
        return buf[:rows * n_voxels].reshape(rows, n_voxels).copy()


def volume(series, shape):
    """[T][N] (x fastest) -> [x, y, z, t]"""
    nx, ny, nz = shape
    return np.ascontiguousarray(np.asarray(series).T.reshape(nz, ny, nx, -1).transpose(2, 1, 0, 3))


def flat(img):
    """[x, y, z(, k)] -> [k][N] in voxel order"""
    if img.ndim == 3:
        img = img[..., None]
    return np.stack([img[..., k].reshape(-1, order="F") for k in range(img.shape[3])])
