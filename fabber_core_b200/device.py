"""Host-side driver of the inner C ABI (include/fabber_cuda.h) through ctypes.

This is plumbing only: it owns device buffers, copies inputs/outputs and calls the hand-written
CUDA library. There is no CPU fallback - if the library cannot be loaded or no GPU is present the
calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import cuda_abi as abi

_LIB = None


class CudaError(RuntimeError):
    pass


def lib():
    """Load libfabber_cuda.so (built in-tree by fabber_core_b200/csrc/Makefile). Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = abi.library_path()
    if not os.path.exists(path):
        raise CudaError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                        % path)
    L = C.CDLL(path)
    L.fabber_cuda_last_error.restype = C.c_char_p
    L.fabber_cuda_malloc.restype = C.c_void_p
    L.fabber_cuda_malloc.argtypes = [C.c_ulonglong]
    L.fabber_cuda_free.argtypes = [C.c_void_p]
    L.fabber_cuda_host_alloc.restype = C.c_void_p
    L.fabber_cuda_host_alloc.argtypes = [C.c_ulonglong]
    L.fabber_cuda_host_free.argtypes = [C.c_void_p]
    for fn in (L.fabber_cuda_memcpy_h2d, L.fabber_cuda_memcpy_d2h):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_void_p]
        fn.restype = C.c_int
    L.fabber_cuda_memset.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_void_p]
    L.fabber_cuda_stream_sync.argtypes = [C.c_void_p]
    L.fabber_cuda_vb_voxelwise.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.fabber_cuda_vb_voxelwise.restype = C.c_int
    L.fabber_cuda_vb_spatial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.fabber_cuda_vb_spatial.restype = C.c_int
    L.fabber_cuda_check_status.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fabber_cuda_check_status.restype = C.c_int
    L.fabber_cuda_model_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fabber_cuda_model_fit.restype = C.c_int
    L.fabber_cuda_measure_fp64_peak.restype = C.c_double
    L.fabber_cuda_measure_fp64_peak.argtypes = [C.c_int]
    L.fabber_cuda_launch_count.restype = C.c_ulonglong
    L.fabber_cuda_gather_voxels.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p]
    L.fabber_cuda_scatter_voxels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_ulonglong,
                                             C.c_void_p, C.c_void_p]
    if L.fabber_cuda_sizeof_problem() != C.sizeof(abi.VbProblem):
        raise CudaError("ctypes mirror of fabber_cuda_vb_problem is out of date")
    if L.fabber_cuda_sizeof_buffers() != C.sizeof(abi.VbBuffers):
        raise CudaError("ctypes mirror of fabber_cuda_vb_buffers is out of date")
    _LIB = L
    return L


def last_error():
    return lib().fabber_cuda_last_error().decode(errors="replace")


def check(rc, what):
    if rc != abi.OK:
        raise CudaError("%s failed (%d): %s" % (what, rc, last_error()))


class DeviceArray(object):
    """A device allocation with a numpy-like shape/dtype; freed on close() or garbage collection."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = lib().fabber_cuda_malloc(self.nbytes)
        if not self.ptr:
            raise CudaError("device allocation of %d bytes failed: %s" % (self.nbytes, last_error()))

    @classmethod
    def from_host(cls, arr, dtype=None, stream=None):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        d = cls(arr.shape, arr.dtype)
        check(lib().fabber_cuda_memcpy_h2d(d.ptr, arr.ctypes.data, d.nbytes, stream), "h2d copy")
        check(lib().fabber_cuda_stream_sync(stream), "sync")
        return d

    def to_host(self, stream=None):
        out = np.empty(self.shape, dtype=self.dtype)
        check(lib().fabber_cuda_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream), "d2h copy")
        check(lib().fabber_cuda_stream_sync(stream), "sync")
        return out

    def close(self):
        if self.ptr:
            lib().fabber_cuda_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VbRun(object):
    """Device-resident buffers of one VB run (inputs + outputs) and the launch call."""

    def __init__(self, spec, n_voxels, spatial=False):
        self.spec = spec
        self.N = int(n_voxels)
        self.spatial = spatial
        P, NN, N = spec.P, spec.NN, self.N
        self.buf = abi.VbBuffers()
        self.out = {
            "mean": DeviceArray((P, N), np.float64),
            "cov": DeviceArray((spec.ncov, N), np.float64),
            "noise": DeviceArray((NN, N), np.float64),
            "free_energy": DeviceArray((N,), np.float64),
            "iterations": DeviceArray((N,), np.int32),
            "status": DeviceArray((N,), np.int32),
        }
        if spec.prob.f_history_len > 0:
            self.out["f_history"] = DeviceArray((spec.prob.f_history_len, N), np.float64)
            self.buf.f_history = self.out["f_history"].ptr
        for k in ("mean", "cov", "noise", "free_energy", "iterations", "status"):
            setattr(self.buf, k, self.out[k].ptr)
        self.inputs = {}
        self.ak = None
        if spatial:
            self.ak = np.zeros((spec.prob.max_iterations + 1, P))
            self.buf.spatial_ak = self.ak.ctypes.data

    def set_data(self, data, stream=None):
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.shape == (self.spec.n_times, self.N)
        self.inputs["data"] = DeviceArray.from_host(data, stream=stream)
        self.buf.data = self.inputs["data"].ptr

    def set_data_device(self, ptr):
        self.buf.data = ptr

    def set_image_prior(self, k, img):
        d = DeviceArray.from_host(img, dtype=np.float64)
        self.inputs["image%d" % k] = d
        self.buf.image_prior[k] = d.ptr

    def set_initial(self, mean=None, cov=None, noise=None, lock_centre=None):
        for name, arr in (("init_mean", mean), ("init_cov", cov), ("init_noise", noise), ("lock_centre", lock_centre)):
            if arr is not None:
                d = DeviceArray.from_host(arr, dtype=np.float64)
                self.inputs[name] = d
                setattr(self.buf, name, d.ptr)

    def set_coords(self, coords):
        d = DeviceArray.from_host(coords, dtype=np.int32)
        self.inputs["coords"] = d
        self.buf.coords = d.ptr

    def launch(self, stream=None):
        prob = self.spec.prob
        prob.n_voxels = self.N
        fn = lib().fabber_cuda_vb_spatial if self.spatial else lib().fabber_cuda_vb_voxelwise
        return fn(C.byref(prob), C.byref(self.buf), stream)

    def launch_range(self, v_begin, v_end, stream=None):
        """voxelwise VB on the voxels [v_begin, v_end) only (fabber_cuda_vb_voxelwise_range)"""
        prob = self.spec.prob
        prob.n_voxels = self.N
        fn = lib().fabber_cuda_vb_voxelwise_range
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        return fn(C.byref(prob), C.byref(self.buf), int(v_begin), int(v_end), stream)

    def sync(self, stream=None):
        check(lib().fabber_cuda_stream_sync(stream), "stream sync")

    def results(self, stream=None):
        out = {k: v.to_host(stream) for k, v in self.out.items()}
        if self.ak is not None:
            out["spatial_ak"] = self.ak.copy()
        return out

    def close(self):
        for d in list(self.out.values()) + list(self.inputs.values()):
            d.close()


def run(spec, data, spatial=False, image_priors=None, coords=None, init_mean=None, init_cov=None,
        init_noise=None, lock_centre=None):
    """Convenience: host arrays in, host arrays out (same signature as the test oracle's run())."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    r = VbRun(spec, data.shape[1], spatial=spatial)
    try:
        r.set_data(data)
        for k, img in (image_priors or {}).items():
            r.set_image_prior(k, img)
        r.set_initial(init_mean, init_cov, init_noise, lock_centre)
        if coords is not None:
            r.set_coords(coords)
        rc = r.launch()
        if rc not in (abi.OK, abi.ERR_BAD_VOXEL):
            raise CudaError("VB launch failed (%d): %s" % (rc, last_error()))
        r.sync()
        out = r.results()
        if rc == abi.OK and not spec.prob.allow_bad_voxels:
            nbad = int(np.count_nonzero(out["status"]))
            if nbad:
                rc = abi.ERR_BAD_VOXEL
        out["rc"] = rc
        out["n_times"] = int(data.shape[0])
        return out
    finally:
        r.close()


def z_slab_parts(coords, nz, n_parts):
    """Cut a z-major voxel list (coords int [3][N], global coordinates) into `n_parts` z-slabs:
    -> list of (v0, v1, own0, own1, z0, z1): part r owns the voxels with z in [z0, z1) = list positions
    [own0, own1) and also holds the ghost planes z0-1 and z1 = positions [v0, v1)."""
    from . import shard

    z = np.asarray(coords[2])
    assert np.all(np.diff(z) >= 0), "voxel list must be z-major (x fastest, then y, then z)"
    first = np.searchsorted(z, np.arange(nz + 2))  # first list position with z >= k
    parts = []
    for r in range(n_parts):
        z0, z1 = shard.z_slab_range(nz, r, n_parts)
        if z1 <= z0:
            raise ValueError("more parts than z-planes")
        own0, own1 = int(first[z0]), int(first[z1])
        v0 = int(first[max(z0 - 1, 0)])
        v1 = int(first[min(z1 + 1, nz)])
        if own1 <= own0:
            raise ValueError("a z-slab holds no voxel of the mask")
        parts.append((v0, v1, own0, own1, z0, z1))
    return parts


class SpatialMultiRun(object):
    """Device-resident buffers of one spatial VB run cut into z-slabs over several GPUs, driven by THIS process
    (fabber_cuda_vb_spatial_multi). `devices`: one ordinal per part (default: round-robin over the visible GPUs;
    repeats put several slabs on one GPU)."""

    def __init__(self, spec, coords, n_parts, devices=None):
        L = lib()
        self.spec = spec
        self.coords = np.ascontiguousarray(coords, dtype=np.int32)
        self.N = self.coords.shape[1]
        spec.prob.n_voxels = self.N
        n_dev = L.fabber_cuda_device_count()
        self.devices = list(devices) if devices is not None else [r % n_dev for r in range(n_parts)]
        self.cuts = z_slab_parts(self.coords, spec.prob.nz, n_parts)
        self.parts = (abi.SlabPart * n_parts)()
        self.keep, self.outs = [], []
        self.ak = np.zeros((spec.prob.max_iterations + 1, spec.P))
        self.prev = L.fabber_cuda_get_device()
        P, NN = spec.P, spec.NN
        for r, (v0, v1, own0, own1, z0, z1) in enumerate(self.cuts):
            check(L.fabber_cuda_set_device(self.devices[r]), "set_device")
            pt = self.parts[r]
            pt.device, pt.v0, pt.v1, pt.own0, pt.own1, pt.own_z0, pt.own_z1 = self.devices[r], v0, v1, own0, own1, z0, z1
            n = v1 - v0
            pt.buf.coords = self._dev(self.coords[:, v0:v1], np.int32)
            o = {"mean": DeviceArray((P, n), np.float64), "cov": DeviceArray((spec.ncov, n), np.float64),
                 "noise": DeviceArray((NN, n), np.float64), "free_energy": DeviceArray((n,), np.float64),
                 "iterations": DeviceArray((n,), np.int32), "status": DeviceArray((n,), np.int32)}
            for k2, v in o.items():
                setattr(pt.buf, k2, v.ptr)
            if r == 0:
                pt.buf.spatial_ak = self.ak.ctypes.data
            self.outs.append(o)
        L.fabber_cuda_set_device(max(self.prev, 0))

    def _dev(self, arr, dtype):
        d = DeviceArray.from_host(np.ascontiguousarray(arr, dtype=dtype))
        self.keep.append(d)
        return d.ptr

    def part_range(self, r):
        """(v0, v1): the columns of the whole voxel list part r holds (own + ghost planes)"""
        return self.cuts[r][0], self.cuts[r][1]

    def set_data(self, data):
        data = np.ascontiguousarray(data, dtype=np.float32)
        for r, cut in enumerate(self.cuts):
            check(lib().fabber_cuda_set_device(self.devices[r]), "set_device")
            self.parts[r].buf.data = self._dev(data[:, cut[0]:cut[1]], np.float32)
        lib().fabber_cuda_set_device(max(self.prev, 0))

    def set_data_device(self, r, ptr):
        """device pointer ON part r's device to its [T][v1 - v0] series"""
        self.parts[r].buf.data = ptr

    def set_inputs(self, image_priors=None, init_mean=None, init_cov=None, init_noise=None, lock_centre=None):
        for r, cut in enumerate(self.cuts):
            v0, v1 = cut[0], cut[1]
            check(lib().fabber_cuda_set_device(self.devices[r]), "set_device")
            buf = self.parts[r].buf
            for k, img in (image_priors or {}).items():
                buf.image_prior[k] = self._dev(np.asarray(img)[v0:v1], np.float64)
            for name, arr in (("init_mean", init_mean), ("init_cov", init_cov), ("init_noise", init_noise),
                              ("lock_centre", lock_centre)):
                if arr is not None:
                    setattr(buf, name, self._dev(np.asarray(arr)[:, v0:v1], np.float64))
        lib().fabber_cuda_set_device(max(self.prev, 0))

    def launch(self):
        """synchronous; returns the library's return code. self.last_ms: device-side duration (max over devices)"""
        L = lib()
        fn = L.fabber_cuda_vb_spatial_multi
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        fn.restype = C.c_int
        rc = fn(C.byref(self.spec.prob), len(self.cuts), C.byref(self.parts))
        L.fabber_cuda_last_multi_ms.restype = C.c_double
        self.last_ms = float(L.fabber_cuda_last_multi_ms())
        return rc

    def results(self):
        spec, N = self.spec, self.N
        full = {"mean": np.zeros((spec.P, N)), "cov": np.zeros((spec.ncov, N)), "noise": np.zeros((spec.NN, N)),
                "free_energy": np.zeros(N), "iterations": np.zeros(N, dtype=np.int32),
                "status": np.zeros(N, dtype=np.int32)}
        for r, (v0, v1, own0, own1, z0, z1) in enumerate(self.cuts):
            check(lib().fabber_cuda_set_device(self.devices[r]), "set_device")
            for k2, v in self.outs[r].items():
                full[k2][..., own0:own1] = v.to_host()[..., own0 - v0:own1 - v0]
        lib().fabber_cuda_set_device(max(self.prev, 0))
        full["spatial_ak"] = self.ak.copy()
        return full

    def close(self):
        for r in range(len(self.outs)):
            lib().fabber_cuda_set_device(self.devices[r])
            for v in self.outs[r].values():
                v.close()
        for d in self.keep:
            d.close()
        self.outs, self.keep = [], []
        lib().fabber_cuda_set_device(max(self.prev, 0))


def run_spatial_multi(spec, data, coords, n_parts, devices=None, image_priors=None, init_mean=None, init_cov=None,
                      init_noise=None, lock_centre=None):
    """Spatial VB of one volume over `n_parts` z-slabs in ONE process (fabber_cuda_vb_spatial_multi): host arrays in,
    host arrays out in the caller's voxel order - same signature and result layout as run(spatial=True)."""
    r = SpatialMultiRun(spec, coords, n_parts, devices)
    try:
        r.set_data(data)
        r.set_inputs(image_priors, init_mean, init_cov, init_noise, lock_centre)
        rc = r.launch()
        if rc not in (abi.OK, abi.ERR_BAD_VOXEL):
            raise CudaError("spatial multi-device run failed (%d): %s" % (rc, last_error()))
        out = r.results()
        if rc == abi.OK and not spec.prob.allow_bad_voxels and np.count_nonzero(out["status"]):
            rc = abi.ERR_BAD_VOXEL
        out["rc"] = rc
        out["n_times"] = int(np.asarray(data).shape[0])
        return out
    finally:
        r.close()
