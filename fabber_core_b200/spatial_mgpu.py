"""Spatial VB over several GPUs: z-slab partition, halo exchange of posterior means, all-reduce of the aK sums.

One process per GPU. Rank r owns the z-planes [z0, z1) of a full nx x ny x nz grid (voxel order is x
fastest, then y, then z, so a slab is a contiguous voxel range) and additionally holds the GHOST planes
z0-1 and z1 of its neighbours. Per iteration the CUDA library (fabber_cuda_vb_spatial_slab,
include/fabber_cuda.h) calls back into this module twice:
  allreduce_sum  the two global sums of SpatialPrior::CalculateaK per parameter (priors.cc:233-343)
  exchange       own boundary planes' means out, ghost planes' means in
and, when a prior couples neighbouring voxels' means (M / m), once per step of the pipelined sweep:
  forward        the freshly swept top-plane means of one block of hyper-planes up to rank r+1
All three are implemented with torch.distributed (NCCL over NVLink on the GPU box; `ThreadComm` emulates the
ranks with threads on one GPU for the single-GPU test-suite).

Exactness: the reference sweeps the voxels sequentially (spatialvb.cc:428-437), so voxel (x, y, z) sees this
iteration's value of (x, y, z-1) and last iteration's value of (x, y, z+1). Both hold across a slab boundary:
the hyper-planes x+y+z = H are cut into blocks, rank r sweeps block s-r at step s and forwards its top plane's
part of that block upwards before rank r+1 starts the same block; the downward halo moves after the sweep.
The multi-GPU result is the one-GPU result (the aK sums differ in summation order only).
"""
import ctypes as C
import threading

import numpy as np

from . import cuda_abi as abi
from . import device, shard

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                          C.c_void_p, C.c_int, C.c_void_p)
FORWARD_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p)


class Slab(C.Structure):
    """mirror of fabber_cuda_slab"""
    _fields_ = [
        ("n_global_voxels", C.c_int),
        ("ghost", C.c_void_p),
        ("send_lo", C.c_void_p), ("send_hi", C.c_void_p), ("recv_lo", C.c_void_p), ("recv_hi", C.c_void_p),
        ("n_send_lo", C.c_int), ("n_send_hi", C.c_int), ("n_recv_lo", C.c_int), ("n_recv_hi", C.c_int),
        ("rank", C.c_int), ("world", C.c_int), ("n_blocks", C.c_int), ("block_planes", C.c_int),
        ("plane_offset", C.c_int),
        ("fwd_send", C.c_void_p), ("fwd_recv", C.c_void_p),
        ("fwd_send_start", C.POINTER(C.c_int)), ("fwd_recv_start", C.POINTER(C.c_int)),
        ("user", C.c_void_p),
        ("allreduce_sum", ALLREDUCE_FN),
        ("exchange", EXCHANGE_FN),
        ("forward", FORWARD_FN),
    ]


class SlabPlan(object):
    """Which planes a rank owns / ghosts, and the index lists of the halo (positions in the local voxel list)."""

    def __init__(self, nx, ny, nz, rank, world, block_planes=None):
        self.nx, self.ny, self.nz, self.rank, self.world = nx, ny, nz, rank, world
        # pipelined sweep: global hyper-planes H = x+y+z in [0, n_hyper) cut into blocks of block_planes.
        # One block of skew per rank boundary: ~8 blocks per rank keeps the pipeline fill near 1/8 of a sweep.
        self.n_hyper = nx + ny + nz - 2
        self.block_planes = int(block_planes) if block_planes else max(1, -(-self.n_hyper // (8 * world)))
        self.n_blocks = -(-self.n_hyper // self.block_planes)
        self.z0, self.z1 = shard.z_slab_range(nz, rank, world)
        if self.z1 - self.z0 < 1:
            raise ValueError("more ranks than z-planes")
        self.ghost_lo = 1 if self.z0 > 0 else 0
        self.ghost_hi = 1 if self.z1 < nz else 0
        self.zlo, self.zhi = self.z0 - self.ghost_lo, self.z1 + self.ghost_hi   # local planes [zlo, zhi)
        self.plane = nx * ny
        self.nz_local = self.zhi - self.zlo
        self.n_local = self.plane * self.nz_local
        self.n_own = self.plane * (self.z1 - self.z0)
        self.n_global = self.plane * nz

    def local_plane(self, z):
        """positions of global plane z in the local voxel list"""
        lz = z - self.zlo
        return np.arange(lz * self.plane, (lz + 1) * self.plane, dtype=np.int32)

    def coords(self):
        idx = np.arange(self.n_local)
        return np.stack([idx % self.nx, (idx // self.nx) % self.ny, idx // self.plane]).astype(np.int32)

    def ghost_mask(self):
        g = np.zeros(self.n_local, dtype=np.uint8)
        if self.ghost_lo:
            g[self.local_plane(self.z0 - 1)] = 1
        if self.ghost_hi:
            g[self.local_plane(self.z1)] = 1
        return g

    def halo_lists(self):
        e = np.zeros(0, dtype=np.int32)
        return (self.local_plane(self.z0) if self.ghost_lo else e,          # send down
                self.local_plane(self.z1 - 1) if self.ghost_hi else e,      # send up
                self.local_plane(self.z0 - 1) if self.ghost_lo else e,      # receive from below
                self.local_plane(self.z1) if self.ghost_hi else e)          # receive from above

    def forward_lists(self):
        """(send, send_start, recv, recv_start): positions of the own top plane / the lower ghost plane grouped
        by block of hyper-planes (block b = [start[b], start[b+1])), ascending position inside a block - the
        sender's and the receiver's lists name the same global voxels in the same order."""
        def grouped(z, present):
            if not present:
                return np.zeros(0, dtype=np.int32), np.zeros(self.n_blocks + 1, dtype=np.int32)
            pos = self.local_plane(z)
            inplane = pos - pos[0]
            H = inplane % self.nx + inplane // self.nx + z
            blk = H // self.block_planes
            order = np.argsort(blk, kind="stable")
            start = np.concatenate([[0], np.cumsum(np.bincount(blk, minlength=self.n_blocks))]).astype(np.int32)
            return pos[order].astype(np.int32), start

        send, send_start = grouped(self.z1 - 1, self.ghost_hi)
        recv, recv_start = grouped(self.z0 - 1, self.ghost_lo)
        return send, send_start, recv, recv_start

    def global_columns(self):
        """voxel range of the WHOLE volume that the local list (own + ghosts) covers"""
        return self.zlo * self.plane, self.zhi * self.plane

    def own_slice(self):
        """slice of the local list that this rank owns"""
        lo = self.ghost_lo * self.plane
        return slice(lo, lo + self.n_own)


class _DevPtr(object):
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _tensor(ptr, n):
    import torch

    return torch.as_tensor(_DevPtr(ptr, n), device="cuda")


class TorchDistComm(object):
    """The real thing: one process per GPU, NCCL through torch.distributed."""

    def __init__(self, rank, world, n_params):
        import torch.distributed as dist

        self.rank, self.world, self.P = rank, world, n_params
        # The aK all-reduce runs on the library's side stream, under sp_noise, while the next point-to-point
        # operations may already be queued on the main stream: it gets a communicator of its own, so the two
        # streams never interleave operations of one NCCL communicator. (Collective call: every rank builds its
        # TorchDistComm at the same point.)
        # FABBER_B200_SLAB_PRIORITY=1 (opt-in, with the library's side stream at high priority): NCCL streams of
        # high priority too, and a communicator for the halo exchange when it runs on the side stream.
        import os

        prio = os.environ.get("FABBER_B200_SLAB_PRIORITY") == "1"
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True) if prio else None
        self.ak_group = dist.new_group(ranks=list(range(world)), pg_options=opts)
        self.side_group = dist.new_group(ranks=list(range(world)), pg_options=opts) if prio else None

    @staticmethod
    def _on(stream):
        """torch.distributed orders its NCCL work behind torch's CURRENT stream (and makes that stream wait for
        the result). The library runs on torch's default stream plus one side stream of its own; making the
        callback's stream current for the call keeps the whole iteration loop asynchronous - no host
        synchronize per callback."""
        import contextlib

        import torch

        if int(stream or 0) == int(torch.cuda.current_stream().cuda_stream):
            return contextlib.nullcontext()
        return torch.cuda.stream(torch.cuda.ExternalStream(int(stream)))

    def allreduce(self, ptr, n, stream):
        import torch.distributed as dist

        t = _tensor(ptr, n)
        with self._on(stream):
            dist.all_reduce(t, group=self.ak_group)
        return 0

    def exchange(self, send_lo, n_slo, send_hi, n_shi, recv_lo, n_rlo, recv_hi, n_rhi, stream):
        import torch
        import torch.distributed as dist

        ops, keep = [], []
        P = self.P
        # on the side stream (every rank takes the same branch in the same iteration): its own communicator
        on_side = int(stream or 0) != int(torch.cuda.current_stream().cuda_stream)
        grp = self.side_group if on_side else None
        if n_slo:
            keep.append(_tensor(send_lo, P * n_slo))
            ops.append(dist.P2POp(dist.isend, keep[-1], self.rank - 1, group=grp))
        if n_rlo:
            keep.append(_tensor(recv_lo, P * n_rlo))
            ops.append(dist.P2POp(dist.irecv, keep[-1], self.rank - 1, group=grp))
        if n_shi:
            keep.append(_tensor(send_hi, P * n_shi))
            ops.append(dist.P2POp(dist.isend, keep[-1], self.rank + 1, group=grp))
        if n_rhi:
            keep.append(_tensor(recv_hi, P * n_rhi))
            ops.append(dist.P2POp(dist.irecv, keep[-1], self.rank + 1, group=grp))
        if ops:
            with self._on(stream):
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        return 0

    def forward(self, step, send, n_send, recv, n_recv, stream):
        import torch.distributed as dist

        if not n_send and not n_recv:
            return 0
        ops, keep = [], []
        if n_send:
            keep.append(_tensor(send, self.P * n_send))
            ops.append(dist.P2POp(dist.isend, keep[-1], self.rank + 1))
        if n_recv:
            keep.append(_tensor(recv, self.P * n_recv))
            ops.append(dist.P2POp(dist.irecv, keep[-1], self.rank - 1))
        with self._on(stream):
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return 0


class ThreadComm(object):
    """Rank emulation for the single-GPU tests: `world` Python threads in one process, each inside its own
    fabber_cuda_vb_spatial_slab call (ctypes releases the GIL), meeting at barriers; buffers are exchanged
    with device-to-device copies. Same callback protocol as TorchDistComm, no NCCL."""

    class Shared(object):
        def __init__(self, world):
            self.barrier = threading.Barrier(world)
            self.slots = {}
            self.lock = threading.Lock()

    def __init__(self, rank, world, n_params, shared):
        self.rank, self.world, self.P, self.sh = rank, world, n_params, shared

    def allreduce(self, ptr, n, stream):
        import torch

        torch.cuda.synchronize()
        with self.sh.lock:
            self.sh.slots[("ar", self.rank)] = _tensor(ptr, n)
        self.sh.barrier.wait()
        if self.rank == 0:
            total = sum(self.sh.slots[("ar", r)].clone() for r in range(self.world))
            for r in range(self.world):
                self.sh.slots[("ar", r)].copy_(total)
            torch.cuda.synchronize()
        self.sh.barrier.wait()
        return 0

    def exchange(self, send_lo, n_slo, send_hi, n_shi, recv_lo, n_rlo, recv_hi, n_rhi, stream):
        import torch

        torch.cuda.synchronize()
        P = self.P
        with self.sh.lock:
            self.sh.slots[("lo", self.rank)] = _tensor(send_lo, P * n_slo) if n_slo else None
            self.sh.slots[("hi", self.rank)] = _tensor(send_hi, P * n_shi) if n_shi else None
        self.sh.barrier.wait()
        if n_rlo:   # from the rank below: what it sends up
            _tensor(recv_lo, P * n_rlo).copy_(self.sh.slots[("hi", self.rank - 1)])
        if n_rhi:   # from the rank above: what it sends down
            _tensor(recv_hi, P * n_rhi).copy_(self.sh.slots[("lo", self.rank + 1)])
        torch.cuda.synchronize()
        self.sh.barrier.wait()
        return 0


    def forward(self, step, send, n_send, recv, n_recv, stream):
        import torch

        torch.cuda.synchronize()
        with self.sh.lock:
            self.sh.slots[("fwd", self.rank)] = _tensor(send, self.P * n_send) if n_send else None
        self.sh.barrier.wait()
        if n_recv:
            _tensor(recv, self.P * n_recv).copy_(self.sh.slots[("fwd", self.rank - 1)])
        torch.cuda.synchronize()
        self.sh.barrier.wait()
        return 0


class SlabRun(object):
    """One rank's slab, set up once and launched any number of times (device buffers, index lists and the
    callback trampolines persist): `launch()` enqueues one whole spatial VB run, `results()` downloads the
    voxels this rank owns."""

    def __init__(self, spec, plan, comm):
        L = device.lib()
        L.fabber_cuda_vb_spatial_slab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fabber_cuda_vb_spatial_slab.restype = C.c_int
        self.L, self.spec, self.plan, self.comm = L, spec, plan, comm
        spec.prob.nx, spec.prob.ny, spec.prob.nz = plan.nx, plan.ny, plan.nz_local
        spec.prob.n_voxels = plan.n_local
        self.run = device.VbRun(spec, plan.n_local, spatial=True)
        self.keep = []
        try:
            self.run.set_coords(plan.coords())
            self.slab = slab = Slab()
            slab.n_global_voxels = plan.n_global
            slab.ghost = self._dev(plan.ghost_mask())
            for name, arr in zip(("send_lo", "send_hi", "recv_lo", "recv_hi"), plan.halo_lists()):
                setattr(slab, name, self._dev(arr))
                setattr(slab, "n_" + name, int(arr.size))
            slab.rank, slab.world = plan.rank, plan.world
            slab.n_blocks, slab.block_planes, slab.plane_offset = plan.n_blocks, plan.block_planes, plan.zlo
            fsend, fsend_start, frecv, frecv_start = plan.forward_lists()
            slab.fwd_send, slab.fwd_recv = self._dev(fsend), self._dev(frecv)
            self.starts = [(C.c_int * len(a))(*a.tolist()) for a in (fsend_start, frecv_start)]
            slab.fwd_send_start = C.cast(self.starts[0], C.POINTER(C.c_int))
            slab.fwd_recv_start = C.cast(self.starts[1], C.POINTER(C.c_int))

            def guarded(fn):
                def call(user, *a):
                    try:
                        return fn(*a)
                    except Exception:   # never let an exception cross the C boundary
                        import traceback

                        traceback.print_exc()
                        return -1
                return call

            slab.allreduce_sum = ALLREDUCE_FN(guarded(comm.allreduce))
            slab.exchange = EXCHANGE_FN(guarded(comm.exchange))
            slab.forward = FORWARD_FN(guarded(comm.forward))
        except Exception:
            self.close()
            raise

    def _dev(self, arr):
        d = device.DeviceArray.from_host(arr if arr.size else np.zeros(1, dtype=arr.dtype))
        self.keep.append(d)
        return d.ptr

    def set_data(self, data_local):
        self.run.set_data(data_local)

    def set_data_device(self, ptr):
        self.run.set_data_device(ptr)

    def launch(self, stream=None):
        """enqueue one run on `stream` (None: the default stream, which is also torch's default current stream,
        so TorchDistComm's NCCL calls order themselves behind the kernels without a host synchronize)"""
        rc = self.L.fabber_cuda_vb_spatial_slab(C.byref(self.spec.prob), C.byref(self.run.buf), C.byref(self.slab),
                                                stream)
        if rc not in (abi.OK, abi.ERR_BAD_VOXEL):
            raise device.CudaError("slab VB failed (%d): %s" % (rc, device.last_error()))
        return rc

    def results(self):
        self.run.sync()
        out = self.run.results()
        own, n = self.plan.own_slice(), self.plan.n_local
        return {k: (v[..., own] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[-1] == n else v)
                for k, v in out.items()}

    def close(self):
        if self.run is not None:
            self.run.close()
            self.run = None
        for d in self.keep:
            d.close()
        self.keep = []


def run_slab(spec, data_local, plan, comm, data_is_device_ptr=False):
    """Run one rank's slab once. data_local: float32 [T][plan.n_local] (host array, or a device pointer when
    data_is_device_ptr). Returns the result dict restricted to the voxels this rank owns."""
    sr = SlabRun(spec, plan, comm)
    try:
        if data_is_device_ptr:
            sr.set_data_device(data_local)
        else:
            sr.set_data(data_local)
        rc = sr.launch()
        res = sr.results()
        res["rc"] = rc
        return res
    finally:
        sr.close()
