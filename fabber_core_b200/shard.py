"""Multi-GPU decomposition of the VB path.

Non-spatial VB: voxels are independent (inference_vb.cc:423-571), so the masked voxel list is cut into
contiguous, balanced ranges - one per rank, one process per GPU - with NO data-path collective; the only
communication is the final gather of the result arrays (torch.distributed, NCCL on GPUs / gloo in the
CPU tests). Spatial VB partitions z-slabs (voxel order is z-major, so a slab is a contiguous range too);
the slab run itself - all-reduced aK sums, the pipelined exact sweep and the halo exchange - lives in
fabber_cuda_vb_spatial_slab / spatial_mgpu.py, see DESIGN.md section 7.
"""
import numpy as np


def voxel_range(n_voxels, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: the first n % world ranks hold one extra voxel."""
    base, extra = divmod(int(n_voxels), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def z_slab_range(nz, rank, world):
    """z-planes [z0, z1) of rank `rank` (spatial mode); same balancing rule as voxel_range."""
    return voxel_range(nz, rank, world)


def shard_columns(array, rank, world):
    """View of the voxel columns (last axis) of `array` that belong to `rank`."""
    lo, hi = voxel_range(array.shape[-1], rank, world)
    return array[..., lo:hi]


def run_sharded(runner, spec_factory, data, rank, world, **kwargs):
    """Run `runner(spec, data_shard, **shard_kwargs)` on this rank's voxel range.

    `kwargs` holding per-voxel arrays (image_priors, init_mean, init_cov, init_noise) are sharded the
    same way. Returns the local result dict."""
    lo, hi = voxel_range(data.shape[1], rank, world)
    local = {}
    for k, v in kwargs.items():
        if v is None:
            local[k] = None
        elif isinstance(v, dict):
            local[k] = {kk: np.ascontiguousarray(np.asarray(vv)[..., lo:hi]) for kk, vv in v.items()}
        else:
            local[k] = np.ascontiguousarray(np.asarray(v)[..., lo:hi])
    return runner(spec_factory(), np.ascontiguousarray(data[:, lo:hi]), **local)


def gather_results(local, n_voxels, rank, world, dst=0):
    """Final gather of per-voxel result arrays ([..., n_local]) to rank `dst` over torch.distributed.
    Returns the full dict on `dst`, None elsewhere. world == 1 needs no process group."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist

    keys = sorted(k for k, v in local.items() if isinstance(v, np.ndarray) and v.ndim >= 1)
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    out = {} if rank == dst else None
    for k in keys:
        arr = np.ascontiguousarray(local[k])
        rows = int(np.prod(arr.shape[:-1])) if arr.ndim > 1 else 1
        counts = [voxel_range(n_voxels, r, world)[1] - voxel_range(n_voxels, r, world)[0] for r in range(world)]
        pad = max(counts)
        send = torch.zeros((rows, pad), dtype=torch.from_numpy(arr.reshape(rows, -1)).dtype, device=device)
        send[:, :arr.shape[-1]] = torch.from_numpy(arr.reshape(rows, -1)).to(device)
        recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send, recv, dst=dst)
        if rank == dst:
            full = np.concatenate([recv[r][:, :counts[r]].cpu().numpy() for r in range(world)], axis=1)
            out[k] = full.reshape(arr.shape[:-1] + (n_voxels,))
    return out
