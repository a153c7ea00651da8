"""World_size 2, gloo: the N > 1 path of non-spatial VB - balanced contiguous voxel ranges, no data-path
collective, one final gather - gives exactly the single-process result. Without a GPU (this container) the
per-rank compute is the oracle, so the sharding / gather logic itself is what is tested; where CUDA is present
(the B200 box) the same ranks run the DEVICE path (device.run) on their shards and the gathered result must be
bit-identical to the one-launch device run."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fabber_core_b200 import cuda_abi as abi
from fabber_core_b200 import shard, synth


def test_voxel_ranges_partition_exactly():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [shard.voxel_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard.z_slab_range(256, 3, 8) == (96, 128)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _runner():
    import torch

    if torch.cuda.is_available():
        from fabber_core_b200 import device

        return device.run
    return oracle.run


def _spec(case, T):
    """the problems whose voxels are independent: VB (white noise, image prior), NLLS, two-echo AR(1)"""
    if case == "vb":
        return abi.ProblemSpec("poly", T, degree=2, prior_types=["N", "N", "I"], need_f=True, convergence="pointzeroone")
    if case == "nlls":
        return abi.ProblemSpec("poly", T, degree=2, method="nlls")
    if case == "ar2":
        return abi.ProblemSpec("poly", T, degree=2, noise="ar", num_echoes=2, ar_cross_terms="dual", need_f=True,
                               max_iterations=4)
    raise ValueError(case)


def _worker(rank, world, port, y, img, tmp, case="vb"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mk = lambda: _spec(case, y.shape[0])
        local = shard.run_sharded(_runner(), mk, y, rank, world, image_priors={2: img} if case == "vb" else None)
        lo, hi = shard.voxel_range(y.shape[1], rank, world)
        assert local["mean"].shape == (3, hi - lo)
        full = shard.gather_results(local, y.shape[1], rank, world, dst=0)
        if rank == 0:
            np.savez(tmp, **{k: v for k, v in full.items()})
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["vb", "nlls", "ar2"])
def test_two_rank_shards_reproduce_the_single_process_run(tmp_path, case):
    y = synth.poly_volume(401 if case == "vb" else 91, 30, 2, seed=51).numpy()  # odd counts: uneven shards
    img = np.linspace(-1e-3, 1e-3, y.shape[1])
    tmp = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), y, img, tmp, case), nprocs=2, join=True)
    got = np.load(tmp)
    ref = _runner()(_spec(case, 30), y, image_priors={2: img} if case == "vb" else None)
    keys = ("mean", "cov", "iterations", "status") + (() if case == "nlls" else ("noise", "free_energy"))
    for k in keys:
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
