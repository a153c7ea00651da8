"""CPU: the z-slab decomposition of the spatial path (fabber_core_b200/spatial_mgpu.py) - ownership, ghost
planes and halo index lists are mutually consistent for any rank count; and a world_size-2 gloo run
exchanges exactly the planes the plan names (the collective protocol of the N > 1 spatial path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fabber_core_b200.spatial_mgpu import SlabPlan


def test_slabs_tile_the_volume_and_halos_match():
    nx, ny, nz = 5, 4, 11
    for world in (1, 2, 3, 4, 11):
        plans = [SlabPlan(nx, ny, nz, r, world) for r in range(world)]
        assert plans[0].z0 == 0 and plans[-1].z1 == nz
        owned = np.zeros(nx * ny * nz, dtype=int)
        for p in plans:
            g0, g1 = p.global_columns()
            glob = np.arange(g0, g1)
            owned[glob[p.own_slice()]] += 1
            assert p.ghost_mask().sum() == (p.ghost_lo + p.ghost_hi) * nx * ny
            assert not p.ghost_mask()[p.own_slice()].any()
            c = p.coords()
            assert c[2].max() == p.nz_local - 1 and c.shape == (3, p.n_local)
        assert np.all(owned == 1)
        for r in range(world - 1):
            lo, hi = plans[r], plans[r + 1]
            # what r sends up is what r+1 receives from below, as global voxel numbers, and vice versa
            up = lo.global_columns()[0] + lo.halo_lists()[1]
            rb = hi.global_columns()[0] + hi.halo_lists()[2]
            assert np.array_equal(up, rb)
            down = hi.global_columns()[0] + hi.halo_lists()[0]
            ra = lo.global_columns()[0] + lo.halo_lists()[3]
            assert np.array_equal(down, ra)


def test_forward_lists_pair_up_block_by_block():
    """pipelined sweep: what rank r forwards for block b is what rank r+1 expects for block b, the blocks
    cover the boundary plane once, and every forwarded voxel's hyper-plane x+y+z lies in its block."""
    nx, ny, nz = 5, 4, 11
    for world, bp in ((2, None), (3, 1), (4, 3), (2, 50)):
        plans = [SlabPlan(nx, ny, nz, r, world, bp) for r in range(world)]
        nb, B = plans[0].n_blocks, plans[0].block_planes
        assert nb * B >= nx + ny + nz - 2 > (nb - 1) * B
        assert plans[0].forward_lists()[2].size == 0 and plans[-1].forward_lists()[0].size == 0
        for r in range(world - 1):
            lo, hi = plans[r], plans[r + 1]
            send, send_start, _, _ = lo.forward_lists()
            _, _, recv, recv_start = hi.forward_lists()
            assert np.array_equal(send_start, recv_start) and send_start[0] == 0 and send_start[-1] == nx * ny
            assert np.array_equal(lo.global_columns()[0] + send, hi.global_columns()[0] + recv)
            assert np.array_equal(np.sort(send), lo.halo_lists()[1])
            g = lo.global_columns()[0] + send
            H = g % nx + (g // nx) % ny + g // (nx * ny)
            for b in range(nb):
                blk = H[send_start[b]:send_start[b + 1]]
                assert np.all(blk // B == b)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nx, ny, nz, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = SlabPlan(nx, ny, nz, rank, world)
        g0, _ = plan.global_columns()
        field = torch.arange(g0, g0 + plan.n_local, dtype=torch.float64) * 1.5   # value = f(global voxel number)
        send_lo, send_hi, recv_lo, recv_hi = [torch.as_tensor(a, dtype=torch.long) for a in plan.halo_lists()]
        field[recv_lo] = -1.0
        field[recv_hi] = -1.0
        ops, bufs = [], {}
        if len(send_lo):
            ops.append(dist.P2POp(dist.isend, field[send_lo].contiguous(), rank - 1))
            bufs["lo"] = torch.empty(len(recv_lo), dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, bufs["lo"], rank - 1))
        if len(send_hi):
            ops.append(dist.P2POp(dist.isend, field[send_hi].contiguous(), rank + 1))
            bufs["hi"] = torch.empty(len(recv_hi), dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, bufs["hi"], rank + 1))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        if "lo" in bufs:
            field[recv_lo] = bufs["lo"]
        if "hi" in bufs:
            field[recv_hi] = bufs["hi"]
        expect = torch.arange(g0, g0 + plan.n_local, dtype=torch.float64) * 1.5
        ok = bool(torch.equal(field, expect))
        sums = torch.tensor([float(rank + 1), 2.0])
        dist.all_reduce(sums)
        ok = ok and sums.tolist() == [sum(range(1, world + 1)), 2.0 * world]
        with open(out % rank, "w") as f:
            f.write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


def test_gloo_halo_exchange_and_allreduce_protocol(tmp_path):
    out = str(tmp_path / "rank%d.txt")
    mp.spawn(_worker, args=(2, _free_port(), 4, 3, 7, out), nprocs=2, join=True)
    assert open(out % 0).read() == "ok" and open(out % 1).read() == "ok"


def test_eight_slabs_schedule_is_a_wavefront():
    """The 8-GPU configuration of bench.py (one volume in eight z-slabs), scaled down: every rank's blocks with
    work form one contiguous run, each rank starts later than the one below it (the stagger DESIGN.md section 7
    measures), and the pipeline needs n_blocks + world - 1 steps."""
    nx, ny, nz, world = 6, 6, 48, 8
    plans = [SlabPlan(nx, ny, nz, r, world) for r in range(world)]
    nb, B = plans[0].n_blocks, plans[0].block_planes
    first = []
    for r, p in enumerate(plans):
        # local hyper-planes of slab r are global planes [zlo, zlo + nx + ny + nz_local - 2]
        lo, hi = p.zlo, p.zlo + nx + ny + p.nz_local - 3
        busy = [b for b in range(nb) if not ((b + 1) * B - 1 < lo or b * B > hi)]
        assert busy == list(range(busy[0], busy[-1] + 1))
        first.append(busy[0] + r)                     # rank r sweeps block b at step b + r
        assert busy[-1] + r <= nb + world - 2
    assert all(b > a for a, b in zip(first, first[1:]))
