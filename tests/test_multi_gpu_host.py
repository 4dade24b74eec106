"""CPU tests of the host-side logic behind the multi-GPU and packed-batch paths:

* tile planning exported by libdeff2d (slab boundary/interior split, packed-batch tile lists);
* the row-slab geometry (`SlabLayout`) and -- with two `gloo` ranks on CPU -- the decomposition
  algebra itself: every rank sweeps its slab (numpy emulation of the device data model, the
  library's own weight tables), exchanges halo rows whenever the next pass of T sweeps would
  need more exact halo rows than are left (every pass with H = T, every m-th with H = m T) and
  all-reduces the boundary-flux sums, exactly as csrc/slab.cu does with NCCL.  The result must equal the
  undecomposed run bit for bit.
"""
import os
import socket

import numpy as np
import pytest

import _emulate as EM
import _oracle as O
import effectivediffusivityfvm_b200 as E
from effectivediffusivityfvm_b200.slab import SlabLayout, partition_rows, _take_rows


def _boxes(tiles, ow, oh):
    return [((int(t) & 0xffff) * ow, (int(t) >> 16) * oh) for t in tiles]


@pytest.mark.parametrize("T", [1, 2, 3, 4, 6, 8])
def test_tile_geometry(T):
    ow, oh, tw, th = E.tile_geometry(T)
    assert (tw, th) == (64, 64)             # square tiles in both thread layouts
    assert oh == th - 2 * T and ow == tw - 2 * ((T + 1) // 2 * 2)
    assert ow % 2 == 0                      # TMA needs 16-byte aligned FP64 box origins


def test_batch_plan_and_tile_list():
    GX, GY = E.batch_plan(256, 256, 512)
    assert GX == 15 and GX * GY >= 512 and GX * (GY - 1) < 512
    assert E.batch_plan(256, 256, 7) == (7, 1)
    assert E.batch_plan(256, 256, 512, limit=32) == (15, 3)
    gx, gy = E.batch_plan(4000, 4000, 50)
    assert gx * gy <= 4 and gx >= 1
    Nx, Ny, GX, GY, T = 40, 56, 5, 3, 4
    ow, oh, _, _ = E.tile_geometry(T)
    rng = np.random.default_rng(1)
    for _ in range(20):
        active = np.flatnonzero(rng.random(GX * GY) < 0.4)
        tiles = E.batch_tile_list(Nx, Ny, GX, GY, active, T)
        NxS, NyS = GX * (Nx + 1) - 1, GY * (Ny + 1) - 1
        cover = np.zeros((NyS + oh, NxS + ow), dtype=np.int32)
        for x0, y0 in _boxes(tiles, ow, oh):
            cover[y0:y0 + oh, x0:x0 + ow] += 1
        assert cover.max() <= 1
        need = np.zeros_like(cover)
        for s in active:
            c0, r0 = (s % GX) * (Nx + 1), (s // GX) * (Ny + 1)
            need[r0:r0 + Ny, c0:c0 + Nx] = 1
            assert np.all(cover[r0:r0 + Ny, c0:c0 + Nx] == 1)          # active images are fully swept
        # no tile is scheduled that touches no active image
        for x0, y0 in _boxes(tiles, ow, oh):
            assert need[y0:y0 + oh, x0:x0 + ow].any()
    assert len(E.batch_tile_list(Nx, Ny, GX, GY, [], T)) == 0


def test_partition_and_layout():
    assert partition_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sum(n for _, n in partition_rows(2007, 8)) == 2007
    L = SlabLayout(2007, 3, 8, amp_y=4, halo=4)
    assert L.halo == 4 and L.halo_src == 1 and L.above == 4 and L.below == 4
    assert L.local_rows == (L.row0 - 4, L.row0 + L.own_rows + 4)
    L0, L7 = SlabLayout(2007, 0, 8, 4, 4), SlabLayout(2007, 7, 8, 4, 4)
    assert L0.above == 0 and L7.below == 0 and L7.local_rows[1] == 2007 * 4
    assert SlabLayout(100, 0, 2, amp_y=3, halo=4).halo == 6        # rounded up to whole source rows
    with pytest.raises(ValueError):
        SlabLayout(8, 0, 4, amp_y=1, halo=4)
    a = np.arange(12).reshape(6, 2)
    assert np.array_equal(_take_rows(a, 4, 9, period=6)[:, 0], [8, 10, 0, 2, 4])


# ----------------------------------------------------------------------------- 2 gloo ranks on CPU

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_worker(rank, world, port, img, nphase, T, nsweeps, out, halo_mult=1):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Ds, Dg, CL, CR = (0.0, 50.0, 0.25, 1.5) if nphase == 3 else (1e-3, 0.0, 0.25, 1.5)
        H, W = img.shape
        Nx, NyG = W, H
        L = SlabLayout(H, rank, world, 1, T * halo_mult)
        valid = L.halo                      # halo rows still exact (csrc/slab.cu: c->halo_valid)
        grid = None
        if nphase == 3:
            grid, _ = E.floodfill((img > 200).astype(np.uint8))
        lo, hi = L.local_rows
        codes = EM.phase_codes(img[lo:hi], nphase, grid=None if grid is None else grid[lo:hi])
        lut, _ = E.build_tables(Ds, 1.0, Dg, Nx, NyG, CL, CR)      # global dx, dy
        pc = EM.pad_codes(codes)
        Ny = hi - lo
        x0 = O.init_x(Nx, NyG, CL, CR)[lo:hi]
        px = EM.pad_field(x0)
        done = 0
        while done < nsweeps:
            t = min(T, nsweeps - done)
            if valid >= t:                   # deep halo: a pass of depth t only consumes t halo rows
                px = EM.sweep(px, pc, lut, Nx, Ny, nsweeps=t)
                valid -= t
                done += t
                continue
            # halo exchange of the current iterate: H whole padded rows each way (csrc/slab.cu)
            valid = L.halo
            reqs = []
            if L.above:
                send = torch.from_numpy(px[1 + L.above:1 + L.above + L.halo].copy())
                recv = torch.empty_like(send)
                reqs += [dist.isend(send, rank - 1), dist.irecv(recv, rank - 1)]
                up = recv
            if L.below:
                last = 1 + L.above + L.own_rows
                send2 = torch.from_numpy(px[last - L.halo:last].copy())
                recv2 = torch.empty_like(send2)
                reqs += [dist.isend(send2, rank + 1), dist.irecv(recv2, rank + 1)]
            for r in reqs:
                r.wait()
            if L.above:
                px[1 + L.above - L.halo:1 + L.above] = up.numpy()
            if L.below:
                px[last:last + L.halo] = recv2.numpy()
        own = px[1 + L.above:1 + L.above + L.own_rows, EM.XOFF:EM.XOFF + Nx]
        # boundary-flux partial sums of the own rows, all-reduced (cuh:1252-1264)
        Dtab = np.array([1.0, Ds, Dg])
        cown = codes[L.above:L.above + L.own_rows]
        half_dx = (1.0 / Nx) / 2.0
        q = torch.tensor([np.sum(Dtab[cown[:, 0] & 3] * (own[:, 0] - CL) / half_dx),
                          np.sum(Dtab[cown[:, -1] & 3] * (CR - own[:, -1]) / half_dx)], dtype=torch.float64)
        dist.all_reduce(q)
        deff = float((q[0] + q[1]) / (2.0 * NyG) / (CR - CL))
        np.save(os.path.join(out, "own%d.npy" % rank), own)
        np.save(os.path.join(out, "deff%d.npy" % rank), np.array([deff]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nphase,T,nsweeps,halo_mult", [(2, 4, 23, 1), (3, 3, 14, 1), (2, 2, 21, 3)])
def test_slab_decomposition_two_gloo_ranks_equals_single_domain(tmp_path, nphase, T, nsweeps, halo_mult):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5 + nphase)
    img = np.choose(rng.integers(0, 3, size=(37, 24)), [0, 150, 255]).astype(np.uint8)
    world = 2
    port = _free_port()
    mp.spawn(_slab_worker, args=(world, port, img, nphase, T, nsweeps, str(tmp_path), halo_mult), nprocs=world, join=True)
    # undecomposed run with the same emulation
    Ds, Dg, CL, CR = (0.0, 50.0, 0.25, 1.5) if nphase == 3 else (1e-3, 0.0, 0.25, 1.5)
    H, W = img.shape
    grid = E.floodfill((img > 200).astype(np.uint8))[0] if nphase == 3 else None
    codes = EM.phase_codes(img, nphase, grid=grid)
    lut, _ = E.build_tables(Ds, 1.0, Dg, W, H, CL, CR)
    px = EM.sweep(EM.pad_field(O.init_x(W, H, CL, CR)), EM.pad_codes(codes), lut, W, H, nsweeps=nsweeps)
    full = px[1:H + 1, EM.XOFF:EM.XOFF + W]
    got = np.concatenate([np.load(tmp_path / ("own%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got, full)                     # overlapped slabs are algebraically exact
    Dtab = np.array([1.0, Ds, Dg])
    half_dx = (1.0 / W) / 2.0
    deff = (np.sum(Dtab[codes[:, 0] & 3] * (full[:, 0] - CL) / half_dx) +
            np.sum(Dtab[codes[:, -1] & 3] * (CR - full[:, -1]) / half_dx)) / (2.0 * H) / (CR - CL)
    d0, d1 = (float(np.load(tmp_path / ("deff%d.npy" % r))[0]) for r in range(world))
    assert d0 == d1                                      # every rank applies the stop rule to the same number
    assert abs(d0 - deff) <= 1e-12 * abs(deff)
