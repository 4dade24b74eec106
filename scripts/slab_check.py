#!/usr/bin/env python
"""Multi-GPU slab check, one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/slab_check.py [--size HxW] [--time]

Every rank loads its slab of a seeded 3-phase (and 2-phase) domain, runs sweeps with NCCL halo
exchange and compares its own rows, Deff and the full reference loop with an undecomposed run of
the same domain on its own GPU.  --time adds a weak-scaling timing of BASELINE config 2 per GPU.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import effectivediffusivityfvm_b200 as E  # noqa: E402
from effectivediffusivityfvm_b200.slab import SlabDomain  # noqa: E402


def blobs(seed, shape, levels=(0, 150, 255), fracs=(0.3, 0.4), smooth=3):
    rng = np.random.default_rng(seed)
    z = rng.random(shape)
    for _ in range(smooth):
        z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
    qs = np.quantile(z, np.cumsum(fracs))
    out = np.full(shape, levels[-1], np.uint8)
    for lv, q in reversed(list(zip(levels[:-1], qs))):
        out[z < q] = lv
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="600x900")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--sweeps", type=int, default=2000)
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--c4", action="store_true", help="strong scaling of BASELINE config 4: one 16384^2 two-phase domain over all ranks")
    ap.add_argument("--detach", action="store_true")
    ap.add_argument("--time-halo", type=int, default=0)
    ap.add_argument("--peer", default="1,0", help="exchange modes to check: 1 = peer-memory push fused into the kernel, 0 = NCCL deep halos")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W = (int(v) for v in args.size.split("x"))
    ok = True
    report = {}
    modes = [bool(int(v)) for v in args.peer.split(",")]
    for peer, (nphase, amp, halo) in (() if args.skip_check else [(pm, cs) for pm in modes for cs in ((3, (2, 2), 4), (2, (1, 1), 16), (3, (1, 1), 12))]):
        img = blobs(11 + nphase, (H, W))
        p = E.default_params(Ds=0.0 if nphase == 3 else 1e-3, Df=1.0, Dg=80.0, amp_x=amp[0], amp_y=amp[1],
                             CL=0.25, CR=1.5, check_every=400)
        ref = E.Deff2D(local)
        ref.set_kernel(2, 4)
        ref.domain_load(img, nphase, p)
        ctx = E.Deff2D(local)
        if halo != 16:
            ctx.set_kernel(2, 4)                               # halo 16 runs the library's default kernel and depth
        dom = SlabDomain(ctx, img, p, rank, world, nphase=nphase, halo=halo, peer=peer)
        L = dom.layout
        for n in (1, 4, 203):
            ref.sweeps(n)
            dom.sweeps(n)
            full = ref.get_field()
            own = dom.own_field()
            same = np.array_equal(own, full[L.row0:L.row0 + L.own_rows], equal_nan=True)
            d_ref, d = ref.flux()[0], dom.flux()
            rel = abs(d - d_ref) / abs(d_ref)
            ok = ok and same and rel < 1e-12
            report["peer%d_p%d_h%d_n%d" % (peer, nphase, halo, n)] = {"field_equal": bool(same), "deff_rel": rel}
        # the reference loop on the decomposed domain: same sweep count, same Deff on every rank
        ref.domain_load(img, nphase, p)
        r_ref = ref.solve(1e-4, 6000)
        ctx2 = ctx
        dom = SlabDomain(ctx2, img, p, rank, world, nphase=nphase, halo=halo, peer=peer)
        r = dom.solve(1e-4, 6000)
        rel = abs(r["deff_raw"] - r_ref["deff_raw"]) / abs(r_ref["deff_raw"])
        ok = ok and r["iters"] == r_ref["iters"] and rel < 1e-12
        report["peer%d_p%d_h%d_solve" % (peer, nphase, halo)] = {"iters": r["iters"], "iters_ref": r_ref["iters"], "deff_rel": rel}
        ref.close()
        ctx.close()
    if not args.skip_check and world >= 2:
        # A peer attach that fails on ONE rank (its slab is thinner than twice the halo) must fail on every rank of the
        # group, not leave the healthy ones waiting in the barrier; afterwards the same context still runs the NCCL exchange.
        rows = 20 * (2 * world) - 1                            # last rank gets 39 rows, the others 40: 2 x halo = 40
        img = blobs(5, (rows, 96))
        p = E.default_params(Ds=1e-3, Df=1.0, CL=0.25, CR=1.5, check_every=400)
        ctx = E.Deff2D(local)
        try:
            SlabDomain(ctx, img, p, rank, world, nphase=2, halo=20, peer=True)
            failed_together = False
        except RuntimeError:
            failed_together = True
        dom = SlabDomain(ctx, img, p, rank, world, nphase=2, halo=20, peer=False)
        ref = E.Deff2D(local)
        ref.domain_load(img, 2, p)
        ref.sweeps(50)
        dom.sweeps(50)
        L = dom.layout
        same = np.array_equal(dom.own_field(), ref.get_field()[L.row0:L.row0 + L.own_rows], equal_nan=True)
        ok = ok and failed_together and same
        report["peer_attach_fails_on_all_ranks"] = {"failed_together": failed_together, "nccl_afterwards_equal": bool(same)}
        ref.close()
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "ok": bool(flag.item()), "rank0": report}), flush=True)
    if args.time:
        img = np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["img00042"]
        p = E.default_params(amp_x=4, amp_y=4)
        ctx = E.Deff2D(local)
        dom = SlabDomain(ctx, img, p, rank, world, weak=True, peer=modes[0], halo=args.time_halo or None)
        if args.detach:
            from effectivediffusivityfvm_b200 import _lib
            _lib.lib().deff2d_slab_peer_detach(ctx._h)         # diagnostic: NCCL path on peer-mapped buffers
        dom.sweeps(400)                                       # warm-up incl. the CUDA-graph capture
        dom.flux()
        dist.barrier()
        torch.cuda.synchronize()
        S = args.sweeps
        stream = torch.cuda.ExternalStream(ctx.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        dom.sweeps(S)
        d = dom.flux()
        e1.record(stream)
        ctx.sync()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"world": world, "weak_glups": dom.global_cells * S / (t.item() * 1e-3) / 1e9,
                              "ms": t.item(), "deff_raw": d}), flush=True)
        ctx.close()
    if args.c4:
        from effectivediffusivityfvm_b200.datasets import c4_image
        img = np.tile(c4_image(4096), (4, 4))                  # periodic generator: a seamless 16384^2 medium
        p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH)
        ctx = E.Deff2D(local)
        dom = SlabDomain(ctx, img, p, rank, world, nphase=2, peer=modes[0])
        dom.sweeps(400)                                       # warm-up incl. the CUDA-graph capture
        dom.flux()
        dist.barrier()
        torch.cuda.synchronize()
        S = args.sweeps
        stream = torch.cuda.ExternalStream(ctx.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        dom.sweeps(S)
        d = dom.flux()
        e1.record(stream)
        ctx.sync()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"world": world, "c4_strong_glups": dom.global_cells * S / (t.item() * 1e-3) / 1e9,
                              "ms": t.item(), "sweeps": S, "deff_raw": d}), flush=True)
        ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
