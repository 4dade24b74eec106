"""Row-slab decomposition of one large domain across the GPUs of a box (one process per GPU).

The reference is single-GPU (`cudaSetDevice(0)`, Deff2D.cuh:908).  Its solve loop
(`JacobiGPU`, cuh:1163-1314) shards naturally by rows: the 5-point stencil only couples
neighbouring rows, and the convergence check needs two sums over the boundary columns.
`SlabLayout` is the pure geometry (which source rows a rank owns and which halo rows it also
holds); `SlabDomain` loads a rank's slab into libdeff2d, initialises the library's NCCL
communicator (unique id broadcast through `torch.distributed`) and drives the same calls as a
single-GPU domain: `sweeps`, `flux`, `solve`.  Halo exchange (ncclSend/ncclRecv of the boundary
rows after every temporally blocked pass) and the flux all-reduce happen inside the library
(csrc/slab.cu).
"""
import numpy as np

from . import api


def partition_rows(nrows, world):
    """Contiguous split of `nrows` source rows over `world` ranks: [(row0, rows), ...]."""
    base, rem = divmod(int(nrows), int(world))
    out, r0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((r0, n))
        r0 += n
    return out


class SlabLayout:
    """Geometry of rank `rank`'s slab of a global image of `H` source rows (amplified by amp_y).

    halo: amplified halo rows held of each neighbour; rounded up to a multiple of amp_y because
    the slab loader thresholds whole source rows (deff2d_domain_load_slab)."""

    def __init__(self, H, rank, world, amp_y=1, halo=16):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank/world")
        self.H, self.rank, self.world, self.amp_y = int(H), int(rank), int(world), int(amp_y)
        self.halo = -(-int(halo) // self.amp_y) * self.amp_y
        self.halo_src = self.halo // self.amp_y
        self.src_row0, self.src_rows = partition_rows(H, world)[rank]
        if world > 1 and self.src_rows * self.amp_y < self.halo:
            raise ValueError("slab of %d rows is thinner than its halo (%d)" % (self.src_rows * self.amp_y, self.halo))
        self.src_above = self.halo_src if rank > 0 else 0
        self.src_below = self.halo_src if rank < world - 1 else 0
        self.ny_global = self.H * self.amp_y
        self.row0 = self.src_row0 * self.amp_y                  # first own amplified row (global)
        self.own_rows = self.src_rows * self.amp_y
        self.above = self.src_above * self.amp_y
        self.below = self.src_below * self.amp_y

    @property
    def local_src(self):
        """(first, last+1) global source rows held locally, halo rows included."""
        return self.src_row0 - self.src_above, self.src_row0 + self.src_rows + self.src_below

    @property
    def local_rows(self):
        """(first, last+1) global amplified rows held locally, halo rows included."""
        return self.row0 - self.above, self.row0 + self.own_rows + self.below


def _take_rows(arr, lo, hi, period=None):
    """Rows [lo, hi) of `arr`; with `period` the array repeats every `period` rows (weak scaling:
    the global domain is `world` vertical copies of one image)."""
    if period is None:
        return np.ascontiguousarray(arr[lo:hi])
    idx = np.arange(lo, hi) % period
    return np.ascontiguousarray(arr[idx])


class SlabDomain:
    """One rank's slab, resident on this rank's GPU.

    img:     the global 8-bit source image (H x W); with weak=True the global domain is `world`
             vertical copies of `img` (every rank then owns one whole copy).
    halo:    amplified halo rows held of each neighbour.  A pass of depth T consumes T of them, so
             the exchange runs every halo / T passes (32 rows, T = 8: once per 4 passes = 32 sweeps).
    pinned:  None (default): FloodFill runs on the device over the whole domain (deff2d_domain_load_slab_global).
             Otherwise the global FloodFill mask (cuh:557-713) at amplified resolution, computed by the caller;
             only the slab's rows of the image and of the mask are uploaded then (deff2d_domain_load_slab).
    """

    def __init__(self, ctx, img, params, rank, world, nphase=3, halo=None, pinned=None, weak=False, nccl_id=None, peer=None):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        # peer=False: NCCL send/recv of deep halos (32 rows, one exchange per 5 passes).  peer=True (ranks of one box):
        # the sweep kernel pushes the boundary rows into the neighbours' halo rows itself (8 halo rows, none
        # recomputed); bit-identical.  Measured (B200, GLUP/s, peer / NCCL): 2 GPUs 1 686 / 1 712 (config 2 stacked) and
        # 1 687 / 1 711 (config 4); 8 GPUs 6 730 / 6 661 and 6 426 / 5 790.  peer=None (default): peer mode from 4 ranks
        # up, and the NCCL exchange if the ranks cannot map each other's memory (they decide together).
        auto = peer is None
        self.peer = (world >= 4) if auto else (bool(peer) and world > 1)
        if nccl_id is None:
            nccl_id = self._broadcast_id()
        ctx.nccl_init(nccl_id, rank, world)
        try:
            self._setup(img, params, nphase, halo, pinned, weak)
        except RuntimeError:
            # the attach failed on every rank together (deff2d_slab_peer_attach reduces the failures over the group)
            if not (auto and self.peer):
                raise
            self.peer = False
            self._setup(img, params, nphase, halo, pinned, weak)

    def _setup(self, img, params, nphase, halo, pinned, weak):
        ctx, rank, world = self.ctx, self.rank, self.world
        Hbase, W = img.shape
        if halo is None:
            halo = 8 if self.peer else 32
        H = Hbase * world if weak else Hbase
        period_src = Hbase if weak else None
        self.layout = L = SlabLayout(H, rank, world, params.amp_y, halo)
        self.Nx = W * params.amp_x
        self.global_cells = self.Nx * L.ny_global
        self.pathflag = None
        if pinned is None:
            # default: every rank uploads the WHOLE source image (1 B per pixel); FloodFill (cuh:557-713) runs on its
            # device over the whole domain -- the flood needs global connectivity, so it is replicated, not sharded --
            # and the slab keeps its rows of the mask.  No host flood, no 1 B/cell mask upload: the same steps as an
            # undecomposed deff2d_domain_load, which is what `reload()` times end to end.
            gimg = np.ascontiguousarray(np.tile(img, (world, 1))) if weak else img
            self._loader, self._load_args = ctx.domain_load_slab_global, (gimg, nphase, params, L.row0, L.own_rows, L.halo)
            self.h2d_bytes = int(gimg.size)
        else:
            lo, hi = L.local_src
            gray = _take_rows(img, lo, hi, period_src)
            a, b = L.local_rows
            pin = _take_rows(np.asarray(pinned, dtype=np.uint8), a, b, None) if nphase == 3 else None
            self._loader, self._load_args = ctx.domain_load_slab, (gray, nphase, params, L.row0, L.ny_global, L.halo, L.src_rows, pin)
            self.h2d_bytes = int(gray.size + (pin.size if pin is not None else 0))
        self._loader(*self._load_args)
        self._attach()
        if pinned is None:
            self.pathflag = ctx.info()["pathflag"]

    def _attach(self):
        """Peer mode: every rank publishes the IPC handles of its buffers, takes its neighbours' and attaches."""
        if not self.peer:
            return
        import torch.distributed as dist
        try:
            mine = self.ctx.slab_peer_export()
        except RuntimeError:
            mine = None                                    # (no IPC handle for these buffers): tell the others
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        if any(h is None for h in handles):
            raise RuntimeError("peer mode: a rank of the group could not export its buffers")
        self.ctx.slab_peer_attach(handles[self.rank - 1] if self.rank > 0 else None,
                                  handles[self.rank + 1] if self.rank < self.world - 1 else None)

    def reload(self):
        """Upload again from the host buffers and reset the iterate to x0: what a fresh solve of the same domain
        costs end to end (image upload, FloodFill, assembly, and in peer mode the handle exchange)."""
        self._loader(*self._load_args)
        self._attach()

    def _broadcast_id(self):
        import torch.distributed as dist
        if self.world == 1 and not (dist.is_available() and dist.is_initialized()):
            return api.nccl_unique_id()
        box = [api.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def sweeps(self, n):
        self.ctx.slab_sweeps(n)

    def flux(self):
        return self.ctx.slab_flux()

    def solve(self, tol, max_iter):
        """The reference loop (cuh:1232-1290) on the decomposed domain; identical on every rank."""
        return self.ctx.solve(tol, max_iter)

    def own_field(self):
        """This rank's own rows of the iterate (own_rows x Nx)."""
        f = self.ctx.get_field()
        L = self.layout
        return f[L.above:L.above + L.own_rows]
