// slab.cu -- multi-GPU row-slab decomposition of one large domain over NCCL (K6).
//
// The reference is single-GPU (cudaSetDevice(0), Deff2D.cuh:908); a 5-point stencil shards
// naturally into horizontal slabs.  One process (or thread) per GPU owns a contiguous band of
// rows plus H halo rows of each neighbour (deff2d_domain_load_slab).  A temporally blocked pass
// of depth T (sweep_tma.cu) leaves the own rows exact and makes T more halo rows stale, so the
// H boundary rows only travel to the neighbours when the next pass would need more exact halo
// rows than are left -- with H = 16 and T = 4 once per 4 passes:
//
//     [exchange: ncclGroup{Send/Recv up, Send/Recv down}] [pass] [pass] [pass] [pass] [exchange] ...
//
// Rows are contiguous (pitch doubles each), so the halo is sent straight from the iterate
// buffer: no pack kernel.  Runs of 16 passes with their exchanges are captured into one CUDA
// graph.  Once per check the two boundary-flux partial sums are all-reduced (ncclAllReduce,
// 2 doubles) and every rank applies the identical stop rule (cuh:1263-1276 via k_check).
// Measured and removed: exchanging after every pass on a second stream beside the interior tiles (boundary tiles in an
// own launch first) -- 1 200 vs 1 341 GLUP/s on 2 GPUs in round 1, the two-launch split costs more than the exchange it
// hides; overlapping only the one exchange per halo cycle the same way (round 2): 1 721 vs 1 716 GLUP/s on 2 GPUs -- no
// gain, the remaining scaling loss is the extra wave of tiles the halo rows add (83 instead of 82 rounds of 148 tiles on
// config 2) -- and it did not complete on 4 GPUs.
//
// Alternative without NCCL in the loop (peer mode, below: deff2d_slab_peer_export / _attach): the neighbours' buffers are
// mapped and the sweep kernel pushes the boundary rows itself; equal at 4 GPUs, 11 % faster at 8 (config 4).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that already loaded NCCL (e.g.
// through torch) shares that copy, and single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "context.h"

namespace deff2d {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommAbort) CommAbort = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

static void nccl_bind(NcclApi &api)
{
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    bool ok = true;
    auto sym = [&](const char *name) { void *p = dlsym(api.handle, name); if (!p) ok = false; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.CommAbort = (decltype(api.CommAbort))sym("ncclCommAbort");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { api.error = "libnccl.so.2 lacks a required symbol"; api.handle = nullptr; }
}

// bound once per process, also when several device threads ask at the same time (multi.cpp)
static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] { nccl_bind(api); });
    return &api;
}

struct SlabGraph {
    cudaGraphExec_t exec = nullptr;
    int T = 0, launches = 0;
    int64_t halo_after = 0;       // c->halo_valid at the end of the captured run
    const void *x0 = nullptr, *x1 = nullptr, *idx = nullptr, *lut = nullptr;
    int gather32 = 0;
    double omega = 0;
    int64_t Ny = 0, above = -1, below = -1;
    bool peer = false;
    const void *peer_up = nullptr, *peer_down = nullptr;
};

// Peer-memory halo exchange (the sweep kernel pushes the boundary rows into the neighbours' halo rows itself):
// what one rank publishes about itself, and what it has mapped of its neighbours.
struct PeerHandle {                  // DEFF2D_PEER_HANDLE_BYTES
    cudaIpcMemHandle_t x[2], sync;
    uint64_t ptr_x[2], ptr_sync;     // the raw device pointers (same-process neighbours use them directly)
    int64_t above, own, below, pitch, Nx;
    int32_t pid, device;
};
static_assert(sizeof(PeerHandle) <= DEFF2D_PEER_HANDLE_BYTES, "peer handle size");

struct PeerSide {
    bool present = false;
    PeerHandle h;                    // what is currently mapped
    double *x[2] = {nullptr, nullptr};
    long long *sync = nullptr;
    bool ipc = false;                // mapped with cudaIpcOpenMemHandle (must be closed), else same-process pointers
};

struct PeerState {
    bool active = false;
    long long *sync = nullptr;       // own block: [0] flag from above, [1] flag from below, [2] boundary tiles done,
                                     //            [3] CTAs done, [4] passes done
    PeerSide up, down;
    DevBuf<uint32_t> tiles;          // per depth T: one interior tile per CTA, the boundary tiles, the other interior tiles
    size_t off[9] = {0};
    int cnt[9] = {0}, nb[9] = {0}, lead[9] = {0};
    int64_t key_Nx = -1, key_Ny = -1, key_above = -1, key_own = -1;
};

struct SlabState {
    ncclComm_t comm = nullptr;
    PeerState peer;
    SlabGraph graph[2];           // one per parity of c->cur at the start of the run
    int rank = 0, nranks = 1;
};

#define NCCLCHECK(call)                                                                      \
    do {                                                                                     \
        ncclResult_t r_ = (call);                                                            \
        if (r_ != ncclSuccess) {                                                             \
            set_error(c, "%s failed: %s", #call, api->GetErrorString(r_));                   \
            return DEFF2D_ERR_NCCL;                                                          \
        }                                                                                    \
    } while (0)

#define CUS(call)                                                                            \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_error(c, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DEFF2D_ERR_CUDA;                                                          \
        }                                                                                    \
    } while (0)

// Halo exchange of the current iterate x[c->cur] on stream `cs`: H whole padded rows to and from
// each neighbour.  Afterwards every local row is exact again.
static int slab_exchange(deff2d_ctx *c, SlabState *s, NcclApi *api, double *buf, cudaStream_t cs)
{
    const int64_t H = std::max(c->halo_above, c->halo_below);
    const bool up = c->halo_above > 0, down = c->halo_below > 0;
    const size_t count = (size_t)H * (size_t)c->pitch;            // doubles per halo block
    NCCLCHECK(api->GroupStart());
    if (up) {
        // own top H rows -> upper neighbour's lower halo; its bottom H own rows -> my upper halo
        NCCLCHECK(api->Send(buf + (size_t)(1 + c->halo_above) * c->pitch, count, ncclDouble, s->rank - 1, s->comm, cs));
        NCCLCHECK(api->Recv(buf + (size_t)(1 + c->halo_above - H) * c->pitch, count, ncclDouble, s->rank - 1, s->comm, cs));
    }
    if (down) {
        const size_t last_own = (size_t)(1 + c->halo_above + c->own_rows);    // padded row after the last own row
        NCCLCHECK(api->Send(buf + (last_own - (size_t)H) * c->pitch, count, ncclDouble, s->rank + 1, s->comm, cs));
        NCCLCHECK(api->Recv(buf + last_own * c->pitch, count, ncclDouble, s->rank + 1, s->comm, cs));
    }
    NCCLCHECK(api->GroupEnd());
    c->launches++;                                                 // the NCCL send/recv kernel
    return DEFF2D_OK;
}

static void peer_close_side(PeerSide &sd)
{
    if (sd.present && sd.ipc) {
        for (int k = 0; k < 2; k++) if (sd.x[k]) cudaIpcCloseMemHandle(sd.x[k]);
        if (sd.sync) cudaIpcCloseMemHandle(sd.sync);
    }
    sd = PeerSide();
}

static int peer_open_side(deff2d_ctx *c, PeerSide &sd, const PeerHandle *h)
{
    if (!h) { peer_close_side(sd); return DEFF2D_OK; }
    if (sd.present && !std::memcmp(&sd.h, h, sizeof(*h))) return DEFF2D_OK;      // same buffers as last time: keep the mapping
    peer_close_side(sd);
    sd.h = *h;
    if (h->pid == (int32_t)getpid()) {
        // a neighbour in this process (multi.cpp: one thread per device): its pointers are valid here once peer access is on
        if (h->device != c->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(h->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                set_error(c, "cudaDeviceEnablePeerAccess(%d) failed: %s", h->device, cudaGetErrorString(e));
                (void)cudaGetLastError();
                return DEFF2D_ERR_CUDA;
            }
            (void)cudaGetLastError();
        }
        sd.x[0] = reinterpret_cast<double *>(h->ptr_x[0]);
        sd.x[1] = reinterpret_cast<double *>(h->ptr_x[1]);
        sd.sync = reinterpret_cast<long long *>(h->ptr_sync);
        sd.ipc = false;
    } else {
        void *p0 = nullptr, *p1 = nullptr, *ps = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p0, h->x[0], cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&p1, h->x[1], cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&ps, h->sync, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error(c, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
            if (p0) cudaIpcCloseMemHandle(p0);
            if (p1) cudaIpcCloseMemHandle(p1);
            return DEFF2D_ERR_CUDA;
        }
        sd.x[0] = static_cast<double *>(p0); sd.x[1] = static_cast<double *>(p1); sd.sync = static_cast<long long *>(ps);
        sd.ipc = true;
    }
    sd.present = true;
    return DEFF2D_OK;
}

// Tile lists of the peer mode for every depth: the tile grid covers the own rows only (the neighbours write the halo
// rows); tiles whose output box touches the first / last H own rows next to a neighbour (they feed the neighbours)
// come right after one interior tile per CTA, then the rest.
static int peer_build_lists(deff2d_ctx *c, PeerState &ps)
{
    if (ps.key_Nx == c->Nx && ps.key_Ny == c->Ny && ps.key_above == c->halo_above && ps.key_own == c->own_rows && ps.tiles.p) return DEFF2D_OK;
    const int64_t H = std::max(c->halo_above, c->halo_below);
    std::vector<uint32_t> all, bd, in;
    for (int T = 1; T <= 8; T++) {
        int ow, oh;
        tma_tile_geometry(c, T, &ow, &oh);
        // the tile grid covers the own rows only, tile row 0 starts at the first own row (the kernel adds `above`)
        const int tiles_x = (int)((c->Nx + ow - 1) / ow), tiles_y = (int)((c->own_rows + oh - 1) / oh);
        bd.clear(); in.clear();
        for (int ty = 0; ty < tiles_y; ty++) {
            const int64_t r0 = (int64_t)ty * oh, r1 = std::min<int64_t>(r0 + oh, c->own_rows);
            const bool b = (c->halo_above > 0 && r0 < H) || (c->halo_below > 0 && r1 > c->own_rows - H);
            for (int tx = 0; tx < tiles_x; tx++) (b ? bd : in).push_back(((uint32_t)ty << 16) | (uint32_t)tx);
        }
        // list: one interior tile per CTA (their loads need no neighbour: the flag wait hides behind them), the boundary
        // tiles, the other interior tiles
        const size_t ncta = (size_t)std::min<int64_t>(c->prop.multiProcessorCount, (int64_t)(bd.size() + in.size()));
        const size_t lead = (in.size() >= ncta) ? ncta : 0;
        ps.off[T] = all.size(); ps.nb[T] = (int)bd.size(); ps.cnt[T] = (int)(bd.size() + in.size()); ps.lead[T] = (int)lead;
        all.insert(all.end(), in.begin(), in.begin() + lead);
        all.insert(all.end(), bd.begin(), bd.end());
        all.insert(all.end(), in.begin() + lead, in.end());
    }
    if (ps.tiles.cap < all.size() || !ps.tiles.p) {
        if (ps.tiles.p) cudaFree(ps.tiles.p);
        ps.tiles.p = nullptr;
        CUS(cudaMalloc((void **)&ps.tiles.p, all.size() * sizeof(uint32_t)));
        ps.tiles.cap = all.size();
    }
    CUS(cudaMemcpyAsync(ps.tiles.p, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CUS(cudaStreamSynchronize(c->stream));
    ps.key_Nx = c->Nx; ps.key_Ny = c->Ny; ps.key_above = c->halo_above; ps.key_own = c->own_rows;
    return DEFF2D_OK;
}

// One pass of depth T in peer mode; flips c->cur.  No NCCL: the kernel pushes and waits itself (sweep_tma.cu, PEER).
static int peer_pass(deff2d_ctx *c, SlabState *s, int T)
{
    PeerState &ps = s->peer;
    const int64_t H = std::max(c->halo_above, c->halo_below);
    PeerArgs pa;
    std::memset(&pa, 0, sizeof(pa));
    for (int b = 0; b < 2; b++) {
        // the neighbour's halo rows that mirror my first / last H own rows: its interior rows [above_p + own_p, +H) / [above_q - H, above_q)
        if (ps.up.present) pa.up[b] = ps.up.x[b] + (1 + ps.up.h.above + ps.up.h.own) * ps.up.h.pitch + DEFF2D_XOFF;
        if (ps.down.present) pa.down[b] = ps.down.x[b] + (1 + ps.down.h.above - H) * ps.down.h.pitch + DEFF2D_XOFF;
    }
    pa.flag_up = ps.up.present ? ps.up.sync + 1 : nullptr;         // I am the lower neighbour of the rank above
    pa.flag_down = ps.down.present ? ps.down.sync + 0 : nullptr;   // ... and the upper neighbour of the rank below
    pa.flag_local = ps.sync;
    pa.counters = reinterpret_cast<unsigned long long *>(ps.sync + 2);
    pa.pass_no = ps.sync + 4;
    pa.pitch = c->pitch;
    pa.above = (int)c->halo_above; pa.own = (int)c->own_rows; pa.H = (int)H; pa.nboundary = ps.nb[T]; pa.Nx = (int)c->Nx;
    pa.lead = ps.lead[T];
    c->store_row0 = c->halo_above; c->store_rows = c->own_rows;
    const int rc = tma_peer_pass(c, T, ps.tiles.p + ps.off[T], ps.cnt[T], &pa);
    c->store_row0 = 0; c->store_rows = 0;
    return rc;
}

// One pass of depth T on a slab; flips c->cur.  Enqueue only (also used under stream capture).  A pass of depth T
// invalidates T more halo rows, so with H halo rows the exchange is only needed every H / T passes -- c->halo_valid
// counts the halo rows that are still exact.  With H = 32, T = 6 the NCCL latency (~25 us per exchange, measured) is
// paid once per 5 passes (~1.1 ms of sweeping on config 2).
static int slab_pass(deff2d_ctx *c, SlabState *s, NcclApi *api, int T)
{
    const bool comm = (c->halo_above > 0 || c->halo_below > 0);
    int rc;
    if (s->peer.active && comm) return peer_pass(c, s, T);
    if (comm && c->halo_valid < T) {
        if ((rc = slab_exchange(c, s, api, c->x[c->cur].p, c->stream))) return rc;
        c->halo_valid = std::max(c->halo_above, c->halo_below);
    }
    if ((rc = tma_pass(c, T, nullptr, 0, c->stream))) return rc;
    c->cur ^= 1;
    if (comm) c->halo_valid -= T;
    return DEFF2D_OK;
}

#define SLAB_GRAPH_PASSES 16

// n sweeps on a slab: passes of depth T = min(tblock, halo rows) with a halo exchange after each.
// Runs of SLAB_GRAPH_PASSES passes -- kernels, events and the NCCL send/recv pairs -- are captured
// once into a CUDA graph and replayed: per pass the host otherwise issues two launches, four
// event operations and an NCCL group (measured: the host then falls behind the GPU).  Enqueue only.
int slab_enqueue_sweeps(deff2d_ctx *c, int64_t n)
{
    SlabState *s = static_cast<SlabState *>(c->slab);
    NcclApi *api = nccl_api();
    if (!s || !s->comm) { set_error(c, "slab sweeps before deff2d_nccl_init"); return DEFF2D_ERR_STATE; }
    const int64_t H = std::max(c->halo_above, c->halo_below);
    if (s->nranks > 1 && (H < 1 || c->own_rows < H)) { set_error(c, "slab needs halo rows >= 1 and own rows >= halo rows"); return DEFF2D_ERR_STATE; }
    const int old_family = c->tile_family;
    c->tile_family = 0;                    // the default thread layout
    int rc = DEFF2D_OK;
    int Tmax = c->tblock > 0 ? c->tblock : 4;
    if (c->kernel == 0) Tmax = c->k2_default_depth;   // default depth; with 32 halo rows an exchange every fifth pass
    if (Tmax > 8) Tmax = 8;
    if (s->nranks > 1 && Tmax > H) Tmax = (int)H;
    while (n > 0 && !rc) {
        const int T = (int)std::min<int64_t>(n, Tmax);
        if (c->use_graphs && n >= (int64_t)T * SLAB_GRAPH_PASSES) {
            SlabGraph &g = s->graph[c->cur];
            const bool valid = g.exec && g.T == T && g.x0 == c->x[0].p && g.x1 == c->x[1].p && g.idx == c->idx16.p &&
                               g.lut == (const void *)c->clut32.p && g.gather32 == c->gather32 && g.omega == c->omega && g.Ny == c->Ny && g.above == c->halo_above && g.below == c->halo_below && g.peer == s->peer.active &&
                               g.peer_up == (const void *)s->peer.up.x[0] && g.peer_down == (const void *)s->peer.down.x[0];
            if (!valid) {
                if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
                // one direct pass pair first: encodes the tensor maps outside the capture
                if ((rc = slab_pass(c, s, api, T)) || (rc = slab_pass(c, s, api, T))) break;
                n -= 2 * T;
                if (n < (int64_t)T * SLAB_GRAPH_PASSES) continue;
                cudaGraph_t graph = nullptr;
                const int64_t launches0 = c->launches;
                cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
                if (e != cudaSuccess) { set_error(c, "cudaStreamBeginCapture failed: %s", cudaGetErrorString(e)); rc = DEFF2D_ERR_CUDA; break; }
                c->halo_valid = 0;                                 // the captured run starts with an exchange: replays do not depend on the state before
                for (int k = 0; k < SLAB_GRAPH_PASSES && !rc; k++) rc = slab_pass(c, s, api, T);
                g.halo_after = c->halo_valid;
                e = cudaStreamEndCapture(c->stream, &graph);
                g.launches = (int)(c->launches - launches0);
                c->launches = launches0;                           // nothing ran yet
                if (rc) { if (graph) cudaGraphDestroy(graph); break; }
                if (e != cudaSuccess || !graph) { set_error(c, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); rc = DEFF2D_ERR_CUDA; break; }
                e = cudaGraphInstantiate(&g.exec, graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) { g.exec = nullptr; set_error(c, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); rc = DEFF2D_ERR_CUDA; break; }
                g.T = T; g.x0 = c->x[0].p; g.x1 = c->x[1].p; g.idx = c->idx16.p; g.lut = c->clut32.p; g.gather32 = c->gather32; g.omega = c->omega;
                g.Ny = c->Ny; g.above = c->halo_above; g.below = c->halo_below; g.peer = s->peer.active;
                g.peer_up = s->peer.up.x[0]; g.peer_down = s->peer.down.x[0];
            }
            cudaError_t e = cudaGraphLaunch(s->graph[c->cur].exec, c->stream);
            if (e != cudaSuccess) { set_error(c, "cudaGraphLaunch failed: %s", cudaGetErrorString(e)); rc = DEFF2D_ERR_CUDA; break; }
            c->launches += s->graph[c->cur].launches;
            c->halo_valid = s->graph[c->cur].halo_after;
            n -= (int64_t)T * SLAB_GRAPH_PASSES;                   // an even number of passes: c->cur unchanged
            continue;
        }
        rc = slab_pass(c, s, api, T);
        n -= T;
    }
    c->tile_family = old_family;
    if (rc) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(c, "slab sweep launch failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
    return DEFF2D_OK;
}

// {Q1, Q2} of this slab's rows -> global sums on every rank (in place in the device state).
int slab_allreduce_q(deff2d_ctx *c)
{
    SlabState *s = static_cast<SlabState *>(c->slab);
    if (!s || !s->comm || s->nranks < 2 || !c->slab_domain) return DEFF2D_OK;
    NcclApi *api = nccl_api();
    NCCLCHECK(api->AllReduce(c->d_state->q, c->d_state->q, 2, ncclDouble, ncclSum, s->comm, c->stream));
    c->launches++;
    return DEFF2D_OK;
}

// a new domain is loaded: the peer mode has to be attached again (flags, counters and tile lists belong to a load)
void slab_peer_reset(deff2d_ctx *c)
{
    if (SlabState *s = static_cast<SlabState *>(c->slab)) s->peer.active = false;
}

bool slab_is_distributed(const deff2d_ctx *c)
{
    const SlabState *s = static_cast<const SlabState *>(c->slab);
    return s && s->comm && s->nranks >= 2 && c->slab_domain;
}

void slab_destroy(deff2d_ctx *c)
{
    SlabState *s = static_cast<SlabState *>(c->slab);
    if (!s) return;
    NcclApi *api = nccl_api();
    for (auto &g : s->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    peer_close_side(s->peer.up);
    peer_close_side(s->peer.down);
    if (s->peer.sync) cudaFree(s->peer.sync);
    if (s->peer.tiles.p) cudaFree(s->peer.tiles.p);
    if (s->comm && api->CommDestroy) api->CommDestroy(s->comm);
    delete s;
    c->slab = nullptr;
}

}  // namespace deff2d

using namespace deff2d;

DEFF2D_EXPORT int deff2d_nccl_unique_id(uint8_t id[DEFF2D_NCCL_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == DEFF2D_NCCL_ID_BYTES, "ncclUniqueId size");
    if (!id) return DEFF2D_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api->handle) { set_error(nullptr, "%s", api->error.c_str()); return DEFF2D_ERR_NCCL; }
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) { set_error(nullptr, "ncclGetUniqueId failed"); return DEFF2D_ERR_NCCL; }
    std::memcpy(id, &u, sizeof(u));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_nccl_init(deff2d_ctx *c, const uint8_t id[DEFF2D_NCCL_ID_BYTES], int rank, int nranks)
{
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return DEFF2D_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api->handle) { set_error(c, "%s", api->error.c_str()); return DEFF2D_ERR_NCCL; }
    slab_destroy(c);
    SlabState *s = new SlabState();
    c->slab = s;
    s->rank = rank; s->nranks = nranks;
    CUS(cudaSetDevice(c->device));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    NCCLCHECK(api->CommInitRank(&s->comm, nranks, u, rank));
    return DEFF2D_OK;
}

// Abort this context's communicator (another rank of the group failed): pending and future NCCL work on it returns
// an error instead of waiting for a peer that will never arrive.  The communicator is gone afterwards.
DEFF2D_EXPORT int deff2d_slab_abort(deff2d_ctx *c)
{
    if (!c) return DEFF2D_ERR_ARG;
    SlabState *s = static_cast<SlabState *>(c->slab);
    NcclApi *api = nccl_api();
    if (!s || !s->comm || !api->CommAbort) return DEFF2D_OK;
    ncclComm_t comm = s->comm;
    s->comm = nullptr;
    api->CommAbort(comm);
    return DEFF2D_OK;
}

// ---- peer-memory halo exchange -----------------------------------------------------------------------------------

DEFF2D_EXPORT int deff2d_slab_peer_export(deff2d_ctx *c, uint8_t handle[DEFF2D_PEER_HANDLE_BYTES])
{
    if (!c || !handle) return DEFF2D_ERR_ARG;
    SlabState *s = static_cast<SlabState *>(c->slab);
    if (!s || !s->comm) { set_error(c, "peer export before deff2d_nccl_init"); return DEFF2D_ERR_STATE; }
    if (!c->loaded || !c->slab_domain) { set_error(c, "peer export needs a loaded slab"); return DEFF2D_ERR_STATE; }
    CUS(cudaSetDevice(c->device));
    PeerState &ps = s->peer;
    if (!ps.sync) CUS(cudaMalloc((void **)&ps.sync, 64));
    PeerHandle h;
    std::memset(&h, 0, sizeof(h));
    CUS(cudaIpcGetMemHandle(&h.x[0], c->x[0].p));
    CUS(cudaIpcGetMemHandle(&h.x[1], c->x[1].p));
    CUS(cudaIpcGetMemHandle(&h.sync, ps.sync));
    h.ptr_x[0] = (uint64_t)(uintptr_t)c->x[0].p; h.ptr_x[1] = (uint64_t)(uintptr_t)c->x[1].p; h.ptr_sync = (uint64_t)(uintptr_t)ps.sync;
    h.above = c->halo_above; h.own = c->own_rows; h.below = c->halo_below; h.pitch = c->pitch; h.Nx = c->Nx;
    h.pid = (int32_t)getpid(); h.device = c->device;
    std::memset(handle, 0, DEFF2D_PEER_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_slab_peer_attach(deff2d_ctx *c, const uint8_t *above, const uint8_t *below)
{
    if (!c) return DEFF2D_ERR_ARG;
    SlabState *s = static_cast<SlabState *>(c->slab);
    NcclApi *api = nccl_api();
    if (!s || !s->comm) { set_error(c, "peer attach before deff2d_nccl_init"); return DEFF2D_ERR_STATE; }
    if (!c->loaded || !c->slab_domain || !s->peer.sync) { set_error(c, "peer attach needs deff2d_slab_peer_export on a loaded slab first"); return DEFF2D_ERR_STATE; }
    CUS(cudaSetDevice(c->device));
    PeerState &ps = s->peer;
    ps.active = false;
    const int64_t H = std::max(c->halo_above, c->halo_below);
    if ((c->halo_above > 0) != (above != nullptr) || (c->halo_below > 0) != (below != nullptr)) {
        set_error(c, "peer attach: a handle is needed exactly where the slab has a neighbour");
        return DEFF2D_ERR_ARG;
    }
    // From here on every rank reaches the all-reduce below, also one whose slab is too thin or whose mapping failed: it
    // reports the failure through the reduced word, and all ranks return an error together (the caller then stays with
    // the NCCL exchange) instead of the healthy ones waiting for it in the barrier.
    PeerHandle hu, hd;
    int rc = DEFF2D_OK;
    if (H < 1 || H > 64 || c->own_rows < 2 * H) { set_error(c, "peer mode needs 1..64 halo rows and at least twice as many own rows"); rc = DEFF2D_ERR_ARG; }
    if (above && !rc) { std::memcpy(&hu, above, sizeof(hu)); if (hu.Nx != c->Nx || hu.pitch != c->pitch || hu.below != H) { set_error(c, "peer attach: the upper neighbour's slab does not match"); rc = DEFF2D_ERR_ARG; } }
    if (below && !rc) { std::memcpy(&hd, below, sizeof(hd)); if (hd.Nx != c->Nx || hd.pitch != c->pitch || hd.above != H) { set_error(c, "peer attach: the lower neighbour's slab does not match"); rc = DEFF2D_ERR_ARG; } }
    if (!rc) rc = peer_open_side(c, ps.up, above ? &hu : nullptr);
    if (!rc) rc = peer_open_side(c, ps.down, below ? &hd : nullptr);
    if (!rc) rc = peer_build_lists(c, ps);
    // flags -1 ("no pass yet"), counters 0, pass 0, [6] = "this rank failed" -- then every rank must have got here before
    // anyone pushes a row into a neighbour (whose load may still be writing its buffers): one all-reduce as a barrier
    const long long init[8] = {-1, -1, 0, 0, 0, 0, rc ? 1 : 0, 0};
    long long failed = 0;
    CUS(cudaMemcpyAsync(ps.sync, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    NCCLCHECK(api->AllReduce(ps.sync + 6, ps.sync + 6, 1, ncclInt64, ncclSum, s->comm, c->stream));
    CUS(cudaMemcpyAsync(&failed, ps.sync + 6, sizeof(failed), cudaMemcpyDeviceToHost, c->stream));
    CUS(cudaStreamSynchronize(c->stream));
    c->launches++;
    if (rc) return rc;
    if (failed) { set_error(c, "peer attach: %lld rank(s) of the group could not map their neighbours", failed); return DEFF2D_ERR_STATE; }
    ps.active = true;
    for (auto &g : s->graph) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_slab_peer_detach(deff2d_ctx *c)
{
    if (!c) return DEFF2D_ERR_ARG;
    SlabState *s = static_cast<SlabState *>(c->slab);
    if (!s) return DEFF2D_OK;
    s->peer.active = false;                       // back to the NCCL exchange; the mappings stay for the next attach
    c->halo_valid = 0;
    for (auto &g : s->graph) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_slab_sweeps(deff2d_ctx *c, int64_t n)
{
    if (!c || n < 0) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    if (cudaSetDevice(c->device) != cudaSuccess) return DEFF2D_ERR_CUDA;
    return slab_enqueue_sweeps(c, n);
}

DEFF2D_EXPORT int deff2d_slab_flux(deff2d_ctx *c, double *deff_raw)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CUS(cudaSetDevice(c->device));
    launch_flux(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->own_first, c->own_rows, c->d_state);
    c->launches++;
    int rc = slab_allreduce_q(c);
    if (rc) return rc;
    CUS(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
    CUS(cudaStreamSynchronize(c->stream));
    if (deff_raw) {
        const double qAvg = (c->h_state->q[0] + c->h_state->q[1]) / (2.0 * (double)c->NyG);   // cuh:1263
        *deff_raw = qAvg / (c->CR - c->CL);                                                  // cuh:1264
    }
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_tile_geometry(int T, int *ow, int *oh, int *tw, int *th)
{
    if (T < 1 || T > 8) return DEFF2D_ERR_ARG;
    int a = 0, b = 0;
    tma_tile_geometry(nullptr, T, &a, &b);
    if (ow) *ow = a;
    if (oh) *oh = b;
    if (tw) *tw = a + 2 * ((T + 1) & ~1);
    if (th) *th = b + 2 * T;
    return DEFF2D_OK;
}
