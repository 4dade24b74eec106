// batch.cu -- K5: packed batch mode.  Many small images are solved per launch by packing them
// into one resident "stack" that the tiled sweep kernel (sweep_tma.cu) walks like a single
// large domain:
//
//   * the stack is a GX x GY grid of slots; slot (gx, gy) holds one image with its interior
//     origin at column gx*(Nx+1), row gy*(Ny+1).  Neighbouring slots share one ghost column
//     (Dirichlet value 1.0, the CL / CR factor lives in the face weight, tables.cpp) and one
//     ghost row (no-flux wall, weight 0), so images never interact: the packed sweep is
//     exactly the reference's per-image sweep (BatchSim / BatchSim3Phase bodies,
//     Deff2D.cuh:1867-2049, 2056-2419), image by image.
//   * every image follows its own copy of the reference loop (cuh:1232-1290): its own sweep
//     counter, checks at its own sweeps 1, 10 001, ..., its own stop rule and its own
//     continuation stage (3-phase pre-conditioning, cuh:1492-1597).  The stage is carried in
//     bits 3-7 of the cell code and selects one of the per-stage weight tables, so images in
//     different stages share a launch.
//   * the host enqueues sweeps up to the next event of any active image (a check or MaxIter),
//     then one k_batch_check launch evaluates boundary flux + stop rule for all images at once
//     and the host reads back 40 bytes per slot.  Tiles whose output touches no active image
//     are dropped from the tile list, finished slots are refilled from the queue.
//
// Nothing here allocates per image: the stack, tables and staging buffers live in the
// context's grow-only arenas.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "context.h"

namespace deff2d {

#define XOFF DEFF2D_XOFF
#define BATCH_MAX_STAGES DEFF2D_MAX_STAGES

struct BatchStage {
    double D[3];            // fluid, solid, gas of the stage
    double tol;
    long long max_iter;
    int precond;            // JacobiGPUPreCond stage: result not recorded as Deff (cuh:1144-1159)
    int pad;
};

struct BatchSlot {
    double deff_old, deff_new, change, conv;   // cuh:1171-1173, 1275
    long long iter;         // iterCount of the current stage
    int stage;
    int status;             // 0 empty, 1 active, 2 finished
    int image;
    int nchecks;
};

struct BatchOut {
    long long iters[BATCH_MAX_STAGES];
    double stage_deff_raw[BATCH_MAX_STAGES];
    double deff_raw, conv;
    unsigned long long phase[3], below150;
    int nstages, pad;
};

struct BatchJob { int slot, image, src, pad; };   // src: position of the image in the staging buffers of this refill

struct BatchGeom {
    int W, H, amp_x, amp_y, Nx, Ny, GX, nphase;
    long long pitch;
    double CL, CR;
    int ce, nstages;
};

struct BatchStages { BatchStage s[BATCH_MAX_STAGES]; };

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_ull(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Threshold + mesh amplification + ghost ring + x0 of every (slot, image) job: the packed
// counterpart of k_init_domain (cuh:1773-1785, 1557-1578, 1730-1734, 750).
__global__ void __launch_bounds__(256)
k_batch_init(BatchGeom g, const BatchJob *__restrict__ jobs, const uint8_t *__restrict__ img_base,
             const uint8_t *__restrict__ grid_base, double *__restrict__ xc, double *__restrict__ xo,
             uint8_t *__restrict__ code, BatchSlot *slots, BatchOut *outs)
{
    const BatchJob job = jobs[blockIdx.y];
    const int gx = job.slot % g.GX, gy = job.slot / g.GX;
    const long long col0 = (long long)gx * (g.Nx + 1), row0 = (long long)gy * (g.Ny + 1);
    const uint8_t *img = img_base + (size_t)job.src * g.W * g.H;
    const uint8_t *grid = grid_base ? grid_base + (size_t)job.src * g.Nx * g.Ny : nullptr;
    const long long wp = g.Nx + 2, total = (long long)(g.Ny + 2) * wp;
    unsigned long long cnt[3] = {0, 0, 0}, below = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const long long r = k / wp;
        const long long i = r - 1, j = (k - r * wp) - 1;
        const long long idx = (row0 + r) * g.pitch + col0 + j + XOFF;
        unsigned cd = DEFF2D_PHASE_GHOST;
        double v0 = 0.0, v1 = 0.0;
        if (j == -1 || j == g.Nx) { v0 = 1.0; v1 = 1.0; }                 // Dirichlet ghost column
        else if (i >= 0 && i < g.Ny) {
            const unsigned char p = img[(i / g.amp_y) * g.W + (int)(j / g.amp_x)];      // cuh:1778
            if (g.nphase == 2) cd = (p < 150) ? DEFF2D_PHASE_FLUID : DEFF2D_PHASE_SOLID;   // cuh:1779
            else cd = (p > 200) ? DEFF2D_PHASE_SOLID : ((p < 50) ? DEFF2D_PHASE_GAS : DEFF2D_PHASE_FLUID);   // cuh:1565-1576
            cnt[cd]++;
            if (grid) {
                const unsigned char gv = grid[i * g.Nx + j];
                if (gv == 1 || gv == 2) cd |= DEFF2D_CODE_PINNED;                      // cuh:750
            }
            v0 = __dadd_rn(__dmul_rn(__ddiv_rn((double)j, (double)g.Nx), __dsub_rn(g.CR, g.CL)), g.CL);   // cuh:1732
        }
        code[idx] = (uint8_t)cd;
        xc[idx] = v0;
        xo[idx] = v1;
    }
    // calcPorosity's count (cuh:399-405) over the source pixels
    const long long npix = (long long)g.W * g.H;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < npix; k += (long long)gridDim.x * blockDim.x)
        below += (img[k] < 150) ? 1u : 0u;
    BatchOut *o = outs + job.slot;              // zeroed by k_batch_clear
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const unsigned long long s = warp_sum_ull(cnt[k]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&o->phase[k], s);
    }
    const unsigned long long sb = warp_sum_ull(below);
    if ((threadIdx.x & 31) == 0 && sb) atomicAdd(&o->below150, sb);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        BatchSlot s;
        s.deff_old = 5; s.deff_new = 1; s.change = 100.0; s.conv = 0;          // cuh:1171-1173
        s.iter = 0; s.stage = 0; s.status = 1; s.image = job.image; s.nchecks = 0;
        slots[job.slot] = s;
    }
}

// the result record of every slot that is about to receive a new image
__global__ void k_batch_clear(const BatchJob *__restrict__ jobs, int njobs, BatchOut *outs)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= njobs) return;
    BatchOut z;
    memset(&z, 0, sizeof(z));
    outs[jobs[k].slot] = z;
}

// After `nsweeps` more sweeps of every active slot: the reference's check (boundary flux,
// Deff, signed change, cuh:1243-1276), its loop condition (cuh:1232) and the drivers' stage
// sequence (cuh:1492-1597), for all slots of `active` in one launch.  The flux sums use the same
// thread-to-row assignment and reduction order as k_flux (kernels.cu), so a packed image gets
// bit-identical Deff values to a single-image solve.
__global__ void __launch_bounds__(1024)
k_batch_check(BatchGeom g, BatchStages stages, BatchSlot *slots, BatchOut *outs, const int *__restrict__ active,
              long long nsweeps, const double *__restrict__ x, uint8_t *code, uint16_t *idx16)
{
    __shared__ double s1[32], s2[32];
    __shared__ int sh_restage;
    const int slot = active[blockIdx.x];
    BatchSlot S = slots[slot];
    if (S.status != 1) return;
    const long long it = S.iter + nsweeps;
    const BatchStage st = stages.s[S.stage];
    const bool do_check = ((it - 1) % g.ce == 0);                   // cuh:1243 on the pre-increment counter
    const int gx = slot % g.GX, gy = slot / g.GX;
    const long long col0 = (long long)gx * (g.Nx + 1), row0 = (long long)gy * (g.Ny + 1);
    if (threadIdx.x == 0) sh_restage = -1;
    double q1 = 0, q2 = 0;
    if (do_check) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        const double half_dx = (1.0 / (double)g.Nx) / 2.0;          // dx / 2.0, cuh:1256
        for (long long r = threadIdx.x; r < g.Ny; r += blockDim.x) {
            const long long base = (row0 + r + 1) * g.pitch + col0 + XOFF;
            const unsigned cl = code[base], cr = code[base + g.Nx - 1];
            const double Dl = st.D[cl & 3u], Dr = st.D[cr & 3u];
            double tl = Dl * (x[base] - g.CL) / half_dx;
            double tr = Dr * (g.CR - x[base + g.Nx - 1]) / half_dx;
            if (Dl == 0.0 && !(cl & 4u)) tl = nan;                  // quirk Q13, see k_flux
            if (Dr == 0.0 && !(cr & 4u)) tr = nan;
            q1 += tl;
            q2 += tr;
        }
        q1 = warp_sum_d(q1);
        q2 = warp_sum_d(q2);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) { s1[wid] = q1; s2[wid] = q2; }
        __syncthreads();
        if (wid == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            q1 = lane < nw ? s1[lane] : 0.0;
            q2 = lane < nw ? s2[lane] : 0.0;
            q1 = warp_sum_d(q1);
            q2 = warp_sum_d(q2);
        }
    }
    if (threadIdx.x == 0) {
        bool stop = false;
        if (do_check) {
            const double qAvg = (q1 + q2) / (2.0 * (double)g.Ny);           // cuh:1263
            const double deffNew = qAvg / (g.CR - g.CL);                    // cuh:1264
            const double change = (S.deff_old - deffNew) / (S.deff_old);    // cuh:1265
            S.deff_new = deffNew;
            S.change = change;
            S.conv = change;                                                // cuh:1275
            S.deff_old = deffNew;                                           // cuh:1273
            S.nchecks++;
            stop = !(st.tol < fabs(change));                                // cuh:1232 (NaN ends the loop)
        }
        S.iter = it;
        if (stop || it >= st.max_iter) {
            BatchOut *o = outs + slot;
            o->iters[S.stage] = it;
            o->stage_deff_raw[S.stage] = S.deff_new;
            if (!st.precond) { o->deff_raw = S.deff_new; o->conv = S.conv; }   // cuh:1309-1311 vs cuh:1144-1159
            o->nstages = S.stage + 1;
            if (S.stage + 1 < g.nstages) {
                S.stage++;
                S.iter = 0; S.deff_old = 5; S.deff_new = 1; S.change = 100.0; S.nchecks = 0;
                sh_restage = S.stage;
            } else {
                S.status = 2;
            }
        }
        slots[slot] = S;
    }
    __syncthreads();
    const int rs = sh_restage;
    if (rs >= 0) {      // next continuation stage: only the weight table changes (cuh:1524), selected by code bits 3-7
        const long long total = (long long)g.Ny * g.Nx;
        for (long long k = threadIdx.x; k < total; k += blockDim.x) {
            const long long i = k / g.Nx, j = k - i * g.Nx;
            const long long idx = (row0 + i + 1) * g.pitch + col0 + j + XOFF;
            code[idx] = (uint8_t)((code[idx] & 7u) | ((unsigned)rs << 3));
            idx16[idx] = (uint16_t)((idx16[idx] & 0xc3ffu) | ((unsigned)rs << 10));
        }
    }
}

// ------------------------------------------------------------------------------------------ host side

struct BatchState {
    DevBuf<BatchSlot> slots;
    DevBuf<BatchOut> outs;
    DevBuf<BatchJob> jobs;
    DevBuf<int> active;
    DevBuf<uint32_t> tiles;
    DevBuf<int> ff_flags;
    int *h_ff_flags = nullptr;
    size_t h_ff_flags_cap = 0;
    uint8_t *h_in = nullptr;              // pinned staging of the images of one refill
    size_t h_in_cap = 0;
    BatchOut *h_out = nullptr;            // pinned: the record of one finished slot
    size_t h_out_cap = 0;
    BatchSlot *h_slots = nullptr;
    BatchJob *h_jobs = nullptr;
    int *h_active = nullptr;
    uint32_t *h_tiles = nullptr;
    size_t h_slots_cap = 0, h_jobs_cap = 0, h_active_cap = 0, h_tiles_cap = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

template <typename T>
static int dev_ensure(deff2d_ctx *c, DevBuf<T> &b, size_t n)
{
    if (b.cap >= n && b.p) return DEFF2D_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    const size_t want = n + n / 8 + 64;
    if (cudaMalloc((void **)&b.p, want * sizeof(T)) != cudaSuccess) {
        b.p = nullptr;
        (void)cudaGetLastError();
        set_error(c, "cudaMalloc of %zu bytes failed", want * sizeof(T));
        return DEFF2D_ERR_ALLOC;
    }
    b.cap = want;
    return DEFF2D_OK;
}

template <typename T>
static int host_ensure(deff2d_ctx *c, T *&p, size_t &cap, size_t n)
{
    if (cap >= n && p) return DEFF2D_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 8 + 64;
    if (cudaMallocHost((void **)&p, want * sizeof(T)) != cudaSuccess) {
        p = nullptr;
        (void)cudaGetLastError();
        set_error(c, "cudaMallocHost of %zu bytes failed", want * sizeof(T));
        return DEFF2D_ERR_ALLOC;
    }
    cap = want;
    return DEFF2D_OK;
}

void batch_destroy(deff2d_ctx *c)
{
    BatchState *b = static_cast<BatchState *>(c->batch);
    if (!b) return;
    if (b->slots.p) cudaFree(b->slots.p);
    if (b->outs.p) cudaFree(b->outs.p);
    if (b->jobs.p) cudaFree(b->jobs.p);
    if (b->active.p) cudaFree(b->active.p);
    if (b->tiles.p) cudaFree(b->tiles.p);
    if (b->ff_flags.p) cudaFree(b->ff_flags.p);
    if (b->h_ff_flags) cudaFreeHost(b->h_ff_flags);
    if (b->h_in) cudaFreeHost(b->h_in);
    if (b->h_out) cudaFreeHost(b->h_out);
    if (b->h_slots) cudaFreeHost(b->h_slots);
    if (b->h_jobs) cudaFreeHost(b->h_jobs);
    if (b->h_active) cudaFreeHost(b->h_active);
    if (b->h_tiles) cudaFreeHost(b->h_tiles);
    if (b->e0) cudaEventDestroy(b->e0);
    if (b->e1) cudaEventDestroy(b->e1);
    delete b;
    c->batch = nullptr;
}

#define CUB(call)                                                                            \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_error(c, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DEFF2D_ERR_CUDA;                                                          \
        }                                                                                    \
    } while (0)

// The stage sequence of the reference drivers for one image (host.cpp: stage_list) in the form k_batch_check reads.
static int build_stages(const deff2d_params *p, BatchStages *st, double *stageD, int *nst)
{
    if (p->mode != DEFF2D_MODE_2PH_BATCH && p->mode != DEFF2D_MODE_3PH) return 1;
    StageSpec spec[BATCH_MAX_STAGES];
    const int n = stage_list(p, spec, BATCH_MAX_STAGES);
    if (n < 1) return 1;
    for (int k = 0; k < n; k++) {
        st->s[k].D[0] = spec[k].Df; st->s[k].D[1] = spec[k].Ds; st->s[k].D[2] = spec[k].Dg;
        st->s[k].tol = spec[k].tol; st->s[k].max_iter = spec[k].max_iter; st->s[k].precond = spec[k].precond; st->s[k].pad = 0;
        stageD[k] = spec[k].stageD;
        // the packed loop needs at least one sweep per stage (cuh:1232)
        if (!(spec[k].tol < 100.0) || spec[k].max_iter < 1) return 1;
    }
    *nst = n;
    return 0;
}

// Slot grid for `count` images of Nx x Ny cells: rows of slots about 4096 cells wide, at most
// ~48 M cells resident (2 x 8 + 1 bytes per cell).
void batch_plan(int64_t Nx, int64_t Ny, int count, int limit, int *GX, int *GY)
{
    int gx = (int)std::max<int64_t>(1, 4096 / (Nx + 1));
    int64_t max_slots = std::max<int64_t>(1, ((int64_t)48 << 20) / (Nx * Ny));
    if (max_slots > 65535) max_slots = 65535;            // one k_batch_init launch covers every slot (gridDim.y)
    if (limit > 0 && limit < max_slots) max_slots = limit;
    int slots = (int)std::min<int64_t>(count, max_slots);
    if (gx > slots) gx = slots;
    int gy = (slots + gx - 1) / gx;
    *GX = gx;
    *GY = gy;
}

// Tiles (output box ow x oh) whose output intersects the interior of an active slot.
void batch_tile_list(int64_t Nx, int64_t Ny, int GX, const int *active, int nactive, int ow, int oh, int tiles_x,
                     int tiles_y, std::vector<uint8_t> &mark, std::vector<uint32_t> &out)
{
    mark.assign((size_t)tiles_x * tiles_y, 0);
    for (int a = 0; a < nactive; a++) {
        const int gx = active[a] % GX, gy = active[a] / GX;
        const int64_t c0 = (int64_t)gx * (Nx + 1), r0 = (int64_t)gy * (Ny + 1);
        const int tx0 = (int)(c0 / ow), tx1 = (int)((c0 + Nx - 1) / ow);
        const int ty0 = (int)(r0 / oh), ty1 = (int)((r0 + Ny - 1) / oh);
        for (int ty = ty0; ty <= ty1 && ty < tiles_y; ty++)
            for (int tx = tx0; tx <= tx1 && tx < tiles_x; tx++) mark[(size_t)ty * tiles_x + tx] = 1;
    }
    out.clear();
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++)
            if (mark[(size_t)ty * tiles_x + tx]) out.push_back(((uint32_t)ty << 16) | (uint32_t)tx);
}

// The packed batch solve as a stream: images are pulled through `fetch` when a slot is free and every finished image
// is handed to `done` at once -- the caller can decode ahead on its own threads while the GPU sweeps, and write each
// result row when it exists (the reference keeps all rows until the end, cuh:2051).
//   fetch(user, k, dst, wait): W*H pixels of image k into dst.  0 ok; 1 not ready yet (only when wait == 0: the
//       solve goes on with the slots it has); 2 no image k and none after it (the batch ends at k); < 0 error.
//   done(user, k, result, field): image k is finished; field: NULL or its concentration map.  Non-zero aborts.
static int batch_stream(deff2d_ctx *c, int count, int W, int H, const deff2d_params *p, deff2d_batch_fetch_fn fetch,
                        deff2d_batch_done_fn done, void *user, bool want_fields, const BatchStages &stages,
                        const double *stageD, int nstages, int *solved)
{
    CUB(cudaSetDevice(c->device));                       // before anything is created: the caller's thread may sit on another device
    BatchState *b = static_cast<BatchState *>(c->batch);
    if (!b) {
        b = new BatchState();
        c->batch = b;
        CUB(cudaEventCreate(&b->e0));
        CUB(cudaEventCreate(&b->e1));
    }
    const int nphase = (p->mode == DEFF2D_MODE_3PH) ? 3 : 2;
    const int64_t Nx = (int64_t)W * p->amp_x, Ny = (int64_t)H * p->amp_y;
    const int64_t cells = Nx * Ny;
    const size_t npix = (size_t)W * H;
    int GX, GY;
    batch_plan(Nx, Ny, count, c->batch_max_slots, &GX, &GY);
    const int nslots = GX * GY;
    const int T = 6;                                     // sweeps per HBM pass: interface-rich small images amortise the
                                                         // per-tile weight gather best at depth 6 (measured 699 / 663 / 684 GLUP/s at T = 6 / 7 / 8 on 64 config-3 images)

    // FloodFill (cuh:557-713): PathFlag always, pinned mask in 3-phase.  On the device, all images of a refill in the
    // same launches (floodfill.cu); on the host with deff2d_set_floodfill(ctx, 1)
    const bool strict = p->strict_reference != 0;        // see domain_load_impl (context.cu)
    const int ff_thr = (nphase == 3) ? 200 : (strict ? 150 : 149);        // cuh:1368, cuh:1695
    const bool ff_host = c->floodfill_mode == 1 || Nx < 2;

    // ---- resident stack -----------------------------------------------------------------------
    c->loaded = false;
    c->slab_domain = false;
    c->halo_above = c->halo_below = 0;
    c->Nx = (int64_t)GX * (Nx + 1) - 1;
    c->Ny = (int64_t)GY * (Ny + 1) - 1;
    c->NxG = Nx; c->NyG = Ny;
    c->pitch = ((c->Nx + 2 * XOFF) + 15) / 16 * 16;
    c->rows = c->Ny + 2;
    c->ghost_period = Nx + 1;
    c->own_first = 0; c->own_rows = c->Ny;
    c->nphase = nphase;
    c->CL = p->CL; c->CR = p->CR;
    c->omega = (p->omega > 0) ? p->omega : 2.0 / 3.0;
    c->check_every = (p->check_every > 0) ? p->check_every : 10000;
    c->solver = 0;
    c->cur = 0;
    const size_t stack_cells = (size_t)c->rows * (size_t)c->pitch;
    int rc;
    auto grow = [&](auto &buf, size_t n) -> int {
        if (buf.cap >= n && buf.p) return DEFF2D_OK;
        if (buf.p) { cudaFree(buf.p); buf.p = nullptr; buf.cap = 0; }
        using E = std::remove_reference_t<decltype(*buf.p)>;
        if (cudaMalloc((void **)&buf.p, n * sizeof(E)) != cudaSuccess) {
            buf.p = nullptr;
            (void)cudaGetLastError();
            set_error(c, "cudaMalloc of %zu bytes failed", n * sizeof(E));
            return DEFF2D_ERR_ALLOC;
        }
        buf.cap = n;
        return DEFF2D_OK;
    };
    if ((rc = grow(c->x[0], stack_cells)) || (rc = grow(c->x[1], stack_cells)) || (rc = grow(c->code, stack_cells)) ||
        (rc = grow(c->idx16, stack_cells))) return rc;
    // staging of one refill (at most every slot at once): source images and their FloodFill states / pinned masks
    if ((rc = grow(c->img, npix * (size_t)nslots)) || (rc = grow(c->grid, (size_t)cells * (size_t)nslots))) return rc;
    if ((rc = grow(c->lut, (size_t)nstages * DEFF2D_LUT_ENTRIES * 4)) || (rc = grow(c->dead, (size_t)nstages * DEFF2D_LUT_ENTRIES)) ||
        (rc = grow(c->clut, (size_t)nstages * DEFF2D_CLUT_ENTRIES * 4)) || (rc = grow(c->clut32, (size_t)nstages * DEFF2D_CLUT_ENTRIES * 8))) return rc;
    c->lut_stages = nstages;
    if (want_fields && (rc = grow(c->dense, (size_t)cells))) return rc;
    if ((rc = dev_ensure(c, b->slots, (size_t)nslots)) || (rc = dev_ensure(c, b->outs, (size_t)nslots)) ||
        (rc = dev_ensure(c, b->jobs, (size_t)nslots)) || (rc = dev_ensure(c, b->active, (size_t)nslots)) ||
        (rc = dev_ensure(c, b->ff_flags, (size_t)std::min(nslots, 32768) + 1))) return rc;
    if ((rc = host_ensure(c, b->h_slots, b->h_slots_cap, (size_t)nslots)) || (rc = host_ensure(c, b->h_jobs, b->h_jobs_cap, (size_t)nslots)) ||
        (rc = host_ensure(c, b->h_active, b->h_active_cap, (size_t)nslots)) ||
        (rc = host_ensure(c, b->h_ff_flags, b->h_ff_flags_cap, (size_t)std::min(nslots, 32768) + 1)) ||
        (rc = host_ensure(c, b->h_in, b->h_in_cap, npix * (size_t)nslots)) || (rc = host_ensure(c, b->h_out, b->h_out_cap, 1))) return rc;

    cudaStream_t s = c->stream;
    CUB(cudaMemsetAsync(c->x[0].p, 0, stack_cells * sizeof(double), s));
    CUB(cudaMemsetAsync(c->x[1].p, 0, stack_cells * sizeof(double), s));
    CUB(cudaMemsetAsync(c->code.p, DEFF2D_PHASE_GHOST, stack_cells, s));
    CUB(cudaMemsetAsync(b->slots.p, 0, (size_t)nslots * sizeof(BatchSlot), s));
    {
        std::vector<double> lut((size_t)nstages * DEFF2D_LUT_ENTRIES * 4);
        std::vector<uint8_t> dead((size_t)nstages * DEFF2D_LUT_ENTRIES);
        std::vector<double> clut((size_t)nstages * DEFF2D_CLUT_ENTRIES * 4);
        for (int k = 0; k < nstages; k++) {
            build_tables(stages.s[k].D, Nx, Ny, c->CL, c->CR, c->omega, lut.data() + (size_t)k * DEFF2D_LUT_ENTRIES * 4,
                         dead.data() + (size_t)k * DEFF2D_LUT_ENTRIES);
            compact_table(lut.data() + (size_t)k * DEFF2D_LUT_ENTRIES * 4, clut.data() + (size_t)k * DEFF2D_CLUT_ENTRIES * 4, nphase);
        }
        std::vector<uint32_t> clut32((size_t)nstages * DEFF2D_CLUT_ENTRIES * 8);
        split_table(clut.data(), clut32.data(), nstages);
        CUB(cudaMemcpyAsync(c->clut.p, clut.data(), clut.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        CUB(cudaMemcpyAsync(c->clut32.p, clut32.data(), clut32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        CUB(cudaMemcpyAsync(c->lut.p, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        CUB(cudaMemcpyAsync(c->dead.p, dead.data(), dead.size(), cudaMemcpyHostToDevice, s));
        CUB(cudaStreamSynchronize(s));
    }

    BatchGeom g;
    g.W = W; g.H = H; g.amp_x = p->amp_x; g.amp_y = p->amp_y; g.Nx = (int)Nx; g.Ny = (int)Ny; g.GX = GX; g.nphase = nphase;
    g.pitch = c->pitch; g.CL = c->CL; g.CR = c->CR; g.ce = c->check_every; g.nstages = nstages;

    const int old_family = c->tile_family, old_tblock = c->tblock;
    struct Restore {                                      // every exit path puts the caller's kernel selection back
        deff2d_ctx *c; int fam, tb;
        ~Restore() { c->tile_family = fam; c->tblock = tb; c->tile_list = nullptr; c->tile_count = 0; }
    } restore_guard{c, old_family, old_tblock};
    c->tile_family = 0;                                  // the default thread layout
    c->gather32 = 0;                                     // 8-byte weight gathers (the statistic that selects the other kind is per domain load)
    // tile grids of the pass depths in use (T and the remainders 1..T-1)
    int ow[9], oh[9], tx_n[9], ty_n[9];
    size_t tiles_cap = 0;
    for (int t = 1; t <= T; t++) {
        tma_tile_geometry(c, t, &ow[t], &oh[t]);
        tx_n[t] = (int)((c->Nx + ow[t] - 1) / ow[t]);
        ty_n[t] = (int)((c->Ny + oh[t] - 1) / oh[t]);
        if (ty_n[t] > 0xffff || tx_n[t] > 0xffff) { set_error(c, "packed batch: tile grid too large"); return DEFF2D_ERR_ARG; }
        tiles_cap += (size_t)tx_n[t] * ty_n[t];
    }
    if ((rc = dev_ensure(c, b->tiles, tiles_cap)) || (rc = host_ensure(c, b->h_tiles, b->h_tiles_cap, tiles_cap))) return rc;
    size_t tile_off[9] = {0};
    int tile_cnt[9] = {0};

    std::vector<int> slot_image((size_t)nslots, -1);     // host mirror: image in each slot (-1 empty)
    std::vector<long long> slot_iter((size_t)nslots, 0);
    std::vector<int> slot_stage((size_t)nslots, 0), slot_pathflag((size_t)nslots, 0);
    std::vector<double> slot_solve_ms((size_t)nslots, 0.0), slot_total_ms((size_t)nslots, 0.0);
    std::vector<uint8_t> mark, host_mask;
    std::vector<uint32_t> tl;
    std::vector<double> field_host;
    if (want_fields) field_host.resize((size_t)cells);
    int next_image = 0, nactive = 0, done_images = 0;
    bool active_changed = true;
    c->tblock = T;
    // measured on 256 x 256 images: 595-643 GLUP/s cluster-resident against 695 with the tiled kernel, so packed batches
    // only go cluster-resident on request (deff2d_set_resident(ctx, 2))
    const bool use_resident = c->resident_mode == 2 && resident_eligible(c, Nx, Ny);

    while (done_images < count) {
        // ---- refill empty slots from the stream ------------------------------------------------
        int njobs = 0, in_flight = 0;
        for (int sl = 0; sl < nslots; sl++) if (slot_image[(size_t)sl] >= 0) in_flight++;
        for (int sl = 0; sl < nslots && next_image < count; sl++) {
            if (slot_image[(size_t)sl] >= 0) continue;
            const int fr = fetch(user, next_image, b->h_in + npix * (size_t)njobs, (in_flight + njobs == 0) ? 1 : 0);
            if (fr == 1) break;                              // not decoded yet: go on sweeping what is resident
            if (fr == 2) { count = next_image; break; }      // the stream ends here
            if (fr != 0) { set_error(c, "packed batch: image %d could not be fetched (%d)", next_image, fr); return fr < 0 ? fr : DEFF2D_ERR_IO; }
            b->h_jobs[njobs].slot = sl;
            b->h_jobs[njobs].image = next_image;
            b->h_jobs[njobs].src = njobs;
            b->h_jobs[njobs].pad = 0;
            slot_image[(size_t)sl] = next_image;
            slot_iter[(size_t)sl] = 0;
            slot_stage[(size_t)sl] = 0;
            slot_solve_ms[(size_t)sl] = slot_total_ms[(size_t)sl] = 0;
            next_image++;
            njobs++;
            active_changed = true;
        }
        if (njobs) {
            CUB(cudaMemcpyAsync(c->img.p, b->h_in, npix * (size_t)njobs, cudaMemcpyHostToDevice, s));
            if (ff_host) {
                host_mask.resize((size_t)cells * (size_t)njobs);
                std::atomic<int> next(0);
                const int nthreads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
                auto work = [&]() {
                    for (;;) {
                        const int k = next.fetch_add(1);
                        if (k >= njobs) break;
                        uint8_t *gm = host_mask.data() + (size_t)k * cells;
                        const uint8_t *src = b->h_in + npix * (size_t)k;
                        for (int64_t i = 0; i < Ny; i++) {
                            const uint8_t *srow = src + (size_t)(i / p->amp_y) * W;
                            uint8_t *gr = gm + (size_t)i * Nx;
                            for (int64_t j = 0; j < Nx; j++) gr[j] = srow[j / p->amp_x] > ff_thr;
                        }
                        slot_pathflag[(size_t)b->h_jobs[k].slot] = floodfill(gm, Nx, Ny, strict);
                    }
                };
                std::vector<std::thread> pool;
                for (int t = 1; t < nthreads && t < njobs; t++) pool.emplace_back(work);
                work();
                for (auto &t : pool) t.join();
                if (nphase == 3) CUB(cudaMemcpyAsync(c->grid.p, host_mask.data(), (size_t)cells * (size_t)njobs, cudaMemcpyHostToDevice, s));
            } else {
                std::vector<int> pf((size_t)njobs, 0);
                for (int k0 = 0; k0 < njobs; k0 += 32768) {
                    const int nk = std::min(32768, njobs - k0);
                    if ((rc = floodfill_device_batch(c, c->img.p + (size_t)k0 * npix, W, H, p->amp_x, p->amp_y, ff_thr,
                                                     c->grid.p + (size_t)k0 * cells, Nx, Ny, nk, b->ff_flags.p, b->h_ff_flags,
                                                     pf.data() + k0, nullptr, strict))) return rc;
                }
                for (int k = 0; k < njobs; k++) slot_pathflag[(size_t)b->h_jobs[k].slot] = pf[(size_t)k];
            }
            CUB(cudaMemcpyAsync(b->jobs.p, b->h_jobs, (size_t)njobs * sizeof(BatchJob), cudaMemcpyHostToDevice, s));
            k_batch_clear<<<(njobs + 255) / 256, 256, 0, s>>>(b->jobs.p, njobs, b->outs.p);
            int bx = (int)std::min<int64_t>(((Ny + 2) * (Nx + 2) + 255) / 256, 64);
            k_batch_init<<<dim3((unsigned)bx, (unsigned)njobs), 256, 0, s>>>(g, b->jobs.p, c->img.p, nphase == 3 ? c->grid.p : nullptr,
                                                                            c->x[c->cur].p, c->x[c->cur ^ 1].p, c->code.p,
                                                                            b->slots.p, b->outs.p);
            // the table indices of the new images (and of their neighbours' shared ghost ring); the whole stack is
            // rebuilt -- 3 B per cell, once per refill, against >= 10 000 sweeps between refills
            launch_build_idx(s, c->code.p, c->idx16.p, c->Nx, c->Ny, c->pitch, c->ghost_period, nphase, nullptr);
            c->launches += 3;
            CUB(cudaStreamSynchronize(s));                   // the staging buffer is free for the next refill
        }
        if (active_changed) {
            nactive = 0;
            for (int sl = 0; sl < nslots; sl++)
                if (slot_image[(size_t)sl] >= 0) b->h_active[nactive++] = sl;
            if (nactive == 0) break;                         // the stream ended (fetch returned 2) and nothing is resident
            CUB(cudaMemcpyAsync(b->active.p, b->h_active, (size_t)nactive * sizeof(int), cudaMemcpyHostToDevice, s));
            size_t off = 0;
            for (int t = 1; t <= T; t++) {
                batch_tile_list(Nx, Ny, GX, b->h_active, nactive, ow[t], oh[t], tx_n[t], ty_n[t], mark, tl);
                tile_off[t] = off;
                tile_cnt[t] = (int)tl.size();
                std::memcpy(b->h_tiles + off, tl.data(), tl.size() * sizeof(uint32_t));
                off += tl.size();
            }
            CUB(cudaMemcpyAsync(b->tiles.p, b->h_tiles, off * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            active_changed = false;
        }
        // ---- sweeps up to the next event of any active image (check or MaxIter) ---------------
        long long n = -1;
        const long long ce = c->check_every;
        for (int a = 0; a < nactive; a++) {
            const int sl = b->h_active[a];
            const long long it = slot_iter[(size_t)sl];
            const BatchStage &st = stages.s[slot_stage[(size_t)sl]];
            const long long to_check = ((it % ce == 0) ? it : (it / ce + 1) * ce) + 1 - it;
            const long long to_max = st.max_iter - it;
            const long long m = std::min(to_check, to_max);
            if (n < 0 || m < n) n = m;
        }
        if (n < 1) { set_error(c, "packed batch: internal scheduling error"); return DEFF2D_ERR_STATE; }
        CUB(cudaEventRecord(b->e0, s));
        if (use_resident) {
            // images of up to 256 x 256 cells: one cluster per image keeps it on chip for all n sweeps (resident.cu)
            if ((rc = resident_sweeps(c, n, Nx, Ny, GX, b->active.p, nactive))) return rc;
        } else {
            if (n >= T && (rc = tma_passes(c, T, n / T, b->tiles.p + tile_off[T], tile_cnt[T]))) return rc;
            if (const int t = (int)(n % T)) {
                if ((rc = tma_pass(c, t, b->tiles.p + tile_off[t], tile_cnt[t], s))) return rc;
                c->cur ^= 1;
            }
        }
        k_batch_check<<<nactive, 1024, 0, s>>>(g, stages, b->slots.p, b->outs.p, b->active.p, n, c->x[c->cur].p, c->code.p, c->idx16.p);
        c->launches++;
        CUB(cudaEventRecord(b->e1, s));
        CUB(cudaMemcpyAsync(b->h_slots, b->slots.p, (size_t)nslots * sizeof(BatchSlot), cudaMemcpyDeviceToHost, s));
        CUB(cudaStreamSynchronize(s));
        {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { set_error(c, "packed batch launch failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
        }
        float ms = 0;
        CUB(cudaEventElapsedTime(&ms, b->e0, b->e1));
        // the window's device time is shared equally by the images that were swept in it; like the
        // reference's `Time` only the non-PreCond stages count towards solve_ms (cuh:1311)
        for (int a = 0; a < nactive; a++) {
            const int sl = b->h_active[a];
            slot_total_ms[(size_t)sl] += ms / nactive;
            if (!stages.s[slot_stage[(size_t)sl]].precond) slot_solve_ms[(size_t)sl] += ms / nactive;
        }
        // ---- adopt the device's decisions; finished images leave at once -------------------------
        for (int a = 0; a < nactive; a++) {
            const int sl = b->h_active[a];
            const BatchSlot &S = b->h_slots[sl];
            slot_iter[(size_t)sl] = S.iter;
            slot_stage[(size_t)sl] = S.stage;
            if (S.status != 2) continue;
            const int im = slot_image[(size_t)sl];
            CUB(cudaMemcpyAsync(b->h_out, b->outs.p + sl, sizeof(BatchOut), cudaMemcpyDeviceToHost, s));
            if (want_fields) {
                DomainView v = view(c);
                const int gx = sl % GX, gy = sl / GX;
                const int64_t off = ((int64_t)gy * (Ny + 1)) * c->pitch + (int64_t)gx * (Nx + 1);
                v.x_in += off; v.code += off;
                v.Nx = Nx; v.Ny = Ny;
                launch_extract_field(s, v, c->dense.p);
                c->launches++;
                CUB(cudaMemcpyAsync(field_host.data(), c->dense.p, (size_t)cells * sizeof(double), cudaMemcpyDeviceToHost, s));
            }
            CUB(cudaStreamSynchronize(s));
            const BatchOut &o = *b->h_out;
            deff2d_result r;
            std::memset(&r, 0, sizeof(r));
            r.n_cells = cells;
            r.pathflag = slot_pathflag[(size_t)sl];
            if (nphase == 2) r.porosity = accumulate_fraction((int64_t)o.below150, (int64_t)npix);       // cuh:397-405
            else {                                                                                       // calcFracts3D, cuh:411-448 (quirk Q21)
                int64_t ns = 0, nl = 0;
                const double Dfin[3] = {p->Df, p->Ds, p->Dg};
                for (int ph = 0; ph < 3; ph++) {
                    if (Dfin[ph] == p->Ds) ns += (int64_t)o.phase[ph];
                    else if (Dfin[ph] == p->Df) nl += (int64_t)o.phase[ph];
                }
                r.SVF = accumulate_fraction(ns, cells);
                r.LVF = accumulate_fraction(nl, cells);
            }
            r.nstages = o.nstages;
            for (int st = 0; st < o.nstages && st < DEFF2D_MAX_STAGES; st++) {
                r.iters[st] = o.iters[st];
                r.stage_deff_raw[st] = o.stage_deff_raw[st];
                r.stage_D[st] = stageD[st];
                r.total_iters += o.iters[st];
            }
            r.deff_raw = o.deff_raw;
            r.conv = o.conv;
            r.deff = o.deff_raw / p->Df;                        // cuh:2017, cuh:2370
            r.last_df = p->Df;
            r.solve_ms = slot_solve_ms[(size_t)sl];
            r.total_ms = slot_total_ms[(size_t)sl];
            if (done && done(user, im, &r, want_fields ? field_host.data() : nullptr)) {
                set_error(c, "packed batch: aborted by the result callback at image %d", im);
                return DEFF2D_ERR_STATE;
            }
            slot_image[(size_t)sl] = -1;
            done_images++;
            active_changed = true;
        }
    }
    if (solved) *solved = done_images;
    return DEFF2D_OK;
}

struct ArrayStream { const uint8_t *gray; size_t npix, ncell; deff2d_result *results; double *fields; };

static int array_fetch(void *user, int k, uint8_t *dst, int)
{
    const ArrayStream *a = static_cast<const ArrayStream *>(user);
    std::memcpy(dst, a->gray + a->npix * (size_t)k, a->npix);
    return 0;
}

static int array_done(void *user, int k, const deff2d_result *r, const double *field)
{
    const ArrayStream *a = static_cast<const ArrayStream *>(user);
    a->results[k] = *r;
    if (a->fields && field) std::memcpy(a->fields + a->ncell * (size_t)k, field, a->ncell * sizeof(double));
    return 0;
}

static bool batch_covers(const deff2d_params *p, int count, int W, int H, BatchStages *stages, double *stageD, int *nstages)
{
    if (count < 2 || W < 1 || H < 1 || p->amp_x < 1 || p->amp_y < 1) return false;
    if (p->verbose == 1) return false;              // keep the reference's per-image stdout order
    if (p->residual_tol > 0 || p->solver != 0) return false;   // non-parity stop rule / solver: per-image path
    std::memset(stages, 0, sizeof(*stages));
    if (build_stages(p, stages, stageD, nstages)) return false;
    const int64_t Nx = (int64_t)W * p->amp_x, Ny = (int64_t)H * p->amp_y;
    if (Nx * Ny > ((int64_t)16 << 20)) return false;    // large images fill the machine on their own
    return true;
}

// Returns 1 when the packed path does not cover the request (the caller then solves image by
// image), 0 on success, a negative status on error.
int batch_resident_solve(deff2d_ctx *c, const uint8_t *gray, int count, int W, int H, const deff2d_params *p,
                         deff2d_result *results, double *fields)
{
    BatchStages stages;
    double stageD[BATCH_MAX_STAGES];
    int nstages = 0;
    if (!batch_covers(p, count, W, H, &stages, stageD, &nstages)) return 1;
    ArrayStream a{gray, (size_t)W * H, (size_t)W * H * (size_t)p->amp_x * (size_t)p->amp_y, results, fields};
    return batch_stream(c, count, W, H, p, array_fetch, array_done, &a, fields != nullptr, stages, stageD, nstages, nullptr);
}

}  // namespace deff2d

DEFF2D_EXPORT int deff2d_batch_supported(const deff2d_params *p, int W, int H)
{
    if (!p) return 0;
    deff2d::BatchStages stages;
    double stageD[BATCH_MAX_STAGES];
    int nstages = 0;
    return deff2d::batch_covers(p, 2, W, H, &stages, stageD, &nstages) ? 1 : 0;
}

DEFF2D_EXPORT int deff2d_solve_batch_stream(deff2d_ctx *c, int count, int W, int H, const deff2d_params *p,
                                            deff2d_batch_fetch_fn fetch, deff2d_batch_done_fn done, void *user,
                                            int want_fields, int *solved)
{
    if (!c || !p || !fetch || count < 0) return DEFF2D_ERR_ARG;
    if (solved) *solved = 0;
    if (count == 0) return DEFF2D_OK;
    deff2d::BatchStages stages;
    double stageD[BATCH_MAX_STAGES];
    int nstages = 0;
    if (!deff2d::batch_covers(p, std::max(count, 2), W, H, &stages, stageD, &nstages)) {
        deff2d::set_error(c, "solve_batch_stream: these parameters are not covered by the packed batch mode");
        return DEFF2D_ERR_ARG;
    }
    return deff2d::batch_stream(c, count, W, H, p, fetch, done, user, want_fields != 0, stages, stageD, nstages, solved);
}

DEFF2D_EXPORT int deff2d_batch_plan(int64_t Nx, int64_t Ny, int count, int limit, int *GX, int *GY)
{
    if (Nx < 1 || Ny < 1 || count < 1 || !GX || !GY) return DEFF2D_ERR_ARG;
    deff2d::batch_plan(Nx, Ny, count, limit, GX, GY);
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_batch_tile_list(int64_t Nx, int64_t Ny, int GX, int GY, const int *active, int nactive, int T,
                                         uint32_t *tiles, int cap)
{
    if (Nx < 1 || Ny < 1 || GX < 1 || GY < 1 || T < 1 || T > 8 || nactive < 0 || (nactive && !active)) return DEFF2D_ERR_ARG;
    int ow, oh;
    deff2d::tma_tile_geometry(nullptr, T, &ow, &oh);
    const int64_t NxS = (int64_t)GX * (Nx + 1) - 1, NyS = (int64_t)GY * (Ny + 1) - 1;
    const int tx = (int)((NxS + ow - 1) / ow), ty = (int)((NyS + oh - 1) / oh);
    std::vector<uint8_t> mark;
    std::vector<uint32_t> out;
    deff2d::batch_tile_list(Nx, Ny, GX, active, nactive, ow, oh, tx, ty, mark, out);
    if ((int)out.size() > cap) return DEFF2D_ERR_ARG;
    if (tiles) std::memcpy(tiles, out.data(), out.size() * sizeof(uint32_t));
    return (int)out.size();
}

