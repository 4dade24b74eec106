// context.cu -- the C ABI of libdeff2d: persistent per-device context, resident domain,
// the reference solve loop and the three driver flows.  See include/deff2d.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "context.h"

namespace deff2d {

static std::string g_create_error;

void set_error(deff2d_ctx *c, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->error = buf;
    else g_create_error = buf;
}

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_error(c, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DEFF2D_ERR_CUDA;                                                          \
        }                                                                                    \
    } while (0)

template <typename T>
static int ensure(deff2d_ctx *c, DevBuf<T> &b, size_t n)
{
    if (b.cap >= n && b.p) return DEFF2D_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    // grow-only arena: 1/8 slack so that a batch of slightly different sizes does not realloc
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc((void **)&b.p, want * sizeof(T));
    if (e != cudaSuccess) {
        want = n;
        e = cudaMalloc((void **)&b.p, want * sizeof(T));
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        set_error(c, "cudaMalloc of %zu bytes failed: %s", want * sizeof(T), cudaGetErrorString(e));
        (void)cudaGetLastError();
        return DEFF2D_ERR_ALLOC;
    }
    b.cap = want;
    return DEFF2D_OK;
}

DomainView view(const deff2d_ctx *c)
{
    DomainView v;
    v.x_in = c->x[c->cur].p;
    v.x_out = c->x[c->cur ^ 1].p;
    v.code = c->code.p;
    v.lut = c->lut.p;
    v.dead = c->dead.p;
    v.Nx = c->Nx; v.Ny = c->Ny; v.pitch = c->pitch;
    v.om = 1.0 - c->omega;      // (1.0 - w), cuh:89
    return v;
}

static int upload_tables(deff2d_ctx *c)
{
    std::vector<double> lut((size_t)DEFF2D_LUT_ENTRIES * 4);
    std::vector<uint8_t> dead(DEFF2D_LUT_ENTRIES);
    build_tables(c->Dphase, c->NxG, c->NyG, c->CL, c->CR, c->omega, lut.data(), dead.data());
    int rc;
    std::vector<double> clut((size_t)DEFF2D_CLUT_ENTRIES * 4);
    compact_table(lut.data(), clut.data(), c->nphase);
    if ((rc = ensure(c, c->lut, lut.size()))) return rc;
    std::vector<uint32_t> clut32((size_t)DEFF2D_CLUT_ENTRIES * 8);
    split_table(clut.data(), clut32.data(), 1);
    if ((rc = ensure(c, c->clut, clut.size()))) return rc;
    if ((rc = ensure(c, c->clut32, clut32.size()))) return rc;
    if ((rc = ensure(c, c->dead, dead.size()))) return rc;
    c->lut_stages = 1;
    CU(cudaMemcpyAsync(c->clut.p, clut.data(), clut.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->clut32.p, clut32.data(), clut32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    // pageable source: the copy is staged before the call returns, the vectors may die
    CU(cudaMemcpyAsync(c->lut.p, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->dead.p, dead.data(), dead.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DEFF2D_OK;
}

// One sweep (or, with the tiled kernel, up to tblock sweeps) -- enqueue only.
static int enqueue_sweeps(deff2d_ctx *c, int64_t n)
{
    if (c->slab && c->slab_domain) return slab_enqueue_sweeps(c, n);
    while (n > 0) {
        int64_t done = 0;
        // kernel 0 (default): the TMA tiled kernel (64 x 64 tiles, 2 x 8 cells per thread) from 64 cells up, replayed
        // as CUDA graphs.  Depth: 6 sweeps per pass -- measured on B200 with the 2 x 8 layout at T = 4 / 5 / 6 / 7 / 8:
        // config 2 861 / 839 / 878 / 848 / 838 GLUP/s, 4096^2 blob medium 720 / 784 / 825 / 793 / 789, 2048^2 site
        // percolation 562 / 582 / 644 / 650 / 626 (the 4 x 4 layout of round 1: 802 / 776 / 814 / 823 / 822 on config 2).
        const int64_t ncell = c->Nx * c->Ny;
        // up to 256 x 256 cells: the whole domain stays on chip for all n sweeps (resident.cu), one launch
        if (c->kernel == 0 && c->resident_mode != 1 && ncell >= 64 && !c->tile_list && resident_eligible(c, c->Nx, c->Ny)) {
            int rc = resident_sweeps(c, n, c->Nx, c->Ny, 1, nullptr, 1);
            if (rc) return rc;
            break;
        }
        if (c->kernel == 2 || (c->kernel == 0 && ncell >= 64)) {
            int rc = launch_sweep_tma(c, n, c->kernel == 0 ? c->k2_default_depth : c->tblock, &done);
            if (rc) return rc;
        }
        if (done == 0) {
            launch_sweep_simple(c->stream, view(c), nullptr);
            c->launches++;
            c->cur ^= 1;
            done = 1;
        }
        n -= done;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(c, "sweep launch failed: %s", cudaGetErrorString(e));
        return DEFF2D_ERR_CUDA;
    }
    return DEFF2D_OK;
}

static int read_state(deff2d_ctx *c)
{
    CU(cudaMemcpyAsync(c->h_state, c->d_state, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DEFF2D_OK;
}

// The reference loop, cuh:1232-1290: sweep; if (iterCount % check == 0) flux + stop rule;
// iterCount++.  Sweeps between two checks are enqueued back to back with no host
// round trip; the host synchronises once per check to read the stop flag.
int solve_loop(deff2d_ctx *c, double tol, int64_t max_iter, bool verbose_checks, double print_div,
               int64_t *iters_out)
{
    launch_reset_state(c->stream, c->d_state);
    c->launches++;
    if (c->solver == 1) {
        // NON-PARITY mode: Chebyshev-accelerated Jacobi to a relative residual (chebyshev.cu), then Deff of the result
        int64_t sweeps = 0;
        int rc = chebyshev_solve(c, c->residual_tol, max_iter, &sweeps);
        if (rc) return rc;
        const double relres = c->h_state->resid;
        launch_flux(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->own_first, c->own_rows, c->d_state);
        c->launches++;
        if ((rc = read_state(c))) return rc;
        const double qAvg = (c->h_state->q[0] + c->h_state->q[1]) / (2.0 * (double)c->NyG);   // cuh:1263
        c->h_state->deff_new = qAvg / (c->CR - c->CL);                                        // cuh:1264
        c->h_state->resid = relres;
        c->h_state->conv = relres;                  // what this mode converged on: the relative residual reached
        c->h_state->change = relres;
        c->h_state->nchecks = 1;
        c->h_state->trace[0] = c->h_state->deff_new;
        if (verbose_checks)
            std::printf("Chebyshev: %lld sweeps, relative residual %1.3e, Deff = %1.3e\n", (long long)sweeps, relres, c->h_state->deff_new / print_div);
        *iters_out = sweeps;
        return DEFF2D_OK;
    }
    const int64_t ce = c->check_every;
    int64_t iter = 0;
    bool stopped = false;
    // `tol < fabs(100.0)` fails for tol >= 100 (or NaN): the reference then runs no sweep at all
    if (!(tol < 100.0)) { *iters_out = 0; return read_state(c); }
    while (iter < max_iter) {
        const int64_t next_check = (iter % ce == 0) ? iter : (iter / ce + 1) * ce;
        const int64_t last = std::min<int64_t>(next_check, max_iter - 1);
        const int64_t n = last - iter + 1;
        int rc = enqueue_sweeps(c, n);
        if (rc) return rc;
        iter += n;
        if (last == next_check) {
            // residual mode: the Deff-change test is switched off (tol -1 is never reached; NaN still stops)
            const double tol_eff = c->residual_tol > 0 ? -1.0 : tol;
            if (slab_is_distributed(c)) {
                launch_flux(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->own_first, c->own_rows, c->d_state);
                c->launches++;
                if ((rc = slab_allreduce_q(c))) return rc;
                launch_check(c->stream, c->d_state, c->NyG, c->CL, c->CR, tol_eff, last);
                c->launches++;
            } else {
                launch_flux_check(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->NyG, c->own_first, c->own_rows, tol_eff, last, c->d_state);
                c->launches++;
            }
            if ((rc = read_state(c))) return rc;
            if (c->residual_tol > 0 && !c->h_state->stop) {
                CU(cudaMemsetAsync(c->d_scalar, 0, sizeof(double), c->stream));
                launch_residual(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->NyG, c->d_scalar);
                c->launches++;
                CU(cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                c->h_state->resid = *c->h_scalar;
                if (!(*c->h_scalar > c->residual_tol)) { iter = last + 1; stopped = true; break; }
            }
            if (verbose_checks)     // cuh:1270
                std::printf("Iteration = %d, Deff = %1.3e, Deff Change = %1.3e\n", (int)last,
                            c->h_state->deff_new / print_div, c->h_state->change);
            if (c->h_state->stop) { stopped = true; break; }
        }
    }
    if (!stopped) { int rc = read_state(c); if (rc) return rc; }
    *iters_out = iter;
    return DEFF2D_OK;
}

static int domain_load_impl(deff2d_ctx *c, const uint8_t *gray, int W, int Hsrc, int nphase,
                            const deff2d_params *p, int64_t grow0, int64_t img_row0, int64_t NyLocal,
                            int64_t NyG, int64_t own_first, int64_t own_rows, const uint8_t *grid_host,
                            bool run_floodfill, bool global_image = false)
{
    if (!gray || W < 1 || Hsrc < 1 || !p || (nphase != 2 && nphase != 3) || p->amp_x < 1 || p->amp_y < 1) {
        set_error(c, "domain_load: invalid argument");
        return DEFF2D_ERR_ARG;
    }
    CU(cudaSetDevice(c->device));
    slab_peer_reset(c);
    const int64_t Nx = (int64_t)W * p->amp_x;
    const int64_t Ny = NyLocal;
    if (Nx * (NyG > Ny ? NyG : Ny) >= ((int64_t)1 << 40)) { set_error(c, "domain too large"); return DEFF2D_ERR_ARG; }
    c->Nx = Nx; c->Ny = Ny; c->NxG = Nx; c->NyG = NyG;
    c->pitch = ((Nx + 2 * DEFF2D_XOFF) + 15) / 16 * 16;
    c->rows = Ny + 2;
    c->ghost_period = Nx + 1;
    c->tile_list = nullptr; c->tile_count = 0;
    c->own_first = own_first; c->own_rows = own_rows;
    c->nphase = nphase;
    c->CL = p->CL; c->CR = p->CR;
    c->omega = (p->omega > 0) ? p->omega : 2.0 / 3.0;
    c->check_every = (p->check_every > 0) ? p->check_every : 10000;
    c->residual_tol = (p->residual_tol > 0 && !c->slab_domain) ? p->residual_tol : 0;
    if (p->solver != 0 && p->solver != 1) { set_error(c, "unknown solver %d", p->solver); return DEFF2D_ERR_ARG; }
    if (p->solver == 1 && c->slab_domain) { set_error(c, "the Chebyshev solver (solver = 1) runs on single-GPU domains"); return DEFF2D_ERR_ARG; }
    c->solver = p->solver;
    if (c->solver == 1) c->omega = 1.0;          // chebyshev.cu: the tables hold the plain Jacobi weights, the factors come per sweep
    c->Dphase[0] = p->Df; c->Dphase[1] = p->Ds; c->Dphase[2] = p->Dg;
    c->cur = 0;
    const size_t cells = (size_t)c->rows * (size_t)c->pitch;
    int rc;
    if ((rc = ensure(c, c->x[0], cells))) return rc;
    if ((rc = ensure(c, c->x[1], cells))) return rc;
    if ((rc = ensure(c, c->code, cells))) return rc;
    if ((rc = ensure(c, c->idx16, cells))) return rc;
    if ((rc = ensure(c, c->img, (size_t)W * Hsrc))) return rc;
    CU(cudaMemcpyAsync(c->img.p, gray, (size_t)W * Hsrc, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_counts, 0, sizeof(Counts), c->stream));

    // FloodFill on the host (cuh:557-713): PathFlag always; the Grid in {1,2} mask feeds the
    // matrix only in 3-phase (cuh:750).  The solid mask follows the evident intent of
    // cuh:1693-1701 / 1364-1377 for MeshAmp > 1 (quirk Q10: the reference indexes with the
    // un-amplified width there and reads out of bounds).
    const uint8_t *grid_dev = nullptr;
    c->pathflag = 0;
    c->floodfill_passes = 0;
    // strict_reference = 0 (defined behaviour where the reference has quirks): no right-column
    // seeding (Q11) and pixel == 150 is solid in the FloodFill mask as it is in D (Q12: cuh:1695 vs cuh:1779)
    const bool strict = p->strict_reference != 0;
    const int ff_thr = (nphase == 3) ? 200 : (strict ? 150 : 149);
    const bool ff_device = run_floodfill && Nx >= 2 &&
                           (c->floodfill_mode == 2 || (c->floodfill_mode == 0 && Nx * Ny >= ((int64_t)1 << 16)));
    if (global_image) {
        // slab of a decomposed domain, `gray` is the whole global image: the flood needs global connectivity, so every
        // rank floods the whole domain on its own GPU (replicas) and keeps the mask rows of its slab --
        // no host flood, no 1 B/cell mask upload, the same steps as an undecomposed load
        if ((rc = ensure(c, c->grid, (size_t)Nx * NyG))) return rc;
        int pf = 0;
        if ((rc = floodfill_device(c, c->img.p, W, Hsrc, p->amp_x, p->amp_y, ff_thr, c->grid.p, Nx, NyG,
                                   reinterpret_cast<int *>(c->d_scalar), reinterpret_cast<int *>(c->h_scalar), &pf,
                                   &c->floodfill_passes, strict))) return rc;
        c->pathflag = pf;
        if (nphase == 3) grid_dev = c->grid.p + (size_t)grow0 * Nx;
    } else if (ff_device) {
        // label propagation on the device (floodfill.cu): same reachability, no 1 B/cell mask upload
        if ((rc = ensure(c, c->grid, (size_t)Nx * Ny))) return rc;
        int pf = 0;
        if ((rc = floodfill_device(c, c->img.p, W, Hsrc, p->amp_x, p->amp_y, ff_thr, c->grid.p, Nx, Ny,
                                   reinterpret_cast<int *>(c->d_scalar), reinterpret_cast<int *>(c->h_scalar), &pf,
                                   &c->floodfill_passes, strict))) return rc;
        c->pathflag = pf;
        if (nphase == 3) grid_dev = c->grid.p;
    } else if (run_floodfill) {
        const int thr = ff_thr;
        c->h_grid.resize((size_t)Nx * Ny);
        for (int64_t i = 0; i < Ny; i++) {
            const uint8_t *srow = gray + (size_t)(i / p->amp_y) * W;
            uint8_t *g = c->h_grid.data() + (size_t)i * Nx;
            if (p->amp_x == 1) for (int64_t j = 0; j < Nx; j++) g[j] = srow[j] > thr;
            else for (int64_t j = 0; j < Nx; j++) g[j] = srow[j / p->amp_x] > thr;
        }
        c->pathflag = floodfill(c->h_grid.data(), Nx, Ny, strict);
        grid_host = (nphase == 3) ? c->h_grid.data() : nullptr;
    }
    if (grid_host) {
        if ((rc = ensure(c, c->grid, (size_t)Nx * Ny))) return rc;
        CU(cudaMemcpyAsync(c->grid.p, grid_host, (size_t)Nx * Ny, cudaMemcpyHostToDevice, c->stream));
        grid_dev = c->grid.p;
    }
    launch_init_domain(c->stream, c->img.p, W, Hsrc, p->amp_x, p->amp_y, nphase, grow0, img_row0, grid_dev,
                       c->x[0].p, c->x[1].p, c->code.p, Nx, Ny, c->pitch, c->NxG, c->CL, c->CR, own_first,
                       own_rows, c->d_counts);
    launch_build_idx(c->stream, c->code.p, c->idx16.p, Nx, Ny, c->pitch, c->ghost_period, nphase, c->d_counts);
    launch_count_below(c->stream, c->img.p, (int64_t)W * Hsrc, 150, c->d_counts);
    c->launches += 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_counts, c->d_counts, sizeof(Counts), cudaMemcpyDeviceToHost, c->stream));
    if ((rc = upload_tables(c))) return rc;     // synchronises the stream
    // media with a phase interface at most cells (site percolation): two 4-byte gathers per weight instead of one 8-byte
    // gather -- 738 against 646 GLUP/s on the 2048^2 percolation medium, but 863 against 884 on config 2 (same box)
    c->gather32 = (c->h_counts->idx_mixed * 2 > c->h_counts->idx_cells) ? 1 : 0;
    c->src_pixels = (int64_t)W * Hsrc;
    c->porosity = accumulate_fraction((int64_t)c->h_counts->below150, c->src_pixels);   // cuh:397-405
    c->loaded = true;
    return DEFF2D_OK;
}

// calcFracts3D (cuh:411-448) classifies by VALUE equality with DCsolid, then DCfluid (quirk
// Q21), on the D array of the final stage; reproduce from the per-phase cell counts.
static void fracts3(const deff2d_ctx *c, double Ds, double Df, double *SVF, double *LVF)
{
    int64_t ns = 0, nl = 0;
    for (int ph = 0; ph < 3; ph++) {
        const int64_t cnt = (int64_t)c->h_counts->phase[ph];
        if (c->Dphase[ph] == Ds) ns += cnt;
        else if (c->Dphase[ph] == Df) nl += cnt;
    }
    const int64_t total = c->NxG * c->NyG;
    *SVF = accumulate_fraction(ns, total);
    *LVF = accumulate_fraction(nl, total);
}

struct StageTimer {
    deff2d_ctx *c;
    explicit StageTimer(deff2d_ctx *ctx) : c(ctx) { cudaEventRecord(c->ev0, c->stream); }
    double stop()
    {
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        return ms;
    }
};

static int run_stage(deff2d_ctx *c, const deff2d_params *p, deff2d_result *res, double Ds, double Df, double Dg,
                     double stageD, double tol, int64_t max_iter, bool precond, double print_div)
{
    int rc;
    c->Dphase[0] = Df; c->Dphase[1] = Ds; c->Dphase[2] = Dg;
    if ((rc = upload_tables(c))) return rc;
    StageTimer t(c);
    int64_t iters = 0;
    // cuh:1267: per-check lines only for single-image runs
    const bool vchk = p->verbose == 1 && p->mode != DEFF2D_MODE_2PH_BATCH && !c->in_batch;
    if ((rc = solve_loop(c, tol, max_iter, vchk, print_div, &iters))) return rc;
    const double ms = t.stop();
    res->total_ms += ms;
    if (!precond) {                                  // cuh:1309-1311 vs cuh:1144-1159
        res->solve_ms += ms;
        res->deff_raw = c->h_state->deff_new;
        if (c->h_state->nchecks > 0) res->conv = c->h_state->conv;
    }
    const int s = res->nstages;
    if (s < DEFF2D_MAX_STAGES) {
        res->iters[s] = iters;
        res->stage_deff_raw[s] = c->h_state->deff_new;
        res->stage_D[s] = stageD;
    }
    res->nstages++;
    res->total_iters += iters;
    if (p->verbose == 1) std::printf("Iterations taken = %d\n", (int)iters);   // cuh:1797, 1544, 1595, 2012
    return DEFF2D_OK;
}

int solve_image_impl(deff2d_ctx *c, const uint8_t *gray, int W, int H, const deff2d_params *p,
                     deff2d_result *res, double *field, int image_number)
{
    if (!res || !p) { set_error(c, "solve_image: null argument"); return DEFF2D_ERR_ARG; }
    std::memset(res, 0, sizeof(*res));
    if (p->mode < 0 || p->mode > 2) { set_error(c, "solve_image: bad mode %d", p->mode); return DEFF2D_ERR_ARG; }
    // cuh:1672-1675 (SingleSim3Phase lacks the guard, quirk Q23; amp < 1 is meaningless anyway)
    if (p->amp_x < 1 || p->amp_y < 1) {
        std::printf("MeshIncrease has to be an integer greater than 1.\n");
        set_error(c, "MeshIncrease has to be an integer greater than 1.");
        return DEFF2D_ERR_ARG;
    }
    const int nphase = (p->mode == DEFF2D_MODE_3PH) ? 3 : 2;
    const int64_t Ny = (int64_t)H * p->amp_y;
    c->halo_above = c->halo_below = 0; c->grow0 = 0;
    c->slab_domain = false;
    int rc = domain_load_impl(c, gray, W, H, nphase, p, 0, 0, Ny, Ny, 0, Ny, nullptr, true);
    if (rc) return rc;
    res->n_cells = c->NxG * c->NyG;
    res->pathflag = c->pathflag;
    if (nphase == 2) {
        res->porosity = c->porosity;
        if (p->verbose == 1) {      // cuh:1660-1663
            std::cout << "Width = " << W << " Height = " << H << " Channel = " << 1 << std::endl;
            std::cout << "Porosity = " << res->porosity << std::endl;
        }
    }
    cudaMemsetAsync(&c->d_state->conv, 0, sizeof(double), c->stream);

    // the stage sequence of the driver (host.cpp: stage_list -- shared with the packed batch and the multi-GPU driver)
    StageSpec stages[DEFF2D_MAX_STAGES];
    const int nst = stage_list(p, stages, DEFF2D_MAX_STAGES);
    if (nst < 0) { set_error(c, "solve_image: more than %d continuation stages", DEFF2D_MAX_STAGES); return DEFF2D_ERR_ARG; }
    res->last_df = p->Df;
    for (int k = 0; k < nst; k++) {
        const StageSpec &st = stages[k];
        if (p->mode == DEFF2D_MODE_3PH) {
            if (st.precond) {
                if (p->verbose == 1) std::printf("Pre-Cond Stage %d: DCG = %1.3e\n", k + 1, st.Dg);
            } else {
                c->Dphase[0] = p->Df; c->Dphase[1] = p->Ds; c->Dphase[2] = p->Dg;
                fracts3(c, p->Ds, p->Df, &res->SVF, &res->LVF);     // cuh:1582
            }
        }
        if ((rc = run_stage(c, p, res, st.Ds, st.Df, st.Dg, st.stageD, st.tol, st.max_iter, st.precond != 0, p->Df))) return rc;
        if (st.precond) continue;
        res->deff = res->deff_raw / st.Df;                          // cuh:1802, cuh:2017, cuh:1601
        res->last_df = st.Df;
        if (p->verbose == 1) {
            if (p->mode == DEFF2D_MODE_2PH_BATCH)                   // cuh:2020
                std::cout << "Number" << image_number << "DCF = " << st.Df << ", Deff " << res->deff << std::endl;
            else if (!st.defined_q8)                                // cuh:1807, cuh:1607
                std::cout << "DCF = " << st.Df << ", Deff " << res->deff << std::endl;
        }
    }
    if (field) {
        rc = deff2d_domain_get_field(c, field);
        if (rc) return rc;
    }
    return DEFF2D_OK;
}

}  // namespace deff2d

using namespace deff2d;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------

DEFF2D_EXPORT int deff2d_create(deff2d_ctx **out, int device)
{
    if (!out) return DEFF2D_ERR_ARG;
    *out = nullptr;
    deff2d_ctx *c = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(nullptr, "no CUDA device: %s (libdeff2d has no CPU fallback)",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        (void)cudaGetLastError();
        return DEFF2D_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error(nullptr, "device %d out of range (0..%d)", device, ndev - 1); return DEFF2D_ERR_ARG; }
    c = new deff2d_ctx();
    c->device = device;
    auto fail = [&](const char *what, cudaError_t err) {
        set_error(nullptr, "%s failed: %s", what, cudaGetErrorString(err));
        delete c;
        return DEFF2D_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaGetDeviceProperties(&c->prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaMalloc((void **)&c->d_state, sizeof(SolveState))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&c->d_counts, sizeof(Counts))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&c->d_scalar, 64)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMallocHost((void **)&c->h_state, sizeof(SolveState))) != cudaSuccess) return fail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void **)&c->h_counts, sizeof(Counts))) != cudaSuccess) return fail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void **)&c->h_scalar, 64)) != cudaSuccess) return fail("cudaMallocHost", e);
    cudaMemset(c->d_state, 0, sizeof(SolveState));
    std::memset(c->h_state, 0, sizeof(SolveState));
    c->kernel = 0;
    c->tblock = 1;
    *out = c;
    return DEFF2D_OK;
}

DEFF2D_EXPORT void deff2d_destroy(deff2d_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    slab_destroy(c);
    batch_destroy(c);
    resident_destroy(c);
    tma_destroy(c);
    for (int k = 0; k < 2; k++) if (c->x[k].p) cudaFree(c->x[k].p);
    if (c->code.p) cudaFree(c->code.p);
    if (c->idx16.p) cudaFree(c->idx16.p);
    if (c->img.p) cudaFree(c->img.p);
    if (c->grid.p) cudaFree(c->grid.p);
    if (c->lut.p) cudaFree(c->lut.p);
    if (c->clut.p) cudaFree(c->clut.p);
    if (c->clut32.p) cudaFree(c->clut32.p);
    if (c->dead.p) cudaFree(c->dead.p);
    if (c->dense.p) cudaFree(c->dense.p);
    if (c->dense8.p) cudaFree(c->dense8.p);
    cudaFree(c->d_state); cudaFree(c->d_counts); cudaFree(c->d_scalar);
    cudaFreeHost(c->h_state); cudaFreeHost(c->h_counts); cudaFreeHost(c->h_scalar);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev_sync);
    cudaStreamDestroy(c->stream); cudaStreamDestroy(c->comm_stream);
    delete c;     // deliberately no cudaDeviceReset() (reference quirk Q16, cuh:1015)
}

DEFF2D_EXPORT const char *deff2d_last_error(const deff2d_ctx *c)
{
    return c ? c->error.c_str() : g_create_error.c_str();
}

DEFF2D_EXPORT int deff2d_solve_image(deff2d_ctx *c, const uint8_t *gray, int W, int H, const deff2d_params *p,
                                     deff2d_result *res, double *field)
{
    if (!c) return DEFF2D_ERR_ARG;
    CU(cudaSetDevice(c->device));
    c->in_batch = false;
    return solve_image_impl(c, gray, W, H, p, res, field, 0);
}

DEFF2D_EXPORT int deff2d_solve_batch(deff2d_ctx *c, const uint8_t *gray, int count, int W, int H,
                                     const deff2d_params *p, deff2d_result *results, double *fields)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (count < 0 || (count > 0 && (!gray || !results)) || !p) { set_error(c, "solve_batch: invalid argument"); return DEFF2D_ERR_ARG; }
    CU(cudaSetDevice(c->device));
    int rc = batch_resident_solve(c, gray, count, W, H, p, results, fields);
    if (rc != 1) return rc;          // 1: the resident batch kernel does not cover this case
    c->in_batch = true;
    const size_t npix = (size_t)W * H;
    const size_t ncell = npix * (size_t)p->amp_x * (size_t)p->amp_y;
    for (int k = 0; k < count; k++) {
        rc = solve_image_impl(c, gray + npix * k, W, H, p, results + k, fields ? fields + ncell * k : nullptr, k);
        if (rc) { c->in_batch = false; return rc; }
    }
    c->in_batch = false;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_load(deff2d_ctx *c, const uint8_t *gray, int W, int H, int nphase,
                                     const deff2d_params *p)
{
    if (!c || !p) return DEFF2D_ERR_ARG;
    const int64_t Ny = (int64_t)H * p->amp_y;
    c->halo_above = c->halo_below = 0; c->grow0 = 0;
    c->slab_domain = false;
    return domain_load_impl(c, gray, W, H, nphase, p, 0, 0, Ny, Ny, 0, Ny, nullptr, true);
}

DEFF2D_EXPORT int deff2d_domain_load_slab(deff2d_ctx *c, const uint8_t *gray, int W, int Hslab, int nphase,
                                          const deff2d_params *p, int64_t row0, int64_t NyGlobal,
                                          int halo_rows, const uint8_t *pinned)
{
    if (!c || !p) return DEFF2D_ERR_ARG;
    // the slab owns amplified global rows [row0, row0 + Hslab*amp_y); it also holds up to
    // halo_rows amplified rows of each neighbour (fewer at the global top / bottom)
    const int64_t own_rows = (int64_t)Hslab * p->amp_y;
    if (row0 % p->amp_y != 0 || halo_rows % p->amp_y != 0) {
        set_error(c, "slab row0 and halo_rows must be multiples of amp_y");
        return DEFF2D_ERR_ARG;
    }
    const int64_t above = std::min<int64_t>(halo_rows, row0);
    const int64_t below = std::min<int64_t>(halo_rows, NyGlobal - (row0 + own_rows));
    if (below < 0) { set_error(c, "slab exceeds the global domain"); return DEFF2D_ERR_ARG; }
    if ((row0 > 0 && above != halo_rows) || (row0 + own_rows < NyGlobal && below != halo_rows)) {
        set_error(c, "slab: a neighbouring slab is thinner than the %d halo rows", halo_rows);   // the exchange moves whole halo blocks
        return DEFF2D_ERR_ARG;
    }
    const int64_t NyLocal = above + own_rows + below;
    const int Hsrc = (int)(NyLocal / p->amp_y);
    c->halo_above = above; c->halo_below = below; c->grow0 = row0 - above;
    c->slab_domain = true;
    c->halo_valid = std::max(above, below);                  // x0 is exact everywhere
    return domain_load_impl(c, gray, W, Hsrc, nphase, p, row0 - above, (row0 - above) / p->amp_y, NyLocal,
                            NyGlobal, above, own_rows, pinned, false);
}

DEFF2D_EXPORT int deff2d_domain_load_slab_global(deff2d_ctx *c, const uint8_t *gray, int W, int H, int nphase,
                                                 const deff2d_params *p, int64_t row0, int64_t own_rows, int halo_rows)
{
    if (!c || !p || !gray || W < 1 || H < 1 || p->amp_x < 1 || p->amp_y < 1) return DEFF2D_ERR_ARG;
    const int64_t NyGlobal = (int64_t)H * p->amp_y;
    if (row0 < 0 || own_rows < 1 || row0 + own_rows > NyGlobal || halo_rows < 0) {
        set_error(c, "slab rows [%lld, %lld) outside the global domain of %lld rows", (long long)row0, (long long)(row0 + own_rows), (long long)NyGlobal);
        return DEFF2D_ERR_ARG;
    }
    const int64_t above = std::min<int64_t>(halo_rows, row0);
    const int64_t below = std::min<int64_t>(halo_rows, NyGlobal - (row0 + own_rows));
    if ((row0 > 0 && above != halo_rows) || (row0 + own_rows < NyGlobal && below != halo_rows)) {
        set_error(c, "slab: a neighbouring slab is thinner than the %d halo rows", halo_rows);   // the exchange moves whole halo blocks
        return DEFF2D_ERR_ARG;
    }
    c->halo_above = above; c->halo_below = below; c->grow0 = row0 - above;
    c->slab_domain = true;
    c->halo_valid = std::max(above, below);                  // x0 is exact everywhere
    // the device holds the whole source image (1 B per pixel); rows are addressed globally (img_row0 = 0)
    return domain_load_impl(c, gray, W, H, nphase, p, row0 - above, 0, above + own_rows + below, NyGlobal, above, own_rows,
                            nullptr, false, true);
}

DEFF2D_EXPORT int deff2d_domain_set_D(deff2d_ctx *c, double Ds, double Df, double Dg)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    c->Dphase[0] = Df; c->Dphase[1] = Ds; c->Dphase[2] = Dg;
    return upload_tables(c);
}

DEFF2D_EXPORT int deff2d_domain_sweeps(deff2d_ctx *c, int64_t n)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    if (n < 0) return DEFF2D_ERR_ARG;
    CU(cudaSetDevice(c->device));
    return enqueue_sweeps(c, n);
}

DEFF2D_EXPORT int deff2d_domain_sweeps_timed(deff2d_ctx *c, int64_t n, float *ms)
{
    if (!c || !ms) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev0, c->stream));
    int rc = enqueue_sweeps(c, n);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_flux(deff2d_ctx *c, double *deff_raw, double *q)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    launch_flux(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->own_first, c->own_rows, c->d_state);
    c->launches++;
    int rc = read_state(c);
    if (rc) return rc;
    if (q) { q[0] = c->h_state->q[0]; q[1] = c->h_state->q[1]; }
    if (deff_raw) {
        const double qAvg = (c->h_state->q[0] + c->h_state->q[1]) / (2.0 * (double)c->NyG);   // cuh:1263
        *deff_raw = qAvg / (c->CR - c->CL);                                                  // cuh:1264
    }
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_residual(deff2d_ctx *c, double *res)
{
    if (!c || !res) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->d_scalar, 0, sizeof(double), c->stream));
    launch_residual(c->stream, view(c), c->Dphase, c->CL, c->CR, c->NxG, c->NyG, c->d_scalar);
    c->launches++;
    CU(cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *res = *c->h_scalar;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_solve(deff2d_ctx *c, double tol, int64_t max_iter, int64_t *iters,
                                      double *deff_raw, double *conv, double *trace, int trace_cap, int *ntrace)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    int64_t it = 0;
    int rc = solve_loop(c, tol, max_iter, false, 1.0, &it);
    if (rc) return rc;
    if (iters) *iters = it;
    if (deff_raw) *deff_raw = c->h_state->deff_new;
    if (conv) *conv = c->h_state->conv;
    const int nt = c->h_state->nchecks;
    if (ntrace) *ntrace = nt;
    if (trace) for (int k = 0; k < nt && k < trace_cap && k < 256; k++) trace[k] = c->h_state->trace[k];
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_get_field(deff2d_ctx *c, double *field)
{
    if (!c || !field) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->Nx * (size_t)c->Ny;
    int rc = ensure(c, c->dense, n);
    if (rc) return rc;
    launch_extract_field(c->stream, view(c), c->dense.p);
    c->launches++;
    CU(cudaMemcpyAsync(field, c->dense.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_set_field(deff2d_ctx *c, const double *field)
{
    if (!c || !field) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->Nx * (size_t)c->Ny;
    int rc = ensure(c, c->dense, n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->dense.p, field, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DomainView v = view(c);
    v.x_out = v.x_in;                  // write into the current iterate
    launch_inject_field(c->stream, v, c->dense.p);
    c->launches++;
    c->halo_valid = 0;                                       // slab mode: exchange before the next pass
    CU(cudaStreamSynchronize(c->stream));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_get_codes(deff2d_ctx *c, uint8_t *codes)
{
    if (!c || !codes) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)c->Nx * (size_t)c->Ny;
    int rc = ensure(c, c->dense8, n);
    if (rc) return rc;
    launch_extract_codes(c->stream, view(c), c->dense8.p);
    c->launches++;
    CU(cudaMemcpyAsync(codes, c->dense8.p, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_domain_info(deff2d_ctx *c, int64_t *Nx, int64_t *Ny, int *pathflag, double *porosity,
                                     double *SVF, double *LVF)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    if (Nx) *Nx = c->Nx;
    if (Ny) *Ny = c->Ny;
    if (pathflag) *pathflag = c->pathflag;
    if (porosity) *porosity = c->porosity;
    if (SVF || LVF) {
        double s, l;
        fracts3(c, c->Dphase[1], c->Dphase[0], &s, &l);
        if (SVF) *SVF = s;
        if (LVF) *LVF = l;
    }
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_sync(deff2d_ctx *c)
{
    if (!c) return DEFF2D_ERR_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_set_kernel(deff2d_ctx *c, int kernel, int tblock)
{
    if (!c || kernel < 0 || kernel > 4 || tblock < 0 || tblock > 16) return DEFF2D_ERR_ARG;
    c->tile_family = (kernel >= 3) ? kernel : 0;       // 3, 4: the two thread layouts of the tiled kernel by name; 0, 2: the default layout
    if (kernel >= 3) kernel = 2;
    c->kernel = kernel;
    c->tblock = tblock > 0 ? tblock : 1;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_set_graphs(deff2d_ctx *c, int enable)
{
    if (!c) return DEFF2D_ERR_ARG;
    c->use_graphs = enable != 0;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_get_graphs(const deff2d_ctx *c) { return (c && c->use_graphs) ? 1 : 0; }

DEFF2D_EXPORT int deff2d_set_floodfill(deff2d_ctx *c, int mode)
{
    if (!c || mode < 0 || mode > 2) return DEFF2D_ERR_ARG;
    c->floodfill_mode = mode;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_set_resident(deff2d_ctx *c, int mode)
{
    if (!c || mode < 0 || mode > 2) return DEFF2D_ERR_ARG;
    c->resident_mode = mode;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_set_batch_slots(deff2d_ctx *c, int max_slots)
{
    if (!c || max_slots < 0) return DEFF2D_ERR_ARG;
    c->batch_max_slots = max_slots;
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_default_depth(const deff2d_ctx *c) { return c ? c->k2_default_depth : DEFF2D_DEFAULT_DEPTH; }

DEFF2D_EXPORT int64_t deff2d_kernel_launches(const deff2d_ctx *c) { return c ? c->launches : 0; }

DEFF2D_EXPORT void *deff2d_stream(deff2d_ctx *c) { return c ? (void *)c->stream : nullptr; }

DEFF2D_EXPORT int deff2d_domain_buffers(deff2d_ctx *c, void **x_cur, void **x_next, int64_t *pitch, int64_t *rows)
{
    if (!c) return DEFF2D_ERR_ARG;
    if (!c->loaded) { set_error(c, "no domain loaded"); return DEFF2D_ERR_STATE; }
    if (x_cur) *x_cur = c->x[c->cur].p;
    if (x_next) *x_next = c->x[c->cur ^ 1].p;
    if (pitch) *pitch = c->pitch;
    if (rows) *rows = c->rows;
    return DEFF2D_OK;
}
