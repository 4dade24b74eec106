// context.h -- the opaque per-device context behind the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kernels.cuh"

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;     // elements
};

struct deff2d_ctx {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr;
    std::string error;

    // resident domain
    bool loaded = false;
    int64_t Nx = 0, Ny = 0;          // local interior (incl. slab halo rows)
    int64_t NxG = 0, NyG = 0;        // global domain (dx = 1/NxG, dy = 1/NyG)
    int64_t pitch = 0, rows = 0;
    int64_t own_first = 0, own_rows = 0;     // rows this context owns (flux, counts)
    int64_t halo_above = 0, halo_below = 0, grow0 = 0;
    int nphase = 2;
    double Dphase[3] = {1, 0, 0};    // fluid, solid, gas of the current stage
    double CL = 0, CR = 1, omega = 2.0 / 3.0;
    int check_every = 10000;
    double residual_tol = 0;         // > 0: stop on the residual instead of the Deff change (non-parity mode)
    int solver = 0;                  // 1: Chebyshev-accelerated Jacobi (chebyshev.cu, non-parity mode)
    int cur = 0;                     // x[cur] holds the newest iterate
    int pathflag = 0;
    double porosity = 0;
    int64_t src_pixels = 0;
    bool in_batch = false;

    DevBuf<double> x[2];
    DevBuf<uint8_t> code, img, grid, dead, dense8;
    DevBuf<uint16_t> idx16;          // per-cell weight-table index derived from `code` (k_build_idx), read by the tiled sweep
    DevBuf<double> lut, dense;
    DevBuf<double> clut;             // compact per-stage weight tables of the tiled sweep (tables.cpp: compact_table), four planes
    DevBuf<uint32_t> clut32;         // the same tables as 32-bit halves, eight planes per stage (lo W,E,S,N, hi W,E,S,N): what K2 gathers
    int gather32 = 0;                // K2 gathers the weights as 32-bit halves (interface-rich media, set by the domain load)
    int lut_stages = 1;              // stages resident in lut / clut (packed batches: all stages of the mode)
    std::vector<uint8_t> h_grid;

    deff2d::SolveState *d_state = nullptr, *h_state = nullptr;
    deff2d::Counts *d_counts = nullptr, *h_counts = nullptr;
    double *d_scalar = nullptr, *h_scalar = nullptr;

    // kernel selection
    int kernel = 0;                  // 0 default, 1 simple, 2 TMA tiled
    int tblock = 1;
    int tile_family = 0;             // sweep_tma.cu thread layout: 3 = 4 x 4 patches, 4 = 2 x 8 patches, else the default
    int k2_default_family = DEFF2D_DEFAULT_TILE_FAMILY;
    int k2_default_depth = DEFF2D_DEFAULT_DEPTH;   // sweeps per HBM pass of kernel 0
    int64_t launches = 0;

    // TMA tiled sweep state (sweep_tma.cu)
    void *tma = nullptr;
    int64_t ghost_period = 0;        // every ghost_period-th column (from -1) is a Dirichlet ghost column: Nx + 1
                                     // for one domain, image width + 1 in a packed batch
    const uint32_t *tile_list = nullptr;   // device: ty << 16 | tx of the tiles to sweep (NULL: the whole tile grid)
    int tile_count = 0;

    // cluster-resident sweeps (resident.cu)
    void *resident = nullptr;
    int resident_mode = 0;           // 0: single domains of up to 256 x 256 cells run cluster-resident, 1: never, 2: packed batches too

    // packed-batch state (batch.cu)
    void *batch = nullptr;
    int batch_max_slots = 0;         // 0: library default
    int floodfill_mode = 0;          // 0 auto (device for >= 64 K cells), 1 host, 2 device
    int floodfill_passes = 0;        // device passes of the last FloodFill (diagnostic)

    bool use_graphs = true;          // replay long runs of passes as CUDA graphs (sweep_tma.cu: tma_passes)
    int grid_limit = 0;              // > 0: cap on the CTAs of a tiled pass (slab mode leaves SMs to NCCL)

    // multi-GPU slab state (slab.cu)
    void *slab = nullptr;
    int64_t halo_valid = 0;          // slab mode: halo rows that are still exact (a pass of depth T consumes T)
    bool slab_domain = false;        // the resident domain is one slab of a decomposed global domain
    int64_t store_row0 = 0, store_rows = 0;   // rows the tiled sweep's store maps cover (0 rows: all); slab peer mode: the own rows
};

namespace deff2d {

void set_error(deff2d_ctx *c, const char *fmt, ...);
DomainView view(const deff2d_ctx *c);
int solve_loop(deff2d_ctx *c, double tol, int64_t max_iter, bool verbose_checks, double print_div,
               int64_t *iters_out);
int solve_image_impl(deff2d_ctx *c, const uint8_t *gray, int W, int H, const deff2d_params *p,
                     deff2d_result *res, double *field, int image_number);

// sweep_tma.cu: enqueue up to min(n, tblock) sweeps with the TMA tiled kernel; *done = sweeps
// enqueued (0: this domain is not eligible, caller falls back to the streaming kernel).
int launch_sweep_tma(deff2d_ctx *c, int64_t n, int T, int64_t *done);
// one pass of depth T over an explicit tile list on `stream`, without flipping c->cur
int tma_pass(deff2d_ctx *c, int T, const uint32_t *list, int count, cudaStream_t stream);
// npasses passes of depth T on c->stream, flipping c->cur after each (CUDA graphs for long runs)
int tma_passes(deff2d_ctx *c, int T, int64_t npasses, const uint32_t *list, int count);
void tma_tile_geometry(const deff2d_ctx *c, int T, int *ow, int *oh);
int tma_cheb_pass(deff2d_ctx *c, const double tau[8]);
// slab peer mode: one pass with the halo push fused in; peer_args points to a PeerArgs (kernels.cuh)
int tma_peer_pass(deff2d_ctx *c, int T, const uint32_t *list, int count, const void *peer_args);   // chebyshev.cu: 8 Richardson steps with per-sweep factors (omega = 1 table)
void tma_destroy(deff2d_ctx *c);

// slab.cu
int slab_allreduce_q(deff2d_ctx *c);     // no-op unless the resident domain is a slab of a multi-rank group
void slab_peer_reset(deff2d_ctx *c);              // a (re)load ends the peer-memory exchange mode until it is attached again
bool slab_is_distributed(const deff2d_ctx *c);   // the resident domain is one slab of a group of >= 2 ranks
int slab_enqueue_sweeps(deff2d_ctx *c, int64_t n);
void slab_destroy(deff2d_ctx *c);
void batch_destroy(deff2d_ctx *c);

// resident.cu: cluster-resident sweeps (K5) of images of up to 256 x 256 cells
bool resident_eligible(deff2d_ctx *c, int64_t Nx, int64_t Ny);
int resident_sweeps(deff2d_ctx *c, int64_t n, int64_t Nx, int64_t Ny, int GX, const int *active, int nactive);
void resident_destroy(deff2d_ctx *c);

// chebyshev.cu: the opt-in accelerated solver (non-parity): sweeps until the relative residual is <= rtol or max_sweeps
int chebyshev_solve(deff2d_ctx *c, double rtol, int64_t max_sweeps, int64_t *sweeps_out);

// floodfill.cu: FloodFill (cuh:557-713) by label propagation on the device; blocks
int floodfill_device(deff2d_ctx *c, const uint8_t *img, int W, int Hsrc, int amp_x, int amp_y, int thr, uint8_t *st, int64_t Nx,
                     int64_t Ny, int *d_flags, int *h_flags, int *pathflag, int *passes, bool reference_quirk);
// the same for `count` images of a packed batch in the same launches (flags: int[1 + count])
int floodfill_device_batch(deff2d_ctx *c, const uint8_t *img, int W, int Hsrc, int amp_x, int amp_y, int thr, uint8_t *st,
                           int64_t Nx, int64_t Ny, int count, int *d_flags, int *h_flags, int *pathflags, int *passes,
                           bool reference_quirk);

// batch.cu: returns 1 when the resident small-image kernel does not cover the request
int batch_resident_solve(deff2d_ctx *c, const uint8_t *gray, int count, int W, int H,
                         const deff2d_params *p, deff2d_result *results, double *fields);

}  // namespace deff2d
