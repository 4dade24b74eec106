// kernels.cuh -- device-side data model and kernel launchers of libdeff2d (sm_100a).
//
// HBM layout of one resident domain (see DESIGN.md "Data layout"):
//   x[2]   FP64 ping-pong iterate, (Ny+2) rows x pitch doubles.  Interior cell (i,j) lives at
//          [(i+1)*pitch + j + XOFF]; column XOFF-1 and XOFF+Nx are the Dirichlet ghost
//          columns (value 1.0, their weight carries CL / CR), rows 0 and Ny+1 are the no-flux
//          ghost rows (weight 0).  pitch is a multiple of 16 doubles (128 B rows, TMA needs 16 B).
//   code   one byte per cell in the same padded geometry (pitch bytes per row):
//          bits 0-1 phase (0 fluid, 1 solid, 2 gas, 3 ghost), bit 2 pinned (3-phase Grid in {1,2}).
//   idx16  two bytes per cell, same geometry, derived from `code` (k_build_idx): bits 0-9 the
//          cell's slot in the compact weight table (deff2d_internal.h: clut_slot; 1023: ghost or
//          pinned), bits 10-13 its continuation stage, bit 15 "Dirichlet ghost column".  Read by
//          the tiled sweep (K2) together with `clut`, the compact planar form of `lut`.
//   lut    2048 x 4 doubles: sweep weights for every (p, pW, pE, pS, pN, pinned), per stage.
// The reference's A[n][5] + b[n] (48 B/cell, cuh:1396-1397) are never materialised.
#pragma once
#include <cuda_runtime.h>

#include "deff2d_internal.h"

namespace deff2d {

struct DomainView {
    double *x_in;            // iterate read by the sweep
    double *x_out;           // iterate written by the sweep
    const uint8_t *code;
    const double *lut;       // [2048][4]
    const uint8_t *dead;     // [2048]
    int64_t Nx, Ny;          // local interior size
    int64_t pitch;           // elements per padded row (x and code)
    double om;               // 1 - omega
};

// Device-resident convergence state of the reference loop (cuh:1167-1176, 1243-1276).
struct SolveState {
    double q[2];             // Q1, Q2 of the last flux evaluation (cuh:1252-1260)
    double deff_old;         // deffOld, seeded with 5 (cuh:1172)
    double deff_new;         // deffNew (cuh:1171)
    double change;           // percentChange, seeded with 100 (cuh:1173)
    double conv;             // myImg->conv (cuh:1275)
    double resid;            // optional residual diagnostic (cuh:451-494)
    long long stop_iter;     // iterCount at which the stop rule fired (-1: not yet)
    int stop;                // 1: sweeps already enqueued behind the check become no-ops
    int nchecks;
    double trace[256];       // deffNew at every check (diagnostics / parity tests)
};

// PEER variant (slab.cu, peer-memory halo exchange fused into the sweep): a rank sweeps only its own rows; the tiles
// next to a neighbouring slab come early in the tile list (behind one interior tile per CTA, which hides the flag
// wait) and, besides their local store, copy the rows the neighbour needs straight into its halo rows over NVLink
// (peer-mapped memory).  When the last of those tiles is counted the neighbours' flags are raised; every pass waits
// for the neighbours' flags of the pass before ahead of its first boundary tile.
struct PeerArgs {
    double *up[2], *down[2];          // neighbours' iterate buffers (by parity), at halo row 0 / column 0 of the interior; NULL: no neighbour
    long long *flag_up, *flag_down;   // the neighbours' flag words this rank raises (peer memory)
    const long long *flag_local;      // [0]: raised by the upper neighbour, [1]: by the lower one
    unsigned long long *counters;     // boundary tiles pushed and counted in this pass (reset by the thread that counts the last one)
    long long *pass_no;               // passes completed on this rank since the domain was loaded (advanced by the same thread)
    long long pitch;
    int above, own, H, nboundary, Nx;
    int lead;                         // list entries in front of the boundary tiles (one interior tile per CTA, or 0)
};

struct Counts {
    unsigned long long phase[4];   // cells per phase on the amplified grid (own rows only)
    unsigned long long below150;   // source pixels < 150 (calcPorosity, cuh:401)
    unsigned long long pinned;
    unsigned long long idx_cells, idx_mixed;   // k_build_idx: live cells, and those whose four neighbours are not all of the cell's phase
};

// ---- launchers (all asynchronous on `s`) -------------------------------------------------

// threshold + mesh amplification + ghost ring + x0 (cuh:1773-1785, 1557-1578, 1730-1734)
void launch_init_domain(cudaStream_t s, const uint8_t *img, int W, int Hsrc, int amp_x, int amp_y,
                        int nphase, int64_t grow0 /* global amplified row of local interior row 0 */,
                        int64_t img_row0 /* global source row of the first row held in img */,
                        const uint8_t *grid /* Ny*Nx: 0/1/2 from FloodFill, or NULL */,
                        double *x0, double *x1, uint8_t *code, int64_t Nx, int64_t Ny, int64_t pitch,
                        int64_t NxG, double CL, double CR, int64_t own_first, int64_t own_rows,
                        Counts *counts);
// per-cell table index of every padded cell from the codes (see idx16 above)
void launch_build_idx(cudaStream_t s, const uint8_t *code, uint16_t *idx16, int64_t Nx, int64_t Ny, int64_t pitch,
                      int64_t ghost_period, int nphase, Counts *counts /* NULL: no statistics */);
void launch_count_below(cudaStream_t s, const uint8_t *img, int64_t n, int thr, Counts *counts);

// K3: plain streaming sweep, one sweep per HBM pass (cuh:69-92 matrix-free)
void launch_sweep_simple(cudaStream_t s, const DomainView &d, const int *stop);

// K4: boundary-flux sums over local rows [row_first, row_first+nrows) (cuh:1252-1260)
void launch_flux(cudaStream_t s, const DomainView &d, const double Dphase[3], double CL, double CR,
                 int64_t NxG, int64_t row_first, int64_t nrows, SolveState *st);
// K4 fused: flux sums + Deff + stop rule in one launch (single-GPU domains)
void launch_flux_check(cudaStream_t s, const DomainView &d, const double Dphase[3], double CL, double CR, int64_t NxG,
                       int64_t NyG, int64_t row_first, int64_t nrows, double tol, long long iter_index, SolveState *st);
// stop rule of cuh:1263-1276 + cuh:1232 on the (all-reduced) Q1, Q2
void launch_check(cudaStream_t s, SolveState *st, int64_t NyG, double CL, double CR, double tol,
                  long long iter_index);
void launch_reset_state(cudaStream_t s, SolveState *st);

// concentration map as the reference downloads it (cuh:1300), NaN in dead cells (Q13)
void launch_extract_field(cudaStream_t s, const DomainView &d, double *dense);
void launch_inject_field(cudaStream_t s, const DomainView &d, const double *dense);
void launch_extract_codes(cudaStream_t s, const DomainView &d, uint8_t *dense);
// K7: mean |qW - qE + qN - qS| (cuh:451-494)
void launch_residual(cudaStream_t s, const DomainView &d, const double Dphase[3], double CL, double CR,
                     int64_t NxG, int64_t NyG, double *partial /* 1 double, zeroed */);

}  // namespace deff2d
