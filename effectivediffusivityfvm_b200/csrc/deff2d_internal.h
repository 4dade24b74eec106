// deff2d_internal.h -- shared between the host-only and CUDA translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/deff2d.h"

#define DEFF2D_LUT_ENTRIES 2048
#define DEFF2D_CLUT_ENTRIES 1024        // slots per plane of the compact table of the tiled sweep (see clut_slot)
#define DEFF2D_CLUT_INERT 1023u         // ghost and pinned cells: all four weights 0
#define DEFF2D_DEFAULT_TILE_FAMILY 4      // sweep_tma.cu: Family<> (2 x 8 cells per thread, 32 lanes across a 64 x 64 tile): what kernel 0 uses
#define DEFF2D_DEFAULT_DEPTH 6            // sweeps per HBM pass of kernel 0 (measured on B200, see enqueue_sweeps)
#define DEFF2D_XOFF 16            // interior column j lives at padded index j + XOFF
#define DEFF2D_PHASE_FLUID 0
#define DEFF2D_PHASE_SOLID 1
#define DEFF2D_PHASE_GAS 2
#define DEFF2D_PHASE_GHOST 3
#define DEFF2D_CODE_PINNED 4

#define DEFF2D_EXPORT extern "C" __attribute__((visibility("default")))

namespace deff2d {

void build_tables(const double Dphase[3], int64_t Nx, int64_t Ny, double CL, double CR, double omega,
                  double *lut, uint8_t *dead);

// The compact per-stage table read by the tiled sweep: four planes (wW, wE, wS, wN) of
// DEFF2D_CLUT_ENTRIES doubles.  Slots are DENSE for the cells that matter: a cell whose four
// neighbours are real phases has one of 32 (2-phase) or 243 (3-phase) interior slots; cells on the domain edge (a
// ghost neighbour) follow at base + p*256 + (pW | pE<<2 | pS<<4 | pN<<6).  2-phase interior: p*16 + (pW | pE<<1 |
// pS<<2 | pN<<3) -- the gathers of a warp at a phase interface touch two cache lines per plane.  3-phase interior:
// the 243 neighbourhoods RANKED by how common they are in a microstructure (slot_perm: single-phase neighbourhoods
// first, then one differing neighbour, then two adjacent ones, ...) instead of p*81 + base-3 digits, which spreads
// the common ones over six cache lines per plane.
#if defined(__CUDACC__)
#define DEFF2D_HD __host__ __device__
#else
#define DEFF2D_HD
#endif
struct SlotPerm { uint8_t v[256]; };          // dense interior numbering -> ranked slot (a permutation of 0..31 / 0..242)
const SlotPerm &slot_perm(int nphase);        // tables.cpp
DEFF2D_HD inline unsigned clut_slot(unsigned p, unsigned w, unsigned e, unsigned s, unsigned n, bool pinned, int nphase,
                                    const SlotPerm &perm)
{
    if (pinned || p == 3u) return DEFF2D_CLUT_INERT;
    const bool edge = (w == 3u) || (e == 3u) || (s == 3u) || (n == 3u);
    if (nphase == 2) {
        if (!edge) return perm.v[p * 16u + (w | (e << 1) | (s << 2) | (n << 3))];
        return 32u + p * 256u + (w | (e << 2) | (s << 4) | (n << 6));
    }
    if (!edge) return perm.v[p * 81u + (w + 3u * e + 9u * s + 27u * n)];
    return 243u + p * 256u + (w | (e << 2) | (s << 4) | (n << 6));
}
void compact_table(const double *lut, double *clut, int nphase);
// The compact table as 32-bit halves: per stage eight planes of DEFF2D_CLUT_ENTRIES words (low words of wW, wE, wS, wN,
// then their high words).  A 4-byte gather of a warp is conflict-free as long as the slots differ mod 32 (32 L1 banks of
// 4 bytes), an 8-byte gather only if they differ mod 16: on an all-interface medium (site percolation) the 8-byte
// gather needs 3.9 LSU cycles per load, two 4-byte gathers 2.
void split_table(const double *clut, uint32_t *clut32, int nstages);

// One continuation stage of the reference drivers: the coefficients it solves with, its stop rule, and whether it is
// a JacobiGPUPreCond stage (its Deff and time are not reported, cuh:1144-1159 vs cuh:1309-1311).
struct StageSpec {
    double Ds, Df, Dg;
    double stageD;          // the value the stage is named by (DCF of a 2-phase stage, DCG of a 3-phase one)
    double tol;
    int64_t max_iter;
    int precond;
    int defined_q8;         // the extra stage of strict_reference = 0 for quirk Q8 (2-phase single with Df < 10)
};
// The stage sequence of one image for p->mode: SingleSim (cuh:1714, 1759-1817), the BatchSim body (cuh:2004-2017),
// SingleSim3Phase / BatchSim3Phase (cuh:1492-1597).  Returns the number of stages (0: no solve at all, quirk Q8),
// -1 if there are more than `cap`.  The one place this sequence is written down.
int stage_list(const deff2d_params *p, StageSpec *out, int cap);

// FloodFill (cuh:557-713) on a byte grid; returns PathFlag.
// reference_quirk: keep the right-column seeding of cuh:601 (quirk Q11); false = flood from the left column only
int floodfill(uint8_t *grid, int64_t Nx, int64_t Ny, bool reference_quirk = true);

// repeated `+= 1/total` accumulation of the reference (cuh:402, cuh:437)
double accumulate_fraction(int64_t count, int64_t total);

}  // namespace deff2d
