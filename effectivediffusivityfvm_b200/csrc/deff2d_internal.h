// deff2d_internal.h -- shared between the host-only and CUDA translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/deff2d.h"

#define DEFF2D_LUT_ENTRIES 2048
#define DEFF2D_CLUT_ENTRIES 1024        // compact table of the tiled sweep: p*256 + (pW|pE<<2|pS<<4|pN<<6), 768 = inert
#define DEFF2D_CLUT_USED 776             // entries a CTA stages in shared memory (769 rounded up)
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

// The compact per-stage table read by the tiled sweep: entry p*256 + n8 (p < 3, not pinned) is
// lut[p | n8 << 2]; entries >= 768 are zero (ghost and pinned cells: x' = (1-omega) x).
void compact_table(const double *lut, double *clut);
// Slot of compact entry e: the two low bits (which 8-bank group a 32-byte entry occupies in shared
// memory) are XOR-folded with the neighbour phases so that the few entries one warp gathers at a
// phase interface fall into different bank groups.
#define DEFF2D_CLUT_SLOT(e) ((e) ^ ((((e) >> 2) ^ ((e) >> 4) ^ ((e) >> 6) ^ ((e) >> 8)) & 3u))

// FloodFill (cuh:557-713) on a byte grid; returns PathFlag.
int floodfill(uint8_t *grid, int64_t Nx, int64_t Ny);

// repeated `+= 1/total` accumulation of the reference (cuh:402, cuh:437)
double accumulate_fraction(int64_t count, int64_t total);

}  // namespace deff2d
