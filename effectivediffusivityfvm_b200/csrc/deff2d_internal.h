// deff2d_internal.h -- shared between the host-only and CUDA translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/deff2d.h"

#define DEFF2D_LUT_ENTRIES 2048
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

// FloodFill (cuh:557-713) on a byte grid; returns PathFlag.
int floodfill(uint8_t *grid, int64_t Nx, int64_t Ny);

// repeated `+= 1/total` accumulation of the reference (cuh:402, cuh:437)
double accumulate_fraction(int64_t count, int64_t total);

}  // namespace deff2d
