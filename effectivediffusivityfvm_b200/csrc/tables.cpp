// tables.cpp -- host-side coefficient tables of libdeff2d (no CUDA needed).
//
// The reference materialises a 5-diagonal matrix A[n][5] and a right-hand side b[n] on
// the host (DiscretizeMatrix2D, Deff2D.cuh:815-902; DiscretizeMatrix2D_ImpSolid,
// cuh:715-812) and streams 48 B/cell of them through the sweep kernel (cuh:69-92).
// Both are functions of the phase of a cell and of its four neighbours only, so here
// they collapse into one lookup table per continuation stage:
//
//   index = p | pW<<2 | pE<<4 | pS<<6 | pN<<8 | pinned<<10         (11 bits)
//   lut[index] = { wW, wE, wS, wN },   w_f = (omega/A0) * c_f
//
// with phases 0 = fluid, 1 = solid, 2 = gas, 3 = ghost.  A ghost W/E neighbour is the
// Dirichlet face of the first/last column (its "c" is the b contribution of cuh:856 /
// cuh:864 and the ghost cell holds the value 1.0); a ghost S/N neighbour is the no-flux
// wall (c = 0, cuh:875-888).  The sweep then is x' = (1-omega) x + sum_f w_f x_f for
// every cell, boundary or not.
#include "deff2d_internal.h"

#include <cmath>
#include <cstring>

namespace deff2d {

// cuh:347-360
static inline double whm(double w1, double w2, double x1, double x2)
{
    return (w1 + w2) / (w1 / x1 + w2 / x2);
}

void build_tables(const double Dphase[3], int64_t Nx, int64_t Ny, double CL, double CR, double omega,
                  double *lut, uint8_t *dead)
{
    const double dx = 1.0 / (double)Nx, dy = 1.0 / (double)Ny;   // cuh:1682-1683
    for (int idx = 0; idx < DEFF2D_LUT_ENTRIES; idx++) {
        double *w = lut + (size_t)idx * 4;
        w[0] = w[1] = w[2] = w[3] = 0.0;
        if (dead) dead[idx] = 0;
        const int p = idx & 3, pW = (idx >> 2) & 3, pE = (idx >> 4) & 3, pS = (idx >> 6) & 3,
                  pN = (idx >> 8) & 3, pinned = (idx >> 10) & 1;
        // pinned cell: identity row A0 = 1, b = 0 (cuh:750-752) -> x' = (1-omega) x.
        // ghost cell: never updated by the kernels; zero weights keep it inert.
        if (pinned || p == 3) continue;
        const double D = Dphase[p];
        double A0 = 0.0, cW = 0.0, cE = 0.0, cS = 0.0, cN = 0.0;
        // ---- x direction, cuh:849-873 (same expression order) ----
        if (pW == 3) {                       // j == 0
            const double dxe = dx, dxw = dx / 2;
            double ke = (pE == 3) ? 0.0 : whm(dxe / 2, dxe / 2, D, Dphase[pE]);
            const double kw = D;
            cE = ke * dy / dxe;
            A0 += (ke * dy / dxe + kw * dy / dxw);
            cW = CL * kw * dy / dxw;         // b of cuh:856, carried by the ghost value 1.0
        } else if (pE == 3) {                // j == Nx-1
            const double dxw = dx, dxe = dx / 2;
            const double kw = whm(dxw / 2, dxw / 2, D, Dphase[pW]);
            const double ke = D;
            cW = kw * dy / dxw;
            A0 += (ke * dy / dxe + kw * dy / dxw);
            cE = CR * ke * dy / dxe;         // b of cuh:864
        } else {
            const double dxw = dx, dxe = dx;
            const double kw = whm(dxw / 2, dxw / 2, D, Dphase[pW]);
            const double ke = whm(dxe / 2, dxe / 2, D, Dphase[pE]);
            cW = kw * dy / dxw;
            cE = ke * dy / dxe;
            A0 += (ke * dy / dxe + kw * dy / dxw);
        }
        // ---- y direction, cuh:875-897 ----
        if (pN == 3) {                       // i == 0
            const double dys = dy;
            const double ks = (pS == 3) ? 0.0 : whm(dys / 2, dys / 2, Dphase[pS], D);
            cS = ks * dx / dys;
            A0 += (ks * dx / dys);
        } else if (pS == 3) {                // i == Ny-1
            const double dyn = dy;
            const double kn = whm(dyn / 2, dyn / 2, D, Dphase[pN]);
            cN = kn * dx / dyn;
            A0 += kn * dx / dyn;
        } else {
            const double dyn = dy, dys = dy;
            const double kn = whm(dyn / 2, dyn / 2, D, Dphase[pN]);
            const double ks = whm(dys / 2, dys / 2, Dphase[pS], D);
            cS = ks * dx / dys;
            cN = kn * dx / dyn;
            A0 += (kn * dx / dyn + ks * dx / dys);
        }
        if (A0 == 0.0) {                     // quirk Q13: w/0 * 0 = NaN in the reference
            if (dead) dead[idx] = 1;
            continue;
        }
        const double dinv = omega / A0;      // w / A[row*5+0], cuh:89
        w[0] = dinv * cW;
        w[1] = dinv * cE;
        w[2] = dinv * cS;
        w[3] = dinv * cN;
    }
}

void compact_table(const double *lut, double *clut, int nphase)
{
    std::memset(clut, 0, sizeof(double) * 4 * DEFF2D_CLUT_ENTRIES);
    const unsigned np = (nphase == 2) ? 2u : 3u;
    for (unsigned p = 0; p < np; p++)
        for (unsigned n8 = 0; n8 < 256; n8++) {
            const unsigned w = n8 & 3u, e = (n8 >> 2) & 3u, s = (n8 >> 4) & 3u, n = (n8 >> 6) & 3u;
            if ((w != 3u && w >= np) || (e != 3u && e >= np) || (s != 3u && s >= np) || (n != 3u && n >= np)) continue;
            const unsigned slot = clut_slot(p, w, e, s, n, false, nphase);
            const double *src = lut + (size_t)(p | (n8 << 2)) * 4;
            for (int f = 0; f < 4; f++) clut[(size_t)f * DEFF2D_CLUT_ENTRIES + slot] = src[f];
        }
}

}  // namespace deff2d

DEFF2D_EXPORT int deff2d_build_tables(double Ds, double Df, double Dg, int64_t Nx, int64_t Ny, double CL,
                                   double CR, double omega, double *lut, uint8_t *dead)
{
    if (!lut || Nx < 1 || Ny < 1) return DEFF2D_ERR_ARG;
    if (!(omega > 0)) omega = 2.0 / 3.0;
    const double D[3] = {Df, Ds, Dg};
    deff2d::build_tables(D, Nx, Ny, CL, CR, omega, lut, dead);
    return DEFF2D_OK;
}
