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

// Rank of the interior neighbourhoods of a 3-phase domain (deff2d_internal.h): by the number of neighbours in another
// phase, two differing neighbours round a corner (W/S, W/N, E/S, E/N) before two opposite ones, then by the dense number.
static SlotPerm make_perm(int nphase)
{
    SlotPerm out;
    std::memset(out.v, 0, sizeof(out.v));
    const unsigned np = (nphase == 2) ? 2u : 3u;
    struct Item { unsigned key, raw; };
    Item items[243];
    unsigned count = 0;
    for (unsigned p = 0; p < np; p++)
        for (unsigned n = 0; n < np; n++)
            for (unsigned s = 0; s < np; s++)
                for (unsigned e = 0; e < np; e++)
                    for (unsigned w = 0; w < np; w++) {
                        const unsigned raw = (np == 2) ? p * 16u + (w | (e << 1) | (s << 2) | (n << 3)) : p * 81u + (w + 3u * e + 9u * s + 27u * n);
                        const unsigned dw = w != p, de = e != p, ds = s != p, dn = n != p;
                        const unsigned d = dw + de + ds + dn;
                        const unsigned opposite = (d == 2 && ((dw && de) || (ds && dn))) ? 1u : 0u;
                        // within a class: the pattern of differing neighbours, then the centre phase, then what the neighbours are
                        const unsigned pattern = dw | (de << 1) | (ds << 2) | (dn << 3);
                        items[count].key = (d << 24) | (opposite << 23) | (pattern << 16) | (raw & 0xffffu);
                        if (np == 2) items[count].key = (d << 24) | (opposite << 23) | (pattern << 16) | p;
                        items[count].raw = raw;
                        count++;
                    }
    for (unsigned a = 1; a < count; a++) {                 // insertion sort: 243 items, once per process
        const Item it = items[a];
        unsigned b = a;
        while (b > 0 && items[b - 1].key > it.key) { items[b] = items[b - 1]; b--; }
        items[b] = it;
    }
    for (unsigned r = 0; r < count; r++) out.v[items[r].raw] = (uint8_t)r;
    // Two phases: the dense numbering p * 16 + (pW | pE<<1 | pS<<2 | pN<<3) stays -- one cache line per centre phase.
    // Measured on one B200 (same box, A/B): ranked against dense slots 884 / 877 GLUP/s on config 2 (3-phase), but 814 /
    // 828 on the 2-phase 4096^2 blob medium and 686 / 705 on 64 packed config-3 images.
    if (np == 2)
        for (unsigned r = 0; r < count; r++) out.v[r] = (uint8_t)r;
    return out;
}

const SlotPerm &slot_perm(int nphase)
{
    static const SlotPerm perm2 = make_perm(2), perm3 = make_perm(3);
    return nphase == 2 ? perm2 : perm3;
}

void compact_table(const double *lut, double *clut, int nphase)
{
    const SlotPerm &perm = slot_perm(nphase);
    std::memset(clut, 0, sizeof(double) * 4 * DEFF2D_CLUT_ENTRIES);
    const unsigned np = (nphase == 2) ? 2u : 3u;
    for (unsigned p = 0; p < np; p++)
        for (unsigned n8 = 0; n8 < 256; n8++) {
            const unsigned w = n8 & 3u, e = (n8 >> 2) & 3u, s = (n8 >> 4) & 3u, n = (n8 >> 6) & 3u;
            if ((w != 3u && w >= np) || (e != 3u && e >= np) || (s != 3u && s >= np) || (n != 3u && n >= np)) continue;
            const unsigned slot = clut_slot(p, w, e, s, n, false, nphase, perm);
            const double *src = lut + (size_t)(p | (n8 << 2)) * 4;
            for (int f = 0; f < 4; f++) clut[(size_t)f * DEFF2D_CLUT_ENTRIES + slot] = src[f];
        }
}

void split_table(const double *clut, uint32_t *clut32, int nstages)
{
    for (int k = 0; k < nstages; k++)
        for (int f = 0; f < 4; f++)
            for (int e = 0; e < DEFF2D_CLUT_ENTRIES; e++) {
                uint64_t bits;
                std::memcpy(&bits, clut + ((size_t)k * 4 + f) * DEFF2D_CLUT_ENTRIES + e, sizeof(bits));
                clut32[((size_t)k * 8 + f) * DEFF2D_CLUT_ENTRIES + e] = (uint32_t)bits;
                clut32[((size_t)k * 8 + 4 + f) * DEFF2D_CLUT_ENTRIES + e] = (uint32_t)(bits >> 32);
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

DEFF2D_EXPORT int deff2d_compact_table(const double *lut, int nphase, double *clut, uint16_t *slot)
{
    if ((nphase != 2 && nphase != 3) || (clut && !lut)) return DEFF2D_ERR_ARG;
    if (clut) deff2d::compact_table(lut, clut, nphase);
    if (slot) {
        const deff2d::SlotPerm &perm = deff2d::slot_perm(nphase);
        const unsigned np = (unsigned)nphase;
        for (unsigned idx = 0; idx < DEFF2D_LUT_ENTRIES; idx++) {
            const unsigned p = idx & 3u, w = (idx >> 2) & 3u, e = (idx >> 4) & 3u, s = (idx >> 6) & 3u, n = (idx >> 8) & 3u;
            const bool pinned = ((idx >> 10) & 1u) != 0;
            auto bad = [np](unsigned q) { return q != 3u && q >= np; };
            if (bad(p) || bad(w) || bad(e) || bad(s) || bad(n)) { slot[idx] = 0xffffu; continue; }
            slot[idx] = (uint16_t)deff2d::clut_slot(p, w, e, s, n, pinned, nphase, perm);
        }
    }
    return DEFF2D_OK;
}

DEFF2D_EXPORT int deff2d_split_table(const double *clut, int nstages, uint32_t *clut32)
{
    if (!clut || !clut32 || nstages < 1) return DEFF2D_ERR_ARG;
    deff2d::split_table(clut, clut32, nstages);
    return DEFF2D_OK;
}
