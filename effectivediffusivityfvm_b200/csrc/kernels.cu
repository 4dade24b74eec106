// kernels.cu -- hand-written sm_100a kernels of the effective-diffusivity path:
//   init_domain   threshold + mesh amplification + ghost ring + x0     (replaces cuh:1773-1785,
//                 1557-1578, 1730-1734 and, with the LUT, DiscretizeMatrix2D* cuh:715-902)
//   sweep_simple  K3: matrix-free damped-Jacobi sweep, 1 sweep / HBM pass   (cuh:69-92, 1281)
//   flux / check  K4: boundary-flux Deff + the reference stop rule           (cuh:1243-1276)
//   extract/inject/residual helpers
// The TMA-staged, temporally blocked sweep (K2) lives in sweep_tma.cu.
#include "kernels.cuh"

#include <cmath>

namespace deff2d {

#define XOFF DEFF2D_XOFF

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// init_domain: one thread per 16 consecutive padded cells of a row (one uint4 code store,
// eight double2 stores per iterate buffer): coalesced 128-bit stores along x.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_init_domain(const uint8_t *__restrict__ img, int W, int Hsrc, int amp_x, int amp_y, int nphase,
              long long grow0, long long img_row0, const uint8_t *__restrict__ grid,
              double *__restrict__ x0, double *__restrict__ x1, uint8_t *__restrict__ code,
              long long Nx, long long Ny, long long pitch, long long NxG, double CL, double CR,
              long long own_first, long long own_rows, Counts *counts)
{
    const long long groups_per_row = pitch / 16;
    const long long total = (Ny + 2) * groups_per_row;
    unsigned long long cnt[3] = {0, 0, 0}, npinned = 0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long r = g / groups_per_row;             // padded row
        const long long c0 = (g - r * groups_per_row) * 16; // first padded column of the group
        const long long i = r - 1;                          // local interior row
        const bool row_inside = (i >= 0 && i < Ny);
        // source row of this amplified row (cuh:1777: i / MeshIncreaseY)
        long long srow = 0;
        if (row_inside) {
            srow = (grow0 + i) / amp_y - img_row0;   // grow0: global amplified row of local row 0
            if (srow < 0) srow = 0;
            if (srow >= Hsrc) srow = Hsrc - 1;
        }
        const bool own = row_inside && i >= own_first && i < own_first + own_rows;
        unsigned int cw[4] = {0, 0, 0, 0};
        double xv[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const long long c = c0 + k;
            const long long j = c - XOFF;
            unsigned int cd = DEFF2D_PHASE_GHOST;
            double v = 0.0;
            if (row_inside && j >= 0 && j < Nx) {
                const unsigned char p = img[srow * W + (int)(j / amp_x)];   // cuh:1778
                if (nphase == 2) cd = (p < 150) ? DEFF2D_PHASE_FLUID : DEFF2D_PHASE_SOLID;   // cuh:1779
                else cd = (p > 200) ? DEFF2D_PHASE_SOLID
                                    : ((p < 50) ? DEFF2D_PHASE_GAS : DEFF2D_PHASE_FLUID);   // cuh:1565-1576
                if (own) cnt[cd]++;
                if (grid) {
                    const unsigned char gv = grid[i * Nx + j];
                    if (gv == 1 || gv == 2) { cd |= DEFF2D_CODE_PINNED; if (own) npinned++; }   // cuh:750
                }
                // cuh:1732, un-contracted so that x0 is bit-identical to the host expression
                v = __dadd_rn(__dmul_rn(__ddiv_rn((double)j, (double)NxG), __dsub_rn(CR, CL)), CL);
            } else if (row_inside && (j == -1 || j == Nx)) {
                v = 1.0;            // Dirichlet ghost: weight carries CL / CR (tables.cpp)
            }
            cw[k >> 2] |= cd << ((k & 3) * 8);
            xv[k] = v;
        }
        *reinterpret_cast<uint4 *>(code + r * pitch + c0) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
        double2 *o0 = reinterpret_cast<double2 *>(x0 + r * pitch + c0);
        double2 *o1 = reinterpret_cast<double2 *>(x1 + r * pitch + c0);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            o0[k] = make_double2(xv[2 * k], xv[2 * k + 1]);
            // the second buffer only needs the ghost values; interior is overwritten by sweep 1
            const long long j0 = c0 + 2 * k - XOFF, j1 = j0 + 1;
            const double g0 = (row_inside && (j0 == -1 || j0 == Nx)) ? 1.0 : 0.0;
            const double g1 = (row_inside && (j1 == -1 || j1 == Nx)) ? 1.0 : 0.0;
            o1[k] = make_double2(g0, g1);
        }
    }
    if (counts) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const unsigned long long s = warp_sum_u64(cnt[k]);
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(&counts->phase[k], s);
        }
        const unsigned long long sp = warp_sum_u64(npinned);
        if ((threadIdx.x & 31) == 0 && sp) atomicAdd(&counts->pinned, sp);
    }
}

void launch_init_domain(cudaStream_t s, const uint8_t *img, int W, int Hsrc, int amp_x, int amp_y,
                        int nphase, int64_t grow0, int64_t img_row0, const uint8_t *grid, double *x0,
                        double *x1, uint8_t *code, int64_t Nx, int64_t Ny, int64_t pitch, int64_t NxG,
                        double CL, double CR, int64_t own_first, int64_t own_rows, Counts *counts)
{
    const long long total = (Ny + 2) * (pitch / 16);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_init_domain<<<blocks, 256, 0, s>>>(img, W, Hsrc, amp_x, amp_y, nphase, grow0, img_row0, grid, x0, x1,
                                         code, Nx, Ny, pitch, NxG, CL, CR, own_first, own_rows, counts);
}

// ------------------------------------------------------------------------------------------
// build_idx: the compact weight-table index of every padded cell (bits 0-9), its stage (10-13, from
// code bits 3-6) and the ghost-column flag (15).  One thread per 8 cells of a row, one 16-byte
// store.  Runs once per image load; the tiled sweep then reads 2 B per cell instead of
// recomputing the index from 5 code bytes per cell per pass.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_build_idx(const uint8_t *__restrict__ code, uint16_t *__restrict__ idx16, long long Nx, long long Ny,
            long long pitch, long long period, int nphase, const __grid_constant__ SlotPerm perm, Counts *counts)
{
    unsigned cells = 0, mixed = 0;
    const long long groups_per_row = pitch / 8;
    const long long total = (Ny + 2) * groups_per_row;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long r = g / groups_per_row;
        const long long c0 = (g - r * groups_per_row) * 8;
        const uint8_t *row = code + r * pitch;
        const uint8_t *up = (r > 0) ? row - pitch : nullptr;
        const uint8_t *dn = (r < Ny + 1) ? row + pitch : nullptr;
        unsigned out[8];
        unsigned prev = (c0 > 0) ? row[c0 - 1] : 3u;
        unsigned cur = row[c0];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const long long c = c0 + k;
            const unsigned next = (c + 1 < pitch) ? row[c + 1] : 3u;
            const unsigned n = up ? up[c] : 3u, s = dn ? dn[c] : 3u;
            // slot in the compact table (deff2d_internal.h: clut_slot); stage in bits 10-13
            unsigned v = clut_slot(cur & 3u, prev & 3u, next & 3u, s & 3u, n & 3u, (cur & 4u) != 0, nphase, perm);
            v |= ((cur >> 3) & 15u) << 10;
            if ((cur & 3u) != 3u && !(cur & 4u)) {             // live cell: how many of them sit at a phase interface
                cells++;
                const unsigned ph = cur & 3u;
                mixed += ((prev & 3u) != ph) | ((next & 3u) != ph) | ((s & 3u) != ph) | ((n & 3u) != ph);
            }
            const long long j = c - XOFF;                      // interior column
            if (j >= -1 && j <= Nx && (j + 1) % period == 0) v |= 0x8000u;
            out[k] = v;
            prev = cur;
            cur = next;
        }
        *reinterpret_cast<uint4 *>(idx16 + r * pitch + c0) =
            make_uint4(out[0] | (out[1] << 16), out[2] | (out[3] << 16), out[4] | (out[5] << 16), out[6] | (out[7] << 16));
    }
    if (counts) {
        const unsigned long long c2 = warp_sum_u64(cells), m2 = warp_sum_u64(mixed);
        if ((threadIdx.x & 31) == 0 && c2) { atomicAdd(&counts->idx_cells, c2); atomicAdd(&counts->idx_mixed, m2); }
    }
}

void launch_build_idx(cudaStream_t s, const uint8_t *code, uint16_t *idx16, int64_t Nx, int64_t Ny, int64_t pitch,
                      int64_t ghost_period, int nphase, Counts *counts)
{
    const long long total = (Ny + 2) * (pitch / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_build_idx<<<blocks, 256, 0, s>>>(code, idx16, Nx, Ny, pitch, ghost_period, nphase, slot_perm(nphase), counts);
}

// calcPorosity's counting loop (cuh:399-405) as a reduction over the source image
__global__ void __launch_bounds__(256)
k_count_below(const uint8_t *__restrict__ img, long long n, int thr, Counts *counts)
{
    unsigned long long c = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x)
        c += (img[k] < thr) ? 1u : 0u;
    c = warp_sum_u64(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counts->below150, c);
}

void launch_count_below(cudaStream_t s, const uint8_t *img, int64_t n, int thr, Counts *counts)
{
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k_count_below<<<blocks, 256, 0, s>>>(img, n, thr, counts);
}

// ------------------------------------------------------------------------------------------
// K3 sweep_simple.  Each thread updates two x-adjacent cells of one row (16 B loads/stores);
// a block covers 256 columns x SIMPLE_ROWS rows.  Neighbour reuse is left to L1/L2.
// ------------------------------------------------------------------------------------------
#define SIMPLE_ROWS 8

__device__ __forceinline__ void lut_load(const double *__restrict__ lut, unsigned idx, double w[4])
{
    const double2 *p = reinterpret_cast<const double2 *>(lut + (size_t)idx * 4);
    const double2 a = __ldg(p), b = __ldg(p + 1);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

__global__ void __launch_bounds__(128)
k_sweep_simple(DomainView d, const int *__restrict__ stop)
{
    if (stop && *stop) return;
    const long long j = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;   // first of two columns
    if (j >= d.Nx) return;
    const long long i0 = (long long)blockIdx.y * SIMPLE_ROWS;
    const bool two = (j + 1 < d.Nx);
    const double om = d.om;
#pragma unroll 2
    for (int rr = 0; rr < SIMPLE_ROWS; rr++) {
        const long long i = i0 + rr;
        if (i >= d.Ny) break;
        const long long base = (i + 1) * d.pitch + j + XOFF;          // even -> 16 B aligned
        const uint8_t *cc = d.code + base;
        const double *xc = d.x_in + base;
        // phase codes: W, the two cells, E; N and S of both cells
        const unsigned cW = cc[-1] & 3u, c0 = cc[0], c1 = cc[1], cE = cc[2] & 3u;
        const unsigned n0 = cc[-d.pitch] & 3u, n1 = cc[-d.pitch + 1] & 3u;
        const unsigned s0 = cc[d.pitch] & 3u, s1 = cc[d.pitch + 1] & 3u;
        const unsigned idx0 = (c0 & 3u) | (cW << 2) | ((c1 & 3u) << 4) | (s0 << 6) | (n0 << 8) | ((c0 & 4u) << 8) | ((c0 >> 3) << 11);
        const unsigned idx1 = (c1 & 3u) | ((c0 & 3u) << 2) | (cE << 4) | (s1 << 6) | (n1 << 8) | ((c1 & 4u) << 8) | ((c1 >> 3) << 11);
        double w0[4], w1[4];
        lut_load(d.lut, idx0, w0);
        lut_load(d.lut, idx1, w1);
        const double2 c = *reinterpret_cast<const double2 *>(xc);
        const double2 n = *reinterpret_cast<const double2 *>(xc - d.pitch);
        const double2 s = *reinterpret_cast<const double2 *>(xc + d.pitch);
        const double xw = xc[-1], xe = xc[2];
        // x' = (1-w) x + sum_f w_f x_f          (cuh:76-89 with A, b folded into the LUT)
        double r0 = om * c.x;
        r0 = fma(w0[0], xw, r0);
        r0 = fma(w0[1], c.y, r0);
        r0 = fma(w0[2], s.x, r0);
        r0 = fma(w0[3], n.x, r0);
        double r1 = om * c.y;
        r1 = fma(w1[0], c.x, r1);
        r1 = fma(w1[1], xe, r1);
        r1 = fma(w1[2], s.y, r1);
        r1 = fma(w1[3], n.y, r1);
        if (two) *reinterpret_cast<double2 *>(d.x_out + base) = make_double2(r0, r1);
        else d.x_out[base] = r0;     // odd Nx: the partner column is the right Dirichlet ghost
    }
}

void launch_sweep_simple(cudaStream_t s, const DomainView &d, const int *stop)
{
    dim3 block(128);
    dim3 grid((unsigned)((d.Nx + 255) / 256), (unsigned)((d.Ny + SIMPLE_ROWS - 1) / SIMPLE_ROWS));
    k_sweep_simple<<<grid, block, 0, s>>>(d, stop);
}

// ------------------------------------------------------------------------------------------
// K4 flux: Q1 = sum_i D(i,0) (x(i,0)-CL)/(dx/2), Q2 = sum_i D(i,Nx-1) (CR-x(i,Nx-1))/(dx/2)
// (cuh:1252-1260).  One block; every thread owns a fixed strided subset of rows and the
// partials are combined in a fixed order, so the result is deterministic.
// ------------------------------------------------------------------------------------------
struct FluxParams { double D[3]; double CL, CR, half_dx; };
// fused stop rule (single-GPU domains): cuh:1263-1276 and the loop condition of cuh:1232 right behind the flux sums
struct CheckParams { int enable; double two_ny, tol; long long iter_index; };

__device__ __forceinline__ void apply_check(SolveState *st, double two_ny, double CL, double CR, double tol, long long iter_index)
{
    if (st->stop) return;
    const double qAvg = (st->q[0] + st->q[1]) / two_ny;          // cuh:1263
    const double deffNew = qAvg / (CR - CL);                     // cuh:1264
    const double change = (st->deff_old - deffNew) / (st->deff_old);   // cuh:1265
    st->deff_new = deffNew;
    st->change = change;
    st->conv = change;                                           // cuh:1275
    st->deff_old = deffNew;                                      // cuh:1273
    if (st->nchecks < 256) st->trace[st->nchecks] = deffNew;
    st->nchecks++;
    if (!(tol < fabs(change))) {                                 // cuh:1232 (NaN ends the loop)
        st->stop = 1;
        st->stop_iter = iter_index + 1;                          // iterCount after the increment of cuh:1289
    }
}

__global__ void __launch_bounds__(1024)
k_flux(DomainView d, FluxParams fp, CheckParams ck, long long row_first, long long nrows, SolveState *st)
{
    __shared__ double s1[32], s2[32];
    double q1 = 0, q2 = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (long long r = threadIdx.x; r < nrows; r += blockDim.x) {
        const long long i = row_first + r;
        const long long base = (i + 1) * d.pitch + XOFF;
        const unsigned cl = d.code[base], cr = d.code[base + d.Nx - 1];
        const double Dl = fp.D[cl & 3u], Dr = fp.D[cr & 3u];
        double tl = Dl * (d.x_in[base] - fp.CL) / fp.half_dx;
        double tr = Dr * (fp.CR - d.x_in[base + d.Nx - 1]) / fp.half_dx;
        // quirk Q13: a non-pinned boundary cell with D = 0 has A0 = 0, the reference iterate
        // there is NaN and 0 * NaN poisons the sum
        if (Dl == 0.0 && !(cl & 4u)) tl = nan;
        if (Dr == 0.0 && !(cr & 4u)) tr = nan;
        q1 += tl;
        q2 += tr;
    }
    q1 = warp_sum(q1);
    q2 = warp_sum(q2);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s1[wid] = q1; s2[wid] = q2; }
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        q1 = lane < nw ? s1[lane] : 0.0;
        q2 = lane < nw ? s2[lane] : 0.0;
        q1 = warp_sum(q1);
        q2 = warp_sum(q2);
        if (lane == 0) {
            st->q[0] = q1; st->q[1] = q2;
            if (ck.enable) apply_check(st, ck.two_ny, fp.CL, fp.CR, ck.tol, ck.iter_index);
        }
    }
}

static FluxParams flux_params(const double Dphase[3], double CL, double CR, int64_t NxG)
{
    FluxParams fp;
    fp.D[0] = Dphase[0]; fp.D[1] = Dphase[1]; fp.D[2] = Dphase[2];
    fp.CL = CL; fp.CR = CR;
    fp.half_dx = (1.0 / (double)NxG) / 2.0;      // dx / 2.0, cuh:1256
    return fp;
}

void launch_flux(cudaStream_t s, const DomainView &d, const double Dphase[3], double CL, double CR,
                 int64_t NxG, int64_t row_first, int64_t nrows, SolveState *st)
{
    CheckParams ck = {0, 0.0, 0.0, 0};
    k_flux<<<1, 1024, 0, s>>>(d, flux_params(Dphase, CL, CR, NxG), ck, row_first, nrows, st);
}

// K4 fused: boundary flux, Deff, signed change and the stop flag in one launch (single-GPU domains; a decomposed
// domain needs the all-reduce of {Q1, Q2} between the two halves and uses launch_flux + launch_check)
void launch_flux_check(cudaStream_t s, const DomainView &d, const double Dphase[3], double CL, double CR, int64_t NxG,
                       int64_t NyG, int64_t row_first, int64_t nrows, double tol, long long iter_index, SolveState *st)
{
    CheckParams ck = {1, 2.0 * (double)NyG, tol, iter_index};
    k_flux<<<1, 1024, 0, s>>>(d, flux_params(Dphase, CL, CR, NxG), ck, row_first, nrows, st);
}

// the stop rule alone, on all-reduced {Q1, Q2} (multi-GPU slabs)
__global__ void k_check(SolveState *st, double two_ny, double CL, double CR, double tol, long long iter_index)
{
    apply_check(st, two_ny, CL, CR, tol, iter_index);
}

void launch_check(cudaStream_t s, SolveState *st, int64_t NyG, double CL, double CR, double tol,
                  long long iter_index)
{
    k_check<<<1, 1, 0, s>>>(st, 2.0 * (double)NyG, CL, CR, tol, iter_index);
}

__global__ void k_reset_state(SolveState *st)
{
    st->q[0] = st->q[1] = 0;
    st->deff_old = 5;        // cuh:1172
    st->deff_new = 1;        // cuh:1171
    st->change = 100.0;      // cuh:1173
    st->resid = 0;
    st->stop_iter = -1;
    st->stop = 0;
    st->nchecks = 0;
    // conv is deliberately kept: JacobiGPU only overwrites it at a check (cuh:1275)
}

void launch_reset_state(cudaStream_t s, SolveState *st) { k_reset_state<<<1, 1, 0, s>>>(st); }

// ------------------------------------------------------------------------------------------
// field / code extraction
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cell_index(const uint8_t *__restrict__ code, long long base, long long pitch)
{
    const unsigned c = code[base];
    return (c & 3u) | ((code[base - 1] & 3u) << 2) | ((code[base + 1] & 3u) << 4) |
           ((code[base + pitch] & 3u) << 6) | ((code[base - pitch] & 3u) << 8) | ((c & 4u) << 8) |
           ((c >> 3) << 11);       // bits 3-7: continuation stage (packed batches, batch.cu)
}

__global__ void __launch_bounds__(256) k_extract_field(DomainView d, double *__restrict__ dense)
{
    const long long n = d.Nx * d.Ny;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / d.Nx, j = k - i * d.Nx;
        const long long base = (i + 1) * d.pitch + j + XOFF;
        const unsigned idx = cell_index(d.code, base, d.pitch);
        dense[k] = d.dead[idx] ? nan : d.x_in[base];
    }
}

void launch_extract_field(cudaStream_t s, const DomainView &d, double *dense)
{
    const long long n = d.Nx * d.Ny;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_extract_field<<<blocks, 256, 0, s>>>(d, dense);
}

__global__ void __launch_bounds__(256) k_inject_field(DomainView d, const double *__restrict__ dense)
{
    const long long n = d.Nx * d.Ny;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / d.Nx, j = k - i * d.Nx;
        d.x_out[(i + 1) * d.pitch + j + XOFF] = dense[k];
    }
}

void launch_inject_field(cudaStream_t s, const DomainView &d, const double *dense)
{
    const long long n = d.Nx * d.Ny;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_inject_field<<<blocks, 256, 0, s>>>(d, dense);
}

__global__ void __launch_bounds__(256) k_extract_codes(DomainView d, uint8_t *__restrict__ dense)
{
    const long long n = d.Nx * d.Ny;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / d.Nx, j = k - i * d.Nx;
        dense[k] = d.code[(i + 1) * d.pitch + j + XOFF];
    }
}

void launch_extract_codes(cudaStream_t s, const DomainView &d, uint8_t *dense)
{
    const long long n = d.Nx * d.Ny;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_extract_codes<<<blocks, 256, 0, s>>>(d, dense);
}

// ------------------------------------------------------------------------------------------
// K7 residual: the reference's (dead) definition, cuh:451-494 -- note that it scales every
// face, x or y, by dy/dx and harmonic means with dx/2 weights; kept verbatim.
// ------------------------------------------------------------------------------------------
struct ResidParams { double k[3][3]; double D[3]; double CL, CR, dy_over_dx, dy_over_hdx; double inv_n; };

__global__ void __launch_bounds__(256) k_residual(DomainView d, ResidParams rp, double *out)
{
    __shared__ double sm[8];
    const long long n = d.Nx * d.Ny;
    double R = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / d.Nx, j = k - i * d.Nx;
        const long long b = (i + 1) * d.pitch + j + XOFF;
        const unsigned p = d.code[b] & 3u;
        const double c = d.x_in[b];
        double qW, qE, qN, qS;
        if (j == 0) qW = rp.dy_over_hdx * rp.D[p] * (c - rp.CL);
        else qW = rp.dy_over_dx * rp.k[p][d.code[b - 1] & 3u] * (c - d.x_in[b - 1]);
        if (j == d.Nx - 1) qE = rp.dy_over_hdx * rp.D[p] * (rp.CR - c);
        else qE = rp.dy_over_dx * rp.k[p][d.code[b + 1] & 3u] * (d.x_in[b + 1] - c);
        if (i == 0) qN = 0;
        else qN = rp.dy_over_dx * rp.k[d.code[b - d.pitch] & 3u][p] * (c - d.x_in[b - d.pitch]);
        if (i == d.Ny - 1) qS = 0;
        else qS = rp.dy_over_dx * rp.k[d.code[b + d.pitch] & 3u][p] * (d.x_in[b + d.pitch] - c);
        R += fabs(qW - qE + qN - qS);
    }
    R = warp_sum(R);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = R;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += sm[w];
        atomicAdd(out, t * rp.inv_n);
    }
}

void launch_residual(cudaStream_t s, const DomainView &d, const double Dphase[3], double CL, double CR,
                     int64_t NxG, int64_t NyG, double *partial)
{
    ResidParams rp;
    const double dx = 1.0 / (double)NxG, dy = 1.0 / (double)NyG;
    for (int a = 0; a < 3; a++) {
        rp.D[a] = Dphase[a];
        for (int b = 0; b < 3; b++)
            rp.k[a][b] = (dx / 2 + dx / 2) / ((dx / 2) / Dphase[a] + (dx / 2) / Dphase[b]);   // cuh:358
    }
    rp.CL = CL; rp.CR = CR;
    rp.dy_over_dx = dy / dx;
    rp.dy_over_hdx = dy / (dx / 2);
    rp.inv_n = 1.0 / ((double)NxG * (double)NyG);
    const long long n = d.Nx * d.Ny;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_residual<<<blocks, 256, 0, s>>>(d, rp, partial);
}

}  // namespace deff2d
