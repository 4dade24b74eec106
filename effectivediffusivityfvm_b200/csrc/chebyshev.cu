// chebyshev.cu -- opt-in accelerated solver (the reference's own wish list, README.md:73-75).
//
// NON-PARITY MODE.  The reference's answer is "the damped-Jacobi iterate at the sweep where the relative Deff change
// between two checks fell below the tolerance" (Deff2D.cuh:1232-1276); on large or high-contrast domains that iterate is
// far from the solution of the linear system (config 5 stops at MaxIter with the change still at 1.7e-3).  For users who
// want the converged effective diffusivity this solver runs Chebyshev-accelerated Jacobi on the same discretisation:
//
//   x_{k+1} = x_k + tau_k D^-1 (b - A x_k),   tau_k = 1 / (theta - delta cos(pi (2 j_k - 1) / (2 m)))
//
// i.e. the reference's sweep (cuh:69-92) with a per-sweep relaxation factor instead of omega = 2/3: the m factors of a
// cycle are the reciprocal roots of the Chebyshev polynomial on [lambda_min, lambda_max] of D^-1 A, applied in the
// Lebedev-Finogenov order (stable for long cycles).  It needs no second iterate and no inner products, so it runs on
// the temporally blocked tiled kernel (sweep_tma.cu, CHEB variant: 8 steps per HBM pass).  lambda_max = 2 (Gershgorin:
// A is weakly diagonally dominant); lambda_min is estimated adaptively from the residual reduction each cycle achieves
// (Hageman & Young): a too large estimate only slows the lowest modes down, it never diverges.
//
// Stop rule: ||D^-1 (b - A x)||_2 <= rtol * ||D^-1 (b - A x0)||_2.  Single-GPU domains.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "context.h"

namespace deff2d {

#define XOFF DEFF2D_XOFF

// sum over the cells of (sum_f u_f x_f - x)^2 with the omega = 1 weights u: the squared 2-norm of the Jacobi-
// preconditioned residual D^-1 (b - A x) (b lives in the ghost columns, tables.cpp)
__global__ void __launch_bounds__(256) k_presid(DomainView d, double *out)
{
    __shared__ double sm[8];
    const long long n = d.Nx * d.Ny;
    double R = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / d.Nx, j = k - i * d.Nx;
        const long long b = (i + 1) * d.pitch + j + XOFF;
        const uint8_t *cc = d.code + b;
        const unsigned c0 = cc[0];
        const unsigned idx = (c0 & 3u) | ((cc[-1] & 3u) << 2) | ((cc[1] & 3u) << 4) | ((cc[d.pitch] & 3u) << 6) |
                             ((cc[-d.pitch] & 3u) << 8) | ((c0 & 4u) << 8) | ((c0 >> 3) << 11);
        const double *w = d.lut + (size_t)idx * 4;
        const double *xc = d.x_in + b;
        double t = w[0] * xc[-1];
        t = fma(w[1], xc[1], t);
        t = fma(w[2], xc[d.pitch], t);
        t = fma(w[3], xc[-d.pitch], t);
        const double r = t - xc[0];
        R = fma(r, r, R);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) R += __shfl_xor_sync(0xffffffffu, R, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = R;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += sm[w];
        atomicAdd(out, t);
    }
}

static int presid_norm(deff2d_ctx *c, double *norm)
{
    if (cudaMemsetAsync(c->d_scalar, 0, sizeof(double), c->stream) != cudaSuccess) return DEFF2D_ERR_CUDA;
    const long long n = c->Nx * c->Ny;
    int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
    k_presid<<<blocks, 256, 0, c->stream>>>(view(c), c->d_scalar);
    c->launches++;
    if (cudaMemcpyAsync(c->h_scalar, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        set_error(c, "Chebyshev solver: residual evaluation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return DEFF2D_ERR_CUDA;
    }
    *norm = std::sqrt(*c->h_scalar);
    return DEFF2D_OK;
}

// Lebedev-Finogenov ordering of 1..m (m a power of two): theta_1 = {1}; theta_2n = {theta_n(1), 2n + 1 - theta_n(1),
// theta_n(2), 2n + 1 - theta_n(2), ...}.  Applying the Chebyshev roots in this order keeps the partial products bounded.
static void lebedev_order(int m, std::vector<int> &ord)
{
    ord.assign(1, 1);
    for (int n = 1; n < m; n *= 2) {
        std::vector<int> next((size_t)2 * n);
        for (int k = 0; k < n; k++) { next[(size_t)2 * k] = ord[(size_t)k]; next[(size_t)2 * k + 1] = 2 * n + 1 - ord[(size_t)k]; }
        ord.swap(next);
    }
}

static double log_cosh(double t) { return t + std::log1p(std::exp(-2.0 * t)) - std::log(2.0); }

int chebyshev_solve(deff2d_ctx *c, double rtol, int64_t max_sweeps, int64_t *sweeps_out)
{
    *sweeps_out = 0;
    if (c->slab_domain || c->tile_list) { set_error(c, "the Chebyshev solver runs on single-GPU domains"); return DEFF2D_ERR_STATE; }
    if (c->omega != 1.0) { set_error(c, "Chebyshev solver: the weight table must hold omega = 1"); return DEFF2D_ERR_STATE; }
    if (!(rtol > 0)) rtol = 1e-8;
    int rc;
    double r0 = 0;
    if ((rc = presid_norm(c, &r0))) return rc;
    c->h_state->resid = 1.0;
    if (!(r0 > 0) || !std::isfinite(r0)) return DEFF2D_OK;              // already solved (or NaN: nothing to iterate on)
    const double b = 2.0;                                                // lambda_max bound (Gershgorin)
    // first guess for lambda_min: the homogeneous strip, (pi / (2 Nx))^2 / 2 of a 5-point Laplacian scaled by its
    // diagonal, never above 1e-2; heterogeneity only lowers it and the adaptive step below finds out
    double a = std::min(1e-2, 1.25 / ((double)c->NxG * (double)c->NxG));
    double r_prev = r0;
    int64_t done = 0;
    std::vector<int> ord;
    std::vector<double> tau;
    int stalls = 0;
    while (done < max_sweeps) {
        // cycle length: predicted reduction ~1e-2 per cycle on [a, b]; a power of two (ordering), 8 .. 65536
        const double y0 = (b + a) / (b - a), th0 = std::acosh(y0);
        int m = 8;
        while (m < 65536 && log_cosh(m * th0) < std::log(100.0)) m *= 2;
        if (done + m > max_sweeps) {                                     // last, shortened cycle: still a power of two >= 8
            while (m > 8 && done + m > max_sweeps) m /= 2;
            if (done + m > max_sweeps) break;
        }
        lebedev_order(m, ord);
        tau.resize((size_t)m);
        const double theta = 0.5 * (b + a), delta = 0.5 * (b - a);
        for (int k = 0; k < m; k++) tau[(size_t)k] = 1.0 / (theta - delta * std::cos(M_PI * (2.0 * ord[(size_t)k] - 1.0) / (2.0 * m)));
        for (int k = 0; k < m; k += 8)
            if ((rc = tma_cheb_pass(c, tau.data() + k))) return rc;
        done += m;
        double r = 0;
        if ((rc = presid_norm(c, &r))) return rc;
        c->h_state->resid = r / r0;
        if (!std::isfinite(r)) { set_error(c, "Chebyshev solver diverged (non-finite residual)"); return DEFF2D_ERR_STATE; }
        if (r <= rtol * r0) break;
        // adaptive lambda_min (Hageman & Young): the cycle should have reduced the residual by 1 / T_m(y0); if it did
        // worse, the slowest surviving mode sits below `a` at the lambda with T_m(y(lambda)) / T_m(y0) = r / r_prev
        const double rho = r / r_prev;
        const double logZ = std::log(rho) + log_cosh(m * th0);           // log of rho * T_m(y0)
        if (logZ > std::log(1.5)) {
            const double ac = (logZ > 30.0) ? logZ + std::log(2.0) : std::acosh(std::exp(logZ));
            const double y = std::cosh(ac / m);
            double a_new = 0.5 * (b + a - y * (b - a));
            if (!(a_new > 0)) a_new = a * 0.1;                            // the estimate left the interval: shrink boldly
            a = std::max(std::min(a_new, a * 0.9), a * 1e-3);
            stalls = 0;
        } else if (rho > 0.9 && ++stalls > 3) {                           // no progress although the bound says so: round-off floor
            break;
        }
        r_prev = r;
    }
    *sweeps_out = done;
    return DEFF2D_OK;
}

}  // namespace deff2d
