// The on-chip sweep loop of the tiled kernel in isolation (debug aid, GPU box only): no TMA, no prologue, no store.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/sweep_probe scripts/probe/sweep_probe.cu
// One CTA per SM holds a 64-column tile in registers (PX x PY cells per thread, 32 lanes across, one warp per PY rows)
// and runs `iters` passes of T sweeps: W / E halo by shuffles, N / S rows through the planar exchange buffer, one
// __syncthreads per sweep.  Prints SM cycles per sweep.  VARIANT selects the loop structure under test.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include "sweep_gen_2x8.inc"

template <int PX, int PY, int NW, int T, int VARIANT>
__global__ void __launch_bounds__(NW * 32, 1) sweep(const double *__restrict__ init, double *out, long long *cyc, int iters)
{
    constexpr int TH = PY * NW, PW = 32, CELLS = PX * 32 * TH;
    extern __shared__ double P0[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane, r0 = warp * PY;
    const int rN = (r0 > 0) ? r0 - 1 : r0;
    const int rS = (r0 + PY < TH) ? r0 + PY : r0 + PY - 1;
    double x[PY][PX], w[PY][PX][4], omc[PX];
#pragma unroll
    for (int py = 0; py < PY; py++)
#pragma unroll
        for (int px = 0; px < PX; px++) {
            x[py][px] = init[(tid * PY + py) * PX + px];
#pragma unroll
            for (int f = 0; f < 4; f++) w[py][px][f] = init[4096 + ((tid * 4 + f) * PY + py) * PX + px];
        }
#pragma unroll
    for (int px = 0; px < PX; px++) omc[px] = 1.0 / 3.0 + 1e-9 * init[px];

    auto publish_rows = [&](double *pb, const double (&top)[PX], const double (&bot)[PX]) {
#pragma unroll
        for (int px = 0; px < PX; px++) {
            pb[(px * TH + r0) * PW + g] = top[px];
            pb[(px * TH + r0 + PY - 1) * PW + g] = bot[px];
        }
    };
    auto row_update = [&](int py, double hw, double he, const double (&upv)[PX], const double (&dnv)[PX],
                          const double (&cur)[PX], double (&o)[PX]) {
        double left = hw;
#pragma unroll
        for (int px = 0; px < PX; px++) {
            const double c = cur[px];
            const double right = (px == PX - 1) ? he : cur[px + 1];
            double r = omc[px] * c;
            r = fma(w[py][px][0], left, r);
            r = fma(w[py][px][1], right, r);
            r = fma(w[py][px][2], dnv[px], r);
            r = fma(w[py][px][3], upv[px], r);
            o[px] = r;
            left = c;
        }
    };
    // VARIANT 6 / 7: warps whose cells all share one table entry keep 4 weights instead of 4 per cell: consecutive DFMAs
    // then share the weight operand (reuse cache).  7: only the odd warps take that path (mixed tile).
    double u0 = w[0][0][0], u1 = w[0][0][1], u2 = w[0][0][2], u3 = w[0][0][3];
    auto row_update_u = [&](double hw, double he, const double (&upv)[PX], const double (&dnv)[PX],
                            const double (&cur)[PX], double (&o)[PX]) {
        double left = hw;
#pragma unroll
        for (int px = 0; px < PX; px++) {
            const double c = cur[px];
            const double right = (px == PX - 1) ? he : cur[px + 1];
            double r = omc[px] * c;
            r = fma(u0, left, r);
            r = fma(u1, right, r);
            r = fma(u2, dnv[px], r);
            r = fma(u3, upv[px], r);
            o[px] = r;
            left = c;
        }
    };
    auto one_sweep_u = [&](int s) {
        const double *pr = P0 + ((s - 1) & 1) * CELLS;
        double hW[PY], hE[PY], hN[PX], hS[PX];
#pragma unroll
        for (int py = 0; py < PY; py++) {
            hW[py] = __shfl_up_sync(0xffffffffu, x[py][PX - 1], 1);
            hE[py] = __shfl_down_sync(0xffffffffu, x[py][0], 1);
        }
#pragma unroll
        for (int px = 0; px < PX; px++) {
            hN[px] = pr[(px * TH + rN) * PW + g];
            hS[px] = pr[(px * TH + rS) * PW + g];
        }
        double up[PX];
#pragma unroll
        for (int px = 0; px < PX; px++) up[px] = hN[px];
#pragma unroll
        for (int py = 0; py < PY; py++) {
            double cur[PX];
#pragma unroll
            for (int px = 0; px < PX; px++) cur[px] = x[py][px];
            if (py == PY - 1) row_update_u(hW[py], hE[py], up, hS, cur, x[py]);
            else row_update_u(hW[py], hE[py], up, x[py + 1], cur, x[py]);
#pragma unroll
            for (int px = 0; px < PX; px++) up[px] = cur[px];
        }
        publish_rows(P0 + (s & 1) * CELLS, x[0], x[PY - 1]);
        __syncthreads();
    };
    auto one_sweep = [&](int s) {
        const double *pr = P0 + ((s - 1) & 1) * CELLS;
        double hW[PY], hE[PY], hN[PX], hS[PX];
#pragma unroll
        for (int py = 0; py < PY; py++) {
            hW[py] = __shfl_up_sync(0xffffffffu, x[py][PX - 1], 1);
            hE[py] = __shfl_down_sync(0xffffffffu, x[py][0], 1);
        }
#pragma unroll
        for (int px = 0; px < PX; px++) {
            hN[px] = pr[(px * TH + rN) * PW + g];
            hS[px] = pr[(px * TH + rS) * PW + g];
        }
        double up[PX];
#pragma unroll
        for (int px = 0; px < PX; px++) up[px] = hN[px];
#pragma unroll
        for (int py = 0; py < PY; py++) {
            double cur[PX];
#pragma unroll
            for (int px = 0; px < PX; px++) cur[px] = x[py][px];
            if (py == PY - 1) row_update(py, hW[py], hE[py], up, hS, cur, x[py]);
            else row_update(py, hW[py], hE[py], up, x[py + 1], cur, x[py]);
#pragma unroll
            for (int px = 0; px < PX; px++) up[px] = cur[px];
        }
        publish_rows(P0 + (s & 1) * CELLS, x[0], x[PY - 1]);
        __syncthreads();
    };

    // VARIANT 2: "scatter" order -- the products of one source value are consecutive in the instruction stream, so the
    // value stays in the operand-reuse cache (a DFMA with three register-file operands holds the dispatch port for three
    // cycles, with one operand from the reuse cache for two: dfma_probe.cu).  Per-cell order of accumulation: N, W, own, E, S.
    auto scatter_sweep = [&](int s) {
        const double *pr = P0 + ((s - 1) & 1) * CELLS;
        double hW[PY], hE[PY], hN[PX], hS[PX];
#pragma unroll
        for (int py = 0; py < PY; py++) {
            hW[py] = __shfl_up_sync(0xffffffffu, x[py][PX - 1], 1);
            hE[py] = __shfl_down_sync(0xffffffffu, x[py][0], 1);
        }
#pragma unroll
        for (int px = 0; px < PX; px++) {
            hN[px] = pr[(px * TH + rN) * PW + g];
            hS[px] = pr[(px * TH + rS) * PW + g];
        }
        double aprev[PX], acur[PX], anext[PX];
#pragma unroll
        for (int px = 0; px < PX; px++) { acur[px] = w[0][px][3] * hN[px]; aprev[px] = 0.0; anext[px] = 0.0; }
#pragma unroll
        for (int py = 0; py < PY; py++) {
            acur[0] = fma(w[py][0][0], hW[py], acur[0]);
#pragma unroll
            for (int px = 0; px < PX; px++) {
                const double c = x[py][px];
                if (py > 0) aprev[px] = fma(w[py - 1][px][2], c, aprev[px]);           // S of the cell above
                if (py + 1 < PY) anext[px] = w[py + 1][px][3] * c;                       // N of the cell below
                if (px + 1 < PX) acur[px + 1] = fma(w[py][px + 1][0], c, acur[px + 1]);  // W of the right neighbour
                acur[px] = fma(omc[px], c, acur[px]);                                    // own
                if (px > 0) acur[px - 1] = fma(w[py][px - 1][1], c, acur[px - 1]);       // E of the left neighbour
            }
            acur[PX - 1] = fma(w[py][PX - 1][1], hE[py], acur[PX - 1]);
#pragma unroll
            for (int px = 0; px < PX; px++) {
                if (py > 0) x[py - 1][px] = aprev[px];
                aprev[px] = acur[px];
                acur[px] = anext[px];
            }
        }
#pragma unroll
        for (int px = 0; px < PX; px++) x[PY - 1][px] = fma(w[PY - 1][px][2], hS[px], aprev[px]);
        publish_rows(P0 + (s & 1) * CELLS, x[0], x[PY - 1]);
        __syncthreads();
    };

    // VARIANT 4 / 5: scatter order with the W / E shuffles of the next sweep issued as soon as a row is final (carried
    // across the barrier in registers); the sweep loop is not unrolled so that the order below survives scheduling.
    // ORDER 0: rows top-down.  ORDER 1: row 0 (which starts with the N halo from shared memory) after rows 1 and 2.
    double cW[PY], cE[PY];
    auto pipelined_sweep = [&](int s, auto order_tag) {
        constexpr int ORDER = decltype(order_tag)::value;
        const double *pr = P0 + ((s - 1) & 1) * CELLS;
        double *pw = P0 + (s & 1) * CELLS;
        double hN[PX], hS[PX];
#pragma unroll
        for (int px = 0; px < PX; px++) {
            hN[px] = pr[(px * TH + rN) * PW + g];
            hS[px] = pr[(px * TH + rS) * PW + g];
        }
        double acc[PY][PX];
        // products of source row py that go to the row below (its N term starts that row's accumulator)
        auto down = [&](int py) {
#pragma unroll
            for (int px = 0; px < PX; px++) if (py + 1 < PY) acc[py + 1][px] = w[py + 1][px][3] * x[py][px];
        };
        // W halo, then per source value: S of the row above, W of the right neighbour, own, E of the left neighbour; E halo
        auto row = [&](int py, bool with_down) {
            acc[py][0] = fma(w[py][0][0], cW[py], acc[py][0]);
#pragma unroll
            for (int px = 0; px < PX; px++) {
                const double c = x[py][px];
                if (py > 0) acc[py - 1][px] = fma(w[py - 1][px][2], c, acc[py - 1][px]);
                if (with_down && py + 1 < PY) acc[py + 1][px] = w[py + 1][px][3] * c;
                if (px + 1 < PX) acc[py][px + 1] = fma(w[py][px + 1][0], c, acc[py][px + 1]);
                acc[py][px] = fma(omc[px], c, acc[py][px]);
                if (px > 0) acc[py][px - 1] = fma(w[py][px - 1][1], c, acc[py][px - 1]);
            }
            acc[py][PX - 1] = fma(w[py][PX - 1][1], cE[py], acc[py][PX - 1]);
        };
        auto finish = [&](int py) {          // row py is final: next sweep's W / E halo of the neighbours, publish edge rows
            if (py == PY - 1) {
#pragma unroll
                for (int px = 0; px < PX; px++) acc[py][px] = fma(w[py][px][2], hS[px], acc[py][px]);
            }
            cW[py] = __shfl_up_sync(0xffffffffu, acc[py][PX - 1], 1);
            cE[py] = __shfl_down_sync(0xffffffffu, acc[py][0], 1);
            if (py == 0) {
#pragma unroll
                for (int px = 0; px < PX; px++) pw[(px * TH + r0) * PW + g] = acc[0][px];
            }
            if (py == PY - 1) {
#pragma unroll
                for (int px = 0; px < PX; px++) pw[(px * TH + r0 + PY - 1) * PW + g] = acc[py][px];
            }
        };
        if (ORDER == 0) {
#pragma unroll
            for (int px = 0; px < PX; px++) acc[0][px] = w[0][px][3] * hN[px];
#pragma unroll
            for (int py = 0; py < PY; py++) {
                row(py, true);
                if (py > 0) finish(py - 1);
            }
            finish(PY - 1);
        } else {
            down(0);
            down(1);
            // rows 1 and 2 without their S products into rows 0 and 1 ... (those need row 0 / row 1 further along)
            // simple form: start row 1's own terms first, row 0 after
            acc[1][0] = fma(w[1][0][0], cW[1], acc[1][0]);
#pragma unroll
            for (int px = 0; px < PX; px++) {
                const double c = x[1][px];
                if (px + 1 < PX) acc[1][px + 1] = fma(w[1][px + 1][0], c, acc[1][px + 1]);
                acc[1][px] = fma(omc[px], c, acc[1][px]);
                if (px > 0) acc[1][px - 1] = fma(w[1][px - 1][1], c, acc[1][px - 1]);
            }
            acc[1][PX - 1] = fma(w[1][PX - 1][1], cE[1], acc[1][PX - 1]);
#pragma unroll
            for (int px = 0; px < PX; px++) acc[0][px] = w[0][px][3] * hN[px];
            // row 0: W halo, own, E, then S from row 1
            acc[0][0] = fma(w[0][0][0], cW[0], acc[0][0]);
#pragma unroll
            for (int px = 0; px < PX; px++) {
                const double c = x[0][px];
                if (px + 1 < PX) acc[0][px + 1] = fma(w[0][px + 1][0], c, acc[0][px + 1]);
                acc[0][px] = fma(omc[px], c, acc[0][px]);
                if (px > 0) acc[0][px - 1] = fma(w[0][px - 1][1], c, acc[0][px - 1]);
            }
            acc[0][PX - 1] = fma(w[0][PX - 1][1], cE[0], acc[0][PX - 1]);
#pragma unroll
            for (int px = 0; px < PX; px++) acc[0][px] = fma(w[0][px][2], x[1][px], acc[0][px]);
            finish(0);
#pragma unroll
            for (int py = 2; py < PY; py++) {
                row(py, true);
                finish(py - 1);
            }
            finish(PY - 1);
        }
#pragma unroll
        for (int py = 0; py < PY; py++)
#pragma unroll
            for (int px = 0; px < PX; px++) x[py][px] = acc[py][px];
        __syncthreads();
    };
    if (VARIANT == 4 || VARIANT == 5) {
#pragma unroll
        for (int py = 0; py < PY; py++) {
            cW[py] = __shfl_up_sync(0xffffffffu, x[py][PX - 1], 1);
            cE[py] = __shfl_down_sync(0xffffffffu, x[py][0], 1);
        }
    }

    publish_rows(P0, x[0], x[PY - 1]);
    __syncthreads();
    if constexpr (VARIANT == 8 && PX == 2 && PY == 8) {
        // hand-scheduled order from gen_sweep.py (compile with -Xptxas -O1 so that the order survives); two sweeps per
        // loop iteration (xa -> xb -> xa), no copies
        double (&xa)[PY][PX] = x;
        double xb[PY][PX];
        double hWa[PY], hEa[PY], hWb[PY], hEb[PY], hN[PX], hS[PX];
        const double (&fac)[PX] = omc;
#pragma unroll
        for (int py = 0; py < PY; py++) {
            hWa[py] = __shfl_up_sync(0xffffffffu, x[py][PX - 1], 1);
            hEa[py] = __shfl_down_sync(0xffffffffu, x[py][0], 1);
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll 1
            for (int s = 1; s <= T; s += 2) {
                {
                    const double *pr = P0;
                    double *pw = P0 + CELLS;
                    hN[0] = pr[(0 * TH + rN) * PW + g]; hN[1] = pr[(1 * TH + rN) * PW + g];
                    hS[0] = pr[(0 * TH + rS) * PW + g]; hS[1] = pr[(1 * TH + rS) * PW + g];
                    SWEEP_2X8_A
                    __syncthreads();
                }
                {
                    const double *pr = P0 + CELLS;
                    double *pw = P0;
                    hN[0] = pr[(0 * TH + rN) * PW + g]; hN[1] = pr[(1 * TH + rN) * PW + g];
                    hS[0] = pr[(0 * TH + rS) * PW + g]; hS[1] = pr[(1 * TH + rS) * PW + g];
                    SWEEP_2X8_B
                    __syncthreads();
                }
            }
        }
        const long long t1 = clock64();
        double acc = 0;
#pragma unroll
        for (int py = 0; py < PY; py++)
#pragma unroll
            for (int px = 0; px < PX; px++) acc += x[py][px];
        out[blockIdx.x * blockDim.x + tid] = acc;
        if (tid == 0) cyc[blockIdx.x] = t1 - t0;
        return;
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (VARIANT == 0) {
#pragma unroll
            for (int s = 1; s <= T; s++) one_sweep(s);
        } else if (VARIANT == 2) {
#pragma unroll
            for (int s = 1; s <= T; s++) scatter_sweep(s);
        } else if (VARIANT == 6 || (VARIANT == 7 && (warp & 1))) {
#pragma unroll
            for (int s = 1; s <= T; s++) one_sweep_u(s);
        } else if (VARIANT == 7) {
#pragma unroll
            for (int s = 1; s <= T; s++) one_sweep(s);
        } else if (VARIANT == 4) {
#pragma unroll 1
            for (int s = 1; s <= T; s++) pipelined_sweep(s, std::integral_constant<int, 0>());
        } else if (VARIANT == 5) {
#pragma unroll 1
            for (int s = 1; s <= T; s++) pipelined_sweep(s, std::integral_constant<int, 1>());
        } else if (VARIANT == 10) {
#pragma unroll 2
            for (int s = 1; s <= T; s++) scatter_sweep(s);
        } else if (VARIANT == 11) {
#pragma unroll 3
            for (int s = 1; s <= T; s++) scatter_sweep(s);
        } else if (VARIANT == 12) {
#pragma unroll 2
            for (int s = 1; s <= T; s++) one_sweep(s);
        } else if (VARIANT == 13) {
#pragma unroll 3
            for (int s = 1; s <= T; s++) one_sweep(s);
        } else if (VARIANT == 3) {
#pragma unroll 1
            for (int s = 1; s <= T; s++) scatter_sweep(s);
        } else {
#pragma unroll 1
            for (int s = 1; s <= T; s++) one_sweep(s);
        }
    }
    const long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int py = 0; py < PY; py++)
#pragma unroll
        for (int px = 0; px < PX; px++) acc += x[py][px];
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PX, int PY, int NW, int T, int VARIANT>
static void run(const char *name, const double *init, double *out, long long *cyc, int nsm)
{
    const int iters = 500;
    const size_t smem = (size_t)2 * PX * 32 * PY * NW * 8;
    auto k = sweep<PX, PY, NW, T, VARIANT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<nsm, NW * 32, smem>>>(init, out, cyc, 5);
    k<<<nsm, NW * 32, smem>>>(init, out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    const double per_sweep = avg / iters / T;
    printf("%-44s %7.1f cycles / sweep, %6.4f cycles / cell\n", name, per_sweep, per_sweep / (32.0 * PX * PY * NW));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    double *init, *out; long long *cyc;
    const size_t n = 1 << 16;
    cudaMalloc(&init, n * 8); cudaMalloc(&out, (size_t)nsm * 512 * 8); cudaMalloc(&cyc, 256 * 8);
    double *h = (double *)malloc(n * 8);
    for (size_t i = 0; i < n; i++) h[i] = (i < 4096) ? (double)rand() / RAND_MAX : 0.16 + 1e-3 * rand() / RAND_MAX;
    cudaMemcpy(init, h, n * 8, cudaMemcpyHostToDevice);
    run<2, 8, 8, 6, 0>("2x8, 8 warps, T=6 unrolled (the kernel's loop)", init, out, cyc, nsm);
    run<2, 8, 8, 6, 1>("2x8, 8 warps, T=6 not unrolled", init, out, cyc, nsm);
    run<2, 8, 8, 6, 2>("2x8, 8 warps, scatter order, unrolled", init, out, cyc, nsm);
    run<2, 8, 8, 6, 3>("2x8, 8 warps, scatter order, not unrolled", init, out, cyc, nsm);
    run<4, 4, 8, 6, 2>("4x4 (128 columns), scatter order", init, out, cyc, nsm);
    run<2, 8, 8, 6, 4>("2x8, scatter, carried shuffles, rows in order", init, out, cyc, nsm);
    run<2, 8, 8, 6, 5>("2x8, scatter, carried shuffles, row 0 late", init, out, cyc, nsm);
    run<2, 8, 8, 6, 6>("2x8, all warps with 4 uniform weights", init, out, cyc, nsm);
    run<2, 8, 8, 6, 7>("2x8, odd warps with 4 uniform weights", init, out, cyc, nsm);
    run<2, 8, 8, 6, 8>("2x8, hand-scheduled scatter order", init, out, cyc, nsm);
    run<2, 8, 8, 6, 10>("2x8, scatter order, unroll 2", init, out, cyc, nsm);
    run<2, 8, 8, 6, 11>("2x8, scatter order, unroll 3", init, out, cyc, nsm);
    run<2, 8, 8, 6, 12>("2x8, gather order, unroll 2", init, out, cyc, nsm);
    run<2, 8, 8, 6, 13>("2x8, gather order, unroll 3", init, out, cyc, nsm);
    run<2, 6, 12, 6, 0>("2x6, 12 warps, unrolled", init, out, cyc, nsm);
    run<2, 4, 16, 6, 0>("2x4, 16 warps, unrolled", init, out, cyc, nsm);
    run<4, 4, 8, 6, 0>("4x4 (32 lanes: 128 columns), 8 warps", init, out, cyc, nsm);
    run<2, 8, 4, 6, 0>("2x8, 4 warps", init, out, cyc, nsm);
    return 0;
}
