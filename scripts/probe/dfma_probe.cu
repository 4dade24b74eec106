// DFMA issue cost against the number of distinct register operands (debug aid, GPU box only):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/dfma_probe scripts/probe/dfma_probe.cu
// MODE 0: x = fma(x, b, h)        one operand the same register in every instruction
// MODE 1: x = fma(w, y, x)        three distinct register pairs per instruction, no operand shared by neighbours in the stream
// MODE 2: the same products ordered so that consecutive instructions share the multiplicand y (operand-reuse cache)
// MODE 3: as 1, but the multiplier w is shared by consecutive instructions
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int MODE>
__global__ void __launch_bounds__(256, 1) dfma(const double *__restrict__ init, double *out, long long *cyc, int iters)
{
    const int tid = threadIdx.x;
    double x[16], w[32], y[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { x[k] = init[tid + k]; y[k] = init[tid + 64 + k] * 1e-3; }
#pragma unroll
    for (int k = 0; k < 32; k++) w[k] = init[tid + 128 + k] * 0.5;
    const double b = init[7] * 0.5;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = fma(x[k], b, y[(k + r) % 16]);
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = fma(w[(k + 7 * r) % 32], y[(k + 3 * r + 1) % 16], x[k]);
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = fma(w[(k + 7 * r) % 32], y[(k / 4 + r) % 16], x[k]);
        } else {
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = fma(w[(k / 4 + 5 * r) % 32], y[(k + 3 * r + 1) % 16], x[k]);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += x[k];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int warps, const double *init, double *out, long long *cyc, int nsm)
{
    const int iters = 2000;
    dfma<MODE><<<nsm, warps * 32>>>(init, out, cyc, 10);
    dfma<MODE><<<nsm, warps * 32>>>(init, out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm; i++) avg += (double)h[i];
    avg /= nsm;
    // 80 DFMAs per thread and iteration; warps / 4 warps per scheduler
    printf("%-44s warps %d: %6.3f cycles per DFMA and scheduler\n", name, warps, avg / iters / 80.0 / (warps / 4.0));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    double *init, *out; long long *cyc;
    cudaMalloc(&init, 4096 * 8); cudaMalloc(&out, (size_t)nsm * 256 * 8); cudaMalloc(&cyc, 256 * 8);
    double h[4096];
    for (int i = 0; i < 4096; i++) h[i] = (double)rand() / RAND_MAX;
    cudaMemcpy(init, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {4, 8}) {
        run<0>("fma(x, b, y): one fixed operand", w, init, out, cyc, nsm);
        run<1>("fma(w, y, x): three distinct operands", w, init, out, cyc, nsm);
        run<2>("  multiplicand shared by 4 neighbours", w, init, out, cyc, nsm);
        run<3>("  multiplier shared by 4 neighbours", w, init, out, cyc, nsm);
    }
    return 0;
}
