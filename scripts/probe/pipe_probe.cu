// Throughput probe for the instruction mix of the tiled sweep (debug aid, GPU box only):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/pipe_probe scripts/probe/pipe_probe.cu
// One CTA of W warps per SM runs `iters` iterations of an unrolled instruction mix; prints SM cycles per iteration.
// Mixes (per thread and iteration): S 64-bit shuffles, L 8-byte shared loads, P 8-byte shared stores, F DFMAs (16 chains),
// optionally one __syncthreads.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int S, int L, int P, int F, bool BAR>
__global__ void __launch_bounds__(512, 1) mix(double *out, long long *cyc, int iters, double a, double b)
{
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = a * (tid + k);
    long long h[16];
#pragma unroll
    for (int k = 0; k < 16; k++) h[k] = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < S; k++)
            h[k % 16] ^= __double_as_longlong((k & 1) ? __shfl_up_sync(0xffffffffu, x[k % 16], 1) : __shfl_down_sync(0xffffffffu, x[k % 16], 1));
#pragma unroll
        for (int k = 0; k < L; k++) h[(k + 3) % 16] ^= __double_as_longlong(sm[(k * blockDim.x + tid + it) & 8191]);
#pragma unroll
        for (int k = 0; k < F; k++) x[k % 16] = fma(x[k % 16], b, __longlong_as_double(h[(k + 5) % 16]));
        if (F == 0) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = __longlong_as_double(__double_as_longlong(x[k]) + h[k]);
        }
#pragma unroll
        for (int k = 0; k < P; k++) sm[(k * blockDim.x + tid) & 8191] = x[k % 16];
        if (BAR) __syncthreads();
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += x[k] + (double)h[k];
    out[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int S, int L, int P, int F, bool BAR>
static void run(const char *name, int warps, double *out, long long *cyc, int nsm)
{
    const int iters = 2000;
    cudaFuncSetAttribute(mix<S, L, P, F, BAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    mix<S, L, P, F, BAR><<<nsm, warps * 32, 200 * 1024>>>(out, cyc, 10, 1.0, 0.5);
    mix<S, L, P, F, BAR><<<nsm, warps * 32, 200 * 1024>>>(out, cyc, iters, 1.0, 0.5);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int k = 0; k < nsm; k++) avg += (double)h[k];
    avg /= nsm;
    printf("%-34s warps %2d: %8.1f cycles / iteration  (%.2f per warp)\n", name, warps, avg / iters, avg / iters / warps);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount;
    double *out; long long *cyc;
    cudaMalloc(&out, (size_t)nsm * 512 * 8);
    cudaMalloc(&cyc, 256 * 8);
    for (int w : {4, 8, 12, 16}) {
        run<16, 0, 0, 0, false>("16 shfl64", w, out, cyc, nsm);
        run<0, 16, 0, 0, false>("16 lds64", w, out, cyc, nsm);
        run<0, 0, 16, 0, false>("16 sts64", w, out, cyc, nsm);
        run<0, 0, 0, 80, false>("80 dfma", w, out, cyc, nsm);
        run<16, 4, 4, 80, false>("16 shfl64+4 lds+4 sts+80 dfma", w, out, cyc, nsm);
        run<16, 4, 4, 80, true>("  the same + barrier", w, out, cyc, nsm);
        run<8, 8, 8, 80, true>("8 shfl64+8 lds+8 sts+80 dfma+bar", w, out, cyc, nsm);
        run<0, 20, 20, 80, true>("20 lds+20 sts+80 dfma+bar", w, out, cyc, nsm);
        run<16, 4, 4, 0, true>("16 shfl64+4 lds+4 sts+bar", w, out, cyc, nsm);
    }
    return 0;
}
