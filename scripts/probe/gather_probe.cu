// L1 gather cost of the weight-table lookups (debug aid, GPU box only):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/gather_probe scripts/probe/gather_probe.cu
// 8 warps per SM; every thread does 64 dependent-free 8-byte __ldg per iteration from a small table with a given
// pattern of slots over the 32 lanes.  Prints SM cycles per warp-wide load instruction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void __launch_bounds__(256, 1) gather(const double *__restrict__ table, const int *__restrict__ slots, long long *cyc,
                                                  double *out, int iters)
{
    const int lane = threadIdx.x & 31;
    int idx[16];
#pragma unroll
    for (int k = 0; k < 16; k++) idx[k] = slots[k * 32 + lane];
    long long acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const double *q = table + idx[k];
            acc ^= __double_as_longlong(__ldg(q)) ^ __double_as_longlong(__ldg(q + 1024)) ^ __double_as_longlong(__ldg(q + 2048)) ^
                   __double_as_longlong(__ldg(q + 3072));
        }
#pragma unroll
        for (int k = 0; k < 16; k++) idx[k] ^= (int)(acc == 0x123456789LL);      // (never true) keeps the loads inside the loop
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (double)acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nsm = p.multiProcessorCount, iters = 500;
    double *table, *out; int *slots; long long *cyc;
    cudaMalloc(&table, 4096 * 8); cudaMalloc(&out, (size_t)nsm * 256 * 8); cudaMalloc(&slots, 512 * 4); cudaMalloc(&cyc, 256 * 8);
    std::vector<double> h(4096);
    for (auto &v : h) v = (double)rand() / RAND_MAX;
    cudaMemcpy(table, h.data(), 4096 * 8, cudaMemcpyHostToDevice);
    struct Pat { const char *name; int (*f)(int k, int lane); };
    const Pat pats[] = {
        {"1 slot (broadcast)", [](int, int) { return 0; }},
        {"2 slots, same 32 B sector (0, 1)", [](int k, int l) { return ((l * 7 + k) % 5 == 0) ? 1 : 0; }},
        {"2 slots, same line, other sector (0, 8)", [](int k, int l) { return ((l * 7 + k) % 5 == 0) ? 8 : 0; }},
        {"2 slots, two lines (0, 31)", [](int k, int l) { return ((l * 7 + k) % 5 == 0) ? 31 : 0; }},
        {"2 slots, two lines, half / half (0, 31)", [](int, int l) { return (l < 16) ? 0 : 31; }},
        {"4 slots, one line (0, 1, 2, 4)", [](int k, int l) { const int t[4] = {0, 1, 2, 4}; return t[(l * 5 + k) & 3]; }},
        {"4 slots, two lines (0, 1, 31, 30)", [](int k, int l) { const int t[4] = {0, 1, 31, 30}; return t[(l * 5 + k) & 3]; }},
        {"8 slots, one line (0..15 step 2)", [](int k, int l) { return ((l * 5 + k) & 7) * 2; }},
        {"8 slots, two lines", [](int k, int l) { return ((l * 5 + k) & 7) * 4; }},
        {"16 slots, one line", [](int k, int l) { return (l * 5 + k) & 15; }},
        {"16 slots, two lines", [](int k, int l) { return ((l * 5 + k) & 15) * 2; }},
        {"32 slots, two lines (coalesced: slot = lane)", [](int, int l) { return l; }},
        {"32 slots, two lines (permuted)", [](int k, int l) { return (l * 5 + k) & 31; }},
        {"32 lanes over 243 slots (16 lines)", [](int k, int l) { return (l * 37 + k * 11) % 243; }},
        {"mostly slot 0, three lanes elsewhere in the line", [](int k, int l) { return (l == 3) ? 2 : (l == 17) ? 4 : (l == 29) ? 8 + (k & 3) : 0; }},
        {"mostly slot 0, three lanes in three other lines", [](int k, int l) { return (l == 3) ? 40 : (l == 17) ? 100 : (l == 29) ? 200 + (k & 3) : 0; }},
    };
    for (const Pat &pt : pats) {
        int hs[512];
        for (int k = 0; k < 16; k++) for (int l = 0; l < 32; l++) hs[k * 32 + l] = pt.f(k, l);
        cudaMemcpy(slots, hs, sizeof(hs), cudaMemcpyHostToDevice);
        gather<<<nsm, 256>>>(table, slots, cyc, out, 10);
        gather<<<nsm, 256>>>(table, slots, cyc, out, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", pt.name, cudaGetErrorString(e)); return 1; }
        long long hc[256];
        cudaMemcpy(hc, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < nsm; i++) avg += (double)hc[i];
        avg /= nsm;
        printf("%-52s %6.2f cycles per warp load\n", pt.name, avg / iters / (64.0 * 8));
    }
    return 0;
}
