// cluster_sync_probe.cu -- cost of a cluster-wide barrier per iteration on B200, with and without work
// between arrive and wait, and with DSMEM stores to the neighbouring CTA each iteration.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cluster_sync_probe cluster_sync_probe.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int MODE>   // 0: arrive+wait back to back, 1: arrive, work, wait, 2: work only (no barrier), 3: mode 1 + DSMEM stores
__global__ void __launch_bounds__(256, 1) k_probe(int iters, int work, double *out, long long *cycles)
{
    __shared__ double halo[2][512];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), n = cluster.num_blocks();
    double *peer = cluster.map_shared_rank(&halo[0][0], (rank + 1) % n);
    double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
    cluster.sync();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 3) peer[(it & 1) * 512 + threadIdx.x] = a;          // fire-and-forget remote store
        if (MODE != 2) cl_arrive();
        if (MODE >= 1) {
#pragma unroll 1
            for (int k = 0; k < work; k++) a = fma(a, b, c);
        }
        if (MODE != 2) cl_wait();
        if (MODE == 3) a += halo[it & 1][threadIdx.x] * 1e-30;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

template <int MODE>
static void run(int csize, int iters, int work)
{
    double *out;
    long long *cyc, h = 0;
    cudaMalloc(&out, 148 * 256 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
    if (csize > 8) cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int nclusters = 0;
    cfg.gridDim = dim3(csize);
    cudaOccupancyMaxActiveClusters(&nclusters, k_probe<MODE>, &cfg);
    if (nclusters > 148 / csize) nclusters = 148 / csize;
    cfg.gridDim = dim3(csize * (nclusters > 0 ? nclusters : 1));
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_probe<MODE>, iters, work, out, cyc);
    cudaError_t e2 = cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("cluster %2d mode %d work %4d: max active clusters %3d  %8.1f cycles/iter  (%s, %s)\n", csize, MODE, work, nclusters,
           (double)h / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int iters = 20000;
    for (int cs : {1, 2, 4, 8, 16}) {
        run<0>(cs, iters, 0);
        for (int w : {40, 80, 160}) { run<2>(cs, iters, w); run<1>(cs, iters, w); run<3>(cs, iters, w); }
    }
    return 0;
}
