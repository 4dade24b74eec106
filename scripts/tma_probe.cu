// Standalone probe for the TMA building blocks used by sweep_tma.cu (debug aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/tma_probe scripts/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Maps { CUtensorMap ld[2]; CUtensorMap st[2]; CUtensorMap code; };

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ Maps maps, int src, double *out, int cox, int H, int ox, int oy)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = (uint8_t *)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    double *IN = (double *)sm;
    uint8_t *CD = sm + 128 * 32 * 8;
    uint64_t *bar = (uint64_t *)(sm + 128 * 32 * 8 + 4096);
    const CUtensorMap *mi = &maps.ld[src];
    const CUtensorMap *mo = &maps.st[src ^ 1];
    if (threadIdx.x == 0) {
        if (MODE & 16) { asm volatile("prefetch.tensormap [%0];" ::"l"(mi) : "memory"); }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t bytes = 128 * 32 * 8 + ((MODE & 2) ? 4096 : 0);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(IN)), "l"(mi), "r"(ox), "r"(oy), "r"(smem_u32(bar)) : "memory");
        if (MODE & 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(CD)), "l"(&maps.code), "r"(cox), "r"(oy), "r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) out[i] = IN[i] + ((MODE & 2) ? CD[i] * 1000.0 : 0.0);
    __syncthreads();
    if (MODE & 4) {
        for (int i = threadIdx.x; i < 126 * 30; i += blockDim.x) IN[i] = 7.0 + i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(mo), "r"(0), "r"(0), "r"(smem_u32(IN)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv)
{
    int mode = argc > 1 ? atoi(argv[1]) : 1;
    const int Nx = 300, Ny = 40, pitch = 336, rows = Ny + 2;
    double *x0, *x1, *out; uint8_t *code;
    CK(cudaMalloc(&x0, pitch * rows * 8)); CK(cudaMalloc(&x1, pitch * rows * 8)); CK(cudaMalloc(&code, pitch * rows));
    CK(cudaMalloc(&out, 128 * 32 * 8));
    std::vector<double> h(pitch * rows); for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    std::vector<uint8_t> hc(pitch * rows); for (size_t i = 0; i < hc.size(); i++) hc[i] = i % 5;
    CK(cudaMemcpy(x0, h.data(), h.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemset(x1, 0, h.size() * 8));
    CK(cudaMemcpy(code, hc.data(), hc.size(), cudaMemcpyHostToDevice));
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    Maps m;
    auto encode = [&](CUtensorMap *tm, CUtensorMapDataType dt, void *base, uint64_t d0, uint64_t d1, uint64_t pb, uint32_t b0, uint32_t b1) {
        cuuint64_t dims[2] = {d0, d1}; cuuint64_t str[1] = {pb}; cuuint32_t box[2] = {b0, b1}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(tm, dt, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode -> %d\n", (int)r);
    };
    double *xs[2] = {x0, x1};
    for (int b = 0; b < 2; b++) {
        encode(&m.ld[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, xs[b], pitch, rows, pitch * 8, 128, 32);
        encode(&m.st[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, xs[b] + pitch + 16, Nx, Ny, pitch * 8, 126, 30);
    }
    encode(&m.code, CU_TENSOR_MAP_DATA_TYPE_UINT8, code, pitch, rows, pitch, 128, 32);
    size_t smem = 128 * 32 * 8 + 4096 + 64 + 128;
    int ox = argc > 2 ? atoi(argv[2]) : 16, oy = argc > 3 ? atoi(argv[3]) : 1, cox = argc > 4 ? atoi(argv[4]) : ox;
#define RUN(M) case M: CK(cudaFuncSetAttribute(probe<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<M><<<1, 256, smem>>>(m, 0, out, cox, 32, ox, oy); break;
    switch (mode) { RUN(1) RUN(3) RUN(5) RUN(7) RUN(9) RUN(17) RUN(23) RUN(31) default: printf("bad mode\n"); return 2; }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<double> ho(128 * 32);
    CK(cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost));
    printf("mode %d ok: out[0]=%g (expect %g) out[129]=%g (expect %g)\n", mode, ho[0], (double)(oy * pitch + ox), ho[129],
           (double)((oy + 1) * pitch + ox + 1));
    if (mode & 4) {
        std::vector<double> hx(pitch * rows);
        CK(cudaMemcpy(hx.data(), x1, hx.size() * 8, cudaMemcpyDeviceToHost));
        printf("store: x1[interior 0,0]=%g (expect 7) [0,125]=%g (expect 132) [0,126]=%g (expect 0)\n", hx[pitch + 16], hx[pitch + 16 + 125], hx[pitch + 16 + 126]);
    }
    return 0;
}
