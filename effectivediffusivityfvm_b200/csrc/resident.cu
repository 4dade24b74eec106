// resident.cu -- K5: cluster-resident damped-Jacobi sweeps for small domains and packed batches (sm_100a).
//
// The reference solves its ML-dataset images one at a time, one launch + one sync + one device-to-device
// copy per sweep (BatchSim / BatchSim3Phase, Deff2D.cuh:1867-2049, 2056-2419; updateX_SOR cuh:69-92).
// An image of up to 256 x 256 cells fits on chip: a thread-block cluster of up to 4 x 4 CTAs owns one
// image for a whole check interval (10 000 sweeps), one 64 x 64 tile per CTA (SM), every thread a
// 2 x 8 patch of cells and its 64 sweep weights in registers -- the layout of the tiled sweep
// (sweep_tma.cu, family 4), but with nothing recomputed and nothing re-read:
//
//   * load once: patch values and per-cell weight-table indices straight from the packed stack in HBM;
//   * per sweep: W / E halo by warp shuffles, N / S halo between warps through a planar shared-memory
//     buffer (as in K2); on the CTA edges the halo comes from the neighbouring CTA of the cluster:
//     the owner pushes its edge row / column into the neighbour's shared memory with st.async
//     (distributed shared memory) and the data's arrival is counted by the neighbour's mbarrier
//     (complete_tx) -- no cluster-wide barrier in the loop, a CTA only ever waits for its four neighbours;
//   * store once.  HBM traffic: 16 B per cell per launch instead of per sweep, no halo recomputation
//     (overlapped tiling computes 64 x 64 cells to deliver 52 x 52 at depth 6), no per-tile prologue.
//
// The arithmetic per cell is the tiled kernel's (same weights, same FMA order), so results are
// bit-identical to K2 / K3; flux, stop rule and stage logic stay in k_batch_check / k_check.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "context.h"

namespace cg = cooperative_groups;

namespace deff2d {

namespace {

constexpr int RT = 64;               // tile edge (cells) per CTA
constexpr int RPX = 2, RPY = 8;      // cells per thread: 2 wide, 8 tall; a warp spans the tile width
constexpr int RPW = 32;              // columns per plane of the exchange buffer (= lanes)
constexpr int RROWS = RT + 2;        // exchange rows: -1 (N neighbour CTA), 0..63, 64 (S neighbour CTA)
constexpr int P_DOUBLES = 2 * RPX * RROWS * RPW;      // two parities
constexpr int H_DOUBLES = 2 * RT;                     // W (or E) halo-in column, two parities
constexpr size_t RES_SMEM = (size_t)(P_DOUBLES + 2 * H_DOUBLES) * 8 + 64;

struct ResGeom {
    int Nx, Ny;          // image size in cells
    int GX;              // slots per stack row (1: a single domain)
    long long pitch;     // padded row pitch of the stack
};

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_rank(uint32_t local, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f64(uint32_t raddr, double v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(raddr), "l"(__double_as_longlong(v)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// grid (ncx, ncy, images), cluster (ncx, ncy, 1): blockIdx.z selects the image, the cluster is the image
__global__ void __launch_bounds__(256, 1)
k_resident(double *__restrict__ x, const uint16_t *__restrict__ idx16, const double *__restrict__ clut, double om,
           ResGeom g, const int *__restrict__ active, long long nsweeps)
{
    extern __shared__ __align__(16) uint8_t smem[];
    double *const P = reinterpret_cast<double *>(smem);                     // [parity][px][row + 1][lane]
    double *const HW = P + P_DOUBLES;                                       // [parity][row]: column -1 of this tile
    double *const HE = HW + H_DOUBLES;                                      // [parity][row]: column 64 of this tile
    uint64_t *const bar = reinterpret_cast<uint64_t *>(HE + H_DOUBLES);     // [parity]

    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cx = blockIdx.x, cy = blockIdx.y, ncx = gridDim.x, ncy = gridDim.y;
    const int slot = active ? active[blockIdx.z] : 0;
    const int gx = slot % g.GX, gy = slot / g.GX;
    const int c0 = RPX * lane, r0 = RPY * warp;                             // patch origin inside the tile
    const int j0 = RT * cx + c0, i0 = RT * cy + r0;                         // ... inside the image
    // padded stack index of image cell (0, 0)
    const long long base = ((long long)gy * (g.Ny + 1) + 1) * g.pitch + (long long)gx * (g.Nx + 1) + DEFF2D_XOFF;

    // ---- neighbours inside the cluster (x fastest in the cluster rank) --------------------------------------
    const bool hasW = cx > 0, hasE = cx < ncx - 1, hasN = cy > 0, hasS = cy < ncy - 1;
    const uint32_t rankW = (uint32_t)(cy * ncx + cx - 1), rankE = (uint32_t)(cy * ncx + cx + 1);
    const uint32_t rankN = (uint32_t)((cy - 1) * ncx + cx), rankS = (uint32_t)((cy + 1) * ncx + cx);
    const uint32_t expect = 8u * RT * ((hasW ? 1u : 0u) + (hasE ? 1u : 0u) + (hasN ? 1u : 0u) + (hasS ? 1u : 0u));

    // ---- constants in the halo-in buffers where there is no neighbour: Dirichlet ghost columns hold 1.0 (the
    //      face weight carries CL / CR, tables.cpp), no-flux rows carry weight 0 and only need a finite value
    for (int k = tid; k < P_DOUBLES; k += 256) P[k] = 0.0;
    for (int k = tid; k < 2 * H_DOUBLES; k += 256) HW[k] = 1.0;             // HW and HE are contiguous
    if (tid == 0) {
        bar_init(&bar[0], 8);
        bar_init(&bar[1], 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();                                 // every CTA's barriers and constants exist before the first remote store

    // ---- patch values and weights into registers ---------------------------------------------------------------
    double xv[RPY][RPX], w[RPY][RPX][4], omc[RPX];
#pragma unroll
    for (int px = 0; px < RPX; px++) omc[px] = (j0 + px < g.Nx) ? om : 1.0;   // columns right of the image keep 1.0
#pragma unroll
    for (int py = 0; py < RPY; py++)
#pragma unroll
        for (int px = 0; px < RPX; px++) {
            const int i = i0 + py, j = j0 + px;
            const bool in = (i < g.Ny) && (j < g.Nx);
            double v = (j >= g.Nx) ? 1.0 : 0.0;
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            if (in) {
                const long long q = base + (long long)i * g.pitch + j;
                v = x[q];
                const unsigned id = idx16[q];
                const unsigned e = ((id & 0x3c00u) << 2) | (id & 0x3ffu);     // stage * 4096 + slot (planar table)
                a0 = __ldg(clut + e); a1 = __ldg(clut + e + DEFF2D_CLUT_ENTRIES);
                a2 = __ldg(clut + e + 2 * DEFF2D_CLUT_ENTRIES); a3 = __ldg(clut + e + 3 * DEFF2D_CLUT_ENTRIES);
            }
            xv[py][px] = v;
            w[py][px][0] = a0; w[py][px][1] = a1; w[py][px][2] = a2; w[py][px][3] = a3;
        }

    // remote addresses of the buffers this thread feeds (parity 0; parity 1 is a fixed offset further)
    const uint32_t lbar = s32(bar);
    uint32_t rN_addr = 0, rN_bar = 0, rS_addr = 0, rS_bar = 0, rWE_addr = 0, rWE_bar = 0;
    if (warp == 0 && hasN) {          // my top row -> row 64 (index 65) of the CTA above
        rN_addr = map_rank(s32(P + (0 * RROWS + RT + 1) * RPW + lane), rankN);
        rN_bar = map_rank(lbar, rankN);
    }
    if (warp == 7 && hasS) {          // my bottom row -> row -1 (index 0) of the CTA below
        rS_addr = map_rank(s32(P + (0 * RROWS + 0) * RPW + lane), rankS);
        rS_bar = map_rank(lbar, rankS);
    }
    const bool sendW = (lane == 0) && hasW, sendE = (lane == 31) && hasE;
    if (sendW) { rWE_addr = map_rank(s32(HE + r0), rankW); rWE_bar = map_rank(lbar, rankW); }   // my column 0 -> its column 64
    if (sendE) { rWE_addr = map_rank(s32(HW + r0), rankE); rWE_bar = map_rank(lbar, rankE); }   // my column 63 -> its column -1

    // level-s rows and columns that other threads need: the top row is published as soon as it exists, the rest at
    // the end of the sweep.  Local neighbours read the planar buffer P, the neighbouring CTAs get st.async pushes.
    auto publish_top = [&](auto par_c) {
        constexpr int par = decltype(par_c)::value;
        double *pb = P + par * (RPX * RROWS * RPW);
        constexpr uint32_t poff = (uint32_t)par * (RPX * RROWS * RPW * 8), boff = (uint32_t)par * 8;
#pragma unroll
        for (int px = 0; px < RPX; px++) pb[(px * RROWS + r0 + 1) * RPW + lane] = xv[0][px];
        if (rN_bar) {
#pragma unroll
            for (int px = 0; px < RPX; px++) st_async_f64(rN_addr + poff + px * (RROWS * RPW * 8), xv[0][px], rN_bar + boff);
        }
    };
    auto publish_rest = [&](auto par_c) {
        constexpr int par = decltype(par_c)::value;
        double *pb = P + par * (RPX * RROWS * RPW);
        constexpr uint32_t poff = (uint32_t)par * (RPX * RROWS * RPW * 8), boff = (uint32_t)par * 8;
#pragma unroll
        for (int px = 0; px < RPX; px++) pb[(px * RROWS + r0 + RPY) * RPW + lane] = xv[RPY - 1][px];
        if (rS_bar) {
#pragma unroll
            for (int px = 0; px < RPX; px++) st_async_f64(rS_addr + poff + px * (RROWS * RPW * 8), xv[RPY - 1][px], rS_bar + boff);
        }
        // lanes 0 / 31: the tile's edge column to the neighbouring CTA, straight from the patch registers (8-byte
        // pushes need no staging moves; two-row 16-byte pushes with per-lane selects measured 1310 cycles per sweep)
        constexpr uint32_t hoff = (uint32_t)par * (RT * 8);
        if (sendW) {
#pragma unroll
            for (int py = 0; py < RPY; py++) st_async_f64(rWE_addr + hoff + py * 8, xv[py][0], rWE_bar + boff);
        }
        if (sendE) {
#pragma unroll
            for (int py = 0; py < RPY; py++) st_async_f64(rWE_addr + hoff + py * 8, xv[py][RPX - 1], rWE_bar + boff);
        }
        // one arrival per warp once its rows are in P (thread 0's also announces the bytes the neighbouring CTAs push):
        // the wait at the top of the next sweep then covers the local exchange and the remote one -- no CTA barrier
        __syncwarp();
        if (lane == 0) {
            if (warp == 0) bar_expect(&bar[par], expect);
            else bar_arrive(&bar[par]);
        }
    };
    using Par0 = std::integral_constant<int, 0>;
    using Par1 = std::integral_constant<int, 1>;

    // W / E halo of the next sweep: exchanged between neighbouring lanes by shuffles right after a row is final
    // (no barrier involved), so that this LSU work overlaps the FP64 work of the rows below
    double hW[RPY], hE[RPY];
#pragma unroll
    for (int py = 0; py < RPY; py++) {
        hW[py] = __shfl_up_sync(0xffffffffu, xv[py][RPX - 1], 1);
        hE[py] = __shfl_down_sync(0xffffffffu, xv[py][0], 1);
    }
    if (nsweeps > 0) {
        publish_top(Par0{});
        publish_rest(Par0{});
    }
    const int rowN = r0, rowS = r0 + RPY + 1;          // exchange-row indices of the rows above / below the patch
    // one sweep: level s-1 (exchange buffers of parity `par`) -> level s; publishes level s unless it is the last
    auto sweep = [&](auto par_c, long long s) {
        constexpr int par = decltype(par_c)::value;
        using Next = std::integral_constant<int, 1 - par>;
        const bool pub = s < nsweeps;
        double hN[RPX], hS[RPX];
        bar_wait(&bar[par], (uint32_t)(((s - 1) >> 1) & 1));              // level s-1: every warp's rows are in P, the neighbours' edges have landed
        const double *pr = P + par * (RPX * RROWS * RPW);
#pragma unroll
        for (int px = 0; px < RPX; px++) {
            hN[px] = pr[(px * RROWS + rowN) * RPW + lane];
            hS[px] = pr[(px * RROWS + rowS) * RPW + lane];
        }
        // lanes 0 / 31: the halo column comes from the neighbouring CTA (or holds the Dirichlet ghost value)
        if (lane == 0) {
            const double2 *hp = reinterpret_cast<const double2 *>(HW + par * RT + r0);
#pragma unroll
            for (int py = 0; py < RPY; py += 2) { const double2 v = hp[py >> 1]; hW[py] = v.x; hW[py + 1] = v.y; }
        }
        if (lane == 31) {
            const double2 *hp = reinterpret_cast<const double2 *>(HE + par * RT + r0);
#pragma unroll
            for (int py = 0; py < RPY; py += 2) { const double2 v = hp[py >> 1]; hE[py] = v.x; hE[py + 1] = v.y; }
        }
        // in-place update; `up[px]` carries the old value of the row above (same FMA order as K2 / K3)
        double up[RPX];
#pragma unroll
        for (int px = 0; px < RPX; px++) up[px] = hN[px];
#pragma unroll
        for (int py = 0; py < RPY; py++) {
            double left = hW[py];
#pragma unroll
            for (int px = 0; px < RPX; px++) {
                const double c = xv[py][px];
                const double right = (px == RPX - 1) ? hE[py] : xv[py][px + 1];
                const double down = (py == RPY - 1) ? hS[px] : xv[py + 1][px];
                double r = omc[px] * c;
                r = fma(w[py][px][0], left, r);
                r = fma(w[py][px][1], right, r);
                r = fma(w[py][px][2], down, r);
                r = fma(w[py][px][3], up[px], r);
                xv[py][px] = r;
                left = c;
                up[px] = c;
            }
            hW[py] = __shfl_up_sync(0xffffffffu, xv[py][RPX - 1], 1);
            hE[py] = __shfl_down_sync(0xffffffffu, xv[py][0], 1);
            if (py == 0 && pub) publish_top(Next{});
        }
        if (pub) publish_rest(Next{});
    };
    {
        long long s = 1;
        for (; s + 1 <= nsweeps; s += 2) { sweep(Par0{}, s); sweep(Par1{}, s + 1); }
        if (s <= nsweeps) sweep(Par0{}, s);
    }

    // ---- store once ------------------------------------------------------------------------------------------
#pragma unroll
    for (int py = 0; py < RPY; py++)
#pragma unroll
        for (int px = 0; px < RPX; px++) {
            const int i = i0 + py, j = j0 + px;
            if (i < g.Ny && j < g.Nx) x[base + (long long)i * g.pitch + j] = xv[py][px];
        }
    cluster.sync();                                 // no CTA leaves while a neighbour may still push into its shared memory
}

struct ResState {
    int ok_x = 0, ok_y = 0;      // cluster shape whose schedulability has been confirmed (0: none yet)
    bool attr = false;
};

}  // namespace

// Can a domain / image of Nx x Ny cells run cluster-resident on this device?
bool resident_eligible(deff2d_ctx *c, int64_t Nx, int64_t Ny)
{
    if (Nx < 1 || Ny < 1 || Nx > 4 * RT || Ny > 4 * RT) return false;
    if (c->prop.major < 9) return false;
    const int ncx = (int)((Nx + RT - 1) / RT), ncy = (int)((Ny + RT - 1) / RT);
    ResState *rs = static_cast<ResState *>(c->resident);
    if (!rs) { rs = new ResState(); c->resident = rs; }
    if (!rs->attr) {
        if (cudaFuncSetAttribute(k_resident, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_resident, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            (void)cudaGetLastError();
            return false;
        }
        rs->attr = true;
    }
    if (rs->ok_x == ncx && rs->ok_y == ncy) return true;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ncx; attr[0].val.clusterDim.y = (unsigned)ncy; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)ncx, (unsigned)ncy, 1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = RES_SMEM;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, k_resident, &cfg) != cudaSuccess || nclusters < 1) {
        (void)cudaGetLastError();
        return false;
    }
    rs->ok_x = ncx; rs->ok_y = ncy;
    return true;
}

// n sweeps of every image in `active` (device list of slot numbers; NULL: the single slot 0), in place on
// x[c->cur] of the resident stack (image Nx x Ny, GX slots per stack row).  Enqueue only.
int resident_sweeps(deff2d_ctx *c, int64_t n, int64_t Nx, int64_t Ny, int GX, const int *active, int nactive)
{
    if (n < 1 || nactive < 1) return DEFF2D_OK;
    if (!resident_eligible(c, Nx, Ny)) { set_error(c, "resident sweeps: domain %lld x %lld is not eligible", (long long)Nx, (long long)Ny); return DEFF2D_ERR_STATE; }
    const int ncx = (int)((Nx + RT - 1) / RT), ncy = (int)((Ny + RT - 1) / RT);
    ResGeom g;
    g.Nx = (int)Nx; g.Ny = (int)Ny; g.GX = GX; g.pitch = c->pitch;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ncx; attr[0].val.clusterDim.y = (unsigned)ncy; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = RES_SMEM; cfg.stream = c->stream;
    for (int z0 = 0; z0 < nactive; z0 += 65535) {
        const int nz = (nactive - z0 < 65535) ? nactive - z0 : 65535;
        cfg.gridDim = dim3((unsigned)ncx, (unsigned)ncy, (unsigned)nz);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_resident, c->x[c->cur].p, (const uint16_t *)c->idx16.p, (const double *)c->clut.p,
                                           1.0 - c->omega, g, active ? active + z0 : (const int *)nullptr, (long long)n);
        if (e != cudaSuccess) { set_error(c, "resident sweep launch failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
        c->launches++;
    }
    return DEFF2D_OK;
}

void resident_destroy(deff2d_ctx *c)
{
    delete static_cast<ResState *>(c->resident);
    c->resident = nullptr;
}

}  // namespace deff2d
