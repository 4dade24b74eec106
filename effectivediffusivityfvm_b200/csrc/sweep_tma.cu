// sweep_tma.cu -- K2: TMA-staged, temporally blocked damped-Jacobi sweep (sm_100a).
//
// Replaces T launches of the reference's updateX_SOR (Deff2D.cuh:69-92) plus its per-sweep
// device-to-device copy (cuh:1281) by ONE pass over HBM:
//
//   * persistent CTAs walk a static list of (TW x TH) tiles; every tile, halo included, is
//     brought into shared memory by two bulk tensor copies (FP64 iterate + the u16 weight-table
//     index of every cell, cp.async.bulk.tensor.2d, mbarrier complete_tx), double buffered so the next tile's
//     copy overlaps the current tile's sweeps; out-of-range boxes are zero filled by TMA;
//   * each thread owns a PX x PY patch of cells for the whole tile visit: the patch values
//     and its 4 x PX x PY sweep weights (gathered once per tile from the stage's compact planar
//     table with the precomputed per-cell slot; one broadcast read when the whole warp's
//     patches are single-phase) live in registers;
//   * T sweeps run on chip; per sweep a thread gets its W / E halo from the neighbouring lanes
//     by warp shuffles (a warp spans the tile width) and publishes only its top and bottom
//     patch rows to a planar exchange buffer (column-phase planes -> conflict-free LDS/STS),
//     reading the rows of the patches above and below back; one __syncthreads per sweep;
//   * after T sweeps the (TW-2*TE) x (TH-2T) interior is staged in shared memory and written
//     back with one bulk tensor store through a tensor map that covers only the interior of the
//     domain (so the outer ghost ring is never overwritten and edge tiles are clipped by the hardware);
//   * long runs of passes are replayed as CUDA graphs (tma_passes).
//
// Default geometry: square 64 x 64 tiles (Family 3: 16 lanes side by side, two patch rows per warp,
// 4 x 4 cells per thread, 256 threads, one CTA per SM), 8 sweeps per pass.
//
// Overlapped tiling is algebraically identical to T plain sweeps: a cell at distance >= T
// from the tile edge only ever sees values that are exact at each intermediate level.
// Ghost columns (Dirichlet value 1.0) keep their value because their weights are zero and
// their (1-omega) factor is replaced by 1 (per thread column, no per-cell branch).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <vector>

#include "context.h"

namespace deff2d {

// ------------------------------------------------------------------------------------------ PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ------------------------------------------------------------------------------------------ kernel

// CHEB variant (chebyshev.cu): per-sweep relaxation factors of one pass; the weight table then holds omega = 1
struct ChebTaus { double t[8]; };

struct TmaMaps {
    CUtensorMap x_load[2];     // padded iterate buffers, box TW x TH
    CUtensorMap x_store[2];    // interior of the iterate buffers, box OW x OH
    CUtensorMap idx;           // padded per-cell table indices, box IW x TH (u16)
};

// LX lanes of a warp lie side by side along x (32: the warp is one row of patches; 16: two rows
// of 16 patches, which makes square 64 x 64 tiles possible with 4-column patches)
template <int T_, int PX_, int PY_, int NWX_, int NWY_, int LX_ = 32>
struct Cfg {
    static constexpr int T = T_, PX = PX_, PY = PY_, NWX = NWX_, NWY = NWY_, LX = LX_, LY = 32 / LX_;
    static constexpr int TW = LX * PX * NWX, TH = PY * LY * NWY;
    static constexpr int NT = 32 * NWX * NWY;
    // TMA needs the innermost box coordinate 16-byte aligned (measured on B200: an odd FP64
    // column or a u8 column that is not a multiple of 16 raises "illegal instruction").  The x
    // halo is therefore rounded up to an even number of columns, and the index box (u16) is 8
    // elements wider than the tile and starts at the previous multiple of 8.
    static constexpr int TE = (T + 1) & ~1;
    static constexpr int OW = TW - 2 * TE, OH = TH - 2 * T;
    static constexpr int CELLS = TW * TH;
    static constexpr int IW = TW + 8;                            // index box width (u16 elements)
    static constexpr int PLANE_W = TW / PX;                      // columns per phase plane
    // shared memory map (bytes)
    static constexpr size_t IN_BYTES = (size_t)CELLS * 8;
    static constexpr size_t CODE_BYTES = ((size_t)IW * TH * 2 + 127) / 128 * 128;
    static constexpr size_t OFF_IN = 0;                           // IN[2]
    static constexpr size_t OFF_P = OFF_IN + 2 * IN_BYTES;        // P[2] planar exchange
    static constexpr size_t OFF_OUT = OFF_P + 2 * IN_BYTES;       // OUT: dense OW x OH box for the bulk store
    static constexpr size_t OUT_BYTES = ((size_t)OW * OH * 8 + 127) / 128 * 128;
    static constexpr size_t OFF_CODE = OFF_OUT + OUT_BYTES;       // CODE[2]
    static constexpr size_t OFF_BAR = OFF_CODE + 2 * CODE_BYTES;  // 2 mbarriers + tile origins
    static constexpr size_t SMEM = OFF_BAR + 64 + 128;            // + alignment slack
    static_assert(OW > 0 && OH > 0, "tile too small for this temporal depth");
    static_assert(TW <= 256 && TH <= 256, "TMA box dimension limit");
};

// LIST: walk an explicit tile list.
// Measured and removed alternatives: staging the weight table in shared memory (695 vs 704 GLUP/s
// on config 2, 513 vs 543 in packed batches once the table became dense and planar -- gathers
// through L1 win); writing the patches straight to global memory with 16-byte stores instead of
// staging the output box for one bulk tensor store (647 vs 702, 532 vs 547); with two patch rows per
// warp (square tiles), exchanging the row between them by shuffles instead of shared memory
// (763 vs 817, 610 vs 650: a 64-bit shuffle costs more LSU time than an 8-byte shared-memory access);
// two 64-thread named barriers per sweep between neighbouring warps instead of one CTA-wide barrier
// (685 vs 828); two 128-thread CTAs per SM on 64 x 32 tiles (744 at T = 4 vs 827).
template <class C, bool LIST, int VAR>
__global__ void __launch_bounds__(C::NT, 1)
k_sweep_tma(const __grid_constant__ TmaMaps maps, int src, const double *__restrict__ lut, const uint32_t *__restrict__ lut32,
            double om, int tiles_x, int ntiles, const uint32_t *__restrict__ tile_list, const ChebTaus taus,
            const PeerArgs peer)
{
    constexpr int T = C::T, TE = C::TE, PX = C::PX, PY = C::PY, TW = C::TW, TH = C::TH, OW = C::OW, OH = C::OH;
    constexpr int PW = C::PLANE_W, IW = C::IW;
    // NON-PARITY Chebyshev mode: sweep s of the pass is the Richardson step x' = x + tau_s (sum_f u_f x_f - x) with the
    // omega = 1 table u; ghost columns keep their value (factor 0)
    constexpr bool CHEB = (VAR & 2) != 0;
    constexpr bool PEER = (VAR & 1) != 0;     // peer-memory halo exchange fused into the pass (needs LIST)
    constexpr bool G32 = (VAR & 4) != 0;      // weights gathered as 32-bit halves (interface-rich media, see fetch32)
    static_assert(!PEER || LIST, "the peer variant walks a tile list (boundary tiles early)");
    static_assert(C::NWX == 1, "a warp spans the tile width (W / E halo by shuffles)");

    // No integer round trip on the base pointer: the compiler must keep seeing the shared
    // address space (a generic pointer turns every LDS/STS below into a slow generic LD/ST).
    extern __shared__ __align__(1024) uint8_t smem[];
    double *const IN0 = reinterpret_cast<double *>(smem + C::OFF_IN);
    double *const P0 = reinterpret_cast<double *>(smem + C::OFF_P);
    double *const OUT = reinterpret_cast<double *>(smem + C::OFF_OUT);
    uint8_t *const CODE0 = smem + C::OFF_CODE;
    uint64_t *const bar = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
    int *const org = reinterpret_cast<int *>(smem + C::OFF_BAR + 16);   // output-box origin of the tile in buffer b: org[2b], org[2b+1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wx = warp % C::NWX, wy = warp / C::NWX;
    const int g = wx * C::LX + (lane % C::LX);   // column group of this thread (plane column index)
    const int c0 = g * PX;                   // first tile column of the patch
    const int r0 = (wy * C::LY + lane / C::LX) * PY;   // first tile row of the patch

    const CUtensorMap *map_in = &maps.x_load[src];
    const CUtensorMap *map_out = &maps.x_store[src ^ 1];
    constexpr uint32_t TX_BYTES = (uint32_t)(C::IN_BYTES + (size_t)IW * TH * 2);

    // interior-coordinate origin of the OUTPUT box of tile t: the whole tile grid, or -- packed
    // batches and slab boundary/interior splits -- the entries of an explicit tile list
    // `entry`: the tile's list entry (LIST only; thread 0 fetches it a whole tile visit before it is needed, so the
    // L2 latency of the read is not paid in front of the sweeps -- the other seven warps wait for warp 0 at every barrier)
    auto tile_origin = [&](int t, uint32_t entry, int &ox, int &oy) {
        int tx, ty;
        if constexpr (LIST) { tx = (int)(entry & 0xffffu); ty = (int)(entry >> 16); }
        else { ty = t / tiles_x; tx = t - ty * tiles_x; }
        ox = tx * OW; oy = ty * OH;
        if constexpr (PEER) oy += peer.above;        // peer mode: the tile grid starts at the first own row
    };
    auto issue_load = [&](int t, uint32_t entry, int b) {
        int ox, oy;
        tile_origin(t, entry, ox, oy);
        if constexpr (LIST) { org[2 * b] = ox; org[2 * b + 1] = oy; }   // released to the consumers by the arrive below
        mbar_expect_tx(&bar[b], TX_BYTES);
        // padded coordinates of the input box: interior (ox-TE, oy-T) -> (+XOFF, +1); even
        const int xs = ox - TE + DEFF2D_XOFF;
        tma_load_2d(IN0 + b * C::CELLS, map_in, xs, oy - T + 1, &bar[b]);
        tma_load_2d(CODE0 + b * C::CODE_BYTES, &maps.idx, xs & ~7, oy - T + 1, &bar[b]);
    };

    if (tid == 0) {
        prefetch_tmap(map_in); prefetch_tmap(map_out); prefetch_tmap(&maps.idx);
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    int tile = blockIdx.x;
    const double *wtab = lut;
    long long pass = 0;
    int pushed = 0;                             // PEER, thread 0: boundary tiles of this CTA whose pushes are not counted yet
    // One device-scope fence per CTA and pass (count_issue); the thread that counts the last boundary tile of the pass has then
    // observed every other CTA's pushes (its atomic reads their counts), so its system-scope fence + release store
    // publishes all of them: the neighbours' flags show this pass.  It also closes the pass on this rank (counter back
    // to 0, pass number + 1) -- nobody else touches either before the next launch.  A CTA that starts so late that it
    // already reads the new pass number has no boundary tile (they are all counted) and merely waits for flags that
    // its neighbours raise during their current pass.
    // The count is split: count_issue sends the atomic, count_finish -- a tile visit later -- looks at what it returned,
    // so that its round trip to L2 is not waited for in front of a CTA-wide barrier.
    unsigned long long counted = 0;             // what the atomic returned + this CTA's share; 0: nothing pending
    auto count_issue = [&]() {
        __threadfence();
        counted = atomicAdd(peer.counters, (unsigned long long)pushed) + (unsigned long long)pushed;
        pushed = 0;
    };
    auto count_finish = [&]() {
        if (counted == (unsigned long long)peer.nboundary) {
            *reinterpret_cast<volatile unsigned long long *>(peer.counters) = 0ull;
            *reinterpret_cast<volatile long long *>(peer.pass_no) = pass + 1;
            __threadfence_system();             // one system fence, then relaxed flag stores: a release pattern for both flags
            if (peer.flag_up) asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(peer.flag_up), "l"(pass) : "memory");
            if (peer.flag_down) asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(peer.flag_down), "l"(pass) : "memory");
        }
        counted = 0;
    };
    if (tid == 0) {
        uint32_t e0 = 0, e1 = 0;
        if constexpr (LIST) {
            if (tile < ntiles) e0 = __ldg(tile_list + tile);
            if (tile + (int)gridDim.x < ntiles) e1 = __ldg(tile_list + tile + gridDim.x);
        }
        if constexpr (PEER) {
            // The halo rows this pass reads were written by the neighbours' previous pass; the halo rows this pass
            // writes were read by it: both are over once their flags show that pass.  The first `lead` list entries
            // (one per CTA) are tiles away from the neighbours -- they read and write own rows only -- so their loads
            // go out before the flags are looked at.  (Looking at the flags only behind the first tile's prologue and
            // issuing the second tile's load there was slower: 231.6 against 227.3 us per pass.)
            if (peer.lead > 0 && tile < ntiles) issue_load(tile, e0, 0);
            pass = *reinterpret_cast<volatile long long *>(peer.pass_no);
            auto flag_at_least = [&](const long long *f, long long want) {
                long long v;
                do { asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory"); } while (v < want);
            };
            if (peer.up[0]) flag_at_least(peer.flag_local + 0, pass - 1);
            if (peer.down[0]) flag_at_least(peer.flag_local + 1, pass - 1);
            fence_proxy_async();
            if (peer.lead == 0 && tile < ntiles) issue_load(tile, e0, 0);
        } else {
            if (tile < ntiles) issue_load(tile, e0, 0);
        }
        if (tile + (int)gridDim.x < ntiles) issue_load(tile + gridDim.x, e1, 1);
    }

    // clamped neighbour positions (tile-edge patches produce halo garbage that is never stored)
    const int rN = (r0 > 0) ? r0 - 1 : r0;
    const int rS = (r0 + PY < TH) ? r0 + PY : r0 + PY - 1;

    int k = 0;
    for (; tile < ntiles; tile += gridDim.x, k++) {
        const int b = k & 1;
        uint32_t next_entry = 0;               // list entry of the tile after next: in flight during the prologue
        if constexpr (LIST) {
            if (tid == 0 && tile + 2 * (int)gridDim.x < ntiles) next_entry = __ldg(tile_list + tile + 2 * (int)gridDim.x);
        }
        mbar_wait(&bar[b], (uint32_t)((k >> 1) & 1));

        // ---- patch values and weights into registers --------------------------------------
        double x[PY][PX];
        double w[PY][PX][4];
        double omc[PX];
        int ox, oy;
        if constexpr (LIST) { ox = org[2 * b]; oy = org[2 * b + 1]; }      // one global read per tile (thread 0), not 256
        else tile_origin(tile, 0u, ox, oy);
        {
            const double *in = IN0 + b * C::CELLS;
            // table index of every cell of the patch: precomputed per cell (k_build_idx, kernels.cu),
            // bits 0-10 phase neighbourhood, 11-14 continuation stage, 15 "Dirichlet ghost column"
            const uint16_t *id = reinterpret_cast<const uint16_t *>(CODE0 + b * C::CODE_BYTES) + ((ox - TE + DEFF2D_XOFF) & 7);
#pragma unroll
            for (int py = 0; py < PY; py++) {
                if constexpr (PX == 4) {
                    // A lane owns 32 consecutive bytes of the row, so lanes l and l+4 of a quarter warp hit
                    // the same banks when all read their first 16 bytes.  Lanes with bit 2 set read their
                    // second half first: every 16-byte load instruction is then conflict-free, and two
                    // selects per value put the halves back in place.
                    const bool hs = (lane & 4) != 0;
                    const double *rp = in + (r0 + py) * TW + c0;
                    const double2 a = *reinterpret_cast<const double2 *>(rp + (hs ? 2 : 0));
                    const double2 b2 = *reinterpret_cast<const double2 *>(rp + (hs ? 0 : 2));
                    x[py][0] = hs ? b2.x : a.x; x[py][1] = hs ? b2.y : a.y;
                    x[py][2] = hs ? a.x : b2.x; x[py][3] = hs ? a.y : b2.y;
                } else if constexpr (PX % 2 == 0) {       // 16-byte loads: conflict-free for PX == 2
#pragma unroll
                    for (int px = 0; px < PX; px += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(in + (r0 + py) * TW + c0 + px);
                        x[py][px] = v.x; x[py][px + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int px = 0; px < PX; px++) x[py][px] = in[(r0 + py) * TW + c0 + px];
                }
            }
            unsigned idx[PY][PX];
            bool uniform = true;
            unsigned first = 0;
#pragma unroll
            for (int py = 0; py < PY; py++) {
                const uint32_t *p32 = reinterpret_cast<const uint32_t *>(id + (r0 + py) * IW + c0);   // 4-byte aligned: c0, TE, XOFF even
#pragma unroll
                for (int px = 0; px < PX; px += 2) {
                    const uint32_t u = p32[px >> 1];
                    if (py == 0 && px == 0) first = (u & 0xffffu) * 0x10001u;
                    uniform = uniform && (u == first);
                    idx[py][px] = u & 0xffffu;
                    idx[py][px + 1] = u >> 16;
                }
            }
            // a Dirichlet ghost column (column -1 or Nx of the domain; in a packed batch every
            // (Nx+1)-th column separates two images) keeps its value: its weights are 0 (LUT,
            // p == 3) and its (1-omega) factor is 1
#pragma unroll
            for (int px = 0; px < PX; px++) {
                unsigned any = 0;
#pragma unroll
                for (int py = 0; py < PY; py++) any |= idx[py][px];
                omc[px] = CHEB ? ((any & 0x8000u) ? 0.0 : 1.0) : ((any & 0x8000u) ? 1.0 : om);
            }
#pragma unroll
            for (int py = 0; py < PY; py++)
#pragma unroll
                for (int px = 0; px < PX; px++)      // offset into the planar table: stage * 4096 + slot
                    idx[py][px] = ((idx[py][px] & 0x3c00u) << 2) | (idx[py][px] & 0x3ffu);
            // Most patches lie inside one phase (every cell has the same neighbourhood index):
            // one LUT entry then serves all PX*PY cells -- 2 instead of 2*PX*PY 16-byte loads.
            // warp-wide decision: a mixed warp would execute both paths
            auto fetch = [&](unsigned e, double &w0, double &w1, double &w2, double &w3) {
                const double *q = wtab + e;
                w0 = __ldg(q); w1 = __ldg(q + DEFF2D_CLUT_ENTRIES); w2 = __ldg(q + 2 * DEFF2D_CLUT_ENTRIES);
                w3 = __ldg(q + 3 * DEFF2D_CLUT_ENTRIES);
            };
            // Interface-rich media (G32 variant, chosen per domain load: context.cu): two 4-byte gathers per weight from the table of
            // 32-bit halves (eight planes per stage: offset + stage * 4096).  A 4-byte gather of a warp is conflict-free
            // when the slots differ mod 32, an 8-byte gather only when they differ mod 16 (deff2d_internal.h).
            auto fetch32 = [&](unsigned e, double &w0, double &w1, double &w2, double &w3) {
                const uint32_t *q = lut32 + e + (e & ~0xfffu);
                constexpr int N = DEFF2D_CLUT_ENTRIES;
                w0 = __hiloint2double((int)__ldg(q + 4 * N), (int)__ldg(q));
                w1 = __hiloint2double((int)__ldg(q + 5 * N), (int)__ldg(q + N));
                w2 = __hiloint2double((int)__ldg(q + 6 * N), (int)__ldg(q + 2 * N));
                w3 = __hiloint2double((int)__ldg(q + 7 * N), (int)__ldg(q + 3 * N));
            };
            if (__all_sync(0xffffffffu, uniform)) {
                double a0, a1, a2, a3;
                fetch(idx[0][0], a0, a1, a2, a3);
#pragma unroll
                for (int py = 0; py < PY; py++)
#pragma unroll
                    for (int px = 0; px < PX; px++) {
                        w[py][px][0] = a0; w[py][px][1] = a1; w[py][px][2] = a2; w[py][px][3] = a3;
                    }
            } else {
#pragma unroll
                for (int py = 0; py < PY; py++)
#pragma unroll
                    for (int px = 0; px < PX; px++) {
                        if constexpr (G32) fetch32(idx[py][px], w[py][px][0], w[py][px][1], w[py][px][2], w[py][px][3]);
                        else fetch(idx[py][px], w[py][px][0], w[py][px][1], w[py][px][2], w[py][px][3]);
                    }
            }
        }

        // The W / E halo comes from the neighbouring lanes by warp shuffles (a warp spans the tile
        // width), so only the top and bottom patch rows go through shared memory.
        auto publish_rows = [&](double *pb, const double (&top)[PX], const double (&bot)[PX]) {
#pragma unroll
            for (int px = 0; px < PX; px++) {
                pb[(px * TH + r0) * PW + g] = top[px];
                pb[(px * TH + r0 + PY - 1) * PW + g] = bot[px];
            }
        };
        // x' = (1-w) x + wW xW + wE xE + wS xS + wN xN   (cuh:76-89, A and b folded into w): one patch row
        auto row_update = [&](int py, double hw, double he, const double (&upv)[PX], const double (&dnv)[PX],
                              const double (&cur)[PX], double (&out)[PX], const double (&fac)[PX]) {
            double left = hw;
#pragma unroll
            for (int px = 0; px < PX; px++) {
                const double c = cur[px];
                const double right = (px == PX - 1) ? he : cur[px + 1];
                double r;
                if constexpr (CHEB) {
                    double t = w[py][px][0] * left;
                    t = fma(w[py][px][1], right, t);
                    t = fma(w[py][px][2], dnv[px], t);
                    t = fma(w[py][px][3], upv[px], t);
                    r = fma(fac[px], t - c, c);
                } else {
                    r = fac[px] * c;
                    r = fma(w[py][px][0], left, r);
                    r = fma(w[py][px][1], right, r);
                    r = fma(w[py][px][2], dnv[px], r);
                    r = fma(w[py][px][3], upv[px], r);
                }
                out[px] = r;
                left = c;
            }
        };

        publish_rows(P0, x[0], x[PY - 1]);
        // the OUT box of the previous tile must have been read by its bulk store before this
        // tile's last sweep overwrites it (the wait is ordered before the writes by S1)
        if (tid == 0) tma_wait_read0();
        __syncthreads();                       // S1: IN[b], CODE[b] fully consumed; P[0] visible

        // prefetch the tile after next into the buffer just consumed: loads run two tiles ahead
        if (tid == 0) {
            const int nt = tile + 2 * (int)gridDim.x;
            if (nt < ntiles) issue_load(nt, next_entry, b);
        }

        // ---- T sweeps on chip --------------------------------------------------------------
#pragma unroll
        for (int s = 1; s <= T; s++) {
            const double *pr = P0 + ((s - 1) & 1) * C::CELLS;
            double hW[PY], hE[PY], hN[PX], hS[PX];
#pragma unroll
            for (int py = 0; py < PY; py++) {
                // the lanes on the tile edge get their own value back (halo garbage, never stored)
                hW[py] = __shfl_up_sync(0xffffffffu, x[py][PX - 1], 1, C::LX);
                hE[py] = __shfl_down_sync(0xffffffffu, x[py][0], 1, C::LX);
            }
#pragma unroll
            for (int px = 0; px < PX; px++) {
                hN[px] = pr[(px * TH + rN) * PW + g];
                hS[px] = pr[(px * TH + rS) * PW + g];
            }
            // in-place update; `up[px]` carries the old value of the row above
            double up[PX], fac[PX];
#pragma unroll
            for (int px = 0; px < PX; px++) { up[px] = hN[px]; fac[px] = CHEB ? omc[px] * taus.t[s - 1] : omc[px]; }
#pragma unroll
            for (int py = 0; py < PY; py++) {
                double cur[PX];
#pragma unroll
                for (int px = 0; px < PX; px++) cur[px] = x[py][px];
                if (py == PY - 1) row_update(py, hW[py], hE[py], up, hS, cur, x[py], fac);
                else row_update(py, hW[py], hE[py], up, x[py + 1], cur, x[py], fac);
#pragma unroll
                for (int px = 0; px < PX; px++) up[px] = cur[px];
            }
            if (s < T) {
                publish_rows(P0 + (s & 1) * C::CELLS, x[0], x[PY - 1]);
                __syncthreads();
            }
        }

        // ---- interior back to HBM through one bulk tensor store ---------------------------
        {
            double *out = OUT;                 // dense OW x OH box
#pragma unroll
            for (int py = 0; py < PY; py++) {
                const int r = r0 + py - T;
                if (r >= 0 && r < OH) {
                    if constexpr (PX == 4) {
                        // same bank argument as for the patch loads: lanes with bit 2 set store their second half first
                        const bool hs = (lane & 4) != 0;
                        const int ca = c0 - TE + (hs ? 2 : 0), cb = c0 - TE + (hs ? 0 : 2);
                        const double2 va = make_double2(hs ? x[py][2] : x[py][0], hs ? x[py][3] : x[py][1]);
                        const double2 vb = make_double2(hs ? x[py][0] : x[py][2], hs ? x[py][1] : x[py][3]);
                        if (ca >= 0 && ca < OW) *reinterpret_cast<double2 *>(out + r * OW + ca) = va;
                        if (cb >= 0 && cb < OW) *reinterpret_cast<double2 *>(out + r * OW + cb) = vb;
                    } else {
#pragma unroll
                        for (int px = 0; px < PX; px += 2) {   // c0, TE and OW are even: pairs never straddle the box edge
                            const int c = c0 + px - TE;
                            if (c >= 0 && c < OW)
                                *reinterpret_cast<double2 *>(out + r * OW + c) = make_double2(x[py][px], x[py][px + 1]);
                        }
                    }
                }
            }
        }
        fence_proxy_async();                   // generic-proxy writes -> visible to the TMA engine
        __syncthreads();
        if (tid == 0) {
            // PEER: the store map covers the own rows only (halo rows belong to the neighbours' pushes)
            tma_store_2d(map_out, ox, PEER ? oy - peer.above : oy, OUT);
            tma_commit();
        }
        if constexpr (PEER) {
            const int entry_no = k * (int)gridDim.x + (int)blockIdx.x;
            if (entry_no >= peer.lead && entry_no < peer.lead + peer.nboundary) {     // the boundary tiles follow the lead tiles in the list
                // rows of the output box among the first / last H own rows -> the neighbour's halo rows, 16 bytes per
                // store; only those rows are walked (at most H of the OH rows of the box)
                const int oyo = oy - peer.above;                                   // first output row, own-row coordinates
                const int ncol2 = min(OW, peer.Nx - ox + 1) >> 1;                  // column pairs inside the domain (the last may be half)
                double *const up = src ? peer.up[0] : peer.up[1], *const down = src ? peer.down[0] : peer.down[1];
                auto push_rows = [&](double *dstbase, int r_lo, int r_hi, int row_shift) {
                    // box rows [r_lo, r_hi) -> neighbour rows (oyo + r - row_shift)
                    const int n = (r_hi - r_lo) * ncol2;
                    for (int e = tid; e < n; e += C::NT) {
                        const int r = r_lo + e / ncol2, cpair = (e % ncol2) * 2;
                        const double2 v = *reinterpret_cast<const double2 *>(OUT + r * OW + cpair);
                        double *q = dstbase + (long long)(oyo + r - row_shift) * peer.pitch + ox + cpair;
                        if (ox + cpair + 1 >= peer.Nx) *q = v.x; else *reinterpret_cast<double2 *>(q) = v;
                    }
                };
                if (up) push_rows(up, max(0, -oyo), min(OH, peer.H - oyo), 0);
                if (down) push_rows(down, max(0, peer.own - peer.H - oyo), min(OH, peer.own - oyo), peer.own - peer.H);
                __syncthreads();               // every thread's push stores are issued (ordered before thread 0's fence below)
                if (tid == 0) pushed++;
            } else if (tid == 0) {
                // a tile after the pushes the rows have long left the SM: the fence in count_issue does not wait for them
                if (counted) count_finish();
                if (pushed > 0) count_issue();
            }
        }
    }
    if constexpr (PEER) {
        if (tid == 0) {
            if (counted) count_finish();
            if (pushed > 0) { count_issue(); count_finish(); }
        }
    }
    if (tid == 0) tma_wait_all0();             // stores complete before the CTA's smem goes away
}

// ------------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    uint64_t version = 0;        // tensor-map generation the graph was captured with
    int T = 0, fam = 0, src = 0, count = 0, grid_limit = 0, gather32 = 0;
    const uint32_t *list = nullptr;
    const void *lut = nullptr;
    double omega = 0;
};

struct TmaState {
    EncodeTiledFn encode = nullptr;
    uint64_t version = 0;        // bumped whenever the tensor maps are re-encoded
    std::vector<GraphEntry> graphs;
    size_t graph_evict = 0;
    TmaMaps maps;
    int cfg_T = 0;               // temporal depth the maps were encoded for
    int ow = 0, oh = 0, tiles_x = 0, tiles_y = 0;
    void *key_x0 = nullptr, *key_x1 = nullptr, *key_code = nullptr;
    int64_t key_Nx = 0, key_Ny = 0, key_pitch = 0;
    int cfg_TH = 0;              // tile height the maps were encoded for
    int64_t key_srow0 = 0, key_srows = 0;   // rows the store maps cover (slab peer mode: the own rows only)
    std::vector<const void *> attr_fns;   // kernels whose dynamic shared-memory limit has been raised
    int max_smem_optin = 0;
};

static int encode_2d(deff2d_ctx *c, TmaState *ts, CUtensorMap *m, CUtensorMapDataType dt, int esize, void *base,
                     uint64_t d0, uint64_t d1, uint64_t pitch_bytes, uint32_t b0, uint32_t b1)
{
    cuuint64_t dims[2] = {d0, d1};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {b0, b1};
    cuuint32_t estr[2] = {1, 1};
    (void)esize;
    CUresult r = ts->encode(m, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(c, "cuTensorMapEncodeTiled failed (CUresult %d) dims %llu x %llu pitch %llu box %u x %u", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)pitch_bytes, b0, b1);
        return DEFF2D_ERR_CUDA;
    }
    return DEFF2D_OK;
}

// tile geometries: square 64 x 64 tiles, 256 threads, 16 cells per thread, one warp per 8 tile rows
//   family 3: 4 x 4 cells per thread, 16 lanes side by side, two patch rows per warp
//   family 4: 2 x 8 cells per thread, 32 lanes side by side: half the shared-memory row exchange of family 3 per
//             sweep (2 + 2 values published and read per thread instead of 4 + 4) for twice the W / E shuffles
template <int T, int F> struct Family;
template <int T> struct Family<T, 3> { using type = Cfg<T, 4, 4, 1, 8, 16>; };
template <int T> struct Family<T, 4> { using type = Cfg<T, 2, 8, 1, 8, 32>; };
// (twelve warps with 2 x 6 cells per thread on 64 x 72 tiles -- 166 registers, three warps per scheduler instead of two:
//  866 / 816 / 612 GLUP/s on config 2 / blob 4096^2 / percolation 2048^2 at depth 6 against 874 / 821 / 644: the sweep is
//  bound by instruction dispatch (a DFMA holds a scheduler's dispatch port for two cycles) and the prologue by the LSU
//  pipe, not by latency, so a third warp buys nothing; scripts/probe/pipe_probe.cu has the measured pipe rates)
// (nine warps on 64 x 72 tiles do not work: registers are allocated per four warps, so 288 threads get 168 each)
// (the 2 x 8 layout with four warps on 64 x 32 tiles and two CTAs per SM, each in its own phase of its own tile: 736
//  GLUP/s at its best depth 4 against 871 -- the per-tile prologue weighs more on half-size tiles than the overlap gains)

static bool attr_done(TmaState *ts, const void *fn)
{
    for (const void *f : ts->attr_fns) if (f == fn) return true;
    ts->attr_fns.push_back(fn);
    return false;
}

template <int T, int F, int VAR>
static int launch_T(deff2d_ctx *c, TmaState *ts, int src, const uint32_t *list, int count, cudaStream_t stream,
                    const ChebTaus &taus = ChebTaus(), const PeerArgs &peer = PeerArgs())
{
    using C = typename Family<T, F>::type;
    auto kern = list ? k_sweep_tma<C, true, VAR> : k_sweep_tma<C, false, VAR>;
    const size_t smem = C::SMEM;
    if (!attr_done(ts, (const void *)kern)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error(c, "cudaFuncSetAttribute(smem %zu) failed: %s", smem, cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
    }
    const int ntiles = list ? count : ts->tiles_x * ts->tiles_y;
    if (ntiles < 1) return DEFF2D_OK;
    int grid = c->prop.multiProcessorCount;
    if (c->grid_limit > 0 && grid > c->grid_limit) grid = c->grid_limit;
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, C::NT, smem, stream>>>(ts->maps, src, c->clut.p, c->clut32.p, 1.0 - c->omega, ts->tiles_x, ntiles, list, taus, peer);
    return DEFF2D_OK;
}

template <int T, int F>
static int prepare_T(deff2d_ctx *c, TmaState *ts)
{
    using C = typename Family<T, F>::type;
    if ((int)C::SMEM > ts->max_smem_optin) { set_error(c, "tile needs %zu B smem > %d", C::SMEM, ts->max_smem_optin); return DEFF2D_ERR_STATE; }
    const int64_t srow0 = c->store_row0, srows = (c->store_rows > 0) ? c->store_rows : c->Ny;
    const bool same = ts->cfg_T == T && ts->cfg_TH == C::TH && ts->key_srow0 == srow0 && ts->key_srows == srows && ts->key_x0 == c->x[0].p && ts->key_x1 == c->x[1].p && ts->key_code == c->idx16.p &&
                      ts->key_Nx == c->Nx && ts->key_Ny == c->Ny && ts->key_pitch == c->pitch;
    if (same) return DEFF2D_OK;
    int rc;
    for (int b = 0; b < 2; b++) {
        if ((rc = encode_2d(c, ts, &ts->maps.x_load[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, c->x[b].p, (uint64_t)c->pitch,
                            (uint64_t)c->rows, (uint64_t)c->pitch * 8, C::TW, C::TH))) return rc;
        if ((rc = encode_2d(c, ts, &ts->maps.x_store[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8,
                            c->x[b].p + (1 + srow0) * c->pitch + DEFF2D_XOFF, (uint64_t)c->Nx, (uint64_t)srows, (uint64_t)c->pitch * 8,
                            C::OW, C::OH))) return rc;
    }
    if (!c->idx16.p) { set_error(c, "tiled sweep: the per-cell table indices have not been built"); return DEFF2D_ERR_STATE; }
    if ((rc = encode_2d(c, ts, &ts->maps.idx, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, c->idx16.p, (uint64_t)c->pitch,
                        (uint64_t)c->rows, (uint64_t)c->pitch * 2, C::IW, C::TH))) return rc;
    ts->version++;
    ts->cfg_T = T;
    ts->cfg_TH = C::TH;
    ts->ow = C::OW; ts->oh = C::OH;
    ts->tiles_x = (int)((c->Nx + C::OW - 1) / C::OW);
    ts->tiles_y = (int)((c->Ny + C::OH - 1) / C::OH);
    ts->key_x0 = c->x[0].p; ts->key_x1 = c->x[1].p; ts->key_code = c->idx16.p;
    ts->key_Nx = c->Nx; ts->key_Ny = c->Ny; ts->key_pitch = c->pitch;
    ts->key_srow0 = srow0; ts->key_srows = srows;
    return DEFF2D_OK;
}

static TmaState *tma_state(deff2d_ctx *c)
{
    TmaState *ts = static_cast<TmaState *>(c->tma);
    if (!ts) {
        ts = new TmaState();
        c->tma = ts;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            (void)cudaGetLastError();
            ts->encode = nullptr;
        } else ts->encode = (EncodeTiledFn)fn;
        cudaDeviceGetAttribute(&ts->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    }
    return ts;
}

// Output-box size of the tiles of temporal depth T (64 x 64 tiles in both families).
static int k2_family(const deff2d_ctx *c) { return (c->tile_family >= 3 && c->tile_family <= 4) ? c->tile_family : c->k2_default_family; }
static int family_th(int fam) { (void)fam; return 64; }      // both layouts use 64 x 64 tiles

void tma_tile_geometry(const deff2d_ctx *c, int T, int *ow, int *oh)
{
    const int te = (T + 1) & ~1;
    *ow = 64 - 2 * te;
    *oh = family_th(c ? k2_family(c) : DEFF2D_DEFAULT_TILE_FAMILY) - 2 * T;
}

// One pass of depth T (1..8) from x[src] into x[src ^ 1] over the tiles of `list` (NULL: the whole
// tile grid) on `stream`.
static int pass_from(deff2d_ctx *c, int T, int src, const uint32_t *list, int count, cudaStream_t stream)
{
    TmaState *ts = tma_state(c);
    if (!ts->encode) { set_error(c, "cuTensorMapEncodeTiled is not available from this driver"); return DEFF2D_ERR_CUDA; }
    if (T < 1 || T > 8) { set_error(c, "temporal depth %d out of range", T); return DEFF2D_ERR_ARG; }
    int rc = DEFF2D_OK;
    const int fam = k2_family(c);
#define DEFF2D_VAR(TT, FF)                                                                                     \
    {                                                                                                          \
        if ((rc = prepare_T<TT, FF>(c, ts))) return rc;                                                        \
        rc = launch_T<TT, FF, 0>(c, ts, src, list, count, stream);                                             \
        if (rc) return rc;                                                                                     \
    }
#define DEFF2D_VAR32(TT)                                                                                       \
    {                                                                                                          \
        if ((rc = prepare_T<TT, 4>(c, ts))) return rc;                                                         \
        rc = launch_T<TT, 4, 4>(c, ts, src, nullptr, 0, stream);                                               \
        if (rc) return rc;                                                                                     \
    }
    // interface-rich single domains and NCCL slabs (no tile list) in the default layout: the 32-bit gather variant
#define DEFF2D_CASE(TT)                                                                   \
    case TT:                                                                              \
        if (fam == 4 && c->gather32 && !list) DEFF2D_VAR32(TT)                            \
        else if (fam == 4) DEFF2D_VAR(TT, 4) else DEFF2D_VAR(TT, 3)                       \
        break;
    switch (T) {
        DEFF2D_CASE(1) DEFF2D_CASE(2) DEFF2D_CASE(3) DEFF2D_CASE(4) DEFF2D_CASE(5) DEFF2D_CASE(6) DEFF2D_CASE(7) DEFF2D_CASE(8)
    }
#undef DEFF2D_CASE
#undef DEFF2D_VAR32
#undef DEFF2D_VAR
    return DEFF2D_OK;
}

// NON-PARITY Chebyshev mode (chebyshev.cu): one pass of 8 Richardson steps with the factors tau[0..7] from x[c->cur] into
// x[c->cur ^ 1] over the whole tile grid; flips c->cur.  The context's table must hold omega = 1.
int tma_cheb_pass(deff2d_ctx *c, const double tau[8])
{
    TmaState *ts = tma_state(c);
    if (!ts->encode) { set_error(c, "cuTensorMapEncodeTiled is not available from this driver"); return DEFF2D_ERR_CUDA; }
    ChebTaus t;
    for (int k = 0; k < 8; k++) t.t[k] = tau[k];
    int rc;
    if ((rc = prepare_T<8, 4>(c, ts))) return rc;
    if ((rc = launch_T<8, 4, 2>(c, ts, c->cur, nullptr, 0, c->stream, t))) return rc;
    c->cur ^= 1;
    c->launches++;
    return DEFF2D_OK;
}

// Slab peer mode (slab.cu): one pass of depth T over `list` (`lead` interior tiles, then the `nboundary` boundary tiles, then the rest) with the halo
// push fused into the kernel; flips c->cur.  c->store_row0 / store_rows select the own rows for the local store.
template <int T>
static int peer_launch(deff2d_ctx *c, TmaState *ts, const uint32_t *list, int count, const PeerArgs &pa)
{
    using C = typename Family<T, 4>::type;
    int rc;
    if ((rc = prepare_T<T, 4>(c, ts))) return rc;
    auto kern = k_sweep_tma<C, true, 1>;
    if (!attr_done(ts, (const void *)kern)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { set_error(c, "cudaFuncSetAttribute(smem %zu) failed: %s", (size_t)C::SMEM, cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
    }
    int grid = c->prop.multiProcessorCount;
    if (grid > count) grid = count;
    kern<<<grid, C::NT, C::SMEM, c->stream>>>(ts->maps, c->cur, c->clut.p, c->clut32.p, 1.0 - c->omega, ts->tiles_x, count, list, ChebTaus(), pa);
    return DEFF2D_OK;
}

int tma_peer_pass(deff2d_ctx *c, int T, const uint32_t *list, int count, const void *peer_args)
{
    TmaState *ts = tma_state(c);
    if (!ts->encode) { set_error(c, "cuTensorMapEncodeTiled is not available from this driver"); return DEFF2D_ERR_CUDA; }
    if (T < 1 || T > 8 || count < 1) { set_error(c, "peer pass: bad depth / tile count"); return DEFF2D_ERR_ARG; }
    const PeerArgs &pa = *static_cast<const PeerArgs *>(peer_args);
    int rc = DEFF2D_OK;
    switch (T) {
    case 1: rc = peer_launch<1>(c, ts, list, count, pa); break;
    case 2: rc = peer_launch<2>(c, ts, list, count, pa); break;
    case 3: rc = peer_launch<3>(c, ts, list, count, pa); break;
    case 4: rc = peer_launch<4>(c, ts, list, count, pa); break;
    case 5: rc = peer_launch<5>(c, ts, list, count, pa); break;
    case 6: rc = peer_launch<6>(c, ts, list, count, pa); break;
    case 7: rc = peer_launch<7>(c, ts, list, count, pa); break;
    default: rc = peer_launch<8>(c, ts, list, count, pa); break;
    }
    if (rc) return rc;
    c->cur ^= 1;
    c->launches++;
    return DEFF2D_OK;
}

// One pass from x[c->cur] into x[c->cur ^ 1]; does not flip c->cur.
int tma_pass(deff2d_ctx *c, int T, const uint32_t *list, int count, cudaStream_t stream)
{
    int rc = pass_from(c, T, c->cur, list, count, stream);
    if (rc) return rc;
    c->launches++;
    return DEFF2D_OK;
}

// `npasses` passes of depth T on c->stream, flipping c->cur after each.  Long runs go through a
// CUDA graph of GRAPH_PASSES kernel nodes (captured once per configuration and replayed): the
// host then issues one launch per 32 passes -- on small domains and in slab mode the per-launch
// host cost (6-8 us measured on the GPU box) otherwise exceeds the kernel time.
#define GRAPH_PASSES 32

int tma_passes(deff2d_ctx *c, int T, int64_t npasses, const uint32_t *list, int count)
{
    TmaState *ts = tma_state(c);
    int rc;
    while (npasses >= GRAPH_PASSES && c->use_graphs) {
        // make sure the tensor maps are current before looking a graph up (re-encoding bumps the version)
        if (ts->cfg_T != T || ts->cfg_TH != family_th(k2_family(c)) || ts->key_srow0 != c->store_row0 || ts->key_srows != ((c->store_rows > 0) ? c->store_rows : c->Ny) || ts->key_x0 != c->x[0].p || ts->key_x1 != c->x[1].p ||
            ts->key_code != c->idx16.p || ts->key_Nx != c->Nx || ts->key_Ny != c->Ny || ts->key_pitch != c->pitch) {
            // a direct pass re-encodes the maps; then the graphs of the old maps are dropped below
            if ((rc = tma_pass(c, T, list, count, c->stream))) return rc;
            c->cur ^= 1;
            npasses--;
            continue;
        }
        GraphEntry *g = nullptr;
        for (auto &e : ts->graphs)
            if (e.exec && e.version == ts->version && e.T == T && e.fam == k2_family(c) && e.src == c->cur && e.list == list &&
                e.count == count && e.grid_limit == c->grid_limit && e.lut == (const void *)c->clut32.p && e.gather32 == c->gather32 && e.omega == c->omega) { g = &e; break; }
        if (!g) {
            // drop stale graphs, then capture GRAPH_PASSES passes
            for (auto &e : ts->graphs)
                if (e.exec && e.version != ts->version) { cudaGraphExecDestroy(e.exec); e.exec = nullptr; }
            GraphEntry *slot = nullptr;
            for (auto &e : ts->graphs) if (!e.exec) { slot = &e; break; }
            if (!slot) {
                if (ts->graphs.size() < 12) { ts->graphs.emplace_back(); slot = &ts->graphs.back(); }
                else { slot = &ts->graphs[ts->graph_evict++ % ts->graphs.size()]; cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
            }
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) { set_error(c, "cudaStreamBeginCapture failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
            rc = DEFF2D_OK;
            for (int k = 0; k < GRAPH_PASSES && !rc; k++) rc = pass_from(c, T, c->cur ^ (k & 1), list, count, c->stream);
            e = cudaStreamEndCapture(c->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess || !graph) { set_error(c, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
            e = cudaGraphInstantiate(&slot->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { slot->exec = nullptr; set_error(c, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
            slot->version = ts->version; slot->T = T; slot->fam = k2_family(c); slot->src = c->cur; slot->list = list;
            slot->count = count; slot->grid_limit = c->grid_limit; slot->lut = c->clut32.p; slot->gather32 = c->gather32; slot->omega = c->omega;
            g = slot;
        }
        cudaError_t e = cudaGraphLaunch(g->exec, c->stream);
        if (e != cudaSuccess) { set_error(c, "cudaGraphLaunch failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
        c->launches += GRAPH_PASSES;
        npasses -= GRAPH_PASSES;          // an even number of passes: c->cur is unchanged
    }
    while (npasses > 0) {
        if ((rc = tma_pass(c, T, list, count, c->stream))) return rc;
        c->cur ^= 1;
        npasses--;
    }
    return DEFF2D_OK;
}

// Up to n sweeps with the context's tile list and depth: whole passes of depth tblock (through
// tma_passes), or one shallower pass for the remainder.  *done = sweeps enqueued.
int launch_sweep_tma(deff2d_ctx *c, int64_t n, int T, int64_t *done)
{
    *done = 0;
    if (T < 1) T = 1;
    if (T > 8) T = 8;
    int rc;
    if (n >= T) {
        const int64_t passes = n / T;
        if ((rc = tma_passes(c, T, passes, c->tile_list, c->tile_count))) return rc;
        *done = passes * T;
    } else {
        if ((rc = tma_pass(c, (int)n, c->tile_list, c->tile_count, c->stream))) return rc;
        c->cur ^= 1;
        *done = n;
    }
    return DEFF2D_OK;
}

void tma_destroy(deff2d_ctx *c)
{
    if (TmaState *ts = static_cast<TmaState *>(c->tma))
        for (auto &e : ts->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    delete static_cast<TmaState *>(c->tma);
    c->tma = nullptr;
}

}  // namespace deff2d
