// floodfill.cu -- FloodFill (Deff2D.cuh:557-713) on the device.
//
// The reference floods on the host with an ordered std::set (O(n log n)); the host FIFO version
// in host.cpp is O(n) but still costs ~0.4 s for the 32 M cells of BASELINE config 2 and sits on
// the end-to-end path of every image (PathFlag always, the pinned mask in 3-phase, cuh:1381).
// Reachability is order independent, so it is computed here by monotone label propagation:
//
//   state byte per cell:  1 solid, 0 reached, 0xFF open and not reached yet, 3 solid but seeded
//                         (reference quirk Q11, cuh:601: while cell (0,0) is solid every
//                         right-column cell, solid or not, is seeded and floods on)
//   k_ff_init   solid mask from the image (> 200 in 3-phase, > 150 in 2-phase; cuh:1368, 1695)
//               with mesh amplification, plus the seeds
//   k_ff_tile   one CTA per 64 x 256 tile: the tile and a one-cell ring (y periodic, cuh:641-665;
//               x closed, cuh:675/687) go to shared memory, reachability is propagated to a
//               local fixed point by alternating column and row sweeps, the tile is written back;
//               launched until no tile changes
//   k_ff_finish unreached open cells -> 2 (cuh:701-708), seeded solids back to 1, PathFlag = a
//               reached cell in the last column (cuh:619-621)
//
// The result is identical, cell for cell, to the host flood (tests/test_gpu_parity.py).
#include <cuda_runtime.h>

#include "context.h"

namespace deff2d {

#define FF_TH 64
#define FF_TW 256
#define FF_PITCH (FF_TW + 2 + 2)      // ring + padding to a multiple of 4

__device__ __forceinline__ bool ff_reached(unsigned v) { return v == 0u || v == 3u; }

// blockIdx.y: image of a packed batch (source images W x Hsrc apart, states Nx * Ny apart)
__global__ void __launch_bounds__(256)
k_ff_init(const uint8_t *__restrict__ img, int W, int Hsrc, int amp_x, int amp_y, int thr, uint8_t *__restrict__ st,
          long long Nx, long long Ny, int reference_quirk)
{
    const long long n = Nx * Ny;
    img += (size_t)blockIdx.y * W * Hsrc;
    st += (size_t)blockIdx.y * n;
    // cuh:601 tests Domain[0]: the quirk is on while cell (0,0) is solid
    const bool quirk = reference_quirk && img[0] > thr;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / Nx, j = k - i * Nx;
        const bool solid = img[(i / amp_y) * W + (int)(j / amp_x)] > thr;
        unsigned v = solid ? 1u : 0xFFu;
        if (j == 0 && !solid) v = 0u;                       // left-column seeds (cuh:589-597)
        if (j == Nx - 1 && quirk) v = solid ? 3u : 0u;      // right-column seeds of the quirk
        st[k] = (uint8_t)v;
    }
}

// blockIdx.z: image of a packed batch
__global__ void __launch_bounds__(256)
k_ff_tile(uint8_t *st, long long Nx, long long Ny, int *changed)
{
    st += (size_t)blockIdx.z * Nx * Ny;
    __shared__ uint8_t s[(FF_TH + 2) * FF_PITCH];
    __shared__ int any_work, blk_changed, iter_changed;
    const long long x0 = (long long)blockIdx.x * FF_TW, y0 = (long long)blockIdx.y * FF_TH;
    const int tid = threadIdx.x;
    if (tid == 0) { any_work = 0; blk_changed = 0; }
    __syncthreads();
    const int rows = (int)((Ny - y0 < FF_TH) ? (Ny - y0) : FF_TH);
    const int cols = (int)((Nx - x0 < FF_TW) ? (Nx - x0) : FF_TW);
    // ---- load tile + ring: local rows -1 .. rows, local columns -1 .. cols ------------------------
    bool has_unreached = false, has_reached = false;
    for (int k = tid; k < (rows + 2) * (FF_TW + 2); k += 256) {
        const int r = k / (FF_TW + 2), c = k - r * (FF_TW + 2);
        if (c > cols + 1) continue;
        long long gy = y0 + r - 1;
        const long long gx = x0 + c - 1;
        if (gy < 0) gy = Ny - 1;                             // periodic in y (cuh:641-665)
        if (gy >= Ny) gy -= Ny;
        unsigned v = 1u;                                     // closed in x (cuh:675, cuh:687)
        if (gx >= 0 && gx < Nx) v = st[gy * Nx + gx];
        s[r * FF_PITCH + c] = (uint8_t)v;
        const bool inner = (r >= 1 && r <= rows && c >= 1 && c <= cols);
        if (inner && v == 0xFFu) has_unreached = true;
        if (ff_reached(v)) has_reached = true;
    }
    if (has_unreached) atomicOr(&any_work, 1);
    if (has_reached) atomicOr(&any_work, 2);
    __syncthreads();
    if (any_work != 3) return;                               // nothing to reach, or nothing to reach it from
    // ---- local fixed point ---------------------------------------------------------------------
    for (int it = 0; it < 4096; it++) {
        if (tid == 0) iter_changed = 0;
        __syncthreads();
        bool ch = false;
        // column sweeps: thread = column, down then up
        if (tid < cols) {
            uint8_t *col = s + 1 + tid;
            bool carry = ff_reached(col[0]);                 // ring row above
            for (int r = 1; r <= rows; r++) {
                const unsigned v = col[r * FF_PITCH];
                if (v == 0xFFu) { if (carry) { col[r * FF_PITCH] = 0; ch = true; } }
                else carry = ff_reached(v);
            }
            carry = ff_reached(col[(rows + 1) * FF_PITCH]);  // ring row below
            for (int r = rows; r >= 1; r--) {
                const unsigned v = col[r * FF_PITCH];
                if (v == 0xFFu) { if (carry) { col[r * FF_PITCH] = 0; ch = true; } }
                else carry = ff_reached(v);
            }
        }
        __syncthreads();
        // row sweeps: 4 threads per row, 64 columns each, right then left
        {
            const int r = tid >> 2, seg = tid & 3;
            if (r < rows) {
                uint8_t *row = s + (r + 1) * FF_PITCH + 1;
                const int c0 = seg * 64, c1 = min(c0 + 64, cols);
                if (c0 < c1) {
                    bool carry = ff_reached(row[c0 - 1]);
                    for (int c = c0; c < c1; c++) {
                        const unsigned v = row[c];
                        if (v == 0xFFu) { if (carry) { row[c] = 0; ch = true; } }
                        else carry = ff_reached(v);
                    }
                    carry = ff_reached(row[c1]);
                    for (int c = c1 - 1; c >= c0; c--) {
                        const unsigned v = row[c];
                        if (v == 0xFFu) { if (carry) { row[c] = 0; ch = true; } }
                        else carry = ff_reached(v);
                    }
                }
            }
        }
        if (ch) { iter_changed = 1; blk_changed = 1; }
        __syncthreads();
        if (!iter_changed) break;
        __syncthreads();
    }
    // ---- write back ----------------------------------------------------------------------------
    if (blk_changed) {
        for (int k = tid; k < rows * FF_TW; k += 256) {
            const int r = k / FF_TW, c = k - r * FF_TW;
            if (c < cols && s[(r + 1) * FF_PITCH + 1 + c] == 0) st[(y0 + r) * Nx + x0 + c] = 0;
        }
        if (tid == 0) atomicOr(changed, 1);
    }
}

// blockIdx.y: image of a packed batch (one PathFlag per image)
__global__ void __launch_bounds__(256)
k_ff_finish(uint8_t *st, long long Nx, long long Ny, int *pathflag)
{
    const long long n = Nx * Ny;
    st += (size_t)blockIdx.y * n;
    pathflag += blockIdx.y;
    int pf = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const unsigned v = st[k];
        const long long j = k % Nx;
        if (j == Nx - 1 && ff_reached(v)) pf = 1;           // cuh:619-621
        if (v == 0xFFu) st[k] = 2;                           // cuh:701-708
        else if (v == 3u) st[k] = 1;
    }
    if (__any_sync(0xffffffffu, pf) && (threadIdx.x & 31) == 0) atomicOr(pathflag, 1);
}

// FloodFill of the amplified solid masks of `count` images (device, W x Hsrc each, back to back) into `st` (device,
// Nx * Ny bytes per image: 0 reached, 1 solid, 2 unreached open).  d_flags: device int[1 + count] scratch (changed
// flag, then one PathFlag per image); h_flags: pinned host int[1 + count].  pathflags[k] = PathFlag of image k.  Blocks.
int floodfill_device_batch(deff2d_ctx *c, const uint8_t *img, int W, int Hsrc, int amp_x, int amp_y, int thr, uint8_t *st,
                           int64_t Nx, int64_t Ny, int count, int *d_flags, int *h_flags, int *pathflags, int *passes,
                           bool reference_quirk)
{
    cudaStream_t s = c->stream;
    const long long n = Nx * Ny;
    if (count < 1) return DEFF2D_OK;
    if (count > 65535) { set_error(c, "device FloodFill: at most 65535 images per call"); return DEFF2D_ERR_ARG; }
    int blocks = (int)std::min<long long>((n + 255) / 256, count > 1 ? 64 : 148 * 16);
    k_ff_init<<<dim3((unsigned)blocks, (unsigned)count), 256, 0, s>>>(img, W, Hsrc, amp_x, amp_y, thr, st, Nx, Ny, reference_quirk ? 1 : 0);
    c->launches++;
    dim3 grid((unsigned)((Nx + FF_TW - 1) / FF_TW), (unsigned)((Ny + FF_TH - 1) / FF_TH), (unsigned)count);
    int total = 0;
    const int burst = 4;                                      // passes per host round trip
    for (;;) {
        cudaMemsetAsync(d_flags, 0, sizeof(int), s);
        for (int k = 0; k < burst; k++) k_ff_tile<<<grid, 256, 0, s>>>(st, Nx, Ny, d_flags);
        c->launches += burst;
        total += burst;
        cudaMemcpyAsync(h_flags, d_flags, sizeof(int), cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error(c, "device FloodFill failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
        if (!h_flags[0]) break;
        if (total > 1000000) { set_error(c, "device FloodFill did not converge"); return DEFF2D_ERR_STATE; }
    }
    cudaMemsetAsync(d_flags + 1, 0, sizeof(int) * (size_t)count, s);
    k_ff_finish<<<dim3((unsigned)blocks, (unsigned)count), 256, 0, s>>>(st, Nx, Ny, d_flags + 1);
    c->launches++;
    cudaMemcpyAsync(h_flags, d_flags, (1 + (size_t)count) * sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error(c, "device FloodFill failed: %s", cudaGetErrorString(e)); return DEFF2D_ERR_CUDA; }
    for (int k = 0; k < count; k++) pathflags[k] = h_flags[1 + k] ? 1 : 0;
    if (passes) *passes = total;
    return DEFF2D_OK;
}

// One image: see floodfill_device_batch.  flags: device / pinned host int[2] scratch.
int floodfill_device(deff2d_ctx *c, const uint8_t *img, int W, int Hsrc, int amp_x, int amp_y, int thr, uint8_t *st, int64_t Nx,
                     int64_t Ny, int *d_flags, int *h_flags, int *pathflag, int *passes, bool reference_quirk)
{
    return floodfill_device_batch(c, img, W, Hsrc, amp_x, amp_y, thr, st, Nx, Ny, 1, d_flags, h_flags, pathflag, passes, reference_quirk);
}

}  // namespace deff2d
