/*
 * deff2d.h -- C ABI of libdeff2d, the B200-native (sm_100a) effective-diffusivity solve.
 *
 * Drop-in boundary for the hot path of adama-wzr/EffectiveDiffusivityFVM
 * (image -> phases -> FVM coefficients -> damped-Jacobi sweeps -> boundary-flux Deff).
 * Plain pointers and sizes only; no torch / CUDA types appear in any signature.
 * Citations: cuh = Deff2DGPU/Deff2D.cuh, cu = Deff2DGPU/Deff2D.cu of the reference.
 *
 * Every function returns 0 on success or a negative deff2d_status; the message of the
 * last error is available from deff2d_last_error().  Nothing here calls getchar() or
 * resets the device (reference quirk Q16, cuh:914, cuh:1015).  A context is not
 * thread-safe: one caller thread per context.  Calls block until their results are on
 * the host unless stated otherwise.  There is NO CPU fallback: without a CUDA device
 * deff2d_create() fails with DEFF2D_ERR_CUDA.
 */
#ifndef DEFF2D_H
#define DEFF2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEFF2D_VERSION 100

typedef enum {
    DEFF2D_OK = 0,
    DEFF2D_ERR_CUDA = -1,     /* CUDA runtime / driver error, or no device */
    DEFF2D_ERR_ARG = -2,      /* invalid argument */
    DEFF2D_ERR_ALLOC = -3,    /* host or device allocation failed */
    DEFF2D_ERR_STATE = -4,    /* call sequence error (e.g. sweeps before a domain is loaded) */
    DEFF2D_ERR_IO = -5,       /* file could not be read / written / decoded */
    DEFF2D_ERR_NCCL = -6
} deff2d_status;

/* Which reference driver's logic to follow (cu:17-50). */
typedef enum {
    DEFF2D_MODE_2PH_SINGLE = 0,   /* SingleSim        cuh:1635-1841  DCF = 100^k continuation  */
    DEFF2D_MODE_2PH_BATCH = 1,    /* BatchSim body    cuh:1867-2049  one solve at Df           */
    DEFF2D_MODE_3PH = 2           /* SingleSim3Phase  cuh:1316-1633  == BatchSim3Phase body    */
} deff2d_mode;

/* Replaces the numeric part of `options` (cuh:18-37). */
typedef struct {
    double Ds, Df, Dg;        /* DCsolid, DCfluid, DCgas                  cuh:20-22 */
    int amp_x, amp_y;         /* MeshIncreaseX / MeshIncreaseY            cuh:23-24 */
    double CL, CR;            /* CLeft, CRight                            cuh:25-26 */
    int64_t max_iter;         /* MAX_ITER                                 cuh:27    */
    double tol;               /* ConvergeCriteria                         cuh:28    */
    int mode;                 /* deff2d_mode                                        */
    int check_every;          /* 0 -> 10000, the reference's hard-coded cadence  cuh:1174 */
    double omega;             /* 0 -> 2/3, the reference's damping               cuh:72   */
    int solver;               /* 0 (default): the reference's damped Jacobi, its iterate and stop rule.  1: NON-PARITY
                               * Chebyshev-accelerated Jacobi (csrc/chebyshev.cu): per-sweep relaxation factors from the
                               * Chebyshev polynomial on the tiled kernel, stop on the true relative residual
                               * (residual_tol, default 1e-8) -- the converged answer, not the reference iterate */
    int verbose;              /* 1: print the reference's per-check / per-stage stdout lines */
    double residual_tol;      /* > 0: NON-PARITY stop rule -- at every check stop when the reference's (dead) Residual
                               * (mean |flux imbalance| per cell, cuh:451-494) is <= residual_tol instead of testing
                               * the relative Deff change (cuh:1232); 0 (default): the reference rule.  Single-GPU domains.
                               * (That Residual is kept verbatim: it scales every face by dy/dx and so plateaus above 0.) */
    int strict_reference;     /* 1 (default): bit-faithful reference quirks.  0: defined behaviour instead --
                               * 2-phase single with Df < 10 runs one stage at Df (Q8, cuh:1714/1761), FloodFill
                               * seeds the left column only (Q11, cuh:601), pixel == 150 is solid in the
                               * FloodFill mask as in D (Q12, cuh:1695 vs 1779) */
} deff2d_params;

#define DEFF2D_MAX_STAGES 16

/* Replaces `simulationInfo` (cuh:39-52) + what the drivers print/write per image. */
typedef struct {
    double porosity;          /* calcPorosity, 2-phase only               cuh:383-408 */
    double SVF, LVF;          /* calcFracts3D, 3-phase only               cuh:411-448 */
    double deff;              /* normalised: deff_raw / DCF               cuh:1802, 1601, 2017 */
    double deff_raw;          /* myImg->deff as JacobiGPU leaves it       cuh:1309 */
    double conv;              /* signed relative change at the last check cuh:1275 */
    int pathflag;             /* FloodFill                                cuh:619-621 */
    int nstages;              /* continuation stages executed (incl. the final one) */
    int64_t iters[DEFF2D_MAX_STAGES];          /* value JacobiGPU returns, per stage */
    double stage_deff_raw[DEFF2D_MAX_STAGES];
    double stage_D[DEFF2D_MAX_STAGES];         /* DCF (2-phase) or DCG_Temp (3-phase) of the stage */
    int64_t total_iters;
    int64_t n_cells;          /* mesh.nElements                           cuh:1681 */
    double solve_ms;          /* device time of the non-PreCond solve loops = CSV `Time`*1000  cuh:1311 */
    double total_ms;          /* device time of everything the call launched */
    double last_df;           /* DCF used by the last stage (CSV `df` column of batch mode, cuh:2034) */
} deff2d_result;

typedef struct deff2d_ctx deff2d_ctx;

/* ---- library / context --------------------------------------------------------------- */

int deff2d_version(void);

/* Replaces initializeGPU / unInitializeGPU (cuh:904-1021): one persistent context per
 * device owns streams, arenas and lookup tables; no per-image allocation or reset. */
int deff2d_create(deff2d_ctx **out, int device);
void deff2d_destroy(deff2d_ctx *ctx);
const char *deff2d_last_error(const deff2d_ctx *ctx);   /* ctx may be NULL: last create error */

void deff2d_default_params(deff2d_params *p);            /* shipped input.txt defaults */

/* ---- the whole path, host buffers in, Deff out ------------------------------------- */

/* One image through the reference driver logic selected by p->mode, covering threshold +
 * mesh amplification (cuh:1773-1785, 1557-1578), FloodFill (cuh:557-713), assembly
 * (cuh:715-902, matrix-free here), the continuation stages (cuh:1759-1817, 1492-1597),
 * the damped-Jacobi loop with the reference stop rule (cuh:1163-1314) and Deff
 * (cuh:1252-1264).  gray: H*W row-major 8-bit pixels as stbi_load(...,1) returns them
 * (cuh:342).  field: NULL or (H*amp_y)*(W*amp_x) doubles receiving the concentration
 * map the reference would write to its CMAP file (cuh:497-524). */
int deff2d_solve_image(deff2d_ctx *ctx, const uint8_t *gray, int W, int H,
                       const deff2d_params *p, deff2d_result *res, double *field);

/* `count` images of identical size packed back to back (count*H*W bytes): the body of
 * BatchSim / BatchSim3Phase (cuh:1867-2049, 2056-2419) for every image.  results: count
 * entries.  fields: NULL or count*(H*amp_y)*(W*amp_x) doubles. */
/* The packed batch solve as a stream.  Images are pulled through `fetch` whenever a slot of the resident stack is
 * free, and every finished image is handed to `done` at once: the caller can decode ahead on its own threads while
 * the GPU sweeps and write each result row when it exists (the reference keeps every row until the end of the batch,
 * cuh:2051, doc 3.6).  Images finish in any order.
 *   fetch(user, k, dst, wait): the W*H pixels of image k into dst.  Return 0 = ok; 1 = not ready yet (allowed only
 *       when wait == 0: the solve goes on with the images it has and asks again); 2 = there is no image k (the batch
 *       ends at k); < 0 = error.  Images are requested in order k = 0, 1, 2, ...
 *   done(user, k, result, field): image k is finished; field is NULL or its concentration map (valid during the
 *       call).  A non-zero return aborts the batch.
 * count: upper bound on the number of images; *solved (optional): how many were solved.  Parameters outside the
 * packed mode (verbose = 1, solver / residual_tol, 2-phase single mode) return an argument error. */
typedef int (*deff2d_batch_fetch_fn)(void *user, int k, uint8_t *dst, int wait);
typedef int (*deff2d_batch_done_fn)(void *user, int k, const deff2d_result *result, const double *field);
int deff2d_solve_batch_stream(deff2d_ctx *ctx, int count, int W, int H, const deff2d_params *p,
                              deff2d_batch_fetch_fn fetch, deff2d_batch_done_fn done, void *user, int want_fields,
                              int *solved);
/* 1 if images of W x H pixels with these parameters go through the packed batch mode, 0 if they are solved one at a
 * time (verbose = 1, solver / residual_tol, 2-phase single mode, images above 16 M cells). */
int deff2d_batch_supported(const deff2d_params *p, int W, int H);
int deff2d_solve_batch(deff2d_ctx *ctx, const uint8_t *gray, int count, int W, int H,
                       const deff2d_params *p, deff2d_result *results, double *fields);

/* One image over several GPUs from one host process (not in the reference, which is
 * single-GPU: cudaSetDevice(0), cuh:908): ctxs[0..nctx) are contexts on different devices; the
 * domain is split into row slabs, one host thread per device, halo exchange and flux all-reduce
 * over NCCL.  Same driver logic, results and `field` layout as deff2d_solve_image. */
int deff2d_solve_image_slabs(deff2d_ctx *const *ctxs, int nctx, const uint8_t *gray, int W, int H,
                             const deff2d_params *p, deff2d_result *res, double *field);

/* ---- device-resident stepping (tests, benchmarks, multi-GPU slabs) ------------------- */

/* Upload an image, threshold/amplify it into the per-cell phase codes, run FloodFill
 * (3-phase: pinned mask; always: PathFlag), set x0 = j/Nx*(CR-CL)+CL (cuh:1730-1734) and
 * build the coefficient tables for (Ds, Df, Dg).  nphase = 2 or 3. */
int deff2d_domain_load(deff2d_ctx *ctx, const uint8_t *gray, int W, int H, int nphase,
                       const deff2d_params *p);
/* Same for a horizontal slab [row0, row0+rows) of a taller global domain of NyGlobal cell
 * rows (multi-GPU row decomposition).  gray holds the slab's own source rows plus
 * `halo_src` source rows above and below where they exist; pinned: NULL or the slab's
 * pinned mask incl. halo rows. */
int deff2d_domain_load_slab(deff2d_ctx *ctx, const uint8_t *gray, int W, int Hslab, int nphase,
                            const deff2d_params *p, int64_t row0, int64_t NyGlobal, int halo_rows,
                            const uint8_t *pinned);
/* The same slab from the WHOLE source image (W x H): the image is uploaded once (1 B per pixel), FloodFill runs on
 * the device over the whole domain (every rank floods its own copy: the flood needs global connectivity), and the
 * slab keeps the rows [row0, row0 + own_rows) plus `halo_rows` rows of each neighbour -- no host flood, no mask
 * upload, the same steps as deff2d_domain_load.  row0, own_rows, halo_rows in amplified rows. */
int deff2d_domain_load_slab_global(deff2d_ctx *ctx, const uint8_t *gray, int W, int H, int nphase,
                                   const deff2d_params *p, int64_t row0, int64_t own_rows, int halo_rows);
/* New continuation stage: only the 3 diffusivities change (cuh:1762, 1524); the phase
 * codes and the iterate stay resident (warm start, cuh:1793). */
int deff2d_domain_set_D(deff2d_ctx *ctx, double Ds, double Df, double Dg);
/* Enqueue `n` damped-Jacobi sweeps (cuh:69-92 + the D2D copy of cuh:1281 as a pointer
 * swap) on the context's stream; returns without synchronising. */
int deff2d_domain_sweeps(deff2d_ctx *ctx, int64_t n);
/* Same, bracketed by CUDA events on the launching stream; blocks; *ms = device time. */
int deff2d_domain_sweeps_timed(deff2d_ctx *ctx, int64_t n, float *ms);
/* Boundary-flux Deff of the current iterate (cuh:1252-1264), un-normalised; blocks.
 * q (optional): {Q1, Q2} partial sums of this domain/slab. */
int deff2d_domain_flux(deff2d_ctx *ctx, double *deff_raw, double *q);
/* Mean |flux imbalance| per cell, the reference's (dead) Residual (cuh:451-494); blocks. */
int deff2d_domain_residual(deff2d_ctx *ctx, double *res);
/* The reference solve loop on the resident domain: sweeps + checks every `check_every`
 * + stop rule (cuh:1232-1290).  Returns the reference's iterCount in *iters. */
int deff2d_domain_solve(deff2d_ctx *ctx, double tol, int64_t max_iter, int64_t *iters,
                        double *deff_raw, double *conv, double *trace, int trace_cap, int *ntrace);
int deff2d_domain_get_field(deff2d_ctx *ctx, double *field);        /* Ny*Nx doubles out */
int deff2d_domain_set_field(deff2d_ctx *ctx, const double *field);  /* Ny*Nx doubles in  */
int deff2d_domain_get_codes(deff2d_ctx *ctx, uint8_t *codes);       /* Ny*Nx bytes out: bits 0-1 phase, bit 2 pinned */
int deff2d_domain_info(deff2d_ctx *ctx, int64_t *Nx, int64_t *Ny, int *pathflag, double *porosity,
                       double *SVF, double *LVF);
int deff2d_sync(deff2d_ctx *ctx);

/* Select the sweep kernel: 0 = library default (TMA-staged tiled kernel, square 64 x 64 tiles, 8
 * sweeps per pass, CUDA-graph replay), 1 = plain streaming kernel (one sweep per HBM pass), 2..5 =
 * the tiled kernel with `tblock` (1..8) sweeps per pass in one of its tile geometries (2: 128 x 32
 * tiles with 2 x 8 cells per thread, 3: 128 x 32 with 4 x 4, 4: 128 x 32 with 2 x 4 and 512 threads,
 * 5: 64 x 64 with 4 x 4 -- the default's).  All give bit-identical iterates.  Tuning / test hook. */
int deff2d_set_kernel(deff2d_ctx *ctx, int kernel, int tblock);
/* Where FloodFill (cuh:557-713) runs for whole-domain loads: 0 = automatic (device from 64 K cells),
 * 1 = host (FIFO flood), 2 = device (label propagation).  Same result either way. */
int deff2d_set_floodfill(deff2d_ctx *ctx, int mode);
/* Sweeps per HBM pass of the default kernel (kernel 0). */
int deff2d_default_depth(const deff2d_ctx *ctx);
/* Cluster-resident sweeps (csrc/resident.cu): 0 = a single domain of up to 256 x 256 cells stays on chip for a
 * whole check interval (default), 1 = never (the tiled kernel runs instead), 2 = the images of a packed batch
 * too (measured slower than the tiled kernel on B200, kept selectable).  Same iterates bit for bit. */
int deff2d_set_resident(deff2d_ctx *ctx, int mode);
/* Packed batch mode (deff2d_solve_batch): at most `max_slots` images resident at a time
 * (0 = library default, sized from the image size); finished images are replaced from the
 * queue.  Tuning / test hook. */
int deff2d_set_batch_slots(deff2d_ctx *ctx, int max_slots);
/* CUDA-graph replay of long runs of passes: 1 (default) on, 0 off.  Tuning hook; the single-process
 * multi-GPU driver switches it off (NCCL cannot capture when its ranks are threads of one process). */
int deff2d_set_graphs(deff2d_ctx *ctx, int enable);
int deff2d_get_graphs(const deff2d_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t deff2d_kernel_launches(const deff2d_ctx *ctx);
/* The CUDA stream (cudaStream_t as an opaque pointer) all work of the context is enqueued on. */
void *deff2d_stream(deff2d_ctx *ctx);
/* Device pointers of the resident iterate buffers and row pitch in doubles (for peer /
 * NCCL halo exchange by the host layer); rows include one ghost row above and below. */
int deff2d_domain_buffers(deff2d_ctx *ctx, void **x_cur, void **x_next, int64_t *pitch, int64_t *rows);

/* ---- multi-GPU row-slab decomposition over NCCL --------------------------------------- */

/* size of the opaque NCCL unique id the ranks must share */
#define DEFF2D_NCCL_ID_BYTES 128
int deff2d_nccl_unique_id(uint8_t id[DEFF2D_NCCL_ID_BYTES]);       /* call on rank 0, broadcast by the host layer */
int deff2d_nccl_init(deff2d_ctx *ctx, const uint8_t id[DEFF2D_NCCL_ID_BYTES], int rank, int nranks);
/* Sweeps on a slab; the halo rows travel to the two neighbours (ncclSend/ncclRecv) whenever the next pass
 * needs more exact halo rows than are left -- enqueue only. */
int deff2d_slab_sweeps(deff2d_ctx *ctx, int64_t n);
/* Global Deff: local {Q1,Q2} -> ncclAllReduce(sum) -> same value on every rank; blocks. */
int deff2d_slab_flux(deff2d_ctx *ctx, double *deff_raw);
/* Peer-memory halo exchange, fused into the sweep kernel (one process or thread per GPU of one box): a rank sweeps
 * only its own rows; the tiles next to a neighbour run early in a pass and copy the rows that neighbour needs straight
 * into its halo rows over NVLink; flags in peer memory order the passes of neighbouring ranks.  No NCCL call between
 * passes, no recomputed halo rows.  After every slab load: each rank exports its handle, the host layer hands every
 * rank the handles of the ranks above / below (NULL at the ends), attach.  The attach is collective (it ends in an
 * all-reduce that doubles as a barrier) and returns an error on EVERY rank if any rank could not map its neighbours or
 * has a slab thinner than 2 x halo rows, so that the group can stay with the NCCL exchange together.  Same bits as the
 * NCCL deep-halo exchange; measured on B200s (GLUP/s, fused / NCCL): 2 GPUs 1 686 / 1 712, 4 GPUs 3 341 / 3 314,
 * 8 GPUs 6 426 / 5 790 on one 16384^2 domain -- the Python host layer uses it from 4 ranks up.  Halo rows: 8 is enough
 * (>= the pass depth, <= 64). */
#define DEFF2D_PEER_HANDLE_BYTES 320
int deff2d_slab_peer_export(deff2d_ctx *ctx, uint8_t handle[DEFF2D_PEER_HANDLE_BYTES]);
int deff2d_slab_peer_attach(deff2d_ctx *ctx, const uint8_t *above, const uint8_t *below);
int deff2d_slab_peer_detach(deff2d_ctx *ctx);      /* back to the NCCL exchange (every rank; needs >= pass-depth halo rows) */
/* Another rank of the group failed: abort this context's communicator so that its pending NCCL work returns
 * instead of waiting for a peer that will never arrive.  deff2d_nccl_init is needed again afterwards. */
int deff2d_slab_abort(deff2d_ctx *ctx);

/* ---- host-side pieces of the path (no GPU needed; CPU-testable) ---------------------- */

/* `count` additions of 1.0/total in double precision, as calcPorosity / calcFracts3D accumulate their fractions
 * (cuh:402, cuh:437) -- the same bits as the loop, without looping over every cell. */
double deff2d_accumulate_fraction(int64_t count, int64_t total);
/* Coefficient tables for one stage, exactly as the library uploads them: lut is
 * 2048*4 doubles -- for the 11-bit index  p | pW<<2 | pE<<4 | pS<<6 | pN<<8 | pinned<<10
 * the four sweep weights (w/A0)*c_f in order W,E,S,N following cuh:815-902 and cuh:89;
 * dead is 2048 bytes (1 where A0 == 0, reference quirk Q13).  Phases: 0 fluid, 1 solid,
 * 2 gas, 3 ghost (Dirichlet face for W/E, no-flux wall for S/N). */
int deff2d_build_tables(double Ds, double Df, double Dg, int64_t Nx, int64_t Ny, double CL, double CR,
                        double omega, double *lut, uint8_t *dead);
/* The compact planar form of `lut` that the tiled and the resident sweep gather from: clut is 4 * 1024 doubles
 * (planes wW, wE, wS, wN), slot[idx] for the 2048 table indices of deff2d_build_tables is the plane position of that
 * neighbourhood (1023: inert -- ghost or pinned cell, all weights 0; 0xffff: a neighbourhood that cannot occur with
 * `nphase` phases).  Interior neighbourhoods are ranked: single-phase first, then one differing neighbour, ... so
 * that the gathers of a warp touch as few cache lines as possible.  Either output may be NULL. */
int deff2d_compact_table(const double *lut, int nphase, double *clut, uint16_t *slot);
/* The compact table(s) as 32-bit halves, the form the tiled sweep gathers from on interface-rich media: per stage eight
 * planes of 1024 words -- the low words of wW, wE, wS, wN, then their high words (clut: nstages * 4 * 1024 doubles in,
 * clut32: nstages * 8 * 1024 words out). */
int deff2d_split_table(const double *clut, int nstages, uint32_t *clut32);
/* FloodFill (cuh:557-713) on a solid mask (1 = solid), incl. the y-periodic wrap and the
 * right-column seeding quirk (cuh:601).  grid: Ny*Nx bytes in/out (unreached non-solid
 * cells become 2).  Returns PathFlag (0/1) or a negative status. */
int deff2d_floodfill(uint8_t *grid, int64_t Nx, int64_t Ny);
/* Tile planning of the TMA kernel, exposed so that the decomposition logic is testable on a CPU box.
 * deff2d_tile_geometry: output box (ow x oh) and input tile (tw x th) of a pass of depth T.
 * deff2d_batch_plan: slot grid of the packed batch mode.  deff2d_batch_tile_list: tiles whose output touches an
 * active slot; `tiles` holds `cap` entries, the function returns the number of tiles or a negative status when
 * `cap` is too small. */
int deff2d_tile_geometry(int T, int *ow, int *oh, int *tw, int *th);
int deff2d_batch_plan(int64_t Nx, int64_t Ny, int count, int limit, int *GX, int *GY);
int deff2d_batch_tile_list(int64_t Nx, int64_t Ny, int GX, int GY, const int *active, int nactive, int T,
                           uint32_t *tiles, int cap);

/* input.txt parser with the reference's quirks (cuh:234-324): case-sensitive "Key: value"
 * lines, numerics parsed as double then cast, unknown keys ignored. */
typedef struct {
    deff2d_params p;
    int nphase;               /* Phases:    cuh:309-310 */
    int batch;                /* RunBatch:  cuh:305-306 */
    int num_images;           /* NumImages: cuh:307-308 */
    int print_cmap;           /* printCMap: cuh:289-290 */
    char input_name[1000];    /* InputName: cuh:275-277 */
    char output_name[1000];   /* OutputName:cuh:285-287 */
    char cmap_name[1000];     /* CMapName:  cuh:292-294 */
    int devices;              /* Devices: N -- extension, unknown to (and ignored by) the reference parser:
                               * split a packed batch over N GPUs; default 1 */
    int field_npy;            /* FieldNpy: 1 -- extension: with printCMap, also write <CMapName>.npy (binary field) */
} deff2d_input;
int deff2d_read_input_file(const char *path, deff2d_input *in);
/* CSV / CMAP writers with the reference's exact formats (cuh:177-232, 497-554). */
int deff2d_write_csv_single(const deff2d_input *in, const deff2d_result *r);
int deff2d_write_csv_batch(const deff2d_input *in, const deff2d_result *r, int count);
/* The same file row by row (index < 0: the header): lets a driver append every image's row as soon
 * as it is solved instead of losing everything on an interruption (reference doc 3.6, cuh:2051). */
int deff2d_append_csv_batch_row(const deff2d_input *in, int index, const deff2d_result *r);
int deff2d_write_cmap(const char *path, const double *field, int64_t Nx, int64_t Ny);
/* The same map as a NumPy .npy file (float64, shape (Ny, Nx)): binary companion of the CMAP text. */
int deff2d_write_field_npy(const char *path, const double *field, int64_t Nx, int64_t Ny);
/* The reference program: read ./input.txt-style file, run the selected driver (cu:17-50),
 * write the same files.  Images are decoded by the library's own readers (PGM/PNG content
 * under any file name; baseline JPEG). */
int deff2d_run_input_file(deff2d_ctx *ctx, const char *path);
/* Decode an image file to 8-bit gray.  *gray is malloc'd (free with deff2d_free). */
int deff2d_load_image(const char *path, uint8_t **gray, int *W, int *H, int *channels);
void deff2d_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* DEFF2D_H */
