/*
 * deff_oracle.c -- CPU restatement of the reference's effective-diffusivity solve.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under effectivediffusivityfvm_b200/ links, imports or
 * executes anything in oracle/.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4); this restatement is pinned against outputs of the reference's
 * own unmodified host code + kernel body executed on CPU threads (oracle/_ref, built by
 * oracle/build_ref.sh from /root/reference) -- see tests/golden/make_golden.py and
 * tests/test_oracle_golden.py -- and against the analytic known-answer cases of the
 * reference documentation (doc section 5.3).
 *
 * Every function cites the reference lines it follows
 * (cuh = /root/reference/Deff2DGPU/Deff2D.cuh).  Arithmetic order is kept identical to
 * the reference so that, compiled with -ffp-contract=off, results are bit-identical to
 * the reference host code compiled the same way.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ options */

typedef struct {
    double Ds, Df, Dg;      /* cuh:20-22 */
    int ampx, ampy;         /* cuh:23-24 */
    double CL, CR;          /* cuh:25-26 */
    long max_iter;          /* cuh:27 */
    double tol;             /* cuh:28 */
    int nphase;             /* cuh:36 */
    int check_every;        /* cuh:1174 (10000 in the reference) */
    double omega;           /* cuh:72 (2/3 in the reference) */
    int verbose;
} orc_opts;

typedef struct {
    double porosity, SVF, LVF;  /* cuh:44-46 */
    double deff;                /* normalised: deff_raw / DCF (cuh:1802, 1601, 2017) */
    double deff_raw;            /* cuh:1309 */
    double conv;                /* cuh:1275 (signed) */
    int pathflag;               /* cuh:50 */
    int nstages;
    long iters[16];             /* sweeps per stage (value returned by JacobiGPU) */
    double stage_deff_raw[16];
    double stage_D[16];         /* the continuation value used at each stage */
    long total_iters;
    double loop_seconds;        /* analogue of gpuTime: non-PreCond loops only (cuh:1311) */
    long nchecks;               /* number of convergence checks recorded in trace */
} orc_result;

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------- small host numerics */

/* cuh:347-360 */
static double whm(double w1, double w2, double x1, double x2)
{
    double H = (w1 + w2) / (w1 / x1 + w2 / x2);
    return H;
}

ORC_API double orc_weighted_harmonic_mean(double w1, double w2, double x1, double x2)
{
    return whm(w1, w2, x1, x2);
}

/* cuh:383-408 : fraction of source pixels < 150, accumulated as += 1/total */
ORC_API double orc_porosity(const unsigned char *img, int W, int H)
{
    double total = (double)H * W;
    double por = 0;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++)
            if (img[i * W + j] < 150) por += 1.0 / total;
    return por;
}

/* cuh:411-448 : SVF/LVF by value equality on the amplified D grid */
ORC_API void orc_fracts3(const double *D, int Nx, int Ny, double DCS, double DCF,
                         double *SVF, double *LVF)
{
    double total = (double)((long)Nx * Ny);
    double s = 0, l = 0;
    for (int i = 0; i < Ny; i++)
        for (int j = 0; j < Nx; j++) {
            if (D[(size_t)i * Nx + j] == DCS) s += 1.0 / total;
            else if (D[(size_t)i * Nx + j] == DCF) l += 1.0 / total;
        }
    *SVF = s;
    *LVF = l;
}

/* 2-phase: cuh:1773-1785 / 1988-2000 ; 3-phase: cuh:1557-1578 */
ORC_API void orc_fill_D(const unsigned char *img, int W, int H, int ampx, int ampy, int nphase,
                        double DCS, double DCF, double DCG, double *D)
{
    int Nx = W * ampx, Ny = H * ampy;
    for (int i = 0; i < Ny; i++) {
        int tr = i / ampy;
        for (int j = 0; j < Nx; j++) {
            int tc = j / ampx;
            unsigned char p = img[tr * W + tc];
            double v;
            if (nphase == 2) v = (p < 150) ? DCF : DCS;
            else v = (p > 200) ? DCS : ((p < 50) ? DCG : DCF);
            D[(size_t)i * Nx + j] = v;
        }
    }
}

/* Solid mask feeding FloodFill.  cuh:1693-1701 (2-phase, >150) and cuh:1364-1377
 * (3-phase, >200).  The reference indexes image and Grid with the un-amplified
 * width (quirk Q10) which reads out of bounds when MeshAmp > 1; for amp == 1 the two
 * are identical.  The oracle implements the evident intent: pixel (i/ampy, j/ampx). */
ORC_API void orc_grid_mask(const unsigned char *img, int W, int H, int ampx, int ampy, int thr,
                           unsigned int *Grid)
{
    int Nx = W * ampx, Ny = H * ampy;
    for (int i = 0; i < Ny; i++)
        for (int j = 0; j < Nx; j++)
            Grid[(size_t)i * Nx + j] = img[(i / ampy) * W + (j / ampx)] > thr ? 1u : 0u;
}

/* cuh:557-713.  Same reachability semantics as the reference's std::set open list
 * (the result of a flood fill does not depend on pop order): 4-connected through
 * non-solid cells, periodic in y (cuh:641-665), not in x (cuh:675, 687); seeds are the
 * non-solid left-column cells (cuh:597-600) plus -- reference bug, quirk Q11 --
 * `if (Domain[indexR == -1])` (cuh:601) which tests Domain[0]: every right-column cell
 * of row r is seeded whenever Domain[0] != 0 at the time row r is examined.
 * PathFlag is set when a popped cell lies in the last column (cuh:619-621).
 * Unreached non-solid cells get Grid = 2 (cuh:701-708). */
ORC_API int orc_floodfill(unsigned int *Grid, int Nx, int Ny, int *pathflag)
{
    size_t n = (size_t)Nx * Ny;
    int *Domain = (int *)malloc(sizeof(int) * n);
    int *queue = (int *)malloc(sizeof(int) * (n + (size_t)Ny + 1));
    if (!Domain || !queue) { free(Domain); free(queue); return -1; }
    size_t qh = 0, qt = 0;
    for (size_t k = 0; k < n; k++) Domain[k] = (Grid[k] == 1) ? 1 : -1;
    for (int row = 0; row < Ny; row++) {
        size_t indexL = (size_t)row * Nx;
        size_t indexR = (size_t)(row + 1) * Nx - 1;
        if (Domain[indexL] == -1) { Domain[indexL] = 0; queue[qt++] = (int)indexL; }
        if (Domain[0]) {                    /* cuh:601: Domain[indexR == -1] == Domain[0] */
            /* the reference inserts unconditionally (solid or not); its std::set
             * de-duplicates keys, a duplicate queue entry here only re-expands a cell,
             * which is idempotent.  The queue has Ny+1 spare slots for that. */
            Domain[indexR] = 0;
            queue[qt++] = (int)indexR;
        }
    }
    int pf = 0;
    while (qh < qt) {
        int idx = queue[qh++];
        int row = idx / Nx, col = idx % Nx;
        if (col == Nx - 1) pf = 1;
        int tr;
        size_t t;
        tr = (row == 0) ? Ny - 1 : row - 1;                     /* north, periodic */
        t = (size_t)tr * Nx + col;
        if (Domain[t] == -1) { Domain[t] = 0; queue[qt++] = (int)t; }
        tr = (row == Ny - 1) ? 0 : row + 1;                     /* south, periodic */
        t = (size_t)tr * Nx + col;
        if (Domain[t] == -1) { Domain[t] = 0; queue[qt++] = (int)t; }
        if (col != 0) {                                         /* west */
            t = (size_t)row * Nx + col - 1;
            if (Domain[t] == -1) { Domain[t] = 0; queue[qt++] = (int)t; }
        }
        if (col != Nx - 1) {                                    /* east */
            t = (size_t)row * Nx + col + 1;
            if (Domain[t] == -1) { Domain[t] = 0; queue[qt++] = (int)t; }
        }
    }
    for (size_t k = 0; k < n; k++)
        if (Domain[k] == -1) Grid[k] = 2;
    /* NB (reference behaviour): a right-column cell seeded through the cuh:601 bug has
     * Domain = 0 even if it is solid; Grid keeps its value 1 for those (only -1 cells
     * are rewritten), so the solid mask is unchanged. */
    free(Domain);
    free(queue);
    if (pf) *pathflag = 1;
    return 0;
}

/* cuh:815-902 (Grid == NULL) and cuh:715-812 (Grid != NULL).  A is AoS [n][5] =
 * (P,W,E,S,N).  Expression order is the reference's. */
ORC_API void orc_discretize(const double *D, double *A, double *b, int Nx, int Ny, double CL,
                            double CR, const unsigned int *Grid)
{
    double dxw, dxe, dys, dyn;
    double kw, ke, ks, kn;
    double dx = 1.0 / Nx, dy = 1.0 / Ny;   /* cuh:1682-1683 */
    for (int i = 0; i < Ny; i++) {
        for (int j = 0; j < Nx; j++) {
            size_t index = (size_t)i * Nx + j;
            b[index] = 0;
            for (int k = 0; k < 5; k++) A[index * 5 + k] = 0;
            if (Grid && (Grid[index] == 1 || Grid[index] == 2)) {   /* cuh:750-752 */
                A[index * 5 + 0] = 1;
                b[index] = 0;
                continue;
            }
            if (j == 0) {
                dxe = dx;
                ke = whm(dxe / 2, dxe / 2, D[index], D[index + 1]);
                dxw = dx / 2;
                kw = D[index];
                A[index * 5 + 2] = -ke * dy / dxe;
                A[index * 5 + 0] += (ke * dy / dxe + kw * dy / dxw);
                b[index] += CL * kw * dy / dxw;
            } else if (j == Nx - 1) {
                dxw = dx;
                kw = whm(dxw / 2, dxw / 2, D[index], D[index - 1]);
                dxe = dx / 2;
                ke = D[index];
                A[index * 5 + 1] = -kw * dy / dxw;
                A[index * 5 + 0] += (ke * dy / dxe + kw * dy / dxw);
                b[index] += CR * ke * dy / dxe;
            } else {
                dxw = dx;
                kw = whm(dxw / 2, dxw / 2, D[index], D[index - 1]);
                dxe = dx;
                ke = whm(dxe / 2, dxe / 2, D[index], D[index + 1]);
                A[index * 5 + 1] = -kw * dy / dxw;
                A[index * 5 + 2] = -ke * dy / dxe;
                A[index * 5 + 0] += (ke * dy / dxe + kw * dy / dxw);
            }
            if (i == 0) {
                dys = dy;
                ks = whm(dys / 2, dys / 2, D[index + Nx], D[index]);
                A[index * 5 + 3] = -ks * dx / dys;
                A[index * 5 + 0] += (ks * dx / dys);
            } else if (i == Ny - 1) {
                dyn = dy;
                kn = whm(dyn / 2, dyn / 2, D[index], D[index - Nx]);
                A[index * 5 + 4] = -kn * dx / dyn;
                A[index * 5 + 0] += kn * dx / dyn;
            } else {
                dyn = dy;
                kn = whm(dyn / 2, dyn / 2, D[index], D[index - Nx]);
                dys = dy;
                ks = whm(dys / 2, dys / 2, D[index + Nx], D[index]);
                A[index * 5 + 3] = -ks * dx / dys;
                A[index * 5 + 4] = -kn * dx / dyn;
                A[index * 5 + 0] += (kn * dx / dyn + ks * dx / dys);
            }
        }
    }
}

/* cuh:69-92 (updateX_SOR) with w = omega; omega = 1 reproduces updateX_V1 up to the
 * reference's `1/A0*(...)` vs `w/A0*(...)` spelling (cuh:115). */
ORC_API void orc_sweep(const double *A, const double *x, const double *b, double *xNew, long n,
                       int Nx, double w)
{
#pragma omp parallel for schedule(static)
    for (long r = 0; r < n; r++) {
        double sigma = 0;
        for (int j = 1; j < 5; j++) {
            double a = A[r * 5 + j];
            if (a != 0) {
                if (j == 1) sigma += a * x[r - 1];
                else if (j == 2) sigma += a * x[r + 1];
                else if (j == 3) sigma += a * x[r + Nx];
                else sigma += a * x[r - Nx];
            }
        }
        xNew[r] = (1.0 - w) * x[r] + w / A[r * 5 + 0] * (b[r] - sigma);
    }
}

/* cuh:1252-1264 */
ORC_API double orc_flux_deff(const double *x, const double *D, int Nx, int Ny, double CL,
                             double CR)
{
    double dx = 1.0 / Nx;
    double Q1 = 0, Q2 = 0;
    for (int j = 0; j < Ny; j++) {
        double mfl = D[(size_t)j * Nx] * (x[(size_t)j * Nx] - CL) / (dx / 2.0);
        double mfr = D[(size_t)(j + 1) * Nx - 1] * (CR - x[(size_t)(j + 1) * Nx - 1]) / (dx / 2.0);
        Q1 += mfl;
        Q2 += mfr;
    }
    double qAvg = (Q1 + Q2) / (2.0 * Ny);
    return qAvg / (CR - CL);
}

/* cuh:451-494 (dead code in the reference; definition of the residual diagnostic) */
ORC_API double orc_residual(int numRows, int numCols, double TL, double TR, const double *cmap,
                            const double *D)
{
    double dx = 1.0 / numCols, dy = 1.0 / numRows;
    double qE, qW, qS, qN, R = 0;
    for (int row = 0; row < numRows; row++) {
        for (int col = 0; col < numCols; col++) {
            size_t c = (size_t)row * numCols + col;
            if (col == 0) {
                qW = dy / (dx / 2) * D[c] * (cmap[c] - TL);
                qE = dy / (dx) * whm(dx / 2, dx / 2, D[c], D[c + 1]) * (cmap[c + 1] - cmap[c]);
            } else if (col == numCols - 1) {
                qW = dy / (dx) * whm(dx / 2, dx / 2, D[c], D[c - 1]) * (cmap[c] - cmap[c - 1]);
                qE = dy / (dx / 2) * D[c] * (TR - cmap[c]);
            } else {
                qW = dy / (dx) * whm(dx / 2, dx / 2, D[c], D[c - 1]) * (cmap[c] - cmap[c - 1]);
                qE = dy / (dx) * whm(dx / 2, dx / 2, D[c], D[c + 1]) * (cmap[c + 1] - cmap[c]);
            }
            if (row == 0) {
                qN = 0;
                qS = dy / dx * whm(dx / 2, dx / 2, D[c + numCols], D[c]) * (cmap[c + numCols] - cmap[c]);
            } else if (row == numRows - 1) {
                qS = 0;
                qN = dy / dx * whm(dx / 2, dx / 2, D[c - numCols], D[c]) * (cmap[c] - cmap[c - numCols]);
            } else {
                qS = dy / dx * whm(dx / 2, dx / 2, D[c + numCols], D[c]) * (cmap[c + numCols] - cmap[c]);
                qN = dy / dx * whm(dx / 2, dx / 2, D[c - numCols], D[c]) * (cmap[c] - cmap[c - numCols]);
            }
            R += fabs(qW - qE + qN - qS);
        }
    }
    return R / ((double)numCols * numRows);
}

/* The solve loop, cuh:1163-1314 (JacobiGPU) / cuh:1024-1160 (JacobiGPUPreCond).
 * x is the warm-start field on entry and the field after `iters` sweeps on exit.
 * trace (optional, capacity trace_cap): un-normalised Deff at every check.
 * Returns iterCount exactly as the reference does. */
ORC_API long orc_jacobi(const double *A, const double *b, double *x, double *x_tmp, const double *D,
                        int Nx, int Ny, double CL, double CR, double tol, long max_iter,
                        int check_every, double omega, double *deff_out, double *conv_out,
                        double *trace, long trace_cap, long *ntrace, double *seconds)
{
    long n = (long)Nx * Ny;
    long iterCount = 0;
    double deffNew = 1, deffOld = 5, percentChange = 100.0;   /* cuh:1171-1173 */
    long nt = 0;
    int have_conv = 0;
    double conv = 0;
    /* device buffers of the reference: d_temp = input, d_x = output; after each sweep
     * d_temp <- d_x (cuh:1281).  Here: pointer roles, then a final copy into x. */
    double *cur = x_tmp, *nxt = x;
    memcpy(x_tmp, x, sizeof(double) * (size_t)n);               /* cuh:1190-1203 */
    double t0 = now_s();
    int swapped = 0;
    while (iterCount < max_iter && tol < fabs(percentChange)) { /* cuh:1232 */
        orc_sweep(A, cur, b, nxt, n, Nx, omega);                /* cuh:1237 */
        if (iterCount % check_every == 0) {                     /* cuh:1243 */
            deffNew = orc_flux_deff(nxt, D, Nx, Ny, CL, CR);
            percentChange = (deffOld - deffNew) / (deffOld);    /* cuh:1265 */
            if (trace && nt < trace_cap) trace[nt] = deffNew;
            nt++;
            deffOld = deffNew;
            conv = percentChange;                               /* cuh:1275 */
            have_conv = 1;
        }
        /* cuh:1281: d_temp <- d_x  == swap roles */
        double *t = cur; cur = nxt; nxt = t;
        swapped ^= 1;
        iterCount++;
    }
    if (seconds) *seconds = now_s() - t0;
    /* newest field is in `cur` (cuh:1300 downloads d_x, which equals d_temp after the
     * last copy).  If no sweep ran, cur == x_tmp == copy of x. */
    if (cur != x) memcpy(x, cur, sizeof(double) * (size_t)n);
    (void)swapped;
    if (deff_out) *deff_out = deffNew;                          /* cuh:1309 */
    if (conv_out && have_conv) *conv_out = conv;
    if (ntrace) *ntrace = nt;
    return iterCount;
}

/* cuh:1730-1734 */
ORC_API void orc_init_x(double *x, int Nx, int Ny, double CL, double CR)
{
    for (int i = 0; i < Ny; i++)
        for (int j = 0; j < Nx; j++)
            x[(size_t)i * Nx + j] = (double)j / Nx * (CR - CL) + CL;
}

/* ------------------------------------------------------------ image drivers */

/* One image through the reference's driver logic.
 *   mode 0: 2-phase single  (SingleSim,        cuh:1635-1841)  DCF = 100^k continuation
 *   mode 1: 2-phase batch   (BatchSim body,    cuh:1867-2049)  one solve at Df
 *   mode 2: 3-phase         (SingleSim3Phase,  cuh:1316-1633 == BatchSim3Phase body)
 * field (optional, Nx*Ny doubles) receives the final concentration map.
 * trace (optional) receives un-normalised Deff at every check of every stage. */
ORC_API int orc_solve_image(const unsigned char *img, int W, int H, const orc_opts *o, int mode,
                            orc_result *res, double *field, double *trace, long trace_cap)
{
    int Nx = W * o->ampx, Ny = H * o->ampy;
    size_t n = (size_t)Nx * Ny;
    memset(res, 0, sizeof(*res));
    int check = o->check_every > 0 ? o->check_every : 10000;
    double omega = o->omega > 0 ? o->omega : 2.0 / 3.0;

    unsigned int *Grid = (unsigned int *)malloc(sizeof(unsigned int) * n);
    double *D = (double *)malloc(sizeof(double) * n);
    double *A = (double *)malloc(sizeof(double) * n * 5);
    double *b = (double *)malloc(sizeof(double) * n);
    double *x = (double *)malloc(sizeof(double) * n);
    double *xt = (double *)malloc(sizeof(double) * n);
    if (!Grid || !D || !A || !b || !x || !xt) {
        free(Grid); free(D); free(A); free(b); free(x); free(xt);
        return -1;
    }
    int pf = 0;
    long nt = 0;
    double secs;

    if (mode == 0 || mode == 1) {
        res->porosity = orc_porosity(img, W, H);                /* cuh:1656 / 1884 */
        orc_grid_mask(img, W, H, o->ampx, o->ampy, 150, Grid);  /* cuh:1693-1701 */
        orc_floodfill(Grid, Nx, Ny, &pf);                       /* cuh:1705 */
        res->pathflag = pf;
        orc_init_x(x, Nx, Ny, o->CL, o->CR);
        if (mode == 1) {
            double DCF = o->Df;
            orc_fill_D(img, W, H, o->ampx, o->ampy, 2, o->Ds, DCF, 0, D);
            orc_discretize(D, A, b, Nx, Ny, o->CL, o->CR, NULL);
            long ntr = 0;
            long it = orc_jacobi(A, b, x, xt, D, Nx, Ny, o->CL, o->CR, o->tol, o->max_iter, check,
                                 omega, &res->deff_raw, &res->conv, trace, trace_cap, &ntr, &secs);
            nt += ntr;
            res->loop_seconds += secs;
            res->iters[0] = it; res->stage_deff_raw[0] = res->deff_raw; res->stage_D[0] = DCF;
            res->nstages = 1; res->total_iters = it;
            res->deff = res->deff_raw / DCF;                    /* cuh:2017 */
        } else {
            double DCF_Max = o->Df;
            double DCF = 10.0f;                                 /* cuh:1714 */
            int count = 1;
            res->deff = 0; /* reference prints an uninitialised value when no stage runs (Q8) */
            while (DCF <= DCF_Max) {                            /* cuh:1761 */
                DCF = pow(100, count);
                if (DCF >= DCF_Max) DCF = DCF_Max;
                orc_fill_D(img, W, H, o->ampx, o->ampy, 2, o->Ds, DCF, 0, D);
                orc_discretize(D, A, b, Nx, Ny, o->CL, o->CR, NULL);
                long ntr = 0;
                long it = orc_jacobi(A, b, x, xt, D, Nx, Ny, o->CL, o->CR, o->tol, o->max_iter,
                                     check, omega, &res->deff_raw, &res->conv,
                                     trace ? trace + nt : NULL, trace ? trace_cap - nt : 0, &ntr,
                                     &secs);
                nt += ntr;
                res->loop_seconds += secs;
                int s = res->nstages;
                if (s < 16) { res->iters[s] = it; res->stage_deff_raw[s] = res->deff_raw; res->stage_D[s] = DCF; }
                res->nstages++;
                res->total_iters += it;
                res->deff = res->deff_raw / DCF;                /* cuh:1802 */
                if (DCF == DCF_Max) break;                      /* cuh:1812 */
                count++;
            }
        }
    } else {
        orc_grid_mask(img, W, H, o->ampx, o->ampy, 200, Grid);  /* cuh:1364-1377 */
        orc_floodfill(Grid, Nx, Ny, &pf);                       /* cuh:1381 */
        res->pathflag = pf;
        orc_init_x(x, Nx, Ny, o->CL, o->CR);                    /* cuh:1402-1408 */
        double DCG_Temp = 10;                                   /* cuh:1492 */
        double tolP = o->tol * 10;                              /* cuh:1501 */
        long maxP = 1000000;                                    /* cuh:1502 */
        while (DCG_Temp < o->Dg) {                              /* cuh:1504 */
            orc_fill_D(img, W, H, o->ampx, o->ampy, 3, o->Ds, o->Df, DCG_Temp, D);
            orc_discretize(D, A, b, Nx, Ny, o->CL, o->CR, Grid);
            long ntr = 0;
            double dr = 0, cv = 0;
            long it = orc_jacobi(A, b, x, xt, D, Nx, Ny, o->CL, o->CR, tolP, maxP, check, omega,
                                 &dr, &cv, trace ? trace + nt : NULL, trace ? trace_cap - nt : 0,
                                 &ntr, &secs);
            nt += ntr;
            int s = res->nstages;
            if (s < 16) { res->iters[s] = it; res->stage_deff_raw[s] = dr; res->stage_D[s] = DCG_Temp; }
            res->nstages++;
            res->total_iters += it;
            DCG_Temp = DCG_Temp * 10;                           /* cuh:1547 */
        }
        orc_fill_D(img, W, H, o->ampx, o->ampy, 3, o->Ds, o->Df, o->Dg, D);
        orc_fracts3(D, Nx, Ny, o->Ds, o->Df, &res->SVF, &res->LVF);  /* cuh:1582 */
        orc_discretize(D, A, b, Nx, Ny, o->CL, o->CR, Grid);
        long ntr = 0;
        long it = orc_jacobi(A, b, x, xt, D, Nx, Ny, o->CL, o->CR, o->tol, o->max_iter, check, omega,
                             &res->deff_raw, &res->conv, trace ? trace + nt : NULL,
                             trace ? trace_cap - nt : 0, &ntr, &secs);
        nt += ntr;
        res->loop_seconds += secs;
        int s = res->nstages;
        if (s < 16) { res->iters[s] = it; res->stage_deff_raw[s] = res->deff_raw; res->stage_D[s] = o->Dg; }
        res->nstages++;
        res->total_iters += it;
        res->deff = res->deff_raw / o->Df;                      /* cuh:1601 */
    }
    res->nchecks = nt;
    if (field) memcpy(field, x, sizeof(double) * n);
    free(Grid); free(D); free(A); free(b); free(x); free(xt);
    return 0;
}

/* Fixed-sweep throughput probe used by bench.py's cpu_baseline leg: assemble once the
 * reference's way, run `sweeps` sweeps of cuh:69-92 over all host threads, return
 * seconds spent in the sweep loop (the analogue of the reference's event time). */
ORC_API double orc_time_sweeps(const unsigned char *img, int W, int H, const orc_opts *o,
                               int three_phase, long sweeps, double *deff_raw_out)
{
    int Nx = W * o->ampx, Ny = H * o->ampy;
    size_t n = (size_t)Nx * Ny;
    unsigned int *Grid = NULL;
    double *D = (double *)malloc(sizeof(double) * n);
    double *A = (double *)malloc(sizeof(double) * n * 5);
    double *b = (double *)malloc(sizeof(double) * n);
    double *x = (double *)malloc(sizeof(double) * n);
    double *xt = (double *)malloc(sizeof(double) * n);
    if (!D || !A || !b || !x || !xt) { free(D); free(A); free(b); free(x); free(xt); return -1; }
    double omega = o->omega > 0 ? o->omega : 2.0 / 3.0;
    if (three_phase) {
        int pf = 0;
        Grid = (unsigned int *)malloc(sizeof(unsigned int) * n);
        orc_grid_mask(img, W, H, o->ampx, o->ampy, 200, Grid);
        orc_floodfill(Grid, Nx, Ny, &pf);
        orc_fill_D(img, W, H, o->ampx, o->ampy, 3, o->Ds, o->Df, o->Dg, D);
    } else {
        orc_fill_D(img, W, H, o->ampx, o->ampy, 2, o->Ds, o->Df, 0, D);
    }
    orc_discretize(D, A, b, Nx, Ny, o->CL, o->CR, Grid);
    /* first-touch the iterate buffers in the same static partition the sweep uses */
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)n; r++) { xt[r] = 0; }
    orc_init_x(x, Nx, Ny, o->CL, o->CR);
    double *cur = x, *nxt = xt;
    double t0 = now_s();
    for (long s = 0; s < sweeps; s++) {
        orc_sweep(A, cur, b, nxt, (long)n, Nx, omega);
        double *t = cur; cur = nxt; nxt = t;
    }
    double secs = now_s() - t0;
    if (deff_raw_out) *deff_raw_out = orc_flux_deff(cur, D, Nx, Ny, o->CL, o->CR);
    free(Grid); free(D); free(A); free(b); free(x); free(xt);
    return secs;
}
