// host.cpp -- host-side pieces of the path that need no GPU: FloodFill, the input.txt
// parser and the CSV / CMAP writers, with the reference's formats and quirks.
#include "deff2d_internal.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <string>
#include <thread>

namespace deff2d {

// FloodFill, cuh:557-713.  grid: 1 = solid, anything else = open.  Reachability through
// open cells from the open cells of the left column; 4-connected; periodic in y
// (cuh:641-665) but not in x (cuh:675, cuh:687).  The reference's open list is an ordered
// std::set; the set of reached cells does not depend on the visiting order, so a flat FIFO
// gives the same result in O(n).  Reference quirk Q11 (cuh:601): the test
// `Domain[indexR == -1]` reads Domain[0], so while cell (0,0) is marked solid every
// right-column cell, solid or not, is seeded as well.  Unreached open cells become 2
// (cuh:701-708).  PathFlag = some visited cell lies in the last column (cuh:619-621).
int floodfill(uint8_t *grid, int64_t Nx, int64_t Ny, bool reference_quirk)
{
    const int64_t n = Nx * Ny;
    // state: 1 solid, 0 reached, 0xFF not reached yet   (Domain of cuh:573-587)
    std::vector<uint8_t> dom((size_t)n);
    for (int64_t k = 0; k < n; k++) dom[(size_t)k] = (grid[k] == 1) ? 1 : 0xFF;
    std::vector<int64_t> queue;
    queue.reserve((size_t)(n / 4 + Ny + 16));
    for (int64_t row = 0; row < Ny; row++) {
        const int64_t iL = row * Nx, iR = (row + 1) * Nx - 1;
        if (dom[(size_t)iL] == 0xFF) { dom[(size_t)iL] = 0; queue.push_back(iL); }
        if (reference_quirk && dom[0] != 0) {    // cuh:601
            dom[(size_t)iR] = 0;
            queue.push_back(iR);
        }
    }
    int pathflag = 0;
    size_t head = 0;
    auto visit = [&](int64_t t) {
        if (dom[(size_t)t] == 0xFF) { dom[(size_t)t] = 0; queue.push_back(t); }
    };
    while (head < queue.size()) {
        const int64_t idx = queue[head++];
        const int64_t row = idx / Nx, col = idx - row * Nx;
        if (col == Nx - 1) pathflag = 1;
        visit(((row == 0) ? Ny - 1 : row - 1) * Nx + col);      // north, periodic
        visit(((row == Ny - 1) ? 0 : row + 1) * Nx + col);      // south, periodic
        if (col != 0) visit(idx - 1);
        if (col != Nx - 1) visit(idx + 1);
    }
    for (int64_t k = 0; k < n; k++)
        if (dom[(size_t)k] == 0xFF) grid[k] = 2;
    return pathflag;
}

// cuh:402 / cuh:437: the reference accumulates fractions as `+= 1.0/total` per cell, so
// e.g. 0.3 prints as 0.299999999999983.  Reproduce the rounding, not the closed form -- but not by
// looping over 268 M cells: while the running sum stays inside one binade every addition adds the
// same multiple of that binade's ulp (the increment rounded to the ulp grid), so whole binades are
// jumped in one step and only the additions that cross a power of two (and the first few, where
// ties occur) are performed for real.  Bit-identical to the loop (tests/test_host_logic.py).
double accumulate_fraction(int64_t count, int64_t total)
{
    const double inc = 1.0 / (double)total;
    double s = 0;
    int64_t left = count;
    for (int k = 0; k < 64 && left > 0; k++, left--) s += inc;      // small sums: ties and exact additions, done literally
    if (left <= 0) return s;
    int ei = 0;
    const double mi = std::frexp(inc, &ei);                           // inc = mi * 2^ei, 0.5 <= mi < 1
    const uint64_t M = (uint64_t)std::ldexp(mi, 53);                  // 53-bit integer mantissa: inc = M * 2^(ei - 53)
    const int einc = ei - 53;
    while (left > 0) {
        int es = 0;
        const double ms = std::frexp(s, &es);
        const int eu = es - 53;                                       // ulp(s) = 2^eu while s stays in [2^(es-1), 2^es)
        const int sh = eu - einc;
        if (sh <= 0 || sh >= 63) { s += inc; left--; continue; }      // exact additions (or inc below half an ulp): literal
        const uint64_t q = M >> sh, r = M & ((1ull << sh) - 1), half = 1ull << (sh - 1);
        if (r == half) { s += inc; left--; continue; }                // a tie in this binade: round-half-even, literal
        const uint64_t d = q + (r > half ? 1 : 0);                    // every addition moves the sum by d ulps
        if (d == 0) return s;                                         // the increment no longer registers
        const uint64_t m = (uint64_t)std::ldexp(ms, 53);              // s = m * 2^eu, 2^52 <= m < 2^53
        const uint64_t room = ((1ull << 53) - 1 - m) / d;             // additions that stay below 2^es
        const uint64_t n = std::min<uint64_t>(room, (uint64_t)left);
        if (n > 0) { s = std::ldexp((double)(m + n * d), eu); left -= (int64_t)n; }
        if (left > 0) { s += inc; left--; }                           // the addition that crosses into the next binade
    }
    return s;
}

int stage_list(const deff2d_params *p, StageSpec *out, int cap)
{
    int n = 0;
    auto add = [&](double Ds, double Df, double Dg, double sd, double tol, int64_t mi, int pre, int q8) {
        if (n >= cap) return false;
        out[n].Ds = Ds; out[n].Df = Df; out[n].Dg = Dg; out[n].stageD = sd; out[n].tol = tol; out[n].max_iter = mi;
        out[n].precond = pre; out[n].defined_q8 = q8;
        n++;
        return true;
    };
    if (p->mode == DEFF2D_MODE_2PH_BATCH) {
        if (!add(p->Ds, p->Df, 0.0, p->Df, p->tol, p->max_iter, 0, 0)) return -1;              // cuh:2004-2017
    } else if (p->mode == DEFF2D_MODE_2PH_SINGLE) {
        const double DCF_Max = p->Df;
        double DCF = 10.0;                                                                       // cuh:1714
        int count = 1;
        if (DCF > DCF_Max && p->strict_reference == 0)                                           // defined behaviour for quirk Q8
            if (!add(p->Ds, DCF_Max, 0.0, DCF_Max, p->tol, p->max_iter, 0, 1)) return -1;
        while (DCF <= DCF_Max) {                                                                 // cuh:1761 (no stage when Df < 10, quirk Q8)
            DCF = std::pow(100, count);                                                          // cuh:1762
            if (DCF >= DCF_Max) DCF = DCF_Max;
            if (!add(p->Ds, DCF, 0.0, DCF, p->tol, p->max_iter, 0, 0)) return -1;
            if (DCF == DCF_Max) break;                                                           // cuh:1812
            count++;
        }
    } else if (p->mode == DEFF2D_MODE_3PH) {
        double DCG_Temp = 10;                                                                    // cuh:1492
        while (DCG_Temp < p->Dg) {                                                               // cuh:1504; tol*10, MAX_ITER 1e6: cuh:1501-1502
            if (!add(p->Ds, p->Df, DCG_Temp, DCG_Temp, p->tol * 10, 1000000, 1, 0)) return -1;
            DCG_Temp = DCG_Temp * 10;                                                            // cuh:1547
        }
        if (!add(p->Ds, p->Df, p->Dg, p->Dg, p->tol, p->max_iter, 0, 0)) return -1;              // cuh:1557-1591
    } else {
        return -1;
    }
    return n;
}

}  // namespace deff2d

DEFF2D_EXPORT int deff2d_floodfill(uint8_t *grid, int64_t Nx, int64_t Ny)
{
    if (!grid || Nx < 1 || Ny < 1) return DEFF2D_ERR_ARG;
    return deff2d::floodfill(grid, Nx, Ny);
}

DEFF2D_EXPORT int deff2d_version(void) { return DEFF2D_VERSION; }

DEFF2D_EXPORT void deff2d_free(void *p) { std::free(p); }

DEFF2D_EXPORT double deff2d_accumulate_fraction(int64_t count, int64_t total)
{
    if (count < 0 || total < 1) return 0.0;
    return deff2d::accumulate_fraction(count, total);
}

DEFF2D_EXPORT void deff2d_default_params(deff2d_params *p)
{
    // Deff2DGPU/input.txt:2-18
    std::memset(p, 0, sizeof(*p));
    p->Ds = 0; p->Df = 1; p->Dg = 1237500;
    p->amp_x = 1; p->amp_y = 1;
    p->CL = 0; p->CR = 1;
    p->max_iter = 500000;
    p->tol = 1e-5;
    p->mode = DEFF2D_MODE_3PH;
    p->check_every = 10000;
    p->omega = 2.0 / 3.0;
    p->solver = 0;
    p->verbose = 0;
    p->residual_tol = 0;
    p->strict_reference = 1;
}

// readInputFile, cuh:234-324.  `sscanf("%s %lf")` per line; keys are case-sensitive and
// include the colon; numeric values go through a double (so "MaxIter: 5e5" works,
// cuh:299-300); unknown lines are ignored.  Unlike the reference the struct starts from
// the shipped defaults instead of uninitialised memory.
DEFF2D_EXPORT int deff2d_read_input_file(const char *path, deff2d_input *in)
{
    if (!path || !in) return DEFF2D_ERR_ARG;
    std::memset(in, 0, sizeof(*in));
    deff2d_default_params(&in->p);
    in->nphase = 3; in->batch = 0; in->num_images = 0; in->print_cmap = 0; in->devices = 1; in->field_npy = 0;
    std::ifstream f(path);
    if (!f.is_open()) return DEFF2D_ERR_IO;
    std::string line;
    char key[1000], name[1000];
    while (std::getline(f, line)) {
        if (line.size() >= sizeof(key)) continue;
        double v = 0;
        key[0] = 0;
        const int got = std::sscanf(line.c_str(), "%999s %lf", key, &v);
        if (got < 1) continue;
        auto is = [&](const char *k) { return std::strcmp(key, k) == 0; };
        auto str = [&](char *dst) {
            if (std::sscanf(line.c_str(), "%999s %999s", key, name) == 2) std::strcpy(dst, name);
        };
        if (is("Ds:")) in->p.Ds = v;
        else if (is("Df:")) in->p.Df = v;
        else if (is("Dg:")) in->p.Dg = v;
        else if (is("MeshAmpX:")) in->p.amp_x = (int)v;
        else if (is("MeshAmpY:")) in->p.amp_y = (int)v;
        else if (is("InputName:")) str(in->input_name);
        else if (is("CR:")) in->p.CR = v;
        else if (is("CL:")) in->p.CL = v;
        else if (is("OutputName:")) str(in->output_name);
        else if (is("printCMap:")) in->print_cmap = (int)v;
        else if (is("CMapName:")) str(in->cmap_name);
        else if (is("Convergence:")) in->p.tol = v;
        else if (is("MaxIter:")) in->p.max_iter = (int64_t)v;
        else if (is("Verbose:")) in->p.verbose = (int)v;
        else if (is("RunBatch:")) in->batch = (int)v;
        else if (is("NumImages:")) in->num_images = (int)v;
        else if (is("Phases:")) in->nphase = (int)v;
        else if (is("Devices:")) in->devices = (int)v < 1 ? 1 : (int)v;
        else if (is("FieldNpy:")) in->field_npy = (int)v;
    }
    if (in->nphase == 3) in->p.mode = DEFF2D_MODE_3PH;
    else in->p.mode = in->batch ? DEFF2D_MODE_2PH_BATCH : DEFF2D_MODE_2PH_SINGLE;
    return DEFF2D_OK;
}

// outputSingle (cuh:177-188) / outputSingle3Phase (cuh:191-202): append header + one row.
DEFF2D_EXPORT int deff2d_write_csv_single(const deff2d_input *in, const deff2d_result *r)
{
    if (!in || !r) return DEFF2D_ERR_ARG;
    FILE *o = std::fopen(in->output_name, "a+");
    if (!o) return DEFF2D_ERR_IO;
    if (in->nphase == 3) {
        std::fprintf(o, "imgNum,SVF,LVF,PathFlag,Deff,Time,nElements,converge,ds,df,dg\n");
        std::fprintf(o, "%s,%f,%f,%d,%1.3e,%f,%d,%1.3e,%1.3e,%1.3e,%1.3e\n", in->input_name, r->SVF, r->LVF,
                     r->pathflag, r->deff, r->solve_ms / 1000, (int)r->n_cells, r->conv, in->p.Ds, in->p.Df,
                     in->p.Dg);
    } else {
        std::fprintf(o, "imgNum,porosity,PathFlag,Deff,Time,nElements,converge,ds,df\n");
        std::fprintf(o, "%s,%f,%d,%f,%f,%d,%f,%f,%f\n", in->input_name, r->porosity, r->pathflag, r->deff,
                     r->solve_ms / 1000, (int)r->n_cells, r->conv, in->p.Ds, in->p.Df);
    }
    std::fclose(o);
    return DEFF2D_OK;
}

// outputBatch (cuh:204-217) / outputBatch3Phase (cuh:219-232): header and rows.  The reference
// buffers all rows and writes them when the last image is done ("if the code is interrupted,
// all progress is lost", doc 3.6); the row-wise entry point lets the driver append each row
// as soon as its image has finished -- the finished file is byte-identical.
static void csv_batch_header(FILE *o, const deff2d_input *in)
{
    if (in->nphase == 3) std::fprintf(o, "imgNum,SVF,LVF,PathFlag,Deff,Time,nElements,converge,ds,df,dg\n");
    else std::fprintf(o, "imgNum,porosity,PathFlag,Deff,Time,nElements,converge,ds,df\n");
}

static void csv_batch_row(FILE *o, const deff2d_input *in, int i, const deff2d_result *r)
{
    if (in->nphase == 3)
        std::fprintf(o, "%d,%f,%f,%d,%1.5e,%f,%d,%1.5e,%1.5e,%1.5e,%1.5e\n", i, r->SVF, r->LVF, r->pathflag, r->deff,
                     r->solve_ms / 1000, (int)r->n_cells, r->conv, in->p.Ds, r->last_df, in->p.Dg);
    else
        std::fprintf(o, "%d,%f,%d,%f,%f,%d,%f,%f,%f\n", i, r->porosity, r->pathflag, r->deff, r->solve_ms / 1000,
                     (int)r->n_cells, r->conv, in->p.Ds, r->last_df);
}

DEFF2D_EXPORT int deff2d_write_csv_batch(const deff2d_input *in, const deff2d_result *r, int count)
{
    if (!in || (!r && count > 0)) return DEFF2D_ERR_ARG;
    FILE *o = std::fopen(in->output_name, "a+");
    if (!o) return DEFF2D_ERR_IO;
    csv_batch_header(o, in);
    for (int i = 0; i < count; i++) csv_batch_row(o, in, i, r + i);
    std::fclose(o);
    return DEFF2D_OK;
}

// Crash-safe variant: index < 0 appends the header, index >= 0 appends (and flushes) row `index`.
DEFF2D_EXPORT int deff2d_append_csv_batch_row(const deff2d_input *in, int index, const deff2d_result *r)
{
    if (!in || (index >= 0 && !r)) return DEFF2D_ERR_ARG;
    FILE *o = std::fopen(in->output_name, "a+");
    if (!o) return DEFF2D_ERR_IO;
    if (index < 0) csv_batch_header(o, in);
    else csv_batch_row(o, in, index, r);
    const bool ok = std::fflush(o) == 0;
    std::fclose(o);
    return ok ? DEFF2D_OK : DEFF2D_ERR_IO;
}

// createCMAP / createCMAPBatch, cuh:497-554: "X,Y,C" then "%d,%d,%1.3e" per cell, y outer,
// x inner, file truncated.  The reference issues one fprintf per cell (~700 MB of text for
// config 2); here blocks of rows are formatted by the host threads in parallel (the text of every
// cell still comes from snprintf with the reference's format, so it is byte-identical) and
// written in order.
DEFF2D_EXPORT int deff2d_write_cmap(const char *path, const double *field, int64_t Nx, int64_t Ny)
{
    if (!path || !field || Nx < 1 || Ny < 1) return DEFF2D_ERR_ARG;
    FILE *o = std::fopen(path, "w+");
    if (!o) return DEFF2D_ERR_IO;
    std::fputs("X,Y,C\n", o);
    const int64_t rows_per_block = std::max<int64_t>(1, (int64_t)(1 << 16) / Nx);
    const int64_t nblocks = (Ny + rows_per_block - 1) / rows_per_block;
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), nblocks));
    bool ok = true;
    // rounds of `nthreads` blocks: format in parallel, write sequentially
    std::vector<std::vector<char>> bufs((size_t)nthreads);
    std::vector<size_t> lens((size_t)nthreads, 0);
    for (int64_t b0 = 0; b0 < nblocks && ok; b0 += nthreads) {
        const int nb = (int)std::min<int64_t>(nthreads, nblocks - b0);
        auto fmt = [&](int t) {
            const int64_t i0 = (b0 + t) * rows_per_block, i1 = std::min<int64_t>(Ny, i0 + rows_per_block);
            std::vector<char> &buf = bufs[(size_t)t];
            buf.resize((size_t)((i1 - i0) * Nx) * 40 + 64);
            size_t pos = 0;
            for (int64_t i = i0; i < i1; i++)
                for (int64_t j = 0; j < Nx; j++)
                    pos += (size_t)std::snprintf(buf.data() + pos, 40, "%d,%d,%1.3e\n", (int)j, (int)i, field[i * Nx + j]);
            lens[(size_t)t] = pos;
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nb; t++) pool.emplace_back(fmt, t);
        fmt(0);
        for (auto &t : pool) t.join();
        for (int t = 0; t < nb && ok; t++) ok = std::fwrite(bufs[(size_t)t].data(), 1, lens[(size_t)t], o) == lens[(size_t)t];
    }
    if (std::fclose(o) != 0) ok = false;
    return ok ? DEFF2D_OK : DEFF2D_ERR_IO;
}

// The concentration map as a NumPy .npy file (format 1.0, little-endian float64, shape (Ny, Nx)):
// 8 bytes per cell instead of ~22 bytes of text, loadable with numpy.load -- the binary companion
// of the CMAP file for post-processing scripts such as the reference's contourC.py.
DEFF2D_EXPORT int deff2d_write_field_npy(const char *path, const double *field, int64_t Nx, int64_t Ny)
{
    if (!path || !field || Nx < 1 || Ny < 1) return DEFF2D_ERR_ARG;
    FILE *o = std::fopen(path, "wb");
    if (!o) return DEFF2D_ERR_IO;
    char dict[128];
    int n = std::snprintf(dict, sizeof(dict), "{'descr': '<f8', 'fortran_order': False, 'shape': (%lld, %lld), }", (long long)Ny, (long long)Nx);
    std::string header(dict, (size_t)n);
    while ((10 + header.size() + 1) % 64 != 0) header.push_back(' ');
    header.push_back('\n');
    const unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(header.size() & 0xff), (unsigned char)(header.size() >> 8)};
    bool ok = std::fwrite(magic, 1, 10, o) == 10 && std::fwrite(header.data(), 1, header.size(), o) == header.size() &&
              std::fwrite(field, sizeof(double), (size_t)(Nx * Ny), o) == (size_t)(Nx * Ny);
    if (std::fclose(o) != 0) ok = false;
    return ok ? DEFF2D_OK : DEFF2D_ERR_IO;
}
