// multi.cpp -- one large image over several GPUs of a box from a single host process: the C++
// counterpart of effectivediffusivityfvm_b200/slab.py.  One host thread per device drives that
// device's context; the domain is split into row slabs (slab.cu: halo exchange and flux
// all-reduce over NCCL inside deff2d_domain_solve), and every rank walks the reference's stage
// sequence (SingleSim / BatchSim body / SingleSim3Phase, Deff2D.cuh:1635-1841, 1867-2049,
// 1316-1633) in lock step -- the all-reduced Deff makes every rank take the same decisions.
#include "deff2d_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <thread>
#include <vector>

using deff2d::StageSpec;

DEFF2D_EXPORT int deff2d_solve_image_slabs(deff2d_ctx *const *ctxs, int nctx, const uint8_t *gray, int W, int H,
                                           const deff2d_params *p, deff2d_result *res, double *field)
{
    if (!ctxs || nctx < 1 || !gray || !p || !res || W < 1 || H < 1) return DEFF2D_ERR_ARG;
    for (int k = 0; k < nctx; k++) if (!ctxs[k]) return DEFF2D_ERR_ARG;
    if (p->mode < 0 || p->mode > 2 || p->amp_x < 1 || p->amp_y < 1) return DEFF2D_ERR_ARG;
    if (nctx == 1) return deff2d_solve_image(ctxs[0], gray, W, H, p, res, field);
    std::memset(res, 0, sizeof(*res));
    const int nphase = (p->mode == DEFF2D_MODE_3PH) ? 3 : 2;
    const int64_t Nx = (int64_t)W * p->amp_x, Ny = (int64_t)H * p->amp_y;
    // halo: 32 amplified rows, in whole source rows; every slab must be at least as thick
    const int halo_src = (32 + p->amp_y - 1) / p->amp_y;
    const int halo = halo_src * p->amp_y;
    int n = nctx;
    while (n > 1 && (H / n) < halo_src) n--;
    if (n == 1) return deff2d_solve_image(ctxs[0], gray, W, H, p, res, field);

    // ---- fractions on the host (source pixels only); FloodFill runs on every rank's device over the whole domain
    //      inside deff2d_domain_load_slab_global (replicated: the flood needs global connectivity) ----------
    res->n_cells = Nx * Ny;
    {
        int64_t below150 = 0, cnt[3] = {0, 0, 0};
        for (size_t k = 0; k < (size_t)W * H; k++) {
            const uint8_t v = gray[k];
            below150 += v < 150;
            if (nphase == 3) cnt[v > 200 ? 1 : (v < 50 ? 2 : 0)]++;
        }
        if (nphase == 2) res->porosity = deff2d::accumulate_fraction(below150, (int64_t)W * H);       // cuh:397-405
        else {                                                                                        // cuh:411-448 (quirk Q21)
            const double Dfin[3] = {p->Df, p->Ds, p->Dg};
            int64_t ns = 0, nl = 0;
            for (int ph = 0; ph < 3; ph++) {
                if (Dfin[ph] == p->Ds) ns += cnt[ph];
                else if (Dfin[ph] == p->Df) nl += cnt[ph];
            }
            const int64_t amp2 = (int64_t)p->amp_x * p->amp_y;
            res->SVF = deff2d::accumulate_fraction(ns * amp2, Nx * Ny);
            res->LVF = deff2d::accumulate_fraction(nl * amp2, Nx * Ny);
        }
    }
    StageSpec spec[DEFF2D_MAX_STAGES];
    const int nst = deff2d::stage_list(p, spec, DEFF2D_MAX_STAGES);
    if (nst < 0) return DEFF2D_ERR_ARG;
    const std::vector<StageSpec> stages(spec, spec + nst);
    uint8_t id[DEFF2D_NCCL_ID_BYTES];
    int rc = deff2d_nccl_unique_id(id);
    if (rc) return rc;

    struct RankOut { int rc = 0; int pathflag = 0; std::vector<int64_t> iters; std::vector<double> deff, ms; double conv = 0; };
    std::vector<RankOut> out((size_t)n);
    // A rank whose slab failed to load must not leave the others waiting in a halo exchange: all ranks
    // meet once after loading and give up together if any of them failed.
    std::mutex mtx;
    std::condition_variable cv;
    int arrived = 0;
    std::atomic<int> failed(0);
    auto meet = [&]() {
        std::unique_lock<std::mutex> lk(mtx);
        if (++arrived == n) cv.notify_all();
        else cv.wait(lk, [&] { return arrived == n; });
    };
    // A failure on one rank in the middle of a stage (set_D, a launch, NCCL) would leave the others blocked in a halo
    // exchange or the flux all-reduce: the failing rank aborts every communicator of the group, which makes the peers'
    // pending NCCL work return, and all ranks leave with an error.
    auto give_up = [&]() {
        if (!failed.exchange(1))
            for (int k = 0; k < n; k++) deff2d_slab_abort(ctxs[k]);
    };
    std::vector<int> graphs_before((size_t)n, 1);
    auto work = [&](int r) {
        RankOut &o = out[(size_t)r];
        deff2d_ctx *c = ctxs[r];
        // NCCL reports an internal error when send/recv of ranks that are threads of one process are
        // captured into CUDA graphs (measured, NCCL 2.28): enqueue the passes directly in this mode
        graphs_before[(size_t)r] = deff2d_get_graphs(c);
        deff2d_set_graphs(c, 0);
        if ((o.rc = deff2d_nccl_init(c, id, r, n))) { failed.store(1); meet(); return; }
        const int s0 = (int)((int64_t)H * r / n), s1 = (int)((int64_t)H * (r + 1) / n);
        const int sa = (r > 0) ? halo_src : 0, sb = (r < n - 1) ? halo_src : 0;
        deff2d_params q = *p;
        q.verbose = 0;
        o.rc = deff2d_domain_load_slab_global(c, gray, W, H, nphase, &q, (int64_t)s0 * p->amp_y, (int64_t)(s1 - s0) * p->amp_y, halo);
        if (!o.rc) o.rc = deff2d_domain_info(c, nullptr, nullptr, &o.pathflag, nullptr, nullptr, nullptr);
        if (o.rc) failed.store(1);
        meet();
        if (failed.load()) { if (!o.rc) o.rc = DEFF2D_ERR_STATE; return; }
        for (const StageSpec &st : stages) {
            if ((o.rc = deff2d_domain_set_D(c, st.Ds, st.Df, st.Dg))) { give_up(); return; }
            int64_t it = 0;
            double d = 0, cv = 0;
            const auto t0 = std::chrono::steady_clock::now();
            if ((o.rc = deff2d_domain_solve(c, st.tol, st.max_iter, &it, &d, &cv, nullptr, 0, nullptr))) { give_up(); return; }
            if (failed.load()) { o.rc = DEFF2D_ERR_STATE; return; }
            o.ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            o.iters.push_back(it);
            o.deff.push_back(d);
            if (!st.precond) o.conv = cv;
        }
        if (field) {
            // own rows of the iterate -> their place in the global map
            const int64_t rows_local = (int64_t)(s1 - s0 + sa + sb) * p->amp_y;
            std::vector<double> loc((size_t)rows_local * Nx);
            if ((o.rc = deff2d_domain_get_field(c, loc.data()))) return;
            std::memcpy(field + (size_t)s0 * p->amp_y * Nx, loc.data() + (size_t)sa * p->amp_y * Nx,
                        (size_t)(s1 - s0) * p->amp_y * Nx * sizeof(double));
        }
    };
    std::vector<std::thread> pool;
    for (int r = 1; r < n; r++) pool.emplace_back(work, r);
    work(0);
    for (auto &t : pool) t.join();
    for (int r = 0; r < n; r++) deff2d_set_graphs(ctxs[r], graphs_before[(size_t)r]);
    for (int r = 0; r < n; r++) if (out[(size_t)r].rc) return out[(size_t)r].rc;
    res->pathflag = out[0].pathflag;

    const RankOut &o0 = out[0];
    res->nstages = (int)stages.size();
    for (size_t k = 0; k < stages.size() && k < DEFF2D_MAX_STAGES; k++) {
        res->iters[k] = o0.iters[k];
        res->stage_deff_raw[k] = o0.deff[k];
        res->stage_D[k] = stages[k].stageD;
        res->total_iters += o0.iters[k];
        double ms = 0;
        for (int r = 0; r < n; r++) ms = std::max(ms, out[(size_t)r].ms[k]);
        res->total_ms += ms;
        if (!stages[k].precond) { res->solve_ms += ms; res->deff_raw = o0.deff[k]; }      // cuh:1309-1311 vs cuh:1144-1159
        if (p->verbose == 1) std::printf("Iterations taken = %d\n", (int)o0.iters[k]);     // cuh:1797, 1544, 1595
    }
    res->conv = o0.conv;
    const double norm = (p->mode == DEFF2D_MODE_3PH) ? p->Df : (stages.empty() ? p->Df : stages.back().Df);
    res->last_df = norm;
    if (!stages.empty()) res->deff = res->deff_raw / norm;                                  // cuh:1802, 1601, 2017
    if (p->verbose == 1 && !stages.empty()) std::cout << "DCF = " << norm << ", Deff " << res->deff << std::endl;
    return DEFF2D_OK;
}
