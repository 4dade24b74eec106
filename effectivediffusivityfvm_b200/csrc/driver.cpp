// driver.cpp -- the reference program's dispatch and four drivers (Deff2D.cu:17-50,
// cuh:1316-2419) on top of the C ABI: same input file, same CSV / CMAP files, same stdout
// lines; the solve itself is the library's.
#include "deff2d_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

// printOptions, cuh:121-175
static int print_options(const deff2d_input *in)
{
    const deff2d_params &p = in->p;
    if (in->batch == 0) {
        std::printf("--------------------------------------\n\n");
        std::printf("Current selected options:\n\n");
        std::printf("--------------------------------------\n");
        std::printf("Number of Phases = %d\n", in->nphase);
        std::printf("DC Fluid = %1.3e\n", p.Df);
        std::printf("DC Solid = %1.3e\n", p.Ds);
        std::printf("DC Gas = %1.3e\n", p.Dg);
        std::printf("Concentration Left = %.2f\n", p.CL);
        std::printf("Concentration Right = %.2f\n", p.CR);
        std::printf("Mesh Amp. X = %d\n", p.amp_x);
        std::printf("Mesh Amp. Y = %d\n", p.amp_y);
        std::printf("Maximum Iterations = %ld\n", (long)p.max_iter);
        std::printf("Convergence = %.10f\n", p.tol);
        std::printf("Name of input image: %s\n", in->input_name);
        std::printf("Name of output file: %s\n", in->output_name);
        if (in->print_cmap == 0) std::printf("Print Concentration Map = False\n");
        else std::printf("Concentration Map Name = %s\n", in->cmap_name);
        std::printf("--------------------------------------\n\n");
    } else if (in->batch == 1) {
        std::printf("--------------------------------------\n\n");
        std::printf("Running Image Batch:\n\n");
        std::printf("Number of Phases = %d\n", in->nphase);
        std::printf("DC Fluid = %1.3e\n", p.Df);
        std::printf("DC Solid = %1.3e\n", p.Ds);
        std::printf("DC Gas = %1.3e\n", p.Dg);
        std::printf("Concentration Left = %.2f\n", p.CL);
        std::printf("Concentration Right = %.2f\n", p.CR);
        std::printf("Mesh Amp. X = %d\n", p.amp_x);
        std::printf("Mesh Amp. Y = %d\n", p.amp_y);
        std::printf("Maximum Iterations = %ld\n", (long)p.max_iter);
        std::printf("Convergence = %.10f\n", p.tol);
        std::printf("Name of output file: %s\n", in->output_name);
        std::printf("Number of files to run: %d\n", in->num_images);
        if (in->print_cmap == 1) std::printf("Printing Concentration Distribution for all images.\n");
        else std::printf("No Concentration maps will be printed.\n");
        std::printf("--------------------------------------\n\n");
    } else {
        std::printf("Options entered are not valid, code will exit.\n");
        return 1;
    }
    return 0;
}

DEFF2D_EXPORT int deff2d_run_input_file(deff2d_ctx *ctx, const char *path)
{
    if (!ctx || !path) return DEFF2D_ERR_ARG;
    deff2d_input in;
    int rc = deff2d_read_input_file(path, &in);
    if (rc) return rc;
    if (in.p.verbose == 1) print_options(&in);                      // cuh:318-322
    else if (in.p.verbose != 0) std::printf("Please enter a value of 0 or 1 for 'verbose'. Default = 0.\n");
    if (in.nphase != 2 && in.nphase != 3) {                         // cu:47-50
        std::printf("Current option entered for Phases is not supported.\n Exiting now. \n");
        return DEFF2D_OK;
    }
    if (in.batch != 0 && in.batch != 1) {                           // cu:27-30
        std::cout << "Error: no valid BatchFlag option, check input file." << std::endl;
        return DEFF2D_OK;
    }
    if (in.batch == 0) {
        // SingleSim (cuh:1635-1841) / SingleSim3Phase (cuh:1316-1633)
        uint8_t *gray = nullptr;
        int W = 0, H = 0, ch = 0;
        rc = deff2d_load_image(in.input_name, &gray, &W, &H, &ch);
        if (rc) { std::printf("Error: could not read image %s\n", in.input_name); return rc; }   // reference: NULL deref (Q24)
        if (ch != 1) {                                              // cuh:1665-1668
            std::printf("Error: please enter a grascale image with 1 channel.\n Current number of channels = %d\n", ch);
            deff2d_free(gray);
            return DEFF2D_ERR_ARG;
        }
        deff2d_result res;
        std::vector<double> field;
        if (in.print_cmap == 1) field.resize((size_t)W * in.p.amp_x * (size_t)H * in.p.amp_y);
        if (in.devices > 1) {
            // one large image over several GPUs: row slabs, one host thread per device (multi.cpp)
            std::vector<deff2d_ctx *> ctxs{ctx};
            for (int d = 1; d < in.devices; d++) {
                deff2d_ctx *cx = nullptr;
                if (deff2d_create(&cx, d) != DEFF2D_OK) break;      // fewer GPUs than asked for: use what is there
                ctxs.push_back(cx);
            }
            rc = deff2d_solve_image_slabs(ctxs.data(), (int)ctxs.size(), gray, W, H, &in.p, &res,
                                          in.print_cmap == 1 ? field.data() : nullptr);
            for (size_t d = 1; d < ctxs.size(); d++) deff2d_destroy(ctxs[d]);
        } else {
            rc = deff2d_solve_image(ctx, gray, W, H, &in.p, &res, in.print_cmap == 1 ? field.data() : nullptr);
        }
        deff2d_free(gray);
        if (rc) return rc;
        if ((rc = deff2d_write_csv_single(&in, &res))) return rc;   // cuh:1821, cuh:1612
        if (in.print_cmap == 1) {                                   // cuh:1825-1827
            rc = deff2d_write_cmap(in.cmap_name, field.data(), (int64_t)W * in.p.amp_x, (int64_t)H * in.p.amp_y);
            if (!rc && in.field_npy == 1)
                rc = deff2d_write_field_npy((std::string(in.cmap_name) + ".npy").c_str(), field.data(), (int64_t)W * in.p.amp_x,
                                            (int64_t)H * in.p.amp_y);
        }
        return rc;
    }
    // BatchSim (cuh:1843-2054) / BatchSim3Phase (cuh:2056-2419): images "%05d.jpg" from 0;
    // results are written once at the end (cuh:2051); only the 3-phase batch writes
    // per-image CMAP_%05d.csv files (cuh:2395-2398, quirk Q17).
    std::vector<deff2d_result> results((size_t)std::max(in.num_images, 0));
    const bool cmap = (in.nphase == 3 && in.print_cmap == 1);
    // Decode every image first (decoding overlaps nothing here, but it decides the path): a batch of
    // equally sized images without per-image verbose output is solved PACKED -- all images resident
    // at once, many per launch (batch.cu) -- and, with the optional `Devices: N` key (ignored by the
    // reference parser, cuh:261-311), split over N GPUs with no communication.
    std::vector<uint8_t> packed;
    int W0 = 0, H0 = 0;
    bool uniform = in.num_images > 0;
    std::vector<std::vector<uint8_t>> singles((size_t)std::max(in.num_images, 0));
    std::vector<int> Ws((size_t)std::max(in.num_images, 0), 0), Hs(Ws), chs(Ws), rcs_load(Ws);
    {
        // decode on the host threads (file order is kept: slot k belongs to "%05d.jpg" % k)
        std::atomic<int> next(0);
        auto load = [&]() {
            for (;;) {
                const int k = next.fetch_add(1);
                if (k >= in.num_images) break;
                char name[100];
                std::snprintf(name, sizeof(name), "%05d.jpg", k);   // cuh:1876
                uint8_t *gray = nullptr;
                rcs_load[(size_t)k] = deff2d_load_image(name, &gray, &Ws[(size_t)k], &Hs[(size_t)k], &chs[(size_t)k]);
                if (!rcs_load[(size_t)k] && chs[(size_t)k] == 1) singles[(size_t)k].assign(gray, gray + (size_t)Ws[(size_t)k] * Hs[(size_t)k]);
                deff2d_free(gray);
            }
        };
        const int nt = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt && t < in.num_images; t++) pool.emplace_back(load);
        load();
        for (auto &t : pool) t.join();
    }
    for (int k = 0; k < in.num_images; k++) {
        if (rcs_load[(size_t)k]) { std::printf("Error: could not read image %05d.jpg\n", k); return rcs_load[(size_t)k]; }
        if (chs[(size_t)k] != 1) {
            std::printf("Error: please enter a grascale image with 1 channel.\n Current number of channels = %d\n", chs[(size_t)k]);
            return DEFF2D_ERR_ARG;
        }
        if (k == 0) { W0 = Ws[0]; H0 = Hs[0]; }
        if (Ws[(size_t)k] != W0 || Hs[(size_t)k] != H0) uniform = false;
    }
    if (uniform && in.p.verbose != 1 && in.num_images >= 2) {
        const size_t npix = (size_t)W0 * H0, ncell = npix * (size_t)in.p.amp_x * (size_t)in.p.amp_y;
        packed.resize(npix * (size_t)in.num_images);
        for (int k = 0; k < in.num_images; k++) std::memcpy(packed.data() + npix * k, singles[(size_t)k].data(), npix);
        singles.clear();
        std::vector<double> fields;
        if (cmap) fields.resize(ncell * (size_t)in.num_images);
        // one context per device; device 0 is the caller's
        std::vector<deff2d_ctx *> ctxs{ctx};
        for (int d = 1; d < in.devices && d < in.num_images; d++) {
            deff2d_ctx *cx = nullptr;
            if (deff2d_create(&cx, d) != DEFF2D_OK) break;          // fewer GPUs than asked for: use what is there
            ctxs.push_back(cx);
        }
        const int nd = (int)ctxs.size();
        std::vector<int> rcs((size_t)nd, 0);
        std::vector<std::thread> pool;
        auto work = [&](int d) {
            const int k0 = (int)((int64_t)in.num_images * d / nd), k1 = (int)((int64_t)in.num_images * (d + 1) / nd);
            deff2d_params p = in.p;
            rcs[(size_t)d] = deff2d_solve_batch(ctxs[(size_t)d], packed.data() + npix * k0, k1 - k0, W0, H0, &p, results.data() + k0,
                                                cmap ? fields.data() + ncell * k0 : nullptr);
        };
        for (int d = 1; d < nd; d++) pool.emplace_back(work, d);
        work(0);
        for (auto &t : pool) t.join();
        for (int d = 1; d < nd; d++) deff2d_destroy(ctxs[(size_t)d]);
        for (int d = 0; d < nd; d++) if (rcs[(size_t)d]) return rcs[(size_t)d];
        if (cmap)
            for (int k = 0; k < in.num_images; k++) {
                char cm[100];
                std::snprintf(cm, sizeof(cm), "CMAP_%05d.csv", k);  // cuh:2396
                if ((rc = deff2d_write_cmap(cm, fields.data() + ncell * k, (int64_t)W0 * in.p.amp_x, (int64_t)H0 * in.p.amp_y))) return rc;
            }
        return deff2d_write_csv_batch(&in, results.data(), in.num_images);
    }
    // one image at a time: keeps the reference's per-image stdout order (Verbose: 1) and handles
    // images of different sizes; every row is appended as soon as its image is solved
    if ((rc = deff2d_append_csv_batch_row(&in, -1, nullptr))) return rc;
    for (int k = 0; k < in.num_images; k++) {
        const int W = Ws[(size_t)k], H = Hs[(size_t)k];
        std::vector<double> field;
        if (cmap) field.resize((size_t)W * in.p.amp_x * (size_t)H * in.p.amp_y);
        deff2d_params p = in.p;
        rc = deff2d_solve_batch(ctx, singles[(size_t)k].data(), 1, W, H, &p, &results[(size_t)k], cmap ? field.data() : nullptr);
        if (rc) return rc;
        if ((rc = deff2d_append_csv_batch_row(&in, k, &results[(size_t)k]))) return rc;
        if (cmap) {
            char cm[100];
            std::snprintf(cm, sizeof(cm), "CMAP_%05d.csv", k);      // cuh:2396
            if ((rc = deff2d_write_cmap(cm, field.data(), (int64_t)W * in.p.amp_x, (int64_t)H * in.p.amp_y))) return rc;
        }
    }
    return DEFF2D_OK;
}
