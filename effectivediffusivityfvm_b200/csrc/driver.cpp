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
#include <condition_variable>
#include <functional>
#include <mutex>
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
    // BatchSim (cuh:1843-2054) / BatchSim3Phase (cuh:2056-2419): images "%05d.jpg" from 0.  The reference decodes,
    // solves and keeps every row until the end (cuh:2051: a crash loses the whole batch, doc 3.6).  Here host threads
    // decode ahead while the GPU sweeps, equally sized images are solved PACKED (many per launch, batch.cu) through
    // the streaming call, every result row is appended as soon as it and all rows before it exist (same file, same
    // order), and the 3-phase CMAP_%05d.csv of an image (cuh:2395-2398; quirk Q17: 2-phase batches write none) is
    // written when that image finishes.  `Devices: N` (ignored by the reference parser, cuh:261-311) splits the
    // images over N GPUs with no communication.
    const int N = std::max(in.num_images, 0);
    const bool cmap = (in.nphase == 3 && in.print_cmap == 1);
    struct Decoded {
        std::vector<uint8_t> px;
        int W = 0, H = 0, ch = 0, rc = 0;
        int state = 0;                       // 0 pending, 1 decoded
    };
    std::vector<Decoded> dec((size_t)N);
    std::mutex mtx;
    std::condition_variable cv;
    std::atomic<int> next_decode(0);
    auto decode_worker = [&]() {
        for (;;) {
            const int k = next_decode.fetch_add(1);                 // in file order: slot k belongs to "%05d.jpg" % k
            if (k >= N) break;
            char name[100];
            std::snprintf(name, sizeof(name), "%05d.jpg", k);       // cuh:1876
            uint8_t *gray = nullptr;
            Decoded d;
            d.rc = deff2d_load_image(name, &gray, &d.W, &d.H, &d.ch);
            if (!d.rc && d.ch == 1) d.px.assign(gray, gray + (size_t)d.W * d.H);
            deff2d_free(gray);
            d.state = 1;
            {
                std::lock_guard<std::mutex> lk(mtx);
                dec[(size_t)k] = std::move(d);
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> decoders;
    {
        const int nt = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
        for (int t = 0; t < nt && t < N; t++) decoders.emplace_back(decode_worker);
    }
    struct JoinAll { std::vector<std::thread> &v; ~JoinAll() { for (auto &t : v) if (t.joinable()) t.join(); } } join_guard{decoders};
    auto wait_decoded = [&](int k) {
        std::unique_lock<std::mutex> lk(mtx);
        cv.wait(lk, [&] { return dec[(size_t)k].state != 0; });
    };
    auto check_image = [&](int k) -> int {                          // the reference's own input checks, per image
        const Decoded &d = dec[(size_t)k];
        if (d.rc) { std::printf("Error: could not read image %05d.jpg\n", k); return d.rc; }   // reference: NULL deref (Q24)
        if (d.ch != 1) {                                            // cuh:1886-1889
            std::printf("Error: please enter a grascale image with 1 channel.\n Current number of channels = %d\n", d.ch);
            return DEFF2D_ERR_ARG;
        }
        return DEFF2D_OK;
    };

    // rows leave in image order as soon as the prefix is complete
    std::vector<deff2d_result> results((size_t)N);
    std::vector<char> have((size_t)N, 0);
    int next_row = 0, io_rc = 0;
    std::mutex out_mtx;
    if ((rc = deff2d_append_csv_batch_row(&in, -1, nullptr))) return rc;       // header (cuh:209, cuh:224)
    auto finish_image = [&](int k, const deff2d_result *r, const double *field, int W, int H) -> int {
        std::lock_guard<std::mutex> lk(out_mtx);
        results[(size_t)k] = *r;
        have[(size_t)k] = 1;
        if (cmap && field) {
            char cm[100];
            std::snprintf(cm, sizeof(cm), "CMAP_%05d.csv", k);      // cuh:2396
            const int e = deff2d_write_cmap(cm, field, (int64_t)W * in.p.amp_x, (int64_t)H * in.p.amp_y);
            if (e && !io_rc) io_rc = e;
        }
        while (next_row < N && have[(size_t)next_row]) {
            const int e = deff2d_append_csv_batch_row(&in, next_row, &results[(size_t)next_row]);
            if (e && !io_rc) io_rc = e;
            next_row++;
        }
        return io_rc;
    };

    int first_unsolved = 0;
    if (N >= 2 && in.p.verbose != 1) {
        wait_decoded(0);
        if ((rc = check_image(0))) return rc;
        const int W0 = dec[0].W, H0 = dec[0].H;
        // one context per device; device 0 is the caller's.  Each device streams a contiguous block of images.
        std::vector<deff2d_ctx *> ctxs{ctx};
        for (int d = 1; d < in.devices && d < N; d++) {
            deff2d_ctx *cx = nullptr;
            if (deff2d_create(&cx, d) != DEFF2D_OK) break;          // fewer GPUs than asked for: use what is there
            ctxs.push_back(cx);
        }
        const int nd = (int)ctxs.size();
        struct Block {
            int k0 = 0, k1 = 0, W0 = 0, H0 = 0, end = -1, rc = 0;
            std::vector<Decoded> *dec = nullptr;
            std::mutex *mtx = nullptr;
            std::condition_variable *cv = nullptr;
            std::function<int(int, const deff2d_result *, const double *)> finish;
            std::function<int(int)> check;
        };
        std::vector<Block> blocks((size_t)nd);
        auto fetch = [](void *user, int k, uint8_t *dst, int wait) -> int {
            Block *b = static_cast<Block *>(user);
            const int g = b->k0 + k;
            if (g >= b->k1) return 2;
            Decoded *d = &(*b->dec)[(size_t)g];
            {
                std::unique_lock<std::mutex> lk(*b->mtx);
                if (d->state == 0) {
                    if (!wait) return 1;                            // not decoded yet: the GPU goes on with what it has
                    b->cv->wait(lk, [&] { return d->state != 0; });
                }
            }
            if (d->rc || d->ch != 1) { b->rc = b->check(g); b->end = g; return b->rc ? b->rc : DEFF2D_ERR_IO; }
            if (d->W != b->W0 || d->H != b->H0) { b->end = g; return 2; }   // another size: the packed run ends here
            std::memcpy(dst, d->px.data(), d->px.size());
            std::vector<uint8_t>().swap(d->px);                     // the pixels now live in the library's staging buffer
            return 0;
        };
        auto done = [](void *user, int k, const deff2d_result *r, const double *field) -> int {
            Block *b = static_cast<Block *>(user);
            return b->finish(b->k0 + k, r, field);
        };
        std::vector<int> rcs((size_t)nd, 0), solved((size_t)nd, 0);
        auto work = [&](int d) {
            Block &b = blocks[(size_t)d];
            b.k0 = (int)((int64_t)N * d / nd); b.k1 = (int)((int64_t)N * (d + 1) / nd);
            b.W0 = W0; b.H0 = H0; b.dec = &dec; b.mtx = &mtx; b.cv = &cv;
            b.finish = [&](int k, const deff2d_result *r, const double *f) { return finish_image(k, r, f, W0, H0); };
            b.check = check_image;
            deff2d_params p = in.p;
            rcs[(size_t)d] = deff2d_solve_batch_stream(ctxs[(size_t)d], b.k1 - b.k0, W0, H0, &p, fetch, done, &b, cmap ? 1 : 0, &solved[(size_t)d]);
        };
        // parameters the packed mode does not cover (e.g. the non-parity solver) fall through to the per-image loop
        const bool packed_ok = deff2d_batch_supported(&in.p, W0, H0) == 1;
        if (packed_ok) {
            std::vector<std::thread> pool;
            for (int d = 1; d < nd; d++) pool.emplace_back(work, d);
            work(0);
            for (auto &t : pool) t.join();
            for (int d = 1; d < nd; d++) deff2d_destroy(ctxs[(size_t)d]);
            for (int d = 0; d < nd; d++) if (rcs[(size_t)d]) return rcs[(size_t)d];
            if (io_rc) return io_rc;
            // images a block could not take (another size) and everything behind them: one at a time below
            first_unsolved = N;
            for (int d = 0; d < nd; d++)
                if (blocks[(size_t)d].k0 + solved[(size_t)d] < blocks[(size_t)d].k1) { first_unsolved = std::min(first_unsolved, blocks[(size_t)d].k0 + solved[(size_t)d]); }
        } else {
            for (int d = 1; d < nd; d++) deff2d_destroy(ctxs[(size_t)d]);
        }
    }
    // one image at a time: keeps the reference's per-image stdout order (Verbose: 1) and handles
    // images of different sizes; every row is appended as soon as its image (and all before it) is solved
    for (int k = first_unsolved; k < N; k++) {
        if (have[(size_t)k]) continue;
        wait_decoded(k);
        if ((rc = check_image(k))) return rc;
        const int W = dec[(size_t)k].W, H = dec[(size_t)k].H;
        std::vector<double> field;
        if (cmap) field.resize((size_t)W * in.p.amp_x * (size_t)H * in.p.amp_y);
        deff2d_params p = in.p;
        deff2d_result r;
        rc = deff2d_solve_batch(ctx, dec[(size_t)k].px.data(), 1, W, H, &p, &r, cmap ? field.data() : nullptr);
        if (rc) return rc;
        if ((rc = finish_image(k, &r, cmap ? field.data() : nullptr, W, H))) return rc;
        std::vector<uint8_t>().swap(dec[(size_t)k].px);
    }
    return io_rc;
}
