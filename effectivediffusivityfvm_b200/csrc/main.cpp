// main.cpp -- `deff2d`: drop-in for the reference executable (Deff2D.cu:3-54): no arguments,
// reads ./input.txt, writes the same CSV / CMAP files.  An optional argument names another
// input file.
#include <cstdio>

#include "../../include/deff2d.h"

int main(int argc, char **argv)
{
    std::fflush(stdout);
    deff2d_ctx *ctx = nullptr;
    int rc = deff2d_create(&ctx, 0);
    if (rc) {
        std::fprintf(stderr, "deff2d: %s\n", deff2d_last_error(nullptr));
        return 1;
    }
    rc = deff2d_run_input_file(ctx, argc > 1 ? argv[1] : "input.txt");
    if (rc) std::fprintf(stderr, "deff2d: error %d: %s\n", rc, deff2d_last_error(ctx));
    deff2d_destroy(ctx);
    return rc ? 1 : 0;
}
