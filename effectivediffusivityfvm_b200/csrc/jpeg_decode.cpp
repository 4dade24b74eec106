// jpeg_decode.cpp -- JPEG ingest (placeholder: not implemented yet; PNG / PGM content under
// any file name is decoded by image_io.cpp).
#include "deff2d_internal.h"

namespace deff2d {
int jpeg_decode_gray(const uint8_t *data, size_t len, std::vector<uint8_t> &out, int *W, int *H, int *ch,
                     std::string &err)
{
    (void)data; (void)len; (void)out; (void)W; (void)H; (void)ch;
    err = "JPEG decoding is not available in this build; supply PNG or PGM content";
    return DEFF2D_ERR_IO;
}
}  // namespace deff2d
