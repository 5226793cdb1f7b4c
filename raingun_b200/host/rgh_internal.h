// Shared declarations of libraingun_host.so's translation units (not installed).
#ifndef RGH_INTERNAL_H
#define RGH_INTERNAL_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/raingun_host.h"

namespace rgh {

// Records the calling thread's last error and returns `code`.
int set_error(int code, const std::string &message);

int jpeg_decode(const uint8_t *data, size_t len, rgh_image *out);
int png_decode(const uint8_t *data, size_t len, rgh_image *out);
int bmp_decode(const uint8_t *data, size_t len, rgh_image *out);
int tga_decode(const uint8_t *data, size_t len, rgh_image *out);
int pnm_decode(const uint8_t *data, size_t len, rgh_image *out);
int gif_decode(const uint8_t *data, size_t len, rgh_image *out);
int png_encode(const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels,
               std::vector<uint8_t> &out);

bool read_file(const char *path, std::vector<uint8_t> &out);

}  // namespace rgh
#endif
