// C ABI glue of libraingun_host.so: errors, allocation, image_open / png_save, CLI option mapping.
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <string>
#include <vector>

#include "rgh_internal.h"

namespace rgh {

static thread_local std::string g_error;

int set_error(int code, const std::string &message) {
    g_error = message;
    return code;
}

bool read_file(const char *path, std::vector<uint8_t> &out) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    out.clear();
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
    const bool ok = !std::ferror(f);
    std::fclose(f);
    return ok;
}

namespace {

std::string lower_ext(const char *path) {
    const char *slash = std::strrchr(path, '/');
    const char *dot = std::strrchr(slash ? slash : path, '.');
    std::string e = dot ? dot + 1 : "";
    for (auto &c : e) c = (char)std::tolower((unsigned char)c);
    return e;
}

// Rust's `str::parse::<u32>()`: optional '+', decimal digits, no overflow.
bool parse_u32(const char *s, uint32_t *out) {
    if (*s == '+') ++s;
    if (!*s) return false;
    uint64_t v = 0;
    for (; *s; ++s) {
        if (*s < '0' || *s > '9') return false;
        v = v * 10 + (uint64_t)(*s - '0');
        if (v > 0xFFFFFFFFull) return false;
    }
    *out = (uint32_t)v;
    return true;
}

}  // namespace
}  // namespace rgh

extern "C" {

const char *rgh_last_error(void) { return rgh::g_error.c_str(); }
void *rgh_alloc(size_t n) { return std::malloc(n ? n : 1); }
void rgh_free(void *p) { std::free(p); }

// Nothing may unwind across the C ABI: allocation failures on absurd (corrupt) dimensions included.
#define RGH_GUARD(expr)                                                                    \
    try {                                                                                  \
        return (expr);                                                                     \
    } catch (const std::bad_alloc &) {                                                     \
        return rgh::set_error(RGH_E_FORMAT, "out of memory (corrupt dimensions?)");        \
    } catch (const std::exception &e) {                                                    \
        return rgh::set_error(RGH_E_FORMAT, std::string("internal error: ") + e.what());   \
    }

int rgh_jpeg_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (!data || !out) return rgh::set_error(RGH_E_INVALID, "rgh_jpeg_decode: null argument");
    std::memset(out, 0, sizeof *out);
    RGH_GUARD(rgh::jpeg_decode(data, len, out))
}

int rgh_png_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (!data || !out) return rgh::set_error(RGH_E_INVALID, "rgh_png_decode: null argument");
    std::memset(out, 0, sizeof *out);
    RGH_GUARD(rgh::png_decode(data, len, out))
}

int rgh_png_encode(const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels, uint8_t **out,
                   size_t *out_len) {
    if (!out || !out_len) return rgh::set_error(RGH_E_INVALID, "rgh_png_encode: null argument");
    std::vector<uint8_t> v;
    const int rc = rgh::png_encode(pixels, width, height, channels, v);
    if (rc != RGH_OK) return rc;
    *out = (uint8_t *)rgh_alloc(v.size());
    if (!*out) return rgh::set_error(RGH_E_IO, "out of memory");
    std::memcpy(*out, v.data(), v.size());
    *out_len = v.size();
    return RGH_OK;
}

int rgh_bmp_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (!data || !out) return rgh::set_error(RGH_E_INVALID, "rgh_bmp_decode: null argument");
    std::memset(out, 0, sizeof *out);
    RGH_GUARD(rgh::bmp_decode(data, len, out))
}

int rgh_tga_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (!data || !out) return rgh::set_error(RGH_E_INVALID, "rgh_tga_decode: null argument");
    std::memset(out, 0, sizeof *out);
    RGH_GUARD(rgh::tga_decode(data, len, out))
}

int rgh_pnm_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (!data || !out) return rgh::set_error(RGH_E_INVALID, "rgh_pnm_decode: null argument");
    std::memset(out, 0, sizeof *out);
    RGH_GUARD(rgh::pnm_decode(data, len, out))
}

int rgh_gif_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (!data || !out) return rgh::set_error(RGH_E_INVALID, "rgh_gif_decode: null argument");
    std::memset(out, 0, sizeof *out);
    RGH_GUARD(rgh::gif_decode(data, len, out))
}

/* image 0.12 `open`: the decoder is chosen by the (case-insensitive) extension. */
int rgh_image_open(const char *path, rgh_image *out) {
    if (!path || !out) return rgh::set_error(RGH_E_INVALID, "rgh_image_open: null argument");
    std::memset(out, 0, sizeof *out);
    const std::string ext = rgh::lower_ext(path);
    int (*decode)(const uint8_t *, size_t, rgh_image *) = nullptr;
    if (ext == "jpg" || ext == "jpeg") decode = rgh::jpeg_decode;
    else if (ext == "png") decode = rgh::png_decode;
    else if (ext == "bmp") decode = rgh::bmp_decode;
    else if (ext == "tga") decode = rgh::tga_decode;
    else if (ext == "gif") decode = rgh::gif_decode;
    else if (ext == "pbm" || ext == "pgm" || ext == "ppm" || ext == "pnm") decode = rgh::pnm_decode;
    if (!decode)
        return rgh::set_error(RGH_E_UNSUPPORTED, "Unsupported image format image/" + ext + " (jpg, jpeg, png, bmp, tga, gif, pbm, pgm, ppm are built)");
    std::vector<uint8_t> data;
    if (!rgh::read_file(path, data)) return rgh::set_error(RGH_E_IO, std::string(std::strerror(errno)) + " (" + path + ")");
    RGH_GUARD(decode(data.data(), data.size(), out))
}

int rgh_png_save(const char *path, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels) {
    if (!path) return rgh::set_error(RGH_E_INVALID, "rgh_png_save: null path");
    std::vector<uint8_t> v;
    const int rc = rgh::png_encode(pixels, width, height, channels, v);
    if (rc != RGH_OK) return rc;
    FILE *f = std::fopen(path, "wb");
    if (!f) return rgh::set_error(RGH_E_IO, std::string("Could not encode image: ") + std::strerror(errno) + " (" + path + ")");
    const bool ok = std::fwrite(v.data(), 1, v.size(), f) == v.size();
    if (std::fclose(f) != 0 || !ok) return rgh::set_error(RGH_E_IO, std::string("Could not encode image: write failed (") + path + ")");
    return RGH_OK;
}

/* src/main.rs:21-96.  clap semantics that matter: `--name value`, `--name=value`, `-w value`,
 * `-wvalue`; `-h` is the height (not help); an argument that `overrides_with` another removes the
 * other from the matches if it came EARLIER on the command line (later arguments win). */
int rgh_cli_parse(int argc, const char *const *argv, rgh_cli_options *out) {
    if (!out || (argc > 0 && !argv)) return rgh::set_error(RGH_E_INVALID, "rgh_cli_parse: null argument");
    bool has_4k = false, has_hd = false, has_draft = false, has_preview = false;
    const char *width = nullptr, *height = nullptr, *output = nullptr, *input = nullptr;
    auto usage = [](const std::string &m) { return rgh::set_error(RGH_E_USAGE, m); };
    bool only_positional = false;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        const char *value = nullptr;
        std::string name;
        if (!only_positional && a[0] == '-' && a[1] == '-' && a[2] == 0) {
            only_positional = true;
            continue;
        }
        if (!only_positional && a[0] == '-' && a[1] == '-') {
            const char *eq = std::strchr(a, '=');
            name = eq ? std::string(a + 2, (size_t)(eq - a - 2)) : std::string(a + 2);
            value = eq ? eq + 1 : nullptr;
        } else if (!only_positional && a[0] == '-' && a[1] != 0) {
            name = a[1] == 'w' ? "width" : a[1] == 'h' ? "height" : a[1] == 'o' ? "output" : std::string("-") + a[1];
            if (a[2] != 0) value = a[2] == '=' ? a + 3 : a + 2;
        } else {
            if (input) return usage(std::string("Found argument '") + a + "' which wasn't expected, or isn't valid in this context");
            input = a;
            continue;
        }
        const bool takes_value = name == "width" || name == "height" || name == "output";
        if (takes_value) {
            if (!value) {
                if (i + 1 >= argc) return usage("The argument '--" + name + "' requires a value but none was supplied");
                value = argv[++i];
            }
            if (name == "width") width = value;
            else if (name == "height") height = value;
            else output = value;
        } else {
            if (value) return usage("The argument '--" + name + "' does not take a value");
            if (name == "4k") {
                has_4k = true;
                has_hd = false;
            } else if (name == "hd") {
                has_hd = true;
                has_4k = false;
            } else if (name == "draft") {
                has_draft = true;
                has_4k = has_hd = false;
                width = height = nullptr;
            } else if (name == "preview") {
                has_preview = true;
            } else {
                return usage("Found argument '" + std::string(a) + "' which wasn't expected, or isn't valid in this context");
            }
        }
    }
    if (!input) return usage("The following required arguments were not provided: <FILE>");
    std::memset(out, 0, sizeof *out);
    out->width = 800;  // RenderOptions::default, src/render.rs:38-46
    out->height = 600;
    out->max_depth_limit = -1;
    out->preview = has_preview ? 1 : 0;
    if (has_draft) out->max_depth_limit = 4;  // src/main.rs:72-81
    else if (has_hd) {
        out->width = 1920;
        out->height = 1080;
    } else if (has_4k) {
        out->width = 3840;
        out->height = 2160;
    }
    if (width && !rgh::parse_u32(width, &out->width)) return usage("Could not parse width");
    if (height && !rgh::parse_u32(height, &out->height)) return usage("Could not parse height");
    if (std::strlen(input) >= sizeof out->input) return usage("input path too long");
    std::strcpy(out->input, input);
    std::string outp;
    if (output) {
        outp = output;
    } else {  // Path::set_extension("png"), src/main.rs:104-113
        const std::string in = input;
        size_t end = in.size();
        while (end > 1 && in[end - 1] == '/') --end;  // trailing separators are not part of the file name
        const size_t slash = in.find_last_of('/', end ? end - 1 : 0);
        const size_t start = slash == std::string::npos ? 0 : slash + 1;
        const std::string file = in.substr(start, end - start);
        if (file.empty() || file == ".." || file == "." || file == "/")
            return rgh::set_error(RGH_E_USAGE, std::string("Could not guess output filename from ") + input);
        const size_t dot = file.find_last_of('.');
        const std::string stem = (dot == std::string::npos || dot == 0) ? file : file.substr(0, dot);
        outp = in.substr(0, start) + stem + ".png";
    }
    if (outp.size() >= sizeof out->output) return usage("output path too long");
    std::strcpy(out->output, outp.c_str());
    return RGH_OK;
}

}  // extern "C"
