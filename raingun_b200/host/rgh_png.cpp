// PNG decode / encode of the native host library (include/raingun_host.h).
//
// Decode replaces `image::open` for .png textures (material.rs:34-47; image 0.12.3 -> png 0.6.2
// with its EXPAND transformation: palettes become RGB, tRNS becomes alpha, 1/2/4-bit grey is
// scaled to 8 bits).  Texture::color only reads r, g, b (material.rs:63-68, color.rs:26-30), so
// grey images are presented as RGB8 and grey+alpha as RGBA8 — what `DynamicImage::get_pixel`
// returns for them.  Encode replaces `ImageBuffer::save` (src/render.rs:58, RGBA8).  PNG is
// lossless, so only the pixel values matter; the DEFLATE stream comes from zlib.
#include <zlib.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rgh_internal.h"

namespace rgh {
namespace {

const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

inline uint32_t be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
inline void put32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Reverses the filter of one scanline in place (`prev` = the reconstructed previous line or null).
bool unfilter(int type, uint8_t *cur, const uint8_t *prev, size_t n, size_t bpp) {
    switch (type) {
        case 0: return true;
        case 1:
            for (size_t i = bpp; i < n; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - bpp]);
            return true;
        case 2:
            if (prev)
                for (size_t i = 0; i < n; ++i) cur[i] = (uint8_t)(cur[i] + prev[i]);
            return true;
        case 3:
            for (size_t i = 0; i < n; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0;
                cur[i] = (uint8_t)(cur[i] + ((a + b) >> 1));
            }
            return true;
        case 4:
            for (size_t i = 0; i < n; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
                cur[i] = (uint8_t)(cur[i] + paeth(a, b, c));
            }
            return true;
        default: return false;
    }
}

}  // namespace

int png_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (len < 8 || std::memcmp(data, kSig, 8) != 0) return set_error(RGH_E_FORMAT, "PNG: bad signature");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_trns = false, seen_iend = false;
    size_t pos = 8;
    while (pos + 12 <= len && !seen_iend) {
        const uint32_t n = be32(data + pos);
        const uint8_t *type = data + pos + 4, *body = data + pos + 8;
        if ((size_t)n > len - pos - 12) return set_error(RGH_E_FORMAT, "PNG: truncated chunk");
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, n + 4);
        if (crc != be32(body + n)) return set_error(RGH_E_FORMAT, "PNG: chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (n != 13) return set_error(RGH_E_FORMAT, "PNG: bad IHDR");
            w = be32(body);
            h = be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || interlace > 1) return set_error(RGH_E_FORMAT, "PNG: bad IHDR methods");
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + n);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + n);
            have_trns = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + n);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_iend = true;
        }
        pos += (size_t)n + 12;
    }
    if (ctype < 0 || w == 0 || h == 0) return set_error(RGH_E_FORMAT, "PNG: missing IHDR");
    int samples;
    switch (ctype) {
        case 0: samples = 1; break;
        case 2: samples = 3; break;
        case 3: samples = 1; break;
        case 4: samples = 2; break;
        case 6: samples = 4; break;
        default: return set_error(RGH_E_FORMAT, "PNG: bad colour type");
    }
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok) return set_error(RGH_E_FORMAT, "PNG: bad bit depth");
    if (depth == 16) return set_error(RGH_E_UNSUPPORTED, "PNG: 16-bit samples (image 0.12 cannot present them as 8-bit pixels)");
    if (ctype == 3 && plte.size() < 3) return set_error(RGH_E_FORMAT, "PNG: palette image without PLTE");
    if ((uint64_t)w * h > (1ull << 31)) return set_error(RGH_E_UNSUPPORTED, "PNG: image too large");

    const size_t bits_pp = (size_t)samples * depth, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    // pass geometry: non-interlaced = one pass covering everything
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1},
                     dy[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t raw_size = 0;
    const int passes = interlace ? 7 : 1;
    for (int p = 0; p < passes; ++p) {
        const uint32_t pw = interlace ? (w + dx[p] - 1 - xs[p]) / dx[p] : w, ph = interlace ? (h + dy[p] - 1 - ys[p]) / dy[p] : h;
        if (pw && ph) raw_size += (size_t)ph * (1 + (pw * bits_pp + 7) / 8);
    }
    std::vector<uint8_t> raw(raw_size);
    uLongf got = (uLongf)raw_size;
    const int zr = uncompress(raw.data(), &got, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || got != raw_size) return set_error(RGH_E_FORMAT, "PNG: bad or short zlib stream");

    const bool alpha = ctype == 4 || ctype == 6 || have_trns;
    const uint32_t oc = alpha ? 4u : 3u;
    uint8_t *px = (uint8_t *)rgh_alloc((size_t)w * h * oc);
    if (!px) return set_error(RGH_E_FORMAT, "PNG: out of memory");
    size_t off = 0;
    const int scale = depth < 8 ? 255 / ((1 << depth) - 1) : 1;
    for (int p = 0; p < passes; ++p) {
        const uint32_t pw = interlace ? (w + dx[p] - 1 - xs[p]) / dx[p] : w, ph = interlace ? (h + dy[p] - 1 - ys[p]) / dy[p] : h;
        if (!pw || !ph) continue;
        const size_t line = (pw * bits_pp + 7) / 8;
        const uint8_t *prev = nullptr;
        for (uint32_t y = 0; y < ph; ++y) {
            uint8_t *cur = raw.data() + off + 1;
            if (!unfilter(raw[off], cur, prev, line, bpp)) {
                rgh_free(px);
                return set_error(RGH_E_FORMAT, "PNG: bad filter type");
            }
            prev = cur;
            off += line + 1;
            const uint32_t oy = interlace ? (uint32_t)ys[p] + y * dy[p] : y;
            for (uint32_t x = 0; x < pw; ++x) {
                const uint32_t ox = interlace ? (uint32_t)xs[p] + x * dx[p] : x;
                uint8_t s[4] = {0, 0, 0, 255};
                if (depth == 8) {
                    for (int k = 0; k < samples; ++k) s[k] = cur[(size_t)x * samples + k];
                } else {  // 1/2/4-bit grey or palette index, MSB first
                    const size_t bit = (size_t)x * depth;
                    s[0] = (uint8_t)((cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1));
                }
                uint8_t r, g, b, a = 255;
                if (ctype == 3) {
                    const size_t idx = s[0];
                    if (idx * 3 + 2 >= plte.size()) {
                        r = g = b = 0;
                    } else {
                        r = plte[idx * 3];
                        g = plte[idx * 3 + 1];
                        b = plte[idx * 3 + 2];
                    }
                    if (have_trns && idx < trns.size()) a = trns[idx];
                } else if (ctype == 0 || ctype == 4) {
                    if (have_trns && ctype == 0 && trns.size() >= 2 && (uint32_t)s[0] == (((uint32_t)trns[0] << 8) | trns[1])) a = 0;
                    r = g = b = (uint8_t)(s[0] * scale);
                    if (ctype == 4) a = s[1];
                } else {
                    r = s[0];
                    g = s[1];
                    b = s[2];
                    if (ctype == 6) a = s[3];
                    else if (have_trns && trns.size() >= 6 && r == trns[1] && g == trns[3] && b == trns[5] && !trns[0] && !trns[2] && !trns[4]) a = 0;
                }
                uint8_t *o = px + ((size_t)oy * w + ox) * oc;
                o[0] = r;
                o[1] = g;
                o[2] = b;
                if (alpha) o[3] = a;
            }
        }
    }
    out->width = w;
    out->height = h;
    out->channels = oc;
    out->reserved = 0;
    out->pixels = px;
    return RGH_OK;
}

int png_encode(const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels, std::vector<uint8_t> &out) {
    if (!pixels || width == 0 || height == 0 || (channels != 1 && channels != 3 && channels != 4))
        return set_error(RGH_E_INVALID, "PNG encode: need L8, RGB8 or RGBA8 pixels and a non-empty image");
    const size_t bpp = channels, line = (size_t)width * bpp;
    // per-row adaptive filter: minimum sum of absolute differences
    std::vector<uint8_t> raw((line + 1) * height), cand(line);
    const std::vector<uint8_t> zero(line, 0);
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t *cur = pixels + (size_t)y * line, *prev = y ? cur - line : zero.data();
        uint8_t *dst = raw.data() + (size_t)y * (line + 1);
        uint64_t best = ~0ull;
        for (int f = 0; f < 5; ++f) {
            uint64_t sum = 0;
            for (size_t i = 0; i < line; ++i) {
                const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
                int pred = 0;
                if (f == 1) pred = a;
                else if (f == 2) pred = b;
                else if (f == 3) pred = (a + b) >> 1;
                else if (f == 4) pred = paeth(a, b, c);
                const uint8_t v = (uint8_t)(cur[i] - pred);
                cand[i] = v;
                sum += v < 128 ? v : 256 - v;
            }
            if (sum < best) {
                best = sum;
                dst[0] = (uint8_t)f;
                std::memcpy(dst + 1, cand.data(), line);
            }
        }
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return set_error(RGH_E_FORMAT, "PNG encode: zlib failure");
    out.clear();
    out.insert(out.end(), kSig, kSig + 8);
    auto chunk = [&](const char *type, const uint8_t *body, size_t n) {
        put32(out, (uint32_t)n);
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        if (n) out.insert(out.end(), body, body + n);
        put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), out.data() + start, (uInt)(n + 4)));
    };
    uint8_t ihdr[13];
    ihdr[0] = (uint8_t)(width >> 24); ihdr[1] = (uint8_t)(width >> 16); ihdr[2] = (uint8_t)(width >> 8); ihdr[3] = (uint8_t)width;
    ihdr[4] = (uint8_t)(height >> 24); ihdr[5] = (uint8_t)(height >> 16); ihdr[6] = (uint8_t)(height >> 8); ihdr[7] = (uint8_t)height;
    ihdr[8] = 8;
    ihdr[9] = channels == 1 ? 0 : (channels == 3 ? 2 : 6);
    ihdr[10] = ihdr[11] = ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), zlen);
    chunk("IEND", nullptr, 0);
    return RGH_OK;
}

}  // namespace rgh
