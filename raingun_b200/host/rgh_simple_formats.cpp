// BMP, TGA, PNM and GIF decoders of the native host library — the simple formats among those
// `image::open` accepts in the reference (image 0.12.3 default features: bmp, tga, ppm besides
// jpeg / png / gif / tiff / webp / ico / hdr; material.rs:42).  Output as the other decoders:
// RGB8 or RGBA8 (or L8 for grey PNM / TGA), rows top to bottom.
//   BMP  BITMAPINFOHEADER family, 1/4/8-bit palettes, 16-bit 5-5-5 / bit fields, 24- and 32-bit,
//        BI_RGB and BI_BITFIELDS (no RLE), bottom-up or top-down
//   TGA  types 1/2/3 and their RLE forms 9/10/11; 8-bit grey, 15/16/24/32-bit colour, 8-bit
//        colour-mapped; either vertical origin
//   PNM  P2/P3/P5/P6 (grey / RGB, ASCII or binary, maxval <= 65535 scaled to 8 bits) and P1/P4 bitmaps
//   GIF  87a / 89a, first frame only (what image 0.12 hands to `DynamicImage`), global / local colour
//        tables, interlace, transparency index -> alpha 0; RGBA8 on a canvas of the logical screen size
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rgh_internal.h"

namespace rgh {
namespace {

struct Reader {
    const uint8_t *d;
    size_t n, p = 0;
    bool ok = true;
    uint8_t u8() {
        if (p < n) return d[p++];
        ok = false;
        return 0;
    }
    uint32_t le16() { uint32_t a = u8(); return a | (u8() << 8); }
    uint32_t le32() { uint32_t a = le16(); return a | (le16() << 16); }
};

bool alloc_image(rgh_image *out, uint32_t w, uint32_t h, uint32_t ch) {
    if (w == 0 || h == 0 || (uint64_t)w * h > (1ull << 28)) return false;
    out->width = w;
    out->height = h;
    out->channels = ch;
    out->reserved = 0;
    out->pixels = (uint8_t *)rgh_alloc((size_t)w * h * ch);
    return out->pixels != nullptr;
}

inline uint8_t scale_bits(uint32_t v, uint32_t mask) {   // extract a bit field and scale it to 0..255
    if (!mask) return 255;
    int shift = 0;
    while (!((mask >> shift) & 1u)) ++shift;
    const uint32_t maxv = mask >> shift;
    return (uint8_t)(((uint64_t)((v & mask) >> shift) * 255u + maxv / 2) / maxv);
}

}  // namespace

int bmp_decode(const uint8_t *data, size_t len, rgh_image *out) {
    Reader r{data, len};
    if (len < 26 || data[0] != 'B' || data[1] != 'M') return set_error(RGH_E_FORMAT, "BMP: bad signature");
    r.p = 10;
    const uint32_t offset = r.le32(), hsize = r.le32();
    int32_t w, h;
    uint32_t bpp, comp = 0, ncolors = 0;
    uint32_t mask[4] = {0, 0, 0, 0};
    if (hsize == 12) {
        w = (int16_t)r.le16();
        h = (int16_t)r.le16();
        r.le16();
        bpp = r.le16();
    } else if (hsize >= 40) {
        w = (int32_t)r.le32();
        h = (int32_t)r.le32();
        r.le16();
        bpp = r.le16();
        comp = r.le32();
        r.le32(); r.le32(); r.le32();
        ncolors = r.le32();
        r.le32();
        if (comp == 3 || comp == 6) {   // BI_BITFIELDS / BI_ALPHABITFIELDS: masks follow (or sit in a V4/V5 header)
            mask[0] = r.le32(); mask[1] = r.le32(); mask[2] = r.le32();
            if (hsize >= 56 || comp == 6) mask[3] = r.le32();
        }
    } else {
        return set_error(RGH_E_UNSUPPORTED, "BMP: unknown header size");
    }
    if (!r.ok) return set_error(RGH_E_FORMAT, "BMP: truncated header");
    if (comp != 0 && comp != 3 && comp != 6) return set_error(RGH_E_UNSUPPORTED, "BMP: RLE / embedded compression");
    const bool top_down = h < 0;
    const uint32_t W = (uint32_t)(w < 0 ? 0 : w), H = (uint32_t)(h < 0 ? -(int64_t)h : h);
    if (bpp != 1 && bpp != 4 && bpp != 8 && bpp != 16 && bpp != 24 && bpp != 32) return set_error(RGH_E_FORMAT, "BMP: bad bit count");
    if (bpp == 16 && comp == 0) { mask[0] = 0x7C00; mask[1] = 0x03E0; mask[2] = 0x001F; }
    if (bpp == 32 && comp == 0) { mask[0] = 0x00FF0000; mask[1] = 0x0000FF00; mask[2] = 0x000000FF; }
    std::vector<uint8_t> pal;
    if (bpp <= 8) {
        const uint32_t entries = ncolors ? ncolors : (1u << bpp), esz = hsize == 12 ? 3u : 4u;
        const size_t pstart = 14 + (size_t)hsize;
        if (entries > 256 || pstart + (size_t)entries * esz > len) return set_error(RGH_E_FORMAT, "BMP: bad palette");
        pal.resize((size_t)entries * 3);
        for (uint32_t i = 0; i < entries; ++i) {
            pal[3 * i] = data[pstart + (size_t)i * esz + 2];
            pal[3 * i + 1] = data[pstart + (size_t)i * esz + 1];
            pal[3 * i + 2] = data[pstart + (size_t)i * esz];
        }
    }
    const size_t stride = (((size_t)W * bpp + 31) / 32) * 4;
    if (W == 0 || H == 0 || offset > len || stride == 0 || (len - offset) / stride < H) return set_error(RGH_E_FORMAT, "BMP: truncated pixel data");
    const bool alpha = mask[3] != 0;
    if (!alloc_image(out, W, H, alpha ? 4u : 3u)) return set_error(RGH_E_FORMAT, "BMP: bad dimensions");
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t *row = data + offset + stride * (size_t)(top_down ? y : H - 1 - y);
        uint8_t *o = out->pixels + (size_t)y * W * out->channels;
        for (uint32_t x = 0; x < W; ++x, o += out->channels) {
            if (bpp <= 8) {
                const uint32_t idx = bpp == 8 ? row[x] : (row[(x * bpp) >> 3] >> (8 - bpp - ((x * bpp) & 7))) & ((1u << bpp) - 1);
                const size_t k = (size_t)idx * 3 + 2 < pal.size() ? (size_t)idx * 3 : 0;
                o[0] = pal.empty() ? 0 : pal[k]; o[1] = pal.empty() ? 0 : pal[k + 1]; o[2] = pal.empty() ? 0 : pal[k + 2];
            } else if (bpp == 24) {
                o[0] = row[3 * x + 2]; o[1] = row[3 * x + 1]; o[2] = row[3 * x];
            } else {
                const uint32_t v = bpp == 16 ? (uint32_t)(row[2 * x] | (row[2 * x + 1] << 8))
                                             : (uint32_t)row[4 * x] | ((uint32_t)row[4 * x + 1] << 8) | ((uint32_t)row[4 * x + 2] << 16) | ((uint32_t)row[4 * x + 3] << 24);
                o[0] = scale_bits(v, mask[0]); o[1] = scale_bits(v, mask[1]); o[2] = scale_bits(v, mask[2]);
                if (alpha) o[3] = scale_bits(v, mask[3]);
            }
        }
    }
    return RGH_OK;
}

int tga_decode(const uint8_t *data, size_t len, rgh_image *out) {
    Reader r{data, len};
    if (len < 18) return set_error(RGH_E_FORMAT, "TGA: truncated header");
    const uint32_t idlen = r.u8(), cmtype = r.u8(), type = r.u8();
    const uint32_t cm_first = r.le16(), cm_len = r.le16(), cm_bits = r.u8();
    r.le16(); r.le16();
    const uint32_t W = r.le16(), H = r.le16(), bpp = r.u8(), desc = r.u8();
    const bool rle = type >= 9;
    const uint32_t base = rle ? type - 8 : type;
    if (base < 1 || base > 3) return set_error(RGH_E_UNSUPPORTED, "TGA: unsupported image type");
    if ((base == 1 && (cmtype != 1 || bpp != 8)) || (base == 3 && bpp != 8) ||
        (base == 2 && bpp != 15 && bpp != 16 && bpp != 24 && bpp != 32))
        return set_error(RGH_E_UNSUPPORTED, "TGA: unsupported pixel depth");
    r.p = 18 + (size_t)idlen;
    auto texel = [](const uint8_t *p, uint32_t bits, uint8_t o[4]) {   // little-endian B,G,R(,A) or 5-5-5
        if (bits == 24 || bits == 32) { o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; o[3] = bits == 32 ? p[3] : 255; }
        else { const uint32_t v = p[0] | (p[1] << 8); o[0] = (uint8_t)(((v >> 10) & 31) * 255 / 31); o[1] = (uint8_t)(((v >> 5) & 31) * 255 / 31); o[2] = (uint8_t)((v & 31) * 255 / 31); o[3] = 255; }
    };
    std::vector<uint8_t> cmap;
    if (cmtype == 1) {
        const uint32_t eb = (cm_bits + 7) / 8;
        if (cm_bits != 15 && cm_bits != 16 && cm_bits != 24 && cm_bits != 32) return set_error(RGH_E_UNSUPPORTED, "TGA: colour map depth");
        if (r.p + (size_t)cm_len * eb > len) return set_error(RGH_E_FORMAT, "TGA: truncated colour map");
        cmap.resize((size_t)(cm_first + cm_len) * 4, 0);
        for (uint32_t i = 0; i < cm_len; ++i) texel(data + r.p + (size_t)i * eb, cm_bits, &cmap[(size_t)(cm_first + i) * 4]);
        r.p += (size_t)cm_len * eb;
    }
    const uint32_t pb = (bpp + 7) / 8;
    const bool has_alpha = (base == 2 && bpp == 32) || (base == 1 && cm_bits == 32);
    const uint32_t ch = base == 3 ? 1u : (has_alpha ? 4u : 3u);
    if (!alloc_image(out, W, H, ch)) return set_error(RGH_E_FORMAT, "TGA: bad dimensions");
    const bool top_origin = (desc & 0x20) != 0, right_origin = (desc & 0x10) != 0;
    const uint64_t total = (uint64_t)W * H;
    uint64_t i = 0;
    uint8_t px[4] = {0, 0, 0, 255}, raw[4] = {0, 0, 0, 0};
    auto fetch = [&]() -> bool {
        if (r.p + pb > len) return false;
        std::memcpy(raw, data + r.p, pb);
        r.p += pb;
        if (base == 3) { px[0] = raw[0]; }
        else if (base == 1) { const size_t k = (size_t)raw[0] * 4; if (k + 3 < cmap.size()) std::memcpy(px, &cmap[k], 4); else { px[0] = px[1] = px[2] = 0; px[3] = 255; } }
        else texel(raw, bpp, px);
        return true;
    };
    auto put = [&](uint64_t idx) {
        uint32_t x = (uint32_t)(idx % W), y = (uint32_t)(idx / W);
        if (!top_origin) y = H - 1 - y;
        if (right_origin) x = W - 1 - x;
        std::memcpy(out->pixels + ((size_t)y * W + x) * ch, px, ch);
    };
    bool ok = true;
    while (i < total && ok) {
        if (!rle) { ok = fetch(); if (ok) put(i++); continue; }
        if (r.p >= len) { ok = false; break; }
        const uint32_t hd = data[r.p++], cnt = (hd & 127u) + 1u;
        if (hd & 128u) { ok = fetch(); for (uint32_t k = 0; ok && k < cnt && i < total; ++k) put(i++); }
        else for (uint32_t k = 0; k < cnt && i < total && ok; ++k) { ok = fetch(); if (ok) put(i++); }
    }
    if (!ok) { rgh_free(out->pixels); out->pixels = nullptr; return set_error(RGH_E_FORMAT, "TGA: truncated pixel data"); }
    return RGH_OK;
}

int pnm_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (len < 3 || data[0] != 'P' || data[1] < '1' || data[1] > '6') return set_error(RGH_E_FORMAT, "PNM: bad signature");
    const int kind = data[1] - '0';
    size_t p = 2;
    auto skip = [&]() {
        for (;;) {
            while (p < len && (data[p] == ' ' || data[p] == '\t' || data[p] == '\r' || data[p] == '\n')) ++p;
            if (p < len && data[p] == '#') { while (p < len && data[p] != '\n') ++p; continue; }
            break;
        }
    };
    auto number = [&](uint32_t &v) -> bool {
        skip();
        if (p >= len || data[p] < '0' || data[p] > '9') return false;
        uint64_t x = 0;
        while (p < len && data[p] >= '0' && data[p] <= '9') { x = x * 10 + (data[p++] - '0'); if (x > 0xFFFFFFFFull) return false; }
        v = (uint32_t)x;
        return true;
    };
    uint32_t W = 0, H = 0, maxv = 1;
    if (!number(W) || !number(H)) return set_error(RGH_E_FORMAT, "PNM: bad header");
    const bool bitmap = kind == 1 || kind == 4, ascii = kind <= 3, rgb = kind == 3 || kind == 6;
    if (!bitmap && (!number(maxv) || maxv == 0 || maxv > 65535)) return set_error(RGH_E_FORMAT, "PNM: bad maxval");
    if (!ascii) { if (p >= len) return set_error(RGH_E_FORMAT, "PNM: truncated"); ++p; }   // exactly one white-space byte before binary data
    const uint32_t ch = rgb ? 3u : 1u;
    if (!alloc_image(out, W, H, ch)) return set_error(RGH_E_FORMAT, "PNM: bad dimensions");
    const size_t count = (size_t)W * H * ch;
    bool ok = true;
    auto scale = [&](uint32_t v) { return (uint8_t)(v >= maxv ? 255 : ((uint64_t)v * 255u + maxv / 2) / maxv); };
    if (ascii) {
        for (size_t i = 0; i < count && ok; ++i) {
            uint32_t v = 0;
            if (bitmap) { skip(); ok = p < len && (data[p] == '0' || data[p] == '1'); if (ok) v = data[p++] == '1'; out->pixels[i] = v ? 0 : 255; }
            else { ok = number(v); out->pixels[i] = scale(v); }
        }
    } else if (bitmap) {
        const size_t stride = (W + 7) / 8;
        ok = p + stride * H <= len;
        for (uint32_t y = 0; ok && y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) out->pixels[(size_t)y * W + x] = (data[p + stride * y + (x >> 3)] >> (7 - (x & 7))) & 1 ? 0 : 255;
    } else {
        const size_t bps = maxv > 255 ? 2 : 1;
        ok = p + count * bps <= len;
        for (size_t i = 0; ok && i < count; ++i)
            out->pixels[i] = bps == 1 ? scale(data[p + i]) : scale(((uint32_t)data[p + 2 * i] << 8) | data[p + 2 * i + 1]);
    }
    if (!ok) { rgh_free(out->pixels); out->pixels = nullptr; return set_error(RGH_E_FORMAT, "PNM: truncated or malformed pixel data"); }
    return RGH_OK;
}

int gif_decode(const uint8_t *data, size_t len, rgh_image *out) {
    if (len < 13 || (std::memcmp(data, "GIF87a", 6) != 0 && std::memcmp(data, "GIF89a", 6) != 0))
        return set_error(RGH_E_FORMAT, "GIF: bad signature");
    Reader r{data, len};
    r.p = 6;
    const uint32_t SW = r.le16(), SH = r.le16(), flags = r.u8();
    r.u8(); r.u8();   // background colour index, pixel aspect ratio
    const uint8_t *gct = nullptr;
    uint32_t gct_n = 0;
    if (flags & 0x80) {
        gct_n = 2u << (flags & 7);
        if (r.p + (size_t)gct_n * 3 > len) return set_error(RGH_E_FORMAT, "GIF: truncated colour table");
        gct = data + r.p;
        r.p += (size_t)gct_n * 3;
    }
    int transparent = -1;
    for (;;) {
        const uint32_t tag = r.u8();
        if (!r.ok || tag == 0x3B) return set_error(RGH_E_FORMAT, "GIF: no image data");
        if (tag == 0x21) {   // extension: only the graphic control extension matters (transparency)
            const uint32_t label = r.u8();
            for (;;) {
                const uint32_t n = r.u8();
                if (!r.ok) return set_error(RGH_E_FORMAT, "GIF: truncated extension");
                if (n == 0) break;
                if (r.p + n > len) return set_error(RGH_E_FORMAT, "GIF: truncated extension");
                if (label == 0xF9 && n >= 4 && (data[r.p] & 1)) transparent = data[r.p + 3];
                r.p += n;
            }
            continue;
        }
        if (tag != 0x2C) return set_error(RGH_E_FORMAT, "GIF: unknown block");
        break;
    }
    const uint32_t fx = r.le16(), fy = r.le16(), fw = r.le16(), fh = r.le16(), iflags = r.u8();
    const uint8_t *ct = gct;
    uint32_t ct_n = gct_n;
    if (iflags & 0x80) {
        ct_n = 2u << (iflags & 7);
        if (r.p + (size_t)ct_n * 3 > len) return set_error(RGH_E_FORMAT, "GIF: truncated colour table");
        ct = data + r.p;
        r.p += (size_t)ct_n * 3;
    }
    const uint32_t min_code = r.u8();
    if (!r.ok || min_code < 1 || min_code > 11 || fw == 0 || fh == 0) return set_error(RGH_E_FORMAT, "GIF: bad image descriptor");
    const uint32_t W = SW ? SW : fw, H = SH ? SH : fh;
    if (!alloc_image(out, W, H, 4)) return set_error(RGH_E_FORMAT, "GIF: bad dimensions");
    std::memset(out->pixels, 0, (size_t)W * H * 4);   // transparent canvas
    // LZW over the concatenated sub-blocks
    std::vector<uint16_t> prefix(4096);
    std::vector<uint8_t> suffix(4096), stack(4097);
    const uint32_t clear = 1u << min_code, eoi = clear + 1;
    uint32_t avail = clear + 2, size = min_code + 1, mask = (1u << size) - 1, old = 0xFFFFFFFFu, first = 0;
    for (uint32_t i = 0; i < clear; ++i) { prefix[i] = 0; suffix[i] = (uint8_t)i; }
    uint64_t bits = 0;
    uint32_t nbits = 0, block = 0;
    const bool interlace = (iflags & 0x40) != 0;
    uint64_t produced = 0;
    const uint64_t total = (uint64_t)fw * fh;
    auto put = [&](uint8_t idx) {
        if (produced >= total) return;
        const uint32_t x = (uint32_t)(produced % fw);
        uint32_t y = (uint32_t)(produced / fw);
        ++produced;
        if (interlace) {   // rows arrive in four passes: 0,8,16.. / 4,12.. / 2,6.. / 1,3..
            const uint32_t n1 = (fh + 7) / 8, n2 = (fh + 3) / 8, n3 = (fh + 1) / 4;
            if (y < n1) y = y * 8;
            else if (y < n1 + n2) y = (y - n1) * 8 + 4;
            else if (y < n1 + n2 + n3) y = (y - n1 - n2) * 4 + 2;
            else y = (y - n1 - n2 - n3) * 2 + 1;
        }
        const uint32_t X = fx + x, Y = fy + y;
        if (X >= W || Y >= H || (int)idx == transparent) return;
        uint8_t *o = out->pixels + ((size_t)Y * W + X) * 4;
        if (ct && idx < ct_n) { o[0] = ct[3 * idx]; o[1] = ct[3 * idx + 1]; o[2] = ct[3 * idx + 2]; }
        o[3] = 255;
    };
    bool done = false;
    while (!done) {
        if (block == 0) {
            block = r.u8();
            if (!r.ok || block == 0) break;
        }
        bits |= (uint64_t)r.u8() << nbits;
        if (!r.ok) break;
        nbits += 8;
        --block;
        while (nbits >= size) {
            uint32_t code = (uint32_t)bits & mask;
            bits >>= size;
            nbits -= size;
            if (code == clear) { avail = clear + 2; size = min_code + 1; mask = (1u << size) - 1; old = 0xFFFFFFFFu; continue; }
            if (code == eoi) { done = true; break; }
            if (old == 0xFFFFFFFFu) {
                if (code >= clear) { done = true; break; }   // corrupt: the first code must be a literal
                put((uint8_t)code);
                old = first = code;
                continue;
            }
            const uint32_t in_code = code;
            uint32_t sp = 0;
            if (code >= avail) {
                if (code > avail) { done = true; break; }     // corrupt stream
                stack[sp++] = (uint8_t)first;
                code = old;
            }
            while (code >= clear && sp < 4096) { stack[sp++] = suffix[code]; code = prefix[code]; }
            first = suffix[code];
            stack[sp++] = (uint8_t)first;
            if (avail < 4096) {
                prefix[avail] = (uint16_t)old;
                suffix[avail] = (uint8_t)first;
                ++avail;
                if ((avail & mask) == 0 && avail < 4096) { ++size; mask = (1u << size) - 1; }
            }
            old = in_code;
            while (sp) put(stack[--sp]);
            if (produced >= total) { done = true; break; }
        }
    }
    return RGH_OK;   // a short stream leaves the remaining pixels transparent, as a streaming decoder would
}

}  // namespace rgh
