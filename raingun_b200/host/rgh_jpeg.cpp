// JPEG decoder of the native host library (include/raingun_host.h: rgh_jpeg_decode).
//
// Replaces `image::open` for .jpg textures (raingun-lib/src/material.rs:34-47).  The reference
// pins image 0.12.3 -> jpeg-decoder 0.1.11 (Cargo.lock), which is NOT vendored under
// /root/reference.  Entropy decoding is fully determined by ITU-T T.81; the three places where
// decoders legitimately differ are restated from jpeg-decoder 0.1.11's published algorithm:
//   * IDCT: the stb_image-style 13-bit fixed-point integer IDCT with de-quantisation fused
//     (column pass keeps 2 extra bits: +512 >> 10; row pass +65536+(128<<17) >> 17, clamp);
//     an all-zero-AC column short-cuts to dc << 2;
//   * chroma upsampling: H1V1 copy; H2V1 triangle filter (3*near+far+2)>>2; H2V2 two-row
//     triangle filter (3*near+far per row, then (3*t1+t0+8)>>4), edges replicated;
//   * colour: YCbCr -> RGB in f32 (1.40200, 0.34414, 0.71414, 1.77200), +0.5, truncate, clamp.
// Whether this restatement equals the reference's decoder is checked the only way the
// reference allows: the textured pixels of examples/test1.png and test3.png
// (tests/test_oracle_golden.py).
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rgh_internal.h"

namespace rgh {
namespace {

const uint8_t kZigZag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman {
    bool present = false;
    // canonical decoding: for each length, the smallest code, and index of its first symbol
    int32_t maxcode[18];
    int32_t valptr[17];
    int32_t mincode[17];
    uint8_t values[256];
    uint8_t look_len[512];  // 9-bit fast path: 0 = longer than 9 bits
    uint8_t look_sym[512];
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int width = 0, height = 0;       // sample dimensions: ceil(W*h/hmax), ceil(H*v/vmax)
    int blocks_w = 0, blocks_h = 0;  // allocated blocks (padded to whole MCUs)
    int dc_table = 0, ac_table = 0;
    std::vector<int16_t> coeffs;  // blocks_w*blocks_h*64, natural (de-zigzagged) order
    std::vector<uint8_t> plane;   // blocks_w*8 x blocks_h*8 samples after the IDCT
    int dc_pred = 0;
};

struct BitReader {
    const uint8_t *p, *end;
    uint64_t bits = 0;
    int count = 0;
    int marker = 0;  // a marker met while filling (stops the fill; zeros are fed from then on)
    bool eof = false;  // the data ended inside an 0xFF sequence: `marker` is a stand-in, not a byte of the file
    void reset(const uint8_t *q) {
        p = q;
        bits = 0;
        count = 0;
        marker = 0;
        eof = false;
    }
    void fill() {
        while (count <= 56) {
            uint32_t b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    while (p < end && *p == 0xFF) ++p;  // fill bytes
                    if (p >= end) eof = true;
                    const uint32_t m = p < end ? *p++ : 0xD9;
                    if (m != 0) {
                        marker = (int)m;
                        b = 0;
                    }
                }
            }
            bits |= (uint64_t)b << (56 - count);
            count += 8;
        }
    }
    inline uint32_t peek(int n) {
        if (count < n) fill();
        return (uint32_t)(bits >> (64 - n));
    }
    inline void skip(int n) {
        bits <<= n;
        count -= n;
    }
    inline uint32_t get(int n) {
        if (n == 0) return 0;
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
    inline int receive_extend(int n) {
        if (n == 0) return 0;
        const int v = (int)get(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
};

struct Decoder {
    const uint8_t *data;
    size_t len;
    size_t pos = 0;
    int width = 0, height = 0;
    bool progressive = false;
    int hmax = 1, vmax = 1;
    int mcus_w = 0, mcus_h = 0;
    std::vector<Component> comps;
    uint16_t quant[4][64];  // natural order
    bool quant_present[4] = {false, false, false, false};
    Huffman dc_tab[4], ac_tab[4];
    int restart_interval = 0;
    bool is_jfif = false;
    int adobe_transform = -1;  // -1: no Adobe marker; 0 unknown (RGB), 1 YCbCr, 2 YCCK
    bool have_frame = false;
    std::string err;
    int err_code = RGH_E_FORMAT;

    bool fail(const std::string &m, int code = RGH_E_FORMAT) {
        if (err.empty()) {
            err = m;
            err_code = code;
        }
        return false;
    }
    bool need(size_t n) { return pos + n <= len ? true : fail("unexpected end of JPEG data"); }
    uint32_t u8() { return data[pos++]; }
    uint32_t u16() {
        const uint32_t v = ((uint32_t)data[pos] << 8) | data[pos + 1];
        pos += 2;
        return v;
    }

    bool build_huffman(Huffman &h, const uint8_t counts[16], const uint8_t *vals, int nvals) {
        h = Huffman();
        std::memcpy(h.values, vals, (size_t)nvals);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            h.valptr[l] = k;
            h.mincode[l] = code;
            const int n = counts[l - 1];
            if (n) {
                if (code + n > (1 << l)) return fail("bad Huffman table (over-subscribed)");
                for (int i = 0; i < n; ++i) {
                    if (l <= 9) {
                        const int first = (code + i) << (9 - l);
                        for (int f = 0; f < (1 << (9 - l)); ++f) {
                            h.look_len[first + f] = (uint8_t)l;
                            h.look_sym[first + f] = vals[k + i];
                        }
                    }
                }
                k += n;
                code += n;
                h.maxcode[l] = code - 1;
            } else {
                h.maxcode[l] = -1;
            }
            code <<= 1;
        }
        h.maxcode[17] = 0x7FFFFFFF;
        h.present = true;
        return true;
    }

    inline int decode_symbol(BitReader &br, const Huffman &h) {
        const uint32_t look = br.peek(16);
        const uint32_t idx = look >> 7;
        if (h.look_len[idx]) {
            br.skip(h.look_len[idx]);
            return h.look_sym[idx];
        }
        for (int l = 10; l <= 16; ++l) {
            const int code = (int)(look >> (16 - l));
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) {
                br.skip(l);
                return h.values[h.valptr[l] + code - h.mincode[l]];
            }
        }
        fail("bad Huffman code");
        return -1;
    }

    // ---------------------------------------------------------------- marker segments
    bool parse_dqt(size_t end) {
        while (pos < end) {
            const uint32_t pq_tq = u8();
            const int pq = pq_tq >> 4, tq = pq_tq & 15;
            if (tq > 3 || pq > 1) return fail("bad DQT");
            if (pos + (pq ? 128u : 64u) > end) return fail("short DQT");
            for (int i = 0; i < 64; ++i) quant[tq][kZigZag[i]] = (uint16_t)(pq ? u16() : u8());
            quant_present[tq] = true;
        }
        return true;
    }
    bool parse_dht(size_t end) {
        while (pos < end) {
            if (pos + 17 > end) return fail("short DHT");
            const uint32_t tc_th = u8();
            const int tc = tc_th >> 4, th = tc_th & 15;
            if (tc > 1 || th > 3) return fail("bad DHT");
            uint8_t counts[16];
            int n = 0;
            for (int i = 0; i < 16; ++i) {
                counts[i] = (uint8_t)u8();
                n += counts[i];
            }
            if (n > 256 || pos + (size_t)n > end) return fail("bad DHT");
            if (!build_huffman(tc ? ac_tab[th] : dc_tab[th], counts, data + pos, n)) return false;
            pos += (size_t)n;
        }
        return true;
    }
    bool parse_sof(int marker, size_t end) {
        if (have_frame) return fail("multiple frames");
        if (marker != 0xC0 && marker != 0xC1 && marker != 0xC2)
            return fail("unsupported JPEG coding process (only baseline, extended sequential and progressive "
                        "Huffman)", RGH_E_UNSUPPORTED);
        progressive = marker == 0xC2;
        if (pos + 6 > end) return fail("short SOF");
        const int precision = (int)u8();
        height = (int)u16();
        width = (int)u16();
        const int nc = (int)u8();
        if (precision != 8) return fail("unsupported sample precision", RGH_E_UNSUPPORTED);
        if (width == 0 || height == 0) return fail("zero image dimension", RGH_E_UNSUPPORTED);
        if ((uint64_t)width * (uint64_t)height > (1ull << 28)) return fail("image larger than 2^28 pixels", RGH_E_UNSUPPORTED);
        if (nc != 1 && nc != 3) return fail("unsupported component count", RGH_E_UNSUPPORTED);
        if (pos + (size_t)nc * 3 > end) return fail("short SOF");
        comps.resize((size_t)nc);
        for (auto &c : comps) {
            c.id = (int)u8();
            const uint32_t hv = u8();
            c.h = hv >> 4;
            c.v = hv & 15;
            c.tq = (int)u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("bad SOF component");
            hmax = c.h > hmax ? c.h : hmax;
            vmax = c.v > vmax ? c.v : vmax;
        }
        if (nc == 1) {  // a single component is never interleaved: its sampling factors are moot
            comps[0].h = comps[0].v = 1;
            hmax = vmax = 1;
        }
        mcus_w = (width + 8 * hmax - 1) / (8 * hmax);
        mcus_h = (height + 8 * vmax - 1) / (8 * vmax);
        for (auto &c : comps) {
            c.width = (width * c.h + hmax - 1) / hmax;
            c.height = (height * c.v + vmax - 1) / vmax;
            c.blocks_w = mcus_w * c.h;
            c.blocks_h = mcus_h * c.v;
            c.coeffs.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
        }
        have_frame = true;
        return true;
    }

    // ---------------------------------------------------------------- entropy-coded segments
    struct Scan {
        int ncomp = 0;
        int ci[4];
        int ss = 0, se = 63, ah = 0, al = 0;
    };

    bool decode_block_baseline(BitReader &br, Component &c, int16_t *blk) {
        const Huffman &dc = dc_tab[c.dc_table], &ac = ac_tab[c.ac_table];
        const int t = decode_symbol(br, dc);
        if (t < 0) return false;
        if (t > 11) return fail("bad DC magnitude");
        c.dc_pred += br.receive_extend(t);
        blk[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            const int rs = decode_symbol(br, ac);
            if (rs < 0) return false;
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) return fail("AC index out of range");
            blk[kZigZag[k]] = (int16_t)br.receive_extend(s);
            ++k;
        }
        return true;
    }

    bool decode_block_dc_first(BitReader &br, Component &c, int16_t *blk, int al) {
        const int t = decode_symbol(br, dc_tab[c.dc_table]);
        if (t < 0) return false;
        if (t > 11) return fail("bad DC magnitude");
        c.dc_pred += br.receive_extend(t);
        blk[0] = (int16_t)(c.dc_pred * (1 << al));
        return true;
    }
    static void decode_block_dc_refine(BitReader &br, int16_t *blk, int al) {
        if (br.get(1)) blk[0] = (int16_t)(blk[0] | (1 << al));
    }
    bool decode_block_ac_first(BitReader &br, Component &c, int16_t *blk, const Scan &s, uint32_t &eobrun) {
        if (eobrun > 0) {
            --eobrun;
            return true;
        }
        const Huffman &ac = ac_tab[c.ac_table];
        for (int k = s.ss; k <= s.se;) {
            const int rs = decode_symbol(br, ac);
            if (rs < 0) return false;
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r == 15) {
                    k += 16;
                    continue;
                }
                eobrun = (1u << r) - 1;
                if (r) eobrun += br.get(r);
                break;
            }
            k += r;
            if (k > 63) return fail("AC index out of range");
            blk[kZigZag[k]] = (int16_t)(br.receive_extend(sz) * (1 << s.al));
            ++k;
        }
        return true;
    }
    bool decode_block_ac_refine(BitReader &br, Component &c, int16_t *blk, const Scan &s, uint32_t &eobrun) {
        const int p1 = 1 << s.al, m1 = -1 * (1 << s.al);
        const Huffman &ac = ac_tab[c.ac_table];
        int k = s.ss;
        if (eobrun == 0) {
            for (; k <= s.se;) {
                const int rs = decode_symbol(br, ac);
                if (rs < 0) return false;
                int r = rs >> 4;
                const int sz = rs & 15;
                int value = 0;
                if (sz == 0) {
                    if (r != 15) {
                        eobrun = (1u << r);
                        if (r) eobrun += br.get(r);
                        break;  // the rest of the band is handled as part of the EOB run
                    }
                } else {
                    if (sz != 1) return fail("bad AC refinement size");
                    value = br.get(1) ? p1 : m1;
                }
                // skip r zero-history coefficients, refining the non-zero ones passed on the way
                for (; k <= s.se; ++k) {
                    int16_t &co = blk[kZigZag[k]];
                    if (co != 0) {
                        if (br.get(1) && (co & p1) == 0) co = (int16_t)(co >= 0 ? co + p1 : co + m1);
                    } else {
                        if (r == 0) {
                            if (value) co = (int16_t)value;
                            ++k;
                            break;
                        }
                        --r;
                    }
                }
            }
        }
        if (eobrun > 0) {
            for (; k <= s.se; ++k) {
                int16_t &co = blk[kZigZag[k]];
                if (co != 0 && br.get(1) && (co & p1) == 0) co = (int16_t)(co >= 0 ? co + p1 : co + m1);
            }
            --eobrun;
        }
        return true;
    }

    bool decode_scan(const Scan &s) {
        BitReader br;
        br.end = data + len;
        br.reset(data + pos);
        for (int i = 0; i < s.ncomp; ++i) comps[(size_t)s.ci[i]].dc_pred = 0;
        uint32_t eobrun = 0;
        const bool interleaved = s.ncomp > 1;
        int units_w, units_h;  // MCUs of this scan
        if (interleaved) {
            units_w = mcus_w;
            units_h = mcus_h;
        } else {
            const Component &c = comps[(size_t)s.ci[0]];
            units_w = (c.width + 7) / 8;
            units_h = (c.height + 7) / 8;
        }
        int expected_rst = 0;
        long mcu_index = 0;
        const long total = (long)units_w * units_h;
        for (int my = 0; my < units_h; ++my) {
            for (int mx = 0; mx < units_w; ++mx, ++mcu_index) {
                if (restart_interval && mcu_index && mcu_index % restart_interval == 0) {
                    // the restart marker: byte-align, expect RSTn
                    br.fill();
                    if (br.marker >= 0xD0 && br.marker <= 0xD7) {
                        if (br.marker != 0xD0 + expected_rst) return fail("restart markers out of order");
                        expected_rst = (expected_rst + 1) & 7;
                        br.reset(br.p);
                    } else {
                        return fail("missing restart marker");
                    }
                    for (int i = 0; i < s.ncomp; ++i) comps[(size_t)s.ci[i]].dc_pred = 0;
                    eobrun = 0;
                }
                for (int i = 0; i < s.ncomp; ++i) {
                    Component &c = comps[(size_t)s.ci[i]];
                    const int bh = interleaved ? c.h : 1, bv = interleaved ? c.v : 1;
                    for (int by = 0; by < bv; ++by) {
                        for (int bx = 0; bx < bh; ++bx) {
                            const int X = mx * bh + bx, Y = my * bv + by;
                            int16_t *blk = &c.coeffs[((size_t)Y * c.blocks_w + X) * 64];
                            bool ok = true;
                            if (!progressive) {
                                ok = decode_block_baseline(br, c, blk);
                            } else if (s.ss == 0) {
                                if (s.ah == 0) ok = decode_block_dc_first(br, c, blk, s.al);
                                else decode_block_dc_refine(br, blk, s.al);
                            } else {
                                ok = s.ah == 0 ? decode_block_ac_first(br, c, blk, s, eobrun)
                                               : decode_block_ac_refine(br, c, blk, s, eobrun);
                            }
                            if (!ok) return false;
                        }
                    }
                }
            }
        }
        (void)total;
        // continue the marker parse after the entropy-coded data
        br.fill();
        if (br.eof) {
            pos = len;   // truncated file: never step back to an earlier 0xFF (that would re-read this scan for ever)
        } else if (br.marker) {
            pos = (size_t)(br.p - data) - 2;   // the byte before the marker code is an 0xFF (the prefix or a fill byte)
        } else {
            // no marker seen yet: scan forward for one
            size_t q = (size_t)(br.p - data);
            while (q + 1 < len && !(data[q] == 0xFF && data[q + 1] != 0 && data[q + 1] != 0xFF)) ++q;
            pos = q;
        }
        return true;
    }

    bool parse_sos(size_t end) {
        if (!have_frame) return fail("SOS before SOF");
        if (pos + 1 > end) return fail("short SOS");
        Scan s;
        s.ncomp = (int)u8();
        if (s.ncomp < 1 || s.ncomp > (int)comps.size() || pos + (size_t)s.ncomp * 2 + 3 > end) return fail("bad SOS");
        for (int i = 0; i < s.ncomp; ++i) {
            const int id = (int)u8();
            const uint32_t t = u8();
            int ci = -1;
            for (size_t j = 0; j < comps.size(); ++j)
                if (comps[j].id == id) ci = (int)j;
            if (ci < 0) return fail("SOS names an unknown component");
            s.ci[i] = ci;
            comps[(size_t)ci].dc_table = t >> 4;
            comps[(size_t)ci].ac_table = t & 15;
            if ((t >> 4) > 3 || (t & 15) > 3) return fail("bad table selector");
        }
        s.ss = (int)u8();
        s.se = (int)u8();
        const uint32_t a = u8();
        s.ah = a >> 4;
        s.al = a & 15;
        if (progressive) {
            if (s.ss > s.se || s.se > 63 || (s.ss == 0 && s.se != 0) || (s.ss != 0 && s.ncomp != 1) || s.al > 13)
                return fail("bad progressive scan parameters");
        } else {
            s.ss = 0;
            s.se = 63;
            s.ah = s.al = 0;
        }
        for (int i = 0; i < s.ncomp; ++i) {
            const Component &c = comps[(size_t)s.ci[i]];
            const bool need_dc = s.ss == 0 && s.ah == 0, need_ac = s.se > 0;
            if (need_dc && !dc_tab[c.dc_table].present) return fail("scan uses an undefined DC Huffman table");
            if (need_ac && !ac_tab[c.ac_table].present) return fail("scan uses an undefined AC Huffman table");
        }
        pos = end;
        return decode_scan(s);
    }

    // ---------------------------------------------------------------- IDCT (stb-style, fused de-quantisation)
    static inline int f2f(float x) { return (int)(x * 4096.0f + 0.5f); }
    static inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

    static void idct_block(const int16_t *co, const uint16_t *q, uint8_t *out, int stride) {
        static const int k0_5411961 = f2f(0.5411961f), kn1_847759065 = f2f(-1.847759065f),
                         k0_765366865 = f2f(0.765366865f), k1_175875602 = f2f(1.175875602f),
                         k0_298631336 = f2f(0.298631336f), k2_053119869 = f2f(2.053119869f),
                         k3_072711026 = f2f(3.072711026f), k1_501321110 = f2f(1.501321110f),
                         kn0_899976223 = f2f(-0.899976223f), kn2_562915447 = f2f(-2.562915447f),
                         kn1_961570560 = f2f(-1.961570560f), kn0_390180644 = f2f(-0.390180644f);
        int tmp[64];
#define RGH_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                      \
    int p2 = (s2), p3 = (s6);                                            \
    int p1 = (p2 + p3) * k0_5411961;                                     \
    int t2 = p1 + p3 * kn1_847759065;                                    \
    int t3 = p1 + p2 * k0_765366865;                                     \
    p2 = (s0);                                                           \
    p3 = (s4);                                                           \
    int t0 = (p2 + p3) * 4096;                                           \
    int t1 = (p2 - p3) * 4096;                                           \
    int x0 = t0 + t3, x3 = t0 - t3, x1 = t1 + t2, x2 = t1 - t2;          \
    t0 = (s7);                                                           \
    t1 = (s5);                                                           \
    t2 = (s3);                                                           \
    t3 = (s1);                                                           \
    p3 = t0 + t2;                                                        \
    int p4 = t1 + t3;                                                    \
    p1 = t0 + t3;                                                        \
    p2 = t1 + t2;                                                        \
    int p5 = (p3 + p4) * k1_175875602;                                   \
    t0 = t0 * k0_298631336;                                              \
    t1 = t1 * k2_053119869;                                              \
    t2 = t2 * k3_072711026;                                              \
    t3 = t3 * k1_501321110;                                              \
    p1 = p5 + p1 * kn0_899976223;                                        \
    p2 = p5 + p2 * kn2_562915447;                                        \
    p3 = p3 * kn1_961570560;                                             \
    p4 = p4 * kn0_390180644;                                             \
    t3 += p1 + p4;                                                       \
    t2 += p2 + p3;                                                       \
    t1 += p2 + p4;                                                       \
    t0 += p1 + p3;
        for (int i = 0; i < 8; ++i) {  // columns
            if (co[i + 8] == 0 && co[i + 16] == 0 && co[i + 24] == 0 && co[i + 32] == 0 && co[i + 40] == 0 &&
                co[i + 48] == 0 && co[i + 56] == 0) {
                const int dc = ((int)co[i] * (int)q[i]) * 4;
                for (int r = 0; r < 8; ++r) tmp[i + 8 * r] = dc;
            } else {
                RGH_IDCT_1D((int)co[i] * q[i], (int)co[i + 8] * q[i + 8], (int)co[i + 16] * q[i + 16],
                            (int)co[i + 24] * q[i + 24], (int)co[i + 32] * q[i + 32], (int)co[i + 40] * q[i + 40],
                            (int)co[i + 48] * q[i + 48], (int)co[i + 56] * q[i + 56])
                x0 += 512;
                x1 += 512;
                x2 += 512;
                x3 += 512;
                tmp[i] = (x0 + t3) >> 10;
                tmp[i + 56] = (x0 - t3) >> 10;
                tmp[i + 8] = (x1 + t2) >> 10;
                tmp[i + 48] = (x1 - t2) >> 10;
                tmp[i + 16] = (x2 + t1) >> 10;
                tmp[i + 40] = (x2 - t1) >> 10;
                tmp[i + 24] = (x3 + t0) >> 10;
                tmp[i + 32] = (x3 - t0) >> 10;
            }
        }
        for (int i = 0; i < 8; ++i) {  // rows
            const int *v = tmp + 8 * i;
            uint8_t *o = out + (size_t)i * stride;
            RGH_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
            x0 += 65536 + (128 << 17);
            x1 += 65536 + (128 << 17);
            x2 += 65536 + (128 << 17);
            x3 += 65536 + (128 << 17);
            o[0] = clamp8((x0 + t3) >> 17);
            o[7] = clamp8((x0 - t3) >> 17);
            o[1] = clamp8((x1 + t2) >> 17);
            o[6] = clamp8((x1 - t2) >> 17);
            o[2] = clamp8((x2 + t1) >> 17);
            o[5] = clamp8((x2 - t1) >> 17);
            o[3] = clamp8((x3 + t0) >> 17);
            o[4] = clamp8((x3 - t0) >> 17);
        }
#undef RGH_IDCT_1D
    }

    bool reconstruct_planes() {
        for (auto &c : comps) {
            if (!quant_present[c.tq]) return fail("frame uses an undefined quantisation table");
            const int stride = c.blocks_w * 8;
            c.plane.assign((size_t)stride * c.blocks_h * 8, 0);
            for (int by = 0; by < c.blocks_h; ++by)
                for (int bx = 0; bx < c.blocks_w; ++bx)
                    idct_block(&c.coeffs[((size_t)by * c.blocks_w + bx) * 64], quant[c.tq],
                               &c.plane[(size_t)by * 8 * stride + (size_t)bx * 8], stride);
            c.coeffs.clear();
            c.coeffs.shrink_to_fit();
        }
        return true;
    }

    // ---------------------------------------------------------------- upsampling (one output row)
    bool upsample_row(const Component &c, int row, std::vector<uint8_t> &out) {
        const int stride = c.blocks_w * 8;
        const bool h1 = c.h == hmax, v1 = c.v == vmax, h2 = c.h * 2 == hmax, v2 = c.v * 2 == vmax;
        const int iw = c.width, ih = c.height;
        if (h1 && v1) {
            std::memcpy(out.data(), &c.plane[(size_t)row * stride], (size_t)width);
            return true;
        }
        if (h2 && v1) {
            const uint8_t *in = &c.plane[(size_t)row * stride];
            if (iw == 1) {
                out[0] = out[1] = in[0];
                return true;
            }
            out[0] = in[0];
            out[1] = (uint8_t)((in[0] * 3u + in[1] + 2u) >> 2);
            for (int i = 1; i < iw - 1; ++i) {
                const uint32_t s = 3u * in[i] + 2u;
                out[(size_t)i * 2] = (uint8_t)((s + in[i - 1]) >> 2);
                out[(size_t)i * 2 + 1] = (uint8_t)((s + in[i + 1]) >> 2);
            }
            out[(size_t)(iw - 1) * 2] = (uint8_t)((in[iw - 1] * 3u + in[iw - 2] + 2u) >> 2);
            out[(size_t)(iw - 1) * 2 + 1] = in[iw - 1];
            return true;
        }
        if (h2 && v2) {
            const float row_near = (float)row / 2.0f;
            // fractional part 0 -> the far row is the previous one, 0.5 -> the next one
            float row_far = row_near + (row_near - std::floor(row_near)) * 3.0f - 0.25f;
            const float last = (float)(ih - 1);
            if (row_far > last) row_far = last;
            const int rn = (int)row_near, rf = row_far < 0.0f ? 0 : (int)row_far;
            const uint8_t *near_ = &c.plane[(size_t)rn * stride], *far_ = &c.plane[(size_t)rf * stride];
            if (iw == 1) {
                out[0] = out[1] = (uint8_t)((3u * near_[0] + far_[0] + 2u) >> 2);
                return true;
            }
            uint32_t t0 = 3u * near_[0] + far_[0];
            uint32_t t1 = 3u * near_[1] + far_[1];
            out[0] = (uint8_t)((t0 + 2u) >> 2);
            out[1] = (uint8_t)((3u * t0 + t1 + 8u) >> 4);
            for (int i = 2; i < iw; ++i) {
                const uint32_t t2 = 3u * near_[i] + far_[i];
                out[(size_t)i * 2 - 2] = (uint8_t)((3u * t1 + t0 + 8u) >> 4);
                out[(size_t)i * 2 - 1] = (uint8_t)((3u * t1 + t2 + 8u) >> 4);
                t0 = t1;
                t1 = t2;
            }
            out[(size_t)iw * 2 - 2] = (uint8_t)((3u * t1 + t0 + 8u) >> 4);
            out[(size_t)iw * 2 - 1] = (uint8_t)((t1 + 2u) >> 2);
            return true;
        }
        return fail("unsupported chroma subsampling ratio (jpeg-decoder 0.1.11 supports 1x1, 2x1 and 2x2)",
                    RGH_E_UNSUPPORTED);
    }

    static inline uint8_t to_u8(float v) {
        const int i = (int)(v + 0.5f);
        return clamp8(i);
    }

    bool emit(rgh_image *img) {
        const int nc = (int)comps.size();
        img->width = (uint32_t)width;
        img->height = (uint32_t)height;
        img->channels = nc == 1 ? 1u : 3u;
        img->reserved = 0;
        img->pixels = (uint8_t *)rgh_alloc((size_t)width * height * img->channels);
        if (!img->pixels) return fail("out of memory");
        const bool ycbcr = nc == 3 && adobe_transform != 0;
        std::vector<std::vector<uint8_t>> rows((size_t)nc);
        for (int i = 0; i < nc; ++i) rows[(size_t)i].assign((size_t)(comps[(size_t)i].width + 1) * 2 + (size_t)width, 0);
        for (int y = 0; y < height; ++y) {
            for (int i = 0; i < nc; ++i)
                if (!upsample_row(comps[(size_t)i], comps[(size_t)i].v == vmax ? y : y, rows[(size_t)i])) {
                    rgh_free(img->pixels);
                    img->pixels = nullptr;
                    return false;
                }
            uint8_t *o = img->pixels + (size_t)y * width * img->channels;
            if (nc == 1) {
                std::memcpy(o, rows[0].data(), (size_t)width);
            } else if (!ycbcr) {
                for (int x = 0; x < width; ++x) {
                    o[3 * x] = rows[0][(size_t)x];
                    o[3 * x + 1] = rows[1][(size_t)x];
                    o[3 * x + 2] = rows[2][(size_t)x];
                }
            } else {
                for (int x = 0; x < width; ++x) {
                    const float Y = (float)rows[0][(size_t)x];
                    const float cb = (float)rows[1][(size_t)x] - 128.0f;
                    const float cr = (float)rows[2][(size_t)x] - 128.0f;
                    const float r = Y + 1.40200f * cr;
                    const float g = Y - 0.34414f * cb - 0.71414f * cr;
                    const float b = Y + 1.77200f * cb;
                    o[3 * x] = to_u8(r);
                    o[3 * x + 1] = to_u8(g);
                    o[3 * x + 2] = to_u8(b);
                }
            }
        }
        return true;
    }

    bool run(rgh_image *img) {
        if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG (no SOI marker)");
        pos = 2;
        bool seen_eoi = false, any_scan = false;
        while (!seen_eoi) {
            // find the next marker
            if (!need(2)) break;
            if (data[pos] != 0xFF) {
                ++pos;
                continue;
            }
            while (pos < len && data[pos] == 0xFF) ++pos;
            if (pos >= len) break;
            const int m = (int)data[pos++];
            if (m == 0 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
            if (m == 0xD9) {
                seen_eoi = true;
                break;
            }
            if (!need(2)) break;
            const size_t seg = u16();
            if (seg < 2 || pos + seg - 2 > len) return fail("bad marker segment length");
            const size_t end = pos + seg - 2;
            bool ok = true;
            switch (m) {
                case 0xDB: ok = parse_dqt(end); break;
                case 0xC4: ok = parse_dht(end); break;
                case 0xC0: case 0xC1: case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA:
                case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    ok = parse_sof(m, end);
                    break;
                case 0xDD:
                    if (end - pos < 2) return fail("short DRI");
                    restart_interval = (int)u16();
                    break;
                case 0xE0:
                    if (end - pos >= 5 && std::memcmp(data + pos, "JFIF\0", 5) == 0) is_jfif = true;
                    break;
                case 0xEE:
                    if (end - pos >= 12 && std::memcmp(data + pos, "Adobe", 5) == 0) adobe_transform = data[pos + 11];
                    break;
                case 0xDA:
                    ok = parse_sos(end);
                    any_scan = ok;
                    if (ok) continue;  // decode_scan positioned `pos` on the next marker
                    break;
                default: break;  // APPn, COM, ...: skipped
            }
            if (!ok) return false;
            pos = end;
        }
        if (!err.empty()) return false;
        if (!have_frame || !any_scan) return fail("JPEG has no image data");
        if (!reconstruct_planes()) return false;
        return emit(img);
    }
};

}  // namespace

int jpeg_decode(const uint8_t *data, size_t len, rgh_image *out) {
    Decoder d;
    d.data = data;
    d.len = len;
    std::memset(d.quant, 0, sizeof d.quant);
    if (!d.run(out)) return set_error(d.err_code, "JPEG: " + d.err);
    return RGH_OK;
}

}  // namespace rgh
