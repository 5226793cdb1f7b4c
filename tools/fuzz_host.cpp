// Mutation fuzzer for libraingun_host's parsers (JPEG, PNG, BMP, TGA, PNM, GIF, YAML scene, CLI), to be built with sanitizers:
//   python tools/fuzz_host_inputs.py /tmp/fz            # writes the seed files from the test bundle
//   g++ -std=c++17 -O1 -g -fwrapv -fsanitize=address,undefined -Iinclude tools/fuzz_host.cpp \
//       raingun_b200/host/rgh_{api,jpeg,png,simple_formats,yaml,scene}.cpp -lz -o /tmp/fz/fuzz && (cd /tmp/fz && ./fuzz 4000)
// Truncations, byte flips, 0xFF runs and junk insertions; any sanitizer report or hang is a bug.
// FUZZ_KEEP_LAST=1 writes each input to last.bin before decoding it (to catch the one that hangs).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "raingun_host.h"
static std::vector<uint8_t> slurp(const char *p) { std::vector<uint8_t> v; FILE *f = fopen(p, "rb"); if (!f) return v; uint8_t b[65536]; size_t n; while ((n = fread(b, 1, sizeof b, f)) > 0) v.insert(v.end(), b, b + n); fclose(f); return v; }
static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static int no_tex(const char *, rgh_image *out, void *) { out->width = out->height = 2; out->channels = 3; out->pixels = (uint8_t *)rgh_alloc(12); memset(out->pixels, 7, 12); return 0; }
int main(int argc, char **argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    const char *imgs[] = {"a.jpg", "prog.jpg", "base420.jpg", "base422.jpg", "grey.jpg", "rgb.png", "pal.png", "rgba.png", "bit.png",
                          "rgb.bmp", "pal.bmp", "rgba.bmp", "rgb.tga", "rle.tga", "pal.tga", "rgb.ppm", "bit.pbm", "pal.gif", "lace.gif"};
    const char *ymls[] = {"test1.yml", "test2.yml", "test3.yml"};
    long ok = 0, bad = 0;
    for (const char *name : imgs) {
        std::vector<uint8_t> orig = slurp(name);
        if (orig.empty()) { printf("missing %s\n", name); continue; }
        int (*decode)(const uint8_t *, size_t, rgh_image *) = strstr(name, ".jpg") ? rgh_jpeg_decode : strstr(name, ".png") ? rgh_png_decode
                                                             : strstr(name, ".bmp") ? rgh_bmp_decode : strstr(name, ".tga") ? rgh_tga_decode
                                                             : strstr(name, ".gif") ? rgh_gif_decode : rgh_pnm_decode;
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> d = orig;
            int kind = rnd() % 4;
            if (kind == 0) d.resize(rnd() % (d.size() + 1));                                   // truncate
            else if (kind == 1) { int k = 1 + rnd() % 8; while (k--) d[rnd() % d.size()] = (uint8_t)rnd(); }   // byte flips
            else if (kind == 2) { size_t a = rnd() % d.size(), n = rnd() % 64; for (size_t i = a; i < a + n && i < d.size(); ++i) d[i] = 0xFF; }
            else { size_t a = rnd() % d.size(); size_t n = rnd() % 256; d.insert(d.begin() + a, n, (uint8_t)rnd()); }  // insert junk
            if (getenv("FUZZ_KEEP_LAST")) { FILE *f = fopen("last.bin", "wb"); fwrite(d.data(), 1, d.size(), f); fclose(f); }
            rgh_image im; memset(&im, 0, sizeof im);
            int rc = d.empty() ? -1 : decode(d.data(), d.size(), &im);
            if (rc == 0) { ++ok; volatile uint8_t x = im.pixels[(size_t)im.width * im.height * im.channels - 1]; (void)x; rgh_free(im.pixels); } else ++bad;
        }
        printf("%s: done\n", name); fflush(stdout);
    }
    for (const char *name : ymls) {
        std::vector<uint8_t> orig = slurp(name);
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> d = orig;
            int kind = rnd() % 4;
            static const char junk[] = " \n\t:-[]{},&*!|>#'\"0123456789.eE+-xyzSphere";
            if (kind == 0) d.resize(rnd() % (d.size() + 1));
            else if (kind == 1) { int k = 1 + rnd() % 8; while (k--) d[rnd() % d.size()] = (uint8_t)junk[rnd() % (sizeof junk - 1)]; }
            else if (kind == 2) { size_t a = rnd() % d.size(), n = rnd() % 40; d.erase(d.begin() + a, d.begin() + (a + n < d.size() ? a + n : d.size())); }
            else { size_t a = rnd() % d.size(); int n = rnd() % 12; while (n--) d.insert(d.begin() + a, (uint8_t)junk[rnd() % (sizeof junk - 1)]); }
            rgh_scene *sc = nullptr;
            int rc = rgh_scene_parse((const char *)d.data(), d.size(), nullptr, no_tex, nullptr, &sc);
            if (rc == 0) { ++ok; const rg_scene_desc *ds = rgh_scene_desc(sc); volatile uint32_t n = ds->n_bodies; (void)n; rgh_scene_destroy(sc); } else ++bad;
        }
        printf("%s: done\n", name); fflush(stdout);
    }
    {   // CLI option mapping: random argument vectors from the real vocabulary plus junk
        static const char *vocab[] = {"-w", "-h", "-o", "--width", "--height", "--output", "--width=12", "--height=", "--4k", "--hd",
                                      "--draft", "--preview", "--", "-", "-w9", "-h=7", "x.yml", "dir/", "..", "", "4294967296", "-3",
                                      "+5", "--bogus", "-q", "a/b.c/d", "\xff\xfe", "--output=o.png"};
        for (int it = 0; it < iters * 4; ++it) {
            const char *av[12];
            int ac = 1 + (int)(rnd() % 10);
            av[0] = "raingun";
            for (int i = 1; i < ac; ++i) av[i] = vocab[rnd() % (sizeof vocab / sizeof vocab[0])];
            rgh_cli_options o;
            if (rgh_cli_parse(ac, av, &o) == 0) { ++ok; volatile size_t n = strlen(o.input) + strlen(o.output); (void)n; } else ++bad;
        }
        printf("cli: done\n");
    }
    printf("ok %ld rejected %ld\n", ok, bad);
    return 0;
}
