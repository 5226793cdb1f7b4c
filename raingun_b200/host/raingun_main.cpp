// `raingun` — the reference CLI's render command on the B200 path (src/main.rs:98-132,
// src/render.rs:49-121): parse the options, load the YAML scene and its textures natively
// (libraingun_host.so), render on the GPU through the C ABI (libraingun_b200.so), save a PNG and
// print the reference's one-line report.  `--preview` has no window here: it drives the streaming
// entry point (Scene::streaming_render -> rg_render_stream) into the collector buffer and reports
// band progress on stderr.  There is no CPU fallback: without a usable sm_100 device the scene
// upload fails and the process exits like the reference does on a panic (status 101).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/raingun_b200.h"
#include "../../include/raingun_host.h"

namespace {

// src/render.rs:233-250
std::string format_duration(long long ms) {
    char buf[64];
    if (ms < 800) std::snprintf(buf, sizeof buf, "%lldms", ms);
    else if (ms < 60000) std::snprintf(buf, sizeof buf, "%.2fs", (double)((float)ms / 1000.0f));
    else {
        const long long minutes = ms / 60000, left = ms - minutes * 60000;
        std::snprintf(buf, sizeof buf, "%lldm %.2fs", minutes, (double)((float)left / 1000.0f));
    }
    return buf;
}

struct Collector {
    std::vector<uint8_t> *image;
    uint32_t width, height, done;
};

int collect_rows(uint32_t y0, uint32_t rows, uint32_t width, const uint8_t *rgba, void *user) {
    Collector *c = (Collector *)user;
    std::memcpy(c->image->data() + (size_t)y0 * width * 4, rgba, (size_t)rows * width * 4);
    c->done += rows;
    std::fprintf(stderr, "\rpreview: %u / %u rows", c->done, c->height);
    if (c->done >= c->height) std::fputc('\n', stderr);
    return 0;
}

[[noreturn]] void die(const char *what, const char *detail, int status) {
    std::fprintf(stderr, "raingun: %s%s%s\n", what, detail && *detail ? ": " : "", detail ? detail : "");
    std::exit(status);
}

}  // namespace

const char kUsage[] =
    "USAGE:\n    raingun [FLAGS] [OPTIONS] <FILE>\n\nFLAGS:\n"
    "        --4k         Renders in 4K resolution. Explicit width/height overrides.\n"
    "        --draft      Renders in 800x600 and lower quality settings.\n"
    "        --hd         Renders in 1080 (HD) resolution. Explicit width/height overrides.\n"
    "        --help       Prints help information\n"
    "        --preview    Streams finished row bands while rendering.\n"
    "        --gpus <N>   Render on N GPUs of this machine (0 or 'all' = every visible GPU; default 1;\n"
    "                     the image is split into row tiles inside the library). B200 path only.\n"
    "    -V, --version    Prints version information\n\nOPTIONS:\n"
    "    -h, --height <PIXELS>      Height of output image.\n"
    "    -o, --output <FILENAME>    Specify filename of the rendered image.\n"
    "    -w, --width <PIXELS>       Width of output image.\n\nARGS:\n"
    "    <FILE>    The scene definition file, in YAML format.\n";

int main(int argc, char **argv) {
    // clap's built-ins (`-h` is taken by --height, so help is long-only): print and exit 0
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--")) break;
        if (!std::strcmp(argv[i], "--help")) {
            std::printf("raingun 0.1.0 (B200 render path)\n\n%s", kUsage);
            return 0;
        }
        if (!std::strcmp(argv[i], "--version") || !std::strcmp(argv[i], "-V")) {
            std::printf("raingun 0.1.0\n");
            return 0;
        }
    }
    // --gpus N is this binary's own option (the reference has one machine's CPU cores instead): take it out
    // of argv before the clap-compatible parser sees it
    int gpus = 1;
    {
        int w = 1;
        for (int i = 1; i < argc; ++i) {
            const bool eq = !std::strncmp(argv[i], "--gpus=", 7);
            if (eq || (!std::strcmp(argv[i], "--gpus") && i + 1 < argc)) {
                const char *v = eq ? argv[i] + 7 : argv[++i];
                gpus = !std::strcmp(v, "all") ? 0 : std::atoi(v);
                if (gpus < 0) gpus = 1;
                continue;
            }
            argv[w++] = argv[i];
        }
        argc = w;
    }
    rgh_cli_options opt;
    if (rgh_cli_parse(argc, argv, &opt) != RGH_OK) {
        std::fprintf(stderr, "error: %s\n\n%s", rgh_last_error(), kUsage);
        return std::strstr(rgh_last_error(), "Could not guess output filename") ? 2 : 1;
    }
    rgh_scene *hs = nullptr;
    if (rgh_scene_load(opt.input, nullptr, &hs) != RGH_OK) die(rgh_last_error(), nullptr, 101);
    if (opt.max_depth_limit >= 0) rgh_scene_limit_depth(hs, (uint32_t)opt.max_depth_limit);

    const char *dev_env = std::getenv("RAINGUN_DEVICE");
    rg_scene *scene = nullptr;
    int create_rc;
    if (gpus == 1) {
        create_rc = rg_scene_create(rgh_scene_desc(hs), dev_env ? std::atoi(dev_env) : 0, &scene);
    } else {
        const int have = rg_device_count();
        if (gpus == 0 || gpus > have) gpus = have;
        std::vector<int32_t> devices;
        for (int k = 0; k < gpus; ++k) devices.push_back(k);
        create_rc = gpus > 0 ? rg_scene_create_multi(rgh_scene_desc(hs), devices.data(), (uint32_t)devices.size(), &scene)
                             : rg_scene_create(rgh_scene_desc(hs), 0, &scene);   // no device: fails with the library's message
    }
    if (create_rc != RG_OK) die("Could not upload the scene", rg_last_error(), 101);

    std::vector<uint8_t> image((size_t)opt.width * opt.height * 4);
    const auto t0 = std::chrono::steady_clock::now();
    int rc;
    rg_stats stats;
    std::memset(&stats, 0, sizeof stats);
    if (opt.preview) {
        Collector c{&image, opt.width, opt.height, 0};
        rc = rg_render_stream(scene, opt.width, opt.height, 0, collect_rows, &c, &stats);
    } else {
        rc = rg_render(scene, opt.width, opt.height, image.data(), &stats);
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (rc != RG_OK) die("render failed", rg_last_error(), 101);
    // Where the reference panics, it writes no image and exits 101; the device counts those pixels instead of
    // unwinding, so the decision is taken here: NaN hit distance (scene.rs:38), `.unwrap()` on a transmission
    // that does not exist (rendering.rs:106), an AABB hit no face claims (bodies.rs:324).
    if (stats.err_nan_distance || stats.err_transmission_none || stats.err_aabb_normal) {
        std::fprintf(stderr,
                     "raingun: the reference would have panicked on this scene: %llu NaN hit distance(s) (scene.rs:38), %llu missing "
                     "transmission(s) (rendering.rs:106), %llu undecidable box normal(s) (bodies.rs:324); no image written\n",
                     (unsigned long long)stats.err_nan_distance, (unsigned long long)stats.err_transmission_none,
                     (unsigned long long)stats.err_aabb_normal);
        return 101;
    }
    if (rgh_png_save(opt.output, image.data(), opt.width, opt.height, 4) != RGH_OK) die(rgh_last_error(), nullptr, 101);
    const auto t2 = std::chrono::steady_clock::now();
    using ms = std::chrono::milliseconds;
    std::printf("%s\t\xE2\x86\x92\t%s\t(%s render, %s write)\n", opt.input, opt.output,
                format_duration(std::chrono::duration_cast<ms>(t1 - t0).count()).c_str(),
                format_duration(std::chrono::duration_cast<ms>(t2 - t1).count()).c_str());
    rg_scene_destroy(scene);
    rgh_scene_destroy(hs);
    return 0;
}
