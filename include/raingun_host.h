/*
 * raingun_host.h — C ABI of the native host side of the drop-in.
 *
 * libraingun_host.so is plain C++17 (no CUDA, no Python): it does what the reference's
 * serde / `image` crates do in front of the render path, and produces the `rg_scene_desc`
 * that include/raingun_b200.h consumes.
 *
 *   YAML scene  -> rg_scene_desc   replaces serde_yaml::from_reader::<Scene>   src/main.rs:117-118
 *                                  (schema: scene.rs:11-31, bodies.rs:13-47, lights.rs:8-26,
 *                                   material.rs:7-54, color.rs:114-160)
 *   JPEG / PNG  -> RGB8 / RGBA8    replaces image::open                        material.rs:34-47
 *                                  (image 0.12.3 -> jpeg-decoder 0.1.11, png 0.6.2; Cargo.lock)
 *   RGBA8       -> PNG file        replaces ImageBuffer::save                  src/render.rs:58
 *   CLI options -> width/height/depth   replaces RenderOptions::from          src/main.rs:66-96
 *
 * Plain pointers and sizes only.  Buffers returned through `uint8_t **` are owned by the
 * library and released with rgh_free().  Every function returns 0 on success and a negative
 * RGH_E_* code on failure; rgh_last_error() describes the last failure of the calling thread.
 */
#ifndef RAINGUN_HOST_H
#define RAINGUN_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "raingun_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RGH_OK = 0,
    RGH_E_INVALID = -1,     /* null pointer / bad argument                               */
    RGH_E_IO = -2,          /* file cannot be opened / read / written                    */
    RGH_E_FORMAT = -3,      /* malformed JPEG / PNG / YAML                               */
    RGH_E_UNSUPPORTED = -4, /* valid file using a feature outside the supported subset   */
    RGH_E_SCHEMA = -5,      /* YAML is well formed but is not a `Scene` (serde error)     */
    RGH_E_USAGE = -6        /* CLI usage error (clap would have exited with status 1)    */
};

/* A decoded image, rows tightly packed, top row first.  channels: 1 (L8), 3 (RGB8), 4 (RGBA8). */
typedef struct rgh_image {
    uint32_t width;
    uint32_t height;
    uint32_t channels;
    uint32_t reserved;
    uint8_t *pixels; /* rgh_free() */
} rgh_image;

/* JPEG decode, restating jpeg-decoder 0.1.11's arithmetic (stb-style integer IDCT with the
 * de-quantisation fused, triangle-filter H2V1 / H2V2 chroma upsampling, f32 YCbCr->RGB):
 * baseline, extended-sequential and progressive Huffman, 8-bit, 1 or 3 components, restart
 * intervals.  Output: L8 (1 component) or RGB8. */
int rgh_jpeg_decode(const uint8_t *data, size_t len, rgh_image *out);

/* PNG decode (8/16-bit, all colour types, non-interlaced and Adam7) to L8 / RGB8 / RGBA8
 * the way `image` presents them to `DynamicImage::get_pixel`; and PNG encode of RGB8/RGBA8
 * (src/render.rs:58 saves the frame as RGBA8 PNG). */
int rgh_png_decode(const uint8_t *data, size_t len, rgh_image *out);
int rgh_png_encode(const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels,
                   uint8_t **out, size_t *out_len);

/* The simple formats among the others image 0.12 opens: BMP (palettes, 16/24/32-bit, bit fields; no
 * RLE), TGA (true-colour, grey, colour-mapped; raw and RLE), PNM (P1-P6).  L8 / RGB8 / RGBA8. */
int rgh_bmp_decode(const uint8_t *data, size_t len, rgh_image *out);
int rgh_tga_decode(const uint8_t *data, size_t len, rgh_image *out);
int rgh_pnm_decode(const uint8_t *data, size_t len, rgh_image *out);
/* GIF 87a / 89a: the first frame (what image 0.12 hands to DynamicImage) as RGBA8 on a canvas of the
 * logical screen size; the transparency index becomes alpha 0. */
int rgh_gif_decode(const uint8_t *data, size_t len, rgh_image *out);

/* image::open: chooses the decoder by file extension, as image 0.12 does (jpg, jpeg, png, bmp,
 * tga, gif, pbm, pgm, ppm, pnm here; tiff, webp, ico, hdr are reported as unsupported). */
int rgh_image_open(const char *path, rgh_image *out);
int rgh_png_save(const char *path, const uint8_t *pixels, uint32_t width, uint32_t height,
                 uint32_t channels);

void rgh_free(void *p);

/* ---- scene ingestion --------------------------------------------------------------------- */
typedef struct rgh_scene rgh_scene;

/* Texture resolver: called once per distinct `image:` path of the YAML.  Must fill `out`
 * (pixels allocated with rgh_alloc) and return 0, or return non-zero ("Could not load texture
 * file", material.rs:43-46).  NULL = rgh_image_open on `texture_root`/path (texture_root NULL or
 * "" = the CWD, as in the reference, material.rs:41-42). */
typedef int (*rgh_texture_cb)(const char *path, rgh_image *out, void *user);
void *rgh_alloc(size_t n);

int rgh_scene_parse(const char *yaml, size_t len, const char *texture_root,
                    rgh_texture_cb loader, void *user, rgh_scene **out);
int rgh_scene_load(const char *path, const char *texture_root, rgh_scene **out);
/* The flattened scene; pointers stay valid until rgh_scene_destroy. */
const rg_scene_desc *rgh_scene_desc(const rgh_scene *scene);
/* src/main.rs:119-123: a limit only ever lowers the scene's own max_recursion_depth. */
void rgh_scene_limit_depth(rgh_scene *scene, uint32_t limit);
/* i-th distinct texture path in first-use order (NULL when out of range). */
const char *rgh_scene_texture_path(const rgh_scene *scene, uint32_t i);
void rgh_scene_destroy(rgh_scene *scene);

/* ---- CLI option mapping (src/main.rs:21-96, src/render.rs:14-31) --------------------------- */
typedef struct rgh_cli_options {
    uint32_t width;            /* default 800  (render.rs RenderOptions::default) */
    uint32_t height;           /* default 600 */
    int32_t max_depth_limit;   /* -1 = none; --draft sets 4 */
    int32_t preview;           /* --preview given */
    char input[4096];
    char output[4096];         /* -o, or input with its extension replaced by .png */
} rgh_cli_options;
int rgh_cli_parse(int argc, const char *const *argv, rgh_cli_options *out);

const char *rgh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* RAINGUN_HOST_H */
