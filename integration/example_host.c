/* A complete C host for the drop-in: what `raingun examples/test1.yml` does, in ~40 lines of C99
 * against the two headers.  Build (from the repository root, after build()):
 *   gcc -std=c99 -Iinclude integration/example_host.c -Lraingun_b200 -lraingun_host -lraingun_b200 \
 *       -Wl,-rpath,$PWD/raingun_b200 -o /tmp/example_host
 *   (cd <dir with examples/ and textures/> && /tmp/example_host examples/test1.yml out.png 800 600)
 * Exit status: 0 ok, 2 usage, 3 scene could not be loaded, 4 no usable GPU / render failed
 * (there is no CPU fallback), 5 PNG could not be written. */
#include <stdio.h>
#include <stdlib.h>

#include "raingun_b200.h"
#include "raingun_host.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s scene.yml out.png [width height]\n", argv[0]);
        return 2;
    }
    const unsigned width = argc > 3 ? (unsigned)atoi(argv[3]) : 800u, height = argc > 4 ? (unsigned)atoi(argv[4]) : 600u;

    rgh_scene *parsed = NULL; /* serde_yaml::from_reader::<Scene> + image::open, src/main.rs:117-118 */
    if (rgh_scene_load(argv[1], NULL, &parsed) != RGH_OK) {
        fprintf(stderr, "%s\n", rgh_last_error());
        return 3;
    }
    rg_scene *scene = NULL; /* the upload layer */
    if (rg_scene_create(rgh_scene_desc(parsed), 0, &scene) != RG_OK) {
        fprintf(stderr, "%s\n", rg_last_error());
        rgh_scene_destroy(parsed);
        return 4;
    }
    unsigned char *rgba = (unsigned char *)malloc((size_t)width * height * 4);
    rg_stats stats;
    int rc = rg_render(scene, width, height, rgba, &stats); /* Scene::render_image, scene.rs:41-43 */
    if (rc != RG_OK) {
        fprintf(stderr, "%s\n", rg_last_error());
    } else {
        printf("%llu rays in %.3f ms on the device (%u kernel launches)\n",
               (unsigned long long)(stats.rays_primary + stats.rays_shadow + stats.rays_reflection + stats.rays_transmission),
               stats.ms_device, stats.gpu_launches);
        if (rgh_png_save(argv[2], rgba, width, height, 4) != RGH_OK) { /* ImageBuffer::save, src/render.rs:58 */
            fprintf(stderr, "%s\n", rgh_last_error());
            rc = 5;
        }
    }
    free(rgba);
    rg_scene_destroy(scene);
    rgh_scene_destroy(parsed);
    return rc == RG_OK ? 0 : (rc == 5 ? 5 : 4);
}
