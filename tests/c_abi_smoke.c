/* A plain C99 consumer of include/grrt.h: scene from a TOML file (host half), grrt_render into host buffers,
 * grrt_tonemap.  Compiled with `cc -std=c99 -pedantic` by tests/test_gpu_cli.py, linked against libgrrt.so only:
 * shows that the boundary is a C ABI (no C++ / torch types) and that a host program needs nothing else.
 *
 *   c_abi_smoke <scene.toml> <resource_root> <width> <height> <n_devices>
 * prints one line: "ok W H rays steps checksum r g b" (checksum = sum of all X, Y, Z; r g b = centre pixel). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "grrt.h"

int main(int argc, char** argv) {
    GrrtCliOpts opts;
    GrrtHostScene* hs = NULL;
    GrrtCtx* ctx = NULL;
    GrrtStats st;
    char err[512];
    const GrrtScene* scene;
    double* xyza;
    uint8_t *cls, *stop, *rgb;
    uint32_t* steps;
    size_t n, i;
    double sum = 0.0;
    int rc, n_devices, d;
    int ids[GRRT_MAX_PEERS];
    uint32_t W, H;

    if (argc < 6) {
        fprintf(stderr, "usage: %s scene.toml resource_root width height n_devices\n", argv[0]);
        return 2;
    }
    grrt_cli_opts_default(&opts);
    opts.width = atol(argv[3]);
    opts.height = atol(argv[4]);
    opts.camera_position[0] = -10.0;
    opts.camera_position[1] = 0.0;
    opts.camera_position[2] = -0.5;
    opts.theta = -3.14159;
    n_devices = atoi(argv[5]);
    if (n_devices < 1 || n_devices > GRRT_MAX_PEERS) return 2;
    rc = grrt_host_scene_from_toml_file(argv[1], argv[2], &opts, &hs, err, sizeof err);
    if (rc != GRRT_OK) {
        fprintf(stderr, "scene: %s\n", err);
        return 1;
    }
    scene = grrt_host_scene_get(hs);
    if (grrt_abi_sizeof(7) != sizeof(GrrtScene) || grrt_abi_sizeof(9) != sizeof(GrrtStats)) {
        fprintf(stderr, "ABI struct size mismatch\n");
        return 1;
    }
    for (d = 0; d < n_devices; ++d) ids[d] = d;
    rc = n_devices == 1 ? grrt_init(0, &ctx) : grrt_init_multi(n_devices, ids, &ctx);
    if (rc != GRRT_OK) {
        fprintf(stderr, "no usable CUDA device (rc %d): libgrrt has no CPU fallback\n", rc);
        return 3;
    }
    W = (uint32_t)scene->camera.cols;
    H = (uint32_t)scene->camera.rows;
    n = (size_t)W * H;
    xyza = (double*)malloc(n * 4 * sizeof(double));
    cls = (uint8_t*)malloc(n);
    stop = (uint8_t*)malloc(n);
    steps = (uint32_t*)malloc(n * sizeof(uint32_t));
    rgb = (uint8_t*)malloc(n * 3);
    if (!xyza || !cls || !stop || !steps || !rgb) return 1;
    rc = grrt_render(ctx, scene, 0, 0, H, W, xyza, cls, stop, steps, &st);
    if (rc != GRRT_OK) {
        fprintf(stderr, "grrt_render: %s\n", grrt_last_error(ctx));
        return 1;
    }
    rc = grrt_tonemap(ctx, xyza, n, GRRT_TONEMAP_REINHARD, 1.0, rgb);
    if (rc != GRRT_OK) {
        fprintf(stderr, "grrt_tonemap: %s\n", grrt_last_error(ctx));
        return 1;
    }
    for (i = 0; i < n; ++i) sum += xyza[4 * i] + xyza[4 * i + 1] + xyza[4 * i + 2];
    i = (size_t)(H / 2) * W + W / 2;
    printf("ok %u %u %llu %llu %.17g %u %u %u\n", W, H, (unsigned long long)(st.rays_base + st.rays_supersample),
           (unsigned long long)st.accepted_steps, sum, rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    free(xyza); free(cls); free(stop); free(steps); free(rgb);
    grrt_destroy(ctx);
    grrt_host_scene_free(hs);
    return 0;
}
