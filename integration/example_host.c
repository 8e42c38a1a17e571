/*
 * integration/example_host.c — plain C99 client of the C ABI (no C++, no torch, no CUDA headers).
 * Built by tests/test_abi.py to prove the boundary is a real C interface:
 *     gcc -std=c99 -Iinclude integration/example_host.c -Lmitsubaer_b200 -lmitsubaer_b200 -o example_host
 * Renders a small frame of a radial GRIN box when a B200 is present; otherwise reports the error string.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mitsubaer_b200.h"

int main(void) {
    const int N = 48;
    if (mer_abi_version() != MER_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }
    float *data = (float *) malloc(sizeof(float) * N * N * N);
    mer_volume_desc vd;
    memset(&vd, 0, sizeof(vd));
    const float pitch = 2.0f / (N - 7); /* unit box sits 3 voxels inside the grid */
    for (int i = 0; i < 3; i++) { vd.res[i] = N; vd.bbox_min[i] = -1.0f - 3 * pitch; vd.bbox_max[i] = 1.0f + 3 * pitch; }
    const float R2 = 3.0f * vd.bbox_max[0] * vd.bbox_max[0];
    for (int z = 0; z < N; z++)
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++) {
                float px = vd.bbox_min[0] + x * pitch, py = vd.bbox_min[1] + y * pitch, pz = vd.bbox_min[2] + z * pitch;
                data[(z * N + y) * N + x] = 2.0f - (px * px + py * py + pz * pz) / R2; /* createRadialRIFWithBox.m */
            }
    mer_rif *rif = NULL;
    if (mer_rif_create(0, &vd, data, MER_RIF_TRICUBIC, &rif) != MER_OK) {
        printf("no GPU path available: %s\n", mer_last_error());
        free(data);
        return 0; /* expected on a machine without a B200: there is no CPU fallback */
    }
    mer_medium_desc md;
    memset(&md, 0, sizeof(md));
    for (int i = 0; i < 3; i++) { md.sigma_s[i] = 3.6f; md.sigma_a[i] = 0.4f; md.shape[i] = -1.0f; md.shape[3 + i] = 1.0f; }
    md.stepsize = 2e-3f; md.medium_sampling_weight = -1.0f; md.strategy = MER_STRATEGY_SINGLE; md.channel = -1;
    md.shape_type = MER_SHAPE_BOX; md.hg_g = 0.9f;
    mer_medium *medium = NULL;
    if (mer_medium_create(&md, rif, NULL, &medium) != MER_OK) { fprintf(stderr, "%s\n", mer_last_error()); return 1; }
    mer_render_desc rd;
    memset(&rd, 0, sizeof(rd));
    rd.width = 64; rd.height = 64; rd.spp_total = 16; rd.sample_stride = 1; rd.seed = 20201201;
    rd.cam_origin[2] = -4.0f; rd.cam_up[1] = 1.0f; rd.fov_deg = 40.0f; rd.filter = MER_FILTER_GAUSSIAN;
    rd.max_depth = -1; rd.rr_depth = 5;
    rd.env_radiance[0] = rd.env_radiance[1] = rd.env_radiance[2] = 1.0f;
    float *film = (float *) calloc((size_t) rd.width * rd.height * 5, sizeof(float));
    mer_render_stats st;
    if (mer_render(medium, &rd, film, &st) != MER_OK) { fprintf(stderr, "%s\n", mer_last_error()); return 1; }
    printf("rendered %llu samples, %llu eikonal steps in %.2f ms (%llu passes)\n", (unsigned long long) st.samples,
           (unsigned long long) st.ray_steps, st.device_ms, (unsigned long long) st.passes);
    mer_medium_destroy(medium);
    mer_rif_destroy(rif);
    free(film);
    free(data);
    return 0;
}
