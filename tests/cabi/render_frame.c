/* A plain C99 client of include/polaris_cuda.h: the call sequence a cgo binding makes, without
 * Python or ctypes in between.  It loads a PLRSCN2 scene dump (go/asset/scene/rawdump.go, polaris_b200/scene.py), renders
 * one frame through NewTracer/Init -> UpdateState x3 -> Trace -> MergeOutput -> SyncFramebuffer
 * (renderer/default.go:70-72,161,188-191) and writes the RGBA8 frame as a binary PPM.
 *
 *   render_frame <scene.plrscn> <width> <height> <spp> <out.ppm> [frustum: 16 floats, eye: 3 floats]
 *
 * TEST CODE (tests/test_cpu_abi.py compiles and links it; tests/test_gpu_parity.py runs it and compares the frame with
 * the one the Python binding renders).  Exit codes: 0 ok, 2 usage, 3 I/O, 10 + pc_status for library errors. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "polaris_cuda.h"

static void *read_section(FILE *f, uint64_t *bytes) {
    void *p;
    if (fread(bytes, 8, 1, f) != 1) return NULL;
    p = malloc(*bytes ? (size_t)*bytes : 1);
    if (*bytes && fread(p, 1, (size_t)*bytes, f) != (size_t)*bytes) { free(p); return NULL; }
    return p;
}

int main(int argc, char **argv) {
    FILE *f;
    char magic[8];
    uint32_t version, w, h, spp, i, n_seeds;
    void *sec[10];
    uint64_t len[10];
    int32_t globals[2];
    float cam[10], eye[3], frustum[16];
    pc_scene_view view;
    pc_tracer *tr = NULL;
    pc_block_request req;
    pc_stats stats;
    uint32_t *seeds;
    uint8_t *rgba;
    uint64_t state = 0x501A2150ull + 2;
    int rc;

    if (argc != 6 + 19) { fprintf(stderr, "usage: render_frame scene w h spp out.ppm <16 frustum floats> <3 eye floats>\n"); return 2; }
    w = (uint32_t)atoi(argv[2]); h = (uint32_t)atoi(argv[3]); spp = (uint32_t)atoi(argv[4]);
    for (i = 0; i < 16; i++) frustum[i] = strtof(argv[6 + i], NULL);
    for (i = 0; i < 3; i++) eye[i] = strtof(argv[22 + i], NULL);
    f = fopen(argv[1], "rb");
    if (!f || fread(magic, 8, 1, f) != 1 || memcmp(magic, "PLRSCN2\0", 8) || fread(&version, 4, 1, f) != 1 || version != 1) return 3;
    for (i = 0; i < 10; i++)
        if (!(sec[i] = read_section(f, &len[i]))) return 3;
    if (fread(globals, 4, 2, f) != 2 || fread(cam, 4, 10, f) != 10) return 3;
    fclose(f);

    memset(&view, 0, sizeof view);
    view.bvh_nodes = sec[0];        view.bvh_nodes_bytes = len[0];
    view.mesh_instances = sec[1];   view.mesh_instances_bytes = len[1];
    view.material_nodes = sec[2];   view.material_nodes_bytes = len[2];
    view.texture_data = sec[3];     view.texture_data_bytes = len[3];
    view.texture_metadata = sec[4]; view.texture_metadata_bytes = len[4];
    view.vertices = sec[5];         view.vertices_bytes = len[5];
    view.normals = sec[6];          view.normals_bytes = len[6];
    view.uvs = sec[7];              view.uvs_bytes = len[7];
    view.material_indices = sec[8]; view.material_indices_bytes = len[8];
    view.emissives = sec[9];        view.emissives_bytes = len[9];
    view.scene_diffuse_mat_index = globals[0];
    view.scene_emissive_mat_index = globals[1];

    if (pc_abi_version() != 1) return 4;
    if ((rc = pc_create(0, "cuda:0", &tr))) { fprintf(stderr, "pc_create: %s\n", pc_last_error(NULL)); return 10 + rc; }
    if ((rc = pc_resize(tr, w, h)) || (rc = pc_upload_scene(tr, &view)) || (rc = pc_set_camera(tr, eye, frustum))) {
        fprintf(stderr, "setup: %s\n", pc_last_error(tr));
        return 10 + rc;
    }
    for (i = 0; i < 10; i++) free(sec[i]);  /* the library copied everything during the calls */

    memset(&req, 0, sizeof req);
    req.frame_w = w; req.frame_h = h; req.block_w = w; req.block_h = h;
    req.samples_per_pixel = spp; req.num_bounces = 5; req.min_bounces_for_rr = 3; req.exposure = 1.2f;
    n_seeds = spp * (1 + req.num_bounces);
    seeds = (uint32_t *)malloc(4 * (size_t)n_seeds);
    for (i = 0; i < n_seeds; i++) {  /* splitmix64, the seed list of SURVEY §8(d) for config 2 */
        uint64_t z;
        state += 0x9E3779B97F4A7C15ull;
        z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        seeds[i] = (uint32_t)z;
    }
    if ((rc = pc_trace(tr, &req, seeds, n_seeds, &stats))) { fprintf(stderr, "pc_trace: %s\n", pc_last_error(tr)); return 10 + rc; }
    if ((rc = pc_merge_output(tr, tr, &req))) return 10 + rc;
    req.accumulated_samples = 0;
    rgba = (uint8_t *)malloc((size_t)w * h * 4);
    if ((rc = pc_sync_framebuffer(tr, &req, rgba))) { fprintf(stderr, "pc_sync_framebuffer: %s\n", pc_last_error(tr)); return 10 + rc; }
    f = fopen(argv[5], "wb");
    if (!f) return 3;
    fprintf(f, "P6\n%u %u\n255\n", w, h);
    for (i = 0; i < w * h; i++) fwrite(rgba + 4 * (size_t)i, 1, 3, f);
    fclose(f);
    printf("%llu query + %llu occlusion rays in %.3f ms on the device, %llu launches\n", (unsigned long long)stats.query_rays,
           (unsigned long long)stats.occlusion_rays, stats.device_time_ns / 1e6, (unsigned long long)stats.kernel_launches);
    pc_destroy(tr);
    free(seeds);
    free(rgba);
    return 0;
}
