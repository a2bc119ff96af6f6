/* The reference renderer's threading contract, driven through the C ABI from plain C + pthreads -- what a Go process with
 * the cgo shim does, without Go (there is no Go toolchain in the image):
 *
 *   renderer/default.go:62-77    one worker per tracer, fed block requests through a channel
 *   renderer/default.go:106-171  renderFrame: Schedule rows -> send a COPY of the block request to every worker -> wait for
 *                                all of them -> primary.SyncFramebuffer(full frame)
 *   renderer/default.go:174-196  jobWorker: Trace(&req); primary.MergeOutput(self, &req) -- every worker calls MergeOutput
 *                                on the SAME primary, concurrently, from its own OS thread
 *   tracer/scheduler.go:50-80    perfect scheduler: rows ~ BlockH / RenderTime of the previous frame, floor, min 1,
 *                                remainder to tracer 0; frame 0 = naive (equal Speed() -> equal rows, :83-106)
 *
 *   render_multi <scene.plrscn> <w> <h> <spp> <tracers> <frames> <out.bin> <16 frustum floats> <3 eye floats>
 *
 * Tracer i lives on CUDA device i % pc_device_count(): on a multi-GPU box the merges cross NVLink (peer loads inside
 * pc_merge_output), on a one-GPU box the same contract runs with every handle on device 0.  Frames accumulate
 * progressively (AccumulatedSamples += spp per frame, renderer/opengl.go:136-171).  out.bin receives what a checker needs
 * to replay the run on ONE handle and compare bit for bit: "PCMULTI1", tracers, frames, rows[frame][tracer] (uint32), then
 * the primary's frame accumulator (w*h float4) and the RGBA8 frame.  Seeds of tracer i in frame f are the splitmix64 list
 * of SURVEY §8(d) started at 0x501A2150 + 7 + 100*i + f.
 *
 * TEST CODE (tests/test_cpu_abi.py compiles it, tests/test_gpu_multi.py runs it).  Exit: 0 ok, 2 usage, 3 I/O,
 * 10 + pc_status for library errors. */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "polaris_cuda.h"

#define MAX_TRACERS 16
#define NUM_BOUNCES 5

typedef struct worker {
    int index;
    pc_tracer *self, *primary;
    pthread_t thread;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int has_job, quit;     /* the "channel" (default.go:62-77): one pending block request */
    pc_block_request req;  /* a COPY, like the reference sends (default.go:128-136) */
    uint32_t frame;
    /* completion (jobCompleteChan, default.go:143-156) */
    int done, status;
    pc_stats stats;
    double trace_s, merge_s;
} worker;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void splitmix(uint64_t state, uint32_t *out, uint32_t n) {
    uint32_t i;
    for (i = 0; i < n; i++) {
        uint64_t z;
        state += 0x9E3779B97F4A7C15ull;
        z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        out[i] = (uint32_t)z;
    }
}

static pthread_mutex_t g_done_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_done_cv = PTHREAD_COND_INITIALIZER;

static void *job_worker(void *arg) {  /* default.go:174-196 */
    worker *w = (worker *)arg;
    for (;;) {
        pc_block_request req;
        uint32_t n_seeds, frame;
        uint32_t *seeds;
        double t0, t1, t2;
        int rc;
        pthread_mutex_lock(&w->mu);
        while (!w->has_job && !w->quit) pthread_cond_wait(&w->cv, &w->mu);
        if (w->quit) { pthread_mutex_unlock(&w->mu); return NULL; }
        req = w->req;
        frame = w->frame;
        w->has_job = 0;
        pthread_mutex_unlock(&w->mu);

        n_seeds = req.samples_per_pixel * (1 + req.num_bounces);
        seeds = (uint32_t *)malloc(4 * (size_t)n_seeds);
        splitmix(0x501A2150ull + 7 + 100 * (uint64_t)w->index + frame, seeds, n_seeds);
        t0 = now_s();
        rc = pc_trace(w->self, &req, seeds, n_seeds, &w->stats);  /* tr.Trace(&blockReq), default.go:188 */
        t1 = now_s();
        if (!rc) rc = pc_merge_output(w->primary, w->self, &req); /* primary.MergeOutput(tr, &blockReq), default.go:191 */
        t2 = now_s();
        free(seeds);
        if (rc) fprintf(stderr, "worker %d: %s\n", w->index, pc_last_error(rc && t1 == t2 ? w->self : w->primary));
        pthread_mutex_lock(&g_done_mu);
        w->status = rc;
        w->trace_s = t1 - t0;
        w->merge_s = t2 - t1;
        w->done = 1;
        pthread_cond_broadcast(&g_done_cv);
        pthread_mutex_unlock(&g_done_mu);
    }
}

static void *read_section(FILE *f, uint64_t *bytes) {
    void *p;
    if (fread(bytes, 8, 1, f) != 1) return NULL;
    p = malloc(*bytes ? (size_t)*bytes : 1);
    if (*bytes && fread(p, 1, (size_t)*bytes, f) != (size_t)*bytes) { free(p); return NULL; }
    return p;
}

/* tracer/scheduler.go:50-80 (perfect) on the previous frame's (BlockH, RenderTime); frame 0: naive with equal speeds */
static void schedule(uint32_t n, uint32_t frame_h, const uint32_t *prev_h, const double *prev_t, int have_prev, uint32_t *rows) {
    uint32_t i, used = 0;
    if (!have_prev) {
        for (i = 0; i < n; i++) { rows[i] = frame_h / n; used += rows[i]; }
    } else {
        double total = 0.0;
        for (i = 0; i < n; i++) total += (double)prev_h[i] / prev_t[i];  /* rows per second */
        for (i = 0; i < n; i++) {
            double share = ((double)prev_h[i] / prev_t[i]) / total;
            rows[i] = (uint32_t)(share * (double)frame_h);  /* floor */
            if (rows[i] < 1) rows[i] = 1;
            used += rows[i];
        }
    }
    if (used <= frame_h) rows[0] += frame_h - used;  /* remainder to the first tracer */
    else {  /* the min-1 clamp can overshoot on tiny frames: take it back from the largest block */
        while (used > frame_h) {
            uint32_t big = 0;
            for (i = 1; i < n; i++) if (rows[i] > rows[big]) big = i;
            rows[big]--; used--;
        }
    }
}

int main(int argc, char **argv) {
    FILE *f;
    char magic[8];
    uint32_t version, w, h, spp, n, frames, i, fr;
    void *sec[10];
    uint64_t len[10];
    int32_t globals[2];
    float cam[10], eye[3], frustum[16];
    pc_scene_view view;
    worker wk[MAX_TRACERS];
    uint32_t rows[MAX_TRACERS], prev_h[MAX_TRACERS];
    double prev_t[MAX_TRACERS];
    uint32_t *all_rows;
    float *acc;
    uint8_t *rgba;
    int rc, nd;
    double t_all0, rays_total = 0.0;

    if (argc != 8 + 19) { fprintf(stderr, "usage: render_multi scene w h spp tracers frames out.bin <16 frustum floats> <3 eye floats>\n"); return 2; }
    w = (uint32_t)atoi(argv[2]); h = (uint32_t)atoi(argv[3]); spp = (uint32_t)atoi(argv[4]);
    n = (uint32_t)atoi(argv[5]); frames = (uint32_t)atoi(argv[6]);
    if (n < 1 || n > MAX_TRACERS || frames < 1 || h < n) return 2;
    for (i = 0; i < 16; i++) frustum[i] = strtof(argv[8 + i], NULL);
    for (i = 0; i < 3; i++) eye[i] = strtof(argv[24 + i], NULL);
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

    nd = pc_device_count();
    if (nd < 1) { fprintf(stderr, "no CUDA device\n"); return 10 + PC_ERR_NO_DEVICE; }
    /* initTracers (default.go:199-253): one tracer per device, every one gets frame dimensions, scene and camera
     * (default.go:70-72); the first is the primary (default.go:255-292) */
    memset(wk, 0, sizeof wk);
    for (i = 0; i < n; i++) {
        char id[32];
        snprintf(id, sizeof id, "cuda:%u", i % (uint32_t)nd);
        if ((rc = pc_create((int)(i % (uint32_t)nd), id, &wk[i].self))) { fprintf(stderr, "pc_create: %s\n", pc_last_error(NULL)); return 10 + rc; }
        if ((rc = pc_resize(wk[i].self, w, h)) || (rc = pc_upload_scene(wk[i].self, &view)) || (rc = pc_set_camera(wk[i].self, eye, frustum))) {
            fprintf(stderr, "setup %u: %s\n", i, pc_last_error(wk[i].self));
            return 10 + rc;
        }
    }
    for (i = 0; i < 10; i++) free(sec[i]);
    for (i = 0; i < n; i++) {
        wk[i].index = (int)i;
        wk[i].primary = wk[0].self;
        pthread_mutex_init(&wk[i].mu, NULL);
        pthread_cond_init(&wk[i].cv, NULL);
        if (pthread_create(&wk[i].thread, NULL, job_worker, &wk[i])) return 3;
    }
    all_rows = (uint32_t *)calloc((size_t)frames * n, 4);
    printf("render_multi: %u tracers on %d device(s), %ux%u, %u spp x %u frames\n", n, nd, w, h, spp, frames);
    t_all0 = now_s();
    for (fr = 0; fr < frames; fr++) {  /* renderFrame (default.go:106-171) */
        pc_block_request req, sync_req;
        uint32_t y = 0;
        double t0 = now_s(), t1, t2, frame_rays = 0.0;
        int failed = 0;
        schedule(n, h, prev_h, prev_t, fr > 0, rows);
        memset(&req, 0, sizeof req);
        req.frame_w = w; req.frame_h = h; req.block_w = w;
        req.samples_per_pixel = spp; req.num_bounces = NUM_BOUNCES; req.min_bounces_for_rr = 3; req.exposure = 1.2f;
        req.accumulated_samples = fr * spp;
        for (i = 0; i < n; i++) {
            req.block_y = y; req.block_h = rows[i];
            y += rows[i];
            all_rows[(size_t)fr * n + i] = rows[i];
            pthread_mutex_lock(&wk[i].mu);
            wk[i].req = req;  /* a copy per worker */
            wk[i].frame = fr;
            wk[i].has_job = 1;
            pthread_cond_signal(&wk[i].cv);
            pthread_mutex_unlock(&wk[i].mu);
        }
        pthread_mutex_lock(&g_done_mu);
        for (;;) {
            uint32_t ready = 0;
            for (i = 0; i < n; i++) ready += wk[i].done ? 1u : 0u;
            if (ready == n) break;
            pthread_cond_wait(&g_done_cv, &g_done_mu);
        }
        for (i = 0; i < n; i++) {
            wk[i].done = 0;
            if (wk[i].status) failed = wk[i].status;
            prev_h[i] = rows[i];
            prev_t[i] = wk[i].trace_s > 1e-9 ? wk[i].trace_s : 1e-9;  /* Stats().RenderTime */
            frame_rays += (double)(wk[i].stats.query_rays + wk[i].stats.occlusion_rays);
        }
        pthread_mutex_unlock(&g_done_mu);
        if (failed) return 10 + failed;
        t1 = now_s();
        memset(&sync_req, 0, sizeof sync_req);  /* the full frame (default.go:159-161) */
        sync_req.frame_w = w; sync_req.frame_h = h; sync_req.block_w = w; sync_req.block_h = h;
        sync_req.samples_per_pixel = spp; sync_req.exposure = 1.2f; sync_req.accumulated_samples = fr * spp;
        if ((rc = pc_sync_framebuffer(wk[0].self, &sync_req, NULL))) { fprintf(stderr, "sync: %s\n", pc_last_error(wk[0].self)); return 10 + rc; }
        t2 = now_s();
        rays_total += frame_rays;
        printf("frame %u: rows", fr);
        for (i = 0; i < n; i++) printf(" %u", rows[i]);
        printf(" | trace ms");
        for (i = 0; i < n; i++) printf(" %.2f", wk[i].trace_s * 1e3);
        printf(" | merge-call ms");
        for (i = 0; i < n; i++) printf(" %.3f", wk[i].merge_s * 1e3);
        printf(" | workers %.2f ms, sync %.2f ms, %.1f Mrays/s\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3, frame_rays / (t2 - t0) / 1e6);
    }
    printf("total: %.1f Mrays/s over %u frames\n", rays_total / (now_s() - t_all0) / 1e6, frames);

    acc = (float *)malloc((size_t)w * h * 16);
    rgba = (uint8_t *)malloc((size_t)w * h * 4);
    if ((rc = pc_read_buffer(wk[0].self, PC_BUF_FRAME_ACCUMULATOR, acc, (uint64_t)w * h * 16)) ||
        (rc = pc_read_buffer(wk[0].self, PC_BUF_FRAME_BUFFER, rgba, (uint64_t)w * h * 4))) {
        fprintf(stderr, "read: %s\n", pc_last_error(wk[0].self));
        return 10 + rc;
    }
    f = fopen(argv[7], "wb");
    if (!f) return 3;
    fwrite("PCMULTI1", 8, 1, f);
    fwrite(&n, 4, 1, f);
    fwrite(&frames, 4, 1, f);
    fwrite(all_rows, 4, (size_t)frames * n, f);
    fwrite(acc, 16, (size_t)w * h, f);
    fwrite(rgba, 4, (size_t)w * h, f);
    fclose(f);
    for (i = 0; i < n; i++) {  /* Close (default.go:177): workers first, the primary last */
        pthread_mutex_lock(&wk[i].mu);
        wk[i].quit = 1;
        pthread_cond_signal(&wk[i].cv);
        pthread_mutex_unlock(&wk[i].mu);
        pthread_join(wk[i].thread, NULL);
    }
    for (i = n; i-- > 0;) pc_destroy(wk[i].self);
    free(all_rows); free(acc); free(rgba);
    return 0;
}
