/*
 * polaris_cuda.h -- C ABI of libpolaris_cuda.so, the B200 (sm_100a) path-tracing
 * backend for polaris.
 *
 * This is the drop-in boundary for the one hot path of the reference: the Go
 * interface `tracer.Tracer` (reference tracer/tracer.go:80-111) as implemented by
 * `tracer/opencl.Tracer` (reference tracer/opencl/tracer.go).  Every entry point
 * below names the reference method it replaces.  A Go `tracer/cuda` package binds
 * these through cgo (see INTEGRATION.md and go/tracer/cuda/); tests and bench.py
 * bind them through ctypes (polaris_b200/_lib.py).
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every host pointer, the
 *    library copies during the call and never retains it (the reference instead
 *    relies on CL_MEM_USE_HOST_PTR, tracer/opencl/device/buffer.go:98-104);
 *  - every function that can fail returns 0 on success or a pc_status code;
 *    pc_last_error() gives the message for the calling handle;
 *  - calls on one handle may come from any OS thread (goroutines migrate); the
 *    library takes a per-handle mutex and does cudaSetDevice on entry;
 *  - all scene structs are the byte layouts of reference
 *    asset/scene/optimized_scene.go:25-165 == tracer/opencl/CL/types.cl.
 */
#ifndef POLARIS_CUDA_H
#define POLARIS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_ABI_VERSION 1

/* ---- status codes (mirror tracer/opencl/errors.go:5-22 where one exists) ---- */
typedef enum pc_status {
    PC_OK = 0,
    PC_ERR_INVALID_ARGUMENT = 1,   /* ErrInvalidChangeData / ErrInvalidOption      */
    PC_ERR_NO_DEVICE = 2,          /* ErrContextCreationFailed                     */
    PC_ERR_ALLOC = 3,              /* ErrAllocatingBuffer                          */
    PC_ERR_COPY_TO_DEVICE = 4,     /* ErrCopyingDataToDevice                       */
    PC_ERR_COPY_TO_HOST = 5,       /* ErrCopyingDataToHost                         */
    PC_ERR_KERNEL = 6,             /* ErrKernelExecutionFailed (sticky: handle dead)*/
    PC_ERR_NO_SCENE_DATA = 7,      /* ErrNoSceneData (errors.go:21)                */
    PC_ERR_NO_FRAME = 8,           /* Trace before FrameDimensions were committed  */
    PC_ERR_UNSUPPORTED_TRACER = 9, /* MergeOutput: "unsupported tracer instance"   */
    PC_ERR_STACK_DEPTH = 10,       /* scene BVH deeper than the traversal stack    */
    PC_ERR_BAD_SCENE = 11,         /* scene view fails validation at upload        */
    PC_ERR_PEER_ACCESS = 12        /* cross-device merge without a P2P path        */
} pc_status;

/* Opaque tracer handle == one tracer/opencl.Tracer bound to one CUDA device. */
typedef struct pc_tracer pc_tracer;

/* tracer.BlockRequest, reference tracer/tracer.go:6-34 (12 x 4 bytes, same order). */
typedef struct pc_block_request {
    uint32_t frame_w, frame_h;
    uint32_t block_x, block_y, block_w, block_h;
    uint32_t samples_per_pixel;
    uint32_t num_bounces;
    uint32_t min_bounces_for_rr;
    float    exposure;
    uint32_t seed;
    uint32_t accumulated_samples;
} pc_block_request;

/* tracer.Stats, reference tracer/tracer.go:37-47 (durations in nanoseconds) plus
 * device-side counters the reference never exposes (pipeline.go:278-282 is unused). */
typedef struct pc_stats {
    uint32_t block_w, block_h;
    uint64_t update_time_ns;
    uint64_t render_time_ns;        /* host wall time of the last pc_trace         */
    uint64_t device_time_ns;        /* CUDA-event time of the last pc_trace        */
    uint64_t query_rays;            /* closest-hit rays launched (primary+indirect)*/
    uint64_t occlusion_rays;        /* any-hit rays launched                       */
    uint64_t kernel_launches;       /* kernels launched by the last pc_trace       */
    /* valid only when PC_OPT_COUNTERS is on: */
    uint64_t nodes_tested;          /* BVH inner-node visits (2 child boxes each)  */
    uint64_t tris_tested;
    uint64_t instances_entered;
    uint64_t shaded_hits;
    uint64_t occlusion_emitted;
    uint64_t indirect_emitted;
    uint64_t unoccluded;
    uint64_t missed_query_rays;
    /* valid only when PC_OPT_KERNEL_TIMERS is on: CUDA-event time and launch count per kernel
     * class of the last pc_trace, see pc_kernel_class */
    uint64_t kernel_time_ns[8];
    uint64_t kernel_count[8];
} pc_stats;

typedef enum pc_kernel_class {
    PC_K_BEGIN_SAMPLE = 0, PC_K_PRIMARY = 1, PC_K_SHADE = 2, PC_K_OCCLUSION = 3, PC_K_QUERY = 4,
    PC_K_TRACE = 5   /* occlusion test + next bounce's closest-hit query in one launch (PC_OPT_FUSE_TRACE) */
} pc_kernel_class;

/* scene.Scene flat buffers, reference asset/scene/optimized_scene.go:167-190, in the
 * order tracer/opencl/buffers.go:180-191 uploads them. Byte sizes, not counts. */
typedef struct pc_scene_view {
    const void *bvh_nodes;        uint64_t bvh_nodes_bytes;        /* 32 B BvhNode          */
    const void *mesh_instances;   uint64_t mesh_instances_bytes;   /* 80 B MeshInstance     */
    const void *material_nodes;   uint64_t material_nodes_bytes;   /* 64 B MaterialNode     */
    const void *texture_data;     uint64_t texture_data_bytes;     /* raw bytes             */
    const void *texture_metadata; uint64_t texture_metadata_bytes; /* 16 B TextureMetadata  */
    const void *vertices;         uint64_t vertices_bytes;         /* float4 per vertex     */
    const void *normals;          uint64_t normals_bytes;          /* float4 per vertex     */
    const void *uvs;              uint64_t uvs_bytes;              /* float2 per vertex     */
    const void *material_indices; uint64_t material_indices_bytes; /* uint32 per triangle   */
    const void *emissives;        uint64_t emissives_bytes;        /* 80 B EmissivePrimitive*/
    int32_t scene_diffuse_mat_index;    /* -1 when the scene has no background material */
    int32_t scene_emissive_mat_index;
} pc_scene_view;

/* pc_read_buffer selectors: the reference's bufferSet (tracer/opencl/buffers.go:21-70). */
typedef enum pc_buffer {
    PC_BUF_RAYS0 = 0, PC_BUF_RAYS1 = 1, PC_BUF_RAYS2 = 2,   /* 32 B Ray               */
    PC_BUF_PATHS = 3,                                       /* 32 B Path              */
    PC_BUF_HIT_FLAGS = 4,                                   /* uint32                 */
    PC_BUF_INTERSECTIONS = 5,                               /* 32 B Intersection      */
    PC_BUF_EMISSIVE_SAMPLES = 6,                            /* float3 in 16 B         */
    PC_BUF_TRACE_ACCUMULATOR = 7,                           /* float3 in 16 B / pixel */
    PC_BUF_FRAME_ACCUMULATOR = 8,
    PC_BUF_FRAME_BUFFER = 9,                                /* RGBA8                  */
    PC_BUF_RAY_COUNTERS = 10                                /* 3 x int32              */
} pc_buffer;

typedef enum pc_option {
    PC_OPT_COUNTERS = 0,        /* 1: count nodes/tris/instances per ray (slower)         */
    PC_OPT_PRIMARY_PACKETS = 1, /* 1: warp-packet traversal for primary rays (what the reference
                                   does on GPUs); 0 (default): per-ray traversal (what it does on
                                   CPU devices, pipeline.go:107-111) -- same hit records bit for
                                   bit, measured faster on every config                       */
    PC_OPT_REFERENCE_ORDER = 2, /* 1: left-first traversal without closest-hit culling,
                                   i.e. literally intersect.cl:184-347; default 0         */
    PC_OPT_USE_GRAPH = 3,       /* 1 (default): replay one CUDA graph per sample          */
    PC_OPT_FIX_Q4 = 4,          /* 1 (default): emissive hits accumulate at pixelIndex;
                                   0: at the ray's path index like pt_integrator.cl:106   */
    PC_OPT_KERNEL_TIMERS = 5,   /* 1: direct launches bracketed by CUDA events on the handle's
                                   stream, per-class times in pc_stats (measurement mode;
                                   implies one sample chain)                                */
    PC_OPT_SAMPLE_CHAINS = 6,   /* 1..8 (default 4): independent sample chains in flight.  Chain c
                                   traces samples c, c+n, ... with its own ray/path/hit state on its
                                   own stream so the launches of different samples overlap; chains
                                   > 0 accumulate separately and are added in chain order at the
                                   end of pc_trace (deterministic; differs from 1 chain only in
                                   float summation order)                                   */
    PC_OPT_FUSE_TRACE = 7,      /* 1 (default): a bounce's occlusion test (+ emissive accumulation) and the
                                   next bounce's closest-hit query run as ONE persistent launch;
                                   results are bit-identical to 0 (two launches)              */
    PC_OPT_SORT_RAYS = 8,       /* 1 (default): k_shade also writes, per tile, a permutation of the emitted occlusion /
                                   indirect rays sorted by (origin octant of the scene, direction octant,
                                   dominant axis) and the traversal kernels walk the rays in that order;
                                   the rays, their order in the buffers and every result stay bit-identical */
    PC_OPT_DEFER_OCCLUSION = 9, /* 1 (default, needs PC_OPT_FUSE_TRACE): a sample's LAST occlusion test (+ emissive accumulation) runs
                                   inside the next sample's primary-ray launch of the same chain instead of as a launch of its
                                   own (a pure tail); the last sample's is flushed at the end of pc_trace.  Bit-identical.       */
    PC_OPT_SAMPLE_SLOTS = 11,   /* samples one set of launches carries per chain: 1..8, 0 (default) = enough for ~4 M paths per launch
                                   (one sample of a 1 Mpx block does not fill the machine).  The ray buffers stay flat, slot after
                                   slot; every sample keeps its own accumulator, seeds and ray numbering, so results are
                                   bit-identical to tracing the samples one by one (tested).                                */
    PC_OPT_TRACE_REFILL = 10    /* schedule of the fused bounce-traversal kernel: 0 = every warp walks fixed 32-ray units, 1 = a warp
                                   refills the lanes whose ray is finished from the queue and splits box steps from triangle
                                   steps (wins on long incoherent walks: instanced / large scenes), -1 (default) = chosen at
                                   pc_upload_scene by the size of the BVH.  Bit-identical either way.                            */
} pc_option;

/* ---- device discovery: device.GetPlatformInfo (tracer/opencl/device/platform.go) ---- */
int pc_abi_version(void);
int pc_device_count(void);
/* name (<= cap bytes incl. NUL), SM count, max SM clock in MHz, and the reference's speed
 * estimate computeUnits*clockMHz/1000 (tracer/opencl/device/device.go:209-222). */
int pc_device_info(int ordinal, char *name, size_t cap, uint32_t *sm_count,
                   uint32_t *clock_mhz, uint32_t *speed);

/* ---- lifecycle: opencl.NewTracer + Tracer.Init (tracer.go:58-117) / Tracer.Close (:120-142) ---- */
int  pc_create(int ordinal, const char *id, pc_tracer **out);
void pc_destroy(pc_tracer *tr);                 /* idempotent on NULL */
const char *pc_id(const pc_tracer *tr);         /* Tracer.Id    (tracer.go:75) */
uint32_t    pc_flags(const pc_tracer *tr);      /* Tracer.Flags (tracer.go:80): Local = 1 */
uint32_t    pc_speed(const pc_tracer *tr);      /* Tracer.Speed (tracer.go:90) */
const char *pc_last_error(const pc_tracer *tr); /* NULL handle: creation error of this thread */

/* ---- Tracer.UpdateState(Synchronous, ...) payloads (tracer.go:150-191) ---- */
/* FrameDimensions -> bufferSet.Resize (buffers.go:127-175) */
int pc_resize(pc_tracer *tr, uint32_t frame_w, uint32_t frame_h);
/* SceneData -> bufferSet.UploadSceneData (buffers.go:177-201); also derives the
 * 16-byte-aligned traversal layout and checks the BVH depth against the stack. */
int pc_upload_scene(pc_tracer *tr, const pc_scene_view *scene);
/* CameraData: camera.Position and camera.Frustrum (tracer.go:176-179);
 * frustum rows are TL, TR, BL, BR as xyzw (asset/scene/camera.go:121-141). */
int pc_set_camera(pc_tracer *tr, const float eye[3], const float frustum[16]);
int pc_set_option(pc_tracer *tr, int option, int value);

/* ---- Tracer.Trace (tracer.go:194-247) ----
 * Traces rows [block_y, block_y+block_h) x frame_w, samples_per_pixel samples.
 * seeds: (1 + num_bounces) per sample in the order the reference draws them from
 * math/rand: the camera seed (tracer.go:222) then one shadeHits seed per bounce
 * (pipeline.go:146).  seeds == NULL: the library draws them from its own generator.
 * Like the reference, req->seed and req->accumulated_samples are updated in place.
 * stats may be NULL. The call returns after the device finished the block. */
int pc_trace(pc_tracer *tr, pc_block_request *req, const uint32_t *seeds,
             size_t n_seeds, pc_stats *stats);
/* Tracer.Stats (tracer.go:145) for the last pc_trace. */
int pc_get_stats(pc_tracer *tr, pc_stats *stats);
/* PC_OPT_KERNEL_TIMERS: the CUDA-event time of EVERY launch of the last pc_trace, in launch order (classes[i] is a
 * pc_kernel_class, us[i] microseconds); *n receives the number of launches, at most cap entries are copied.  What a
 * median per kernel class is computed from (the per-class sums of pc_stats include the cold first sample). */
int pc_get_kernel_timings(pc_tracer *tr, uint32_t *classes, float *us, uint32_t cap, uint32_t *n);

/* ---- Tracer.MergeOutput (tracer.go:279-286): dst.frameAccumulator[rows] += src.traceAccumulator[rows].
 * dst and src may live on different GPUs (peer loads over NVLink). Returns without
 * waiting for completion, like Exec1DNoWait (resources.go:119); safe to call
 * concurrently for one dst. */
int pc_merge_output(pc_tracer *dst, pc_tracer *src, const pc_block_request *req);
/* Same merge when the source rows arrive as a host or device buffer produced by another
 * process (one process per GPU, rows gathered with NCCL): rows points at
 * block_w*block_h float4 values for the block of req. is_device: 0 host, 1 device ptr. */
int pc_merge_rows(pc_tracer *dst, const void *rows, int is_device, const pc_block_request *req);
/* Device pointer + byte length of the block rows of this tracer's trace accumulator
 * (what a gather sends). */
int pc_trace_rows(pc_tracer *tr, const pc_block_request *req, void **device_ptr, uint64_t *bytes);

/* ---- one process per GPU (torchrun / MPI style launch): the reference's tracers share ONE OpenCL context so that the
 * primary's aggregateAccumulator can take a peer device's buffer as its argument (tracer/opencl/device/context.go:11-28,
 * renderer/default.go:227, resources.go:108-124).  Across processes the same reach is a CUDA IPC mapping:
 *   worker : pc_ipc_export(slot) once per slot -> 64-byte handle, shipped to the primary's process by any means;
 *            after each pc_trace, pc_ipc_publish_rows(req, slot) copies the block's rows of the trace accumulator into
 *            export buffer `slot` (same frame-pixel offsets) and returns when they are visible;
 *   primary: pc_ipc_open(handle) once -> a device pointer valid in this process (peer access over NVLink);
 *            pc_merge_rows(primary, (char*)ptr + 16*frame_w*block_y, 1, req) adds the rows with peer loads.
 * Two slots so that a worker can publish pass i+1 while the primary still reads pass i; the caller orders
 * publish -> merge -> next publish of the same slot (bench.py does it with the scheduler's stats all-gather). */
#define PC_IPC_HANDLE_BYTES 64
int pc_ipc_export(pc_tracer *tr, int slot, void *handle64);
int pc_ipc_publish_rows(pc_tracer *tr, const pc_block_request *req, int slot);
int pc_ipc_open(pc_tracer *dst, const void *handle64, void **peer_ptr);
int pc_ipc_close(pc_tracer *dst, void *peer_ptr);

/* Device.WaitForKernels (what SyncFramebuffer starts with, tracer.go:259): returns when every launch enqueued on this
 * handle, merges included, has completed. */
int pc_wait_for_kernels(pc_tracer *tr);

/* ---- Tracer.SyncFramebuffer (tracer.go:250-276): wait, tonemap (hdr.cl:5-28) rows
 * [0, block_h) and optionally copy the RGBA8 frame (frame_w*frame_h*4 bytes) to rgba_out
 * (what SaveFrameBuffer / CopyFrameBufferToOpenGLTexture read, pipeline.go:216-256). */
int pc_sync_framebuffer(pc_tracer *tr, const pc_block_request *req, uint8_t *rgba_out);

/* ---- debug pipeline stages: MonteCarloIntegrator(debugFlags) (pipeline.go:17-30,113-200) + kernels/debug.cl ----
 * opencl.DebugFlag values (1 << iota, iota starting at 1).  PC_DEBUG_FRAMEBUFFER is the reference's
 * SaveFrameBuffer post-process stage (pipeline.go:61-63,216-234): pass rgba_out to pc_sync_framebuffer. */
typedef enum pc_debug_flag {
    PC_DEBUG_PRIMARY_DEPTH = 2, PC_DEBUG_PRIMARY_NORMALS = 4, PC_DEBUG_ALL_EMISSIVE = 8, PC_DEBUG_VISIBLE_EMISSIVE = 16,
    PC_DEBUG_OCCLUDED_EMISSIVE = 32, PC_DEBUG_THROUGHPUT = 64, PC_DEBUG_ACCUMULATOR = 128, PC_DEBUG_FRAMEBUFFER = 256
} pc_debug_flag;
typedef struct pc_debug_frame { uint32_t flag, bounce; } pc_debug_frame;
/* Number of frames pc_trace_debug produces: depth + normals once, the other five once per bounce. */
uint32_t pc_debug_frame_count(uint32_t debug_flags, uint32_t num_bounces);
/* pc_trace with the debug stages switched on.  Every stage renders into the RGBA8 debug buffer
 * (frame_w*frame_h*4 bytes, cleared first, resources.go:362-375) and the reference dumps it to
 * debug-<stage>[-<bounce>].png, overwriting the file every sample: frames_out receives the dumps of the
 * LAST sample back to back, in the order the reference writes them, infos[i] naming stage and bounce of
 * frame i (the Go shim encodes the PNGs).  Direct launches, one sample chain, nothing fused: a debugging
 * aid, not a fast path.  Like the reference's normals stage, matSelectNode runs on the real paths and may
 * set their dispersion bits (debug.cl:94).  PC_ERR_INVALID_ARGUMENT when frames_out / infos are too small. */
int pc_trace_debug(pc_tracer *tr, pc_block_request *req, const uint32_t *seeds, size_t n_seeds,
                   uint32_t debug_flags, uint8_t *frames_out, uint64_t frames_cap_bytes,
                   pc_debug_frame *infos, uint32_t infos_cap, uint32_t *n_frames, pc_stats *stats);

/* ---- scene compilation on the device: the geometry half of asset/compiler/compiler.go:81-231 --
 * top-level BVH over the instances, one BVH per mesh with the SAH sweep of asset/compiler/bvh/bvh_builder.go:124-224,
 * triangles re-ordered into leaf order, instance and emissive records -- with the BVH build running as level-synchronous
 * CUDA kernels (polaris_b200/csrc/pc_bvh_build.cu).  Output: the reference's flat buffers, byte-identical to the host
 * build of libpolaris_scene.so (same split decisions, ties in iteration order), ready for pc_upload_scene.
 * The handle owns the buffers until pc_compiled_free. */
typedef struct pc_raw_mesh {      /* one triangle soup as the wavefront reader leaves it                             */
    const float *vertices;        /* ntris * 9                                                                         */
    const float *normals;         /* ntris * 9                                                                         */
    const float *uvs;             /* ntris * 6                                                                         */
    const int32_t *material;      /* ntris: index into mat_root / mat_emissive                                         */
    uint32_t ntris;
} pc_raw_mesh;
typedef struct pc_raw_instance {  /* asset/scene/reader/wavefront.go:505-523                                           */
    uint32_t mesh_index;
    float inv_transform[16];      /* what compiler.go:191 stores                                                       */
    float bbox_min[3], bbox_max[3], center[3];
} pc_raw_instance;
typedef enum pc_compiled_buffer {
    PC_CB_BVH_NODES = 0, PC_CB_MESH_INSTANCES = 1, PC_CB_EMISSIVES = 2, PC_CB_VERTICES = 3, PC_CB_NORMALS = 4, PC_CB_UVS = 5,
    PC_CB_MATERIAL_INDEX = 6
} pc_compiled_buffer;
void *pc_compile_geometry(int ordinal, const pc_raw_mesh *meshes, uint32_t n_meshes, const pc_raw_instance *insts,
                          uint32_t n_insts, const int32_t *mat_root, const int32_t *mat_emissive, uint32_t n_materials,
                          int32_t env_emissive_node);
/* bvh.Build alone (bvh_builder.go:124-224) over n volumes given as n x 3 float arrays; out_order receives the item order of
 * the leaves (leaf.ldata = -(first index into out_order), rdata = count). */
void *pc_build_bvh(int ordinal, const float *bmin, const float *bmax, const float *center, uint32_t n, int min_leaf_items,
                   uint32_t *out_order);
const char *pc_compiled_error(void *compiled);                     /* NULL when the build succeeded */
int pc_compiled_get(void *compiled, int which, const void **ptr, uint64_t *bytes);
void pc_compiled_depths(void *compiled, int *top_depth, int *mesh_depth);
/* seconds: [0] triangle bounds, [1] BVH builds (wall), [2] pre-order flatten, [3] leaf-order gather, [4] total,
 * [5] the device builder's own share of [1] (uploads + kernels + read-backs); out8 has room for 8 doubles */
void pc_compiled_timing(void *compiled, double *out8);
void pc_compiled_free(void *compiled);

/* ---- pinned host memory: page-lock (and later release) a caller-owned buffer -- scene arrays before pc_upload_scene, the
 * RGBA8 frame handed to pc_sync_framebuffer -- so the copies run at full host-link rate.  Optional; the library never
 * retains host pointers either way.  Registering a range twice is not an error. */
int pc_host_register(void *ptr, uint64_t bytes);
int pc_host_unregister(void *ptr);

/* ---- test / oracle hooks ---- */
int pc_read_buffer(pc_tracer *tr, int which, void *dst, uint64_t bytes);
/* Upload n rays (32 B each) into rays0 and run one intersection kernel on them:
 * mode 0 = rayIntersectionQuery, 1 = rayIntersectionTest, 2 = packet query.
 * out_flags: n uint32; out_hits: n 32-B Intersection records (ignored for mode 1). */
int pc_debug_intersect(pc_tracer *tr, const void *rays, uint32_t n, int mode,
                       uint32_t *out_flags, void *out_hits);
/* Evaluate the device BxDF functions (bxdf.cl:29-105) for n records; see
 * polaris_b200/_lib.py for the 64-byte input / 48-byte output record layouts. */
int pc_debug_bxdf(pc_tracer *tr, const void *in_records, uint32_t n, void *out_records);
/* Device RNG (random_sampler.cl:7-16): n states (uint2) -> draws x n x float2 + final states. */
int pc_debug_rng(pc_tracer *tr, uint32_t *states_inout, uint32_t n, uint32_t draws, float *out);
/* Device tonemap (hdr.cl:5-28) of n float4 accumulators -> n RGBA8. */
int pc_debug_tonemap(pc_tracer *tr, const float *acc, uint32_t n, float sample_weight,
                     float exposure, uint8_t *rgba_out);

#ifdef __cplusplus
}
#endif
#endif /* POLARIS_CUDA_H */
