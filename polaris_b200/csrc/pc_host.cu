// pc_host.cu -- libpolaris_cuda.so: the C ABI of include/polaris_cuda.h over the CUDA kernels.
//
// This is the host half of the reference's tracer/opencl package re-expressed for CUDA:
//   tracer.go     (state machine, Trace / MergeOutput / SyncFramebuffer)   -> pc_trace & co.
//   pipeline.go   (MonteCarloIntegrator bounce loop)                       -> record_sample()
//   resources.go  (one method per kernel launch, argument wiring)          -> the launch_* helpers
//   buffers.go    (bufferSet: Resize / UploadSceneData)                    -> pc_resize / pc_upload_scene
//   device/*.go   (OpenCL wrapper)                                         -> CUDA runtime, one stream per handle
// Differences that matter: no host synchronisation inside a sample (ray counts stay on the device),
// one CUDA graph replayed per sample, seeds are an explicit input.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/polaris_cuda.h"
#include "pc_kernels.cuh"

#ifndef PC_DEFAULT_SORT_RAYS
#define PC_DEFAULT_SORT_RAYS 1
#endif
#ifndef PC_DEFAULT_FUSE_TRACE
#define PC_DEFAULT_FUSE_TRACE 1
#endif
#ifndef PC_DEFAULT_DEFER_OCC
#define PC_DEFAULT_DEFER_OCC 1
#endif
#include "pc_layout.hpp"

using namespace pc;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;  // logical size (what pc_read_buffer may return)
    size_t cap = 0;    // allocated size
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cap = 0;
    }
    cudaError_t alloc(size_t n) {
        if (n == 0) n = 16;  // keep pointers non-null for empty tables
        if (p && n <= cap && n * 2 >= cap) {  // re-uploading a scene of the same size: no cudaFree / cudaMalloc round trip
            bytes = n;
            return cudaSuccess;
        }
        release();
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = cap = n;
        return e;
    }
};

// What the captured per-sample graph depends on.  Block position / size and the camera are NOT part of
// it: they reach the kernels through the TraceParams device struct, so the perfect scheduler moving row
// boundaries every frame, or an interactive camera, never forces a re-instantiation.
struct GraphKey {
    uint32_t nb = 0, rr = 0;
    int counters = 0, packets = 0, reforder = 0, fixq4 = 0, chains = 0, slots = 0, fuse = 0, sort = 0, defer = 0, refill = 0;
    DScene scene{};  // the captured launches carry the scene's device pointers and scalars BY VALUE: same struct, same graph
    const void *seedsPtr = nullptr;
    bool operator==(const GraphKey &o) const {
        return nb == o.nb && rr == o.rr && counters == o.counters && packets == o.packets && reforder == o.reforder &&
               fixq4 == o.fixq4 && chains == o.chains && slots == o.slots && fuse == o.fuse && sort == o.sort && defer == o.defer && refill == o.refill && memcmp(&scene, &o.scene, sizeof(DScene)) == 0 && seedsPtr == o.seedsPtr;
    }
};

// An event on the destination's frame stream marking the end of a merge that reads a source tracer's trace
// accumulator; the source's next Trace waits for it before clearing that accumulator.  Shared: either handle may die first.
struct ReadFence {
    cudaEvent_t ev = nullptr;
    ~ReadFence() { if (ev) cudaEventDestroy(ev); }
};

constexpr int MAX_CHAINS = 8;
constexpr int GRAPH_SAMPLES_PER_CHAIN = 4;  // samples each chain and slot contributes to one replay of the captured graph (batches per chain)
struct Chain {
    DevBuf rays[3], paths, hitFlags, hits, emSamples, ctl, status, permOcc, permInd;
    DevBuf accs[MAX_SLOTS];  // one accumulator per sample slot (chain 0 / slot 0 is the handle's trace accumulator)
    FrameBufs fb{};
    cudaStream_t stream = nullptr;  // chain 0 uses the handle's stream
    cudaEvent_t evJoin = nullptr;
    // rows [dirtyY0, dirtyY1) of this chain's accumulator (chain 0: the trace accumulator) may be non-zero: the next
    // pc_trace clears those and its own block instead of the whole frame (the reference clears W*H every Trace,
    // tracer.go:215; the observable state -- zero outside the traced block -- is the same)
    uint32_t dirtyY0[MAX_SLOTS] = {}, dirtyY1[MAX_SLOTS] = {};
};

}  // namespace

struct pc_tracer {
    std::mutex mu;
    int device = 0;
    std::string id, err;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    bool dead = false;
    // scene
    DevBuf bvh, inst, mats, texData, texMeta, verts, normals, uvs, matIdx, emissives, node64, tri48, inst80, node128;
    DScene sc{};
    bool hasScene = false;
    int stackNeed = 0;
    uint64_t sceneEpoch = 0, cameraEpoch = 0;
    // frame
    uint32_t W = 0, H = 0;
    DevBuf traceAcc, frameAcc, frameBuf, seedsDev, scratch, params, debugBuf;
    // Sample chains.  The samples of a block request are independent given their seeds, and one sample's
    // launches (a few hundred thousand rays each at the BASELINE sizes) leave the machine idle in their
    // tails, so chain c traces samples c, c+nChains, ... with its own ray / path / hit state on its own
    // stream, and the chains' launches overlap.  Chain 0 accumulates into the trace accumulator, every
    // other chain into its own, added to it once at the end of pc_trace in chain order (deterministic).
    Chain chain[MAX_CHAINS];
    size_t rayCap = 0;      // rays PER SAMPLE SLOT the per-chain ray / path / hit state holds: sized by the largest BLOCK
                            // traced so far (+ slack), not by the frame
    int raySlots = 0;       // sample slots the per-chain state holds (rayCap * raySlots rays per buffer)
    int lastSlot = 0;       // slot of the last sample of the last pc_trace (pc_read_buffer)
    uint32_t lastSlotPaths = 0, lastBounces = 1;
    int nChains = 0;        // chains with allocated state
    int lastChain = 0;      // chain that traced the last sample of the last pc_trace (pc_read_buffer)
    size_t statusStride = 0;  // words per bounce
    CameraParams cam{};
    bool hasCamera = false;
    // options
    int optCounters = 0, optPackets = 0, optRefOrder = 0, optGraph = 1, optFixQ4 = 1, optTimers = 0, optChains = 4, optFuse = PC_DEFAULT_FUSE_TRACE, optSort = PC_DEFAULT_SORT_RAYS, optDeferOcc = PC_DEFAULT_DEFER_OCC, optRefill = -1, optSlots = 0;
    bool refillNow = false;   // what optRefill resolves to for the uploaded scene
    size_t innerNodes = 0;
    int occGrid = 0;
    cudaEvent_t evFork = nullptr;
    std::vector<cudaEvent_t> timerEvents;  // pairs, PC_OPT_KERNEL_TIMERS
    std::vector<int> timerClass;
    std::vector<float> timerUs;  // per launch, same order as timerClass (pc_get_kernel_timings)
    // graph cache
    cudaGraphExec_t graphExec = nullptr;
    GraphKey graphKey;
    uint64_t launchesPerSample = 0;
    // frame state (DESIGN.md "frame accumulator reset").  Everything that touches the FRAME accumulator (its
    // one clear per frame, the merges, the tonemap) runs on frameStream under frameMu, so that other tracers' MergeOutput
    // calls neither wait for this tracer's own Trace (the reference enqueues them without waiting, resources.go:119) nor
    // serialise behind it on the device.
    std::mutex frameMu;
    cudaStream_t frameStream = nullptr;
    bool frameOpen = false;           // the frame accumulator was cleared for the frame in progress
    bool ownFirstPassTraced = false;  // this tracer's own first-pass Trace of the open frame has been issued
    std::vector<uint8_t> firstPassRows;  // rows already merged in the open frame's first pass (a repeat means: a new frame)
    // multi-process exchange (pc_ipc_*): two frame-sized export buffers other processes map through CUDA IPC
    DevBuf exportBuf[2];
    std::vector<void *> ipcOpened;
    std::mutex readMu;
    std::vector<std::shared_ptr<ReadFence>> pendingReads;  // merges (possibly on other devices) still reading traceAcc
    std::vector<std::shared_ptr<ReadFence>> fencePool;     // this tracer's fences as a merge destination, recycled
    size_t fenceNext = 0;
    pc_stats stats{};
    uint64_t seedState = 0x501A2150ull;
    int persistentGrid = 0, shadeGrid = 0;
};

namespace {

int fail(pc_tracer *tr, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (tr) tr->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(tr, code, call)                                                                     \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            if ((code) == PC_ERR_KERNEL) (tr)->dead = true;                                    \
            return fail((tr), (code), "%s: %s", #call, cudaGetErrorString(e_));                \
        }                                                                                      \
    } while (0)

int enter(pc_tracer *tr) {
    if (!tr) return PC_ERR_INVALID_ARGUMENT;
    if (tr->dead) return fail(tr, PC_ERR_KERNEL, "tracer is dead after a sticky CUDA error: %s", tr->err.c_str());
    cudaError_t e = cudaSetDevice(tr->device);
    if (e != cudaSuccess) return fail(tr, PC_ERR_NO_DEVICE, "cudaSetDevice(%d): %s", tr->device, cudaGetErrorString(e));
    return 0;
}

int grid_for(size_t n, int block, int cap) {
    size_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

uint32_t next_seed(uint64_t &s) {  // splitmix64
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)z;
}

void drop_graph(pc_tracer *tr) {
    if (tr->graphExec) cudaGraphExecDestroy(tr->graphExec);
    tr->graphExec = nullptr;
}

// ---- one sample of Tracer.Trace: PrimaryRayGenerator + MonteCarloIntegrator (pipeline.go:94-213)
// enqueued on the handle's stream; every launch below is also a node of the per-sample graph.
// PC_OPT_KERNEL_TIMERS: bracket a launch with two events on the launching stream
struct LaunchTimer {
    pc_tracer *tr;
    size_t slot = 0;
    LaunchTimer(pc_tracer *t, int cls) : tr(t) {
        if (!tr->optTimers) return;
        slot = tr->timerClass.size();
        while (tr->timerEvents.size() < 2 * (slot + 1)) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            tr->timerEvents.push_back(e);
        }
        tr->timerClass.push_back(cls);
        cudaEventRecord(tr->timerEvents[2 * slot], tr->stream);
    }
    ~LaunchTimer() {
        if (tr->optTimers) cudaEventRecord(tr->timerEvents[2 * slot + 1], tr->stream);
    }
};

// ---- debug stages (pipeline.go:113-200, resources.go:362-520): pc_trace_debug ----
struct DebugSink {
    uint32_t flags = 0;
    uint8_t *frames = nullptr;
    uint64_t cap = 0;
    pc_debug_frame *infos = nullptr;
    uint32_t infosCap = 0, n = 0;
    bool capture = false;  // the reference overwrites its PNG files every sample: only the last sample's dumps survive
    int overflow = 0;
    uint64_t launches = 0;
    uint32_t sampleIndex = 0;  // tracer.go:240 bumps AccumulatedSamples after every sample
};

// dumpDebugBuffer (pipeline.go:259-277): the debug buffer -> the caller's next frame slot, ordered on the stream
static void debug_dump(pc_tracer *tr, cudaStream_t s, DebugSink &d, uint32_t flag, uint32_t bounce) {
    if (!d.capture) return;
    const uint64_t bytes = (uint64_t)tr->W * tr->H * 4;
    if ((uint64_t)(d.n + 1) * bytes > d.cap || d.n >= d.infosCap) { d.overflow = 1; return; }
    cudaMemcpyAsync(d.frames + (uint64_t)d.n * bytes, tr->debugBuf.p, bytes, cudaMemcpyDeviceToHost, s);
    d.infos[d.n].flag = flag;
    d.infos[d.n].bounce = bounce;
    d.n++;
}

static void debug_stage(pc_tracer *tr, Chain &ch, const pc_block_request &req, DebugSink &d, uint32_t flag, uint32_t bounce, int a) {
    cudaStream_t s = ch.stream;
    const uint32_t px = tr->W * tr->H, n = req.frame_w * req.block_h;
    const FrameBufs &fb = ch.fb;
    const TraceCtl *ctl = (const TraceCtl *)ch.ctl.p;
    uchar4 *out = (uchar4 *)tr->debugBuf.p;
    uint32_t *maxBits = (uint32_t *)((char *)tr->debugBuf.p + (size_t)px * 4);
    const int B = 256, G = (int)((n + B - 1) / B);
    k_debug_clear<<<(px + B - 1) / B, B, 0, s>>>(out, px);  // DebugClearBuffer
    d.launches += 2;
    switch (flag) {
        case PC_DEBUG_PRIMARY_DEPTH: {
            const uint32_t one = 0x3F800000u;  // maxDepth starts at 1.0 (resources.go:397)
            cudaMemcpyAsync(maxBits, &one, 4, cudaMemcpyHostToDevice, s);
            k_debug_max_depth<<<(px + B - 1) / B, B, 0, s>>>(fb.hits, px, maxBits);
            k_debug_depth<<<G, B, 0, s>>>(ctl, a, fb.paths, fb.hitFlags, fb.hits, maxBits, out, n);
            d.launches++;
            break;
        }
        case PC_DEBUG_PRIMARY_NORMALS:
            k_debug_normals<<<G, B, 0, s>>>(tr->sc, ctl, a, fb.rays[a], fb.paths, fb.hitFlags, fb.hits, out, n);
            break;
        case PC_DEBUG_ALL_EMISSIVE: case PC_DEBUG_VISIBLE_EMISSIVE: case PC_DEBUG_OCCLUDED_EMISSIVE:
            k_debug_emissive<<<G, B, 0, s>>>(ctl, fb.rays[2], fb.paths, fb.hitFlags, fb.emissiveSamples,
                                            flag == PC_DEBUG_VISIBLE_EMISSIVE ? 1u : 0u, flag == PC_DEBUG_OCCLUDED_EMISSIVE ? 1u : 0u, out, n);
            break;
        case PC_DEBUG_THROUGHPUT:
            k_debug_throughput<<<G, B, 0, s>>>(fb.paths, out, n);
            break;
        case PC_DEBUG_ACCUMULATOR: {
            const float sampleWeight = 1.0f / (float)(req.accumulated_samples + d.sampleIndex + req.samples_per_pixel);  // resources.go:509
            k_debug_accumulator<<<G, B, 0, s>>>(sampleWeight, fb.traceAcc, out, n);
            break;
        }
    }
    debug_dump(tr, s, d, flag, bounce);
}

// Whether a sample's last occlusion launch is deferred into the next sample's k_primary (see k_primary)
static bool defer_last_occlusion(const pc_tracer *tr, bool dbg) {
    return tr->optFuse && tr->optDeferOcc && !tr->optRefOrder && !tr->optPackets && !tr->optTimers && !dbg;
}
static int last_occlusion_slot(uint32_t nb) { return 1 + 2 * ((int)nb - 1); }  // == the slot the un-deferred launch uses

template <bool COUNT>
void record_sample_t(pc_tracer *tr, Chain &ch, const pc_block_request &req, uint32_t sampleAdvance, uint32_t nSlots, uint32_t slotStride,
                     uint64_t *launches, DebugSink *dbg) {
    cudaStream_t s = ch.stream;
    TraceCtl *ctl = (TraceCtl *)ch.ctl.p;
    const uint32_t *seeds = (const uint32_t *)tr->seedsDev.p;
    unsigned long long *status = (unsigned long long *)ch.status.p;
    const uint32_t nb = req.num_bounces, perSample = 1 + nb;
    const int pg = tr->persistentGrid;
    const int shadeGrid = tr->shadeGrid;
    const TraceParams *params = (const TraceParams *)tr->params.p;
    const FrameBufs &fb = ch.fb;
    uint64_t L = 0;
    const bool defer = defer_last_occlusion(tr, dbg != nullptr);
    const int sortedOcc = tr->optSort && !tr->optRefOrder ? 1 : 0;
    size_t statusWords = tr->statusStride * nb;
    {
        LaunchTimer lt(tr, PC_K_BEGIN_SAMPLE);
        k_begin_sample<<<grid_for(statusWords, 256, 1024), 256, 0, s>>>(ctl, status, statusWords, sampleAdvance, nSlots, slotStride);
    }
    L++;
    int slot = 0;
    // generatePrimaryRays + RayPacketIntersectionQuery / RayIntersectionQuery (pipeline.go:107-111)
    {
    LaunchTimer lt(tr, PC_K_PRIMARY);
    if (tr->optRefOrder)
        k_primary<2, COUNT><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb, ctl, seeds, params, perSample, slot, -1, 0);
    else if (tr->optPackets)
        k_primary<1, COUNT><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb, ctl, seeds, params, perSample, slot, -1, 0);
    else
        k_primary<0, COUNT><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb, ctl, seeds, params, perSample, slot, defer ? last_occlusion_slot(nb) : -1, sortedOcc);
    }
    L++;
    slot++;
    int a = 0;
    const int sorted = tr->optSort && !tr->optRefOrder ? 1 : 0;  // traversal order of the bounce rays (k_shade writes the permutations)
    const uint32_t df = dbg ? dbg->flags : 0u;
    if (df & PC_DEBUG_PRIMARY_DEPTH) debug_stage(tr, ch, req, *dbg, PC_DEBUG_PRIMARY_DEPTH, 0, a);      // pipeline.go:113-119
    if (df & PC_DEBUG_PRIMARY_NORMALS) debug_stage(tr, ch, req, *dbg, PC_DEBUG_PRIMARY_NORMALS, 0, a);  // :120-126
    for (uint32_t bounce = 0; bounce < nb; bounce++) {
        // ShadePrimaryRayMisses / ShadeIndirectRayMisses + ShadeHits (pipeline.go:134-146)
        {
            LaunchTimer lt(tr, PC_K_SHADE);
            shade_launch(COUNT, shadeGrid, s, tr->sc, fb, ctl, seeds, status + (size_t)bounce * tr->statusStride, perSample, bounce,
                         req.min_bounces_for_rr, a, tr->optFixQ4, sorted);
        }
        L++;
        if (df & PC_DEBUG_THROUGHPUT) debug_stage(tr, ch, req, *dbg, PC_DEBUG_THROUGHPUT, bounce, a);  // :151-157
        if (tr->optFuse && !tr->optRefOrder && bounce + 1 < nb && !dbg) {
            // RayIntersectionTest(2) + AccumulateEmissiveSamples(2) and the next bounce's RayIntersectionQuery
            // (pipeline.go:160-165, :203-209) are independent: one persistent launch covers both
            a = 1 - a;
            LaunchTimer lt(tr, PC_K_TRACE);
            if (tr->refillNow) k_trace<COUNT, true><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb, ctl, a, slot, sorted);
            else k_trace<COUNT, false><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb, ctl, a, slot, sorted);
            L++;
            slot += 2;
            continue;
        }
        if (defer && bounce + 1 == nb) break;  // the last occlusion test rides in the next sample's k_primary (or the final flush)
        // RayIntersectionTest(2) + AccumulateEmissiveSamples(2) (pipeline.go:160-165)
        {
        LaunchTimer lt(tr, PC_K_OCCLUSION);
        if (tr->optRefOrder)
            k_occlusion<true, COUNT><<<tr->occGrid, TRAV_BLOCK, 0, s>>>(tr->sc, fb.rays[2], fb.paths, fb.emissiveSamples, fb, 1, dbg ? fb.hitFlags : nullptr, ctl, slot, sorted && !dbg ? fb.permOcc : nullptr);
        else
            k_occlusion<false, COUNT><<<tr->occGrid, TRAV_BLOCK, 0, s>>>(tr->sc, fb.rays[2], fb.paths, fb.emissiveSamples, fb, 1, dbg ? fb.hitFlags : nullptr, ctl, slot, sorted && !dbg ? fb.permOcc : nullptr);
        }
        L++;
        slot++;
        if (df & PC_DEBUG_ALL_EMISSIVE) debug_stage(tr, ch, req, *dbg, PC_DEBUG_ALL_EMISSIVE, bounce, a);            // :170-176
        if (df & PC_DEBUG_VISIBLE_EMISSIVE) debug_stage(tr, ch, req, *dbg, PC_DEBUG_VISIBLE_EMISSIVE, bounce, a);    // :178-184
        if (df & PC_DEBUG_OCCLUDED_EMISSIVE) debug_stage(tr, ch, req, *dbg, PC_DEBUG_OCCLUDED_EMISSIVE, bounce, a);  // :186-192
        if (df & PC_DEBUG_ACCUMULATOR) debug_stage(tr, ch, req, *dbg, PC_DEBUG_ACCUMULATOR, bounce, a);              // :194-200
        if (bounce + 1 < nb) {  // pipeline.go:203-209
            a = 1 - a;
            LaunchTimer lt(tr, PC_K_QUERY);
            if (tr->optRefOrder)
                k_query<true, COUNT><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb.rays[a], fb.hitFlags, fb.hits, ctl, a, slot, nullptr);
            else
                k_query<false, COUNT><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, fb.rays[a], fb.hitFlags, fb.hits, ctl, a, slot, sorted ? fb.permInd : nullptr);
            L++;
            slot++;
        }
    }
    *launches = L;
}

// One BATCH of a chain: nSlots samples (curSample, curSample + slotStride, ...) through one set of launches.
void record_sample(pc_tracer *tr, Chain &ch, const pc_block_request &req, uint32_t sampleAdvance, uint32_t nSlots, uint32_t slotStride,
                   uint64_t *launches, DebugSink *dbg = nullptr) {
    if (tr->optCounters) record_sample_t<true>(tr, ch, req, sampleAdvance, nSlots, slotStride, launches, dbg);
    else record_sample_t<false>(tr, ch, req, sampleAdvance, nSlots, slotStride, launches, dbg);
}

// Enqueue perChain[c] samples on every chain c < nChains: fork the chain streams off the handle's stream,
// let each chain run its samples back to back, join.  Works identically under stream capture (the fork /
// join events become graph dependencies and the chains become parallel branches of the graph).
// perChain[c] = samples of chain c, traced as batches of up to `slots` samples per set of launches.  *launchTotal
// accumulates the number of launches.
int enqueue_chains(pc_tracer *tr, const pc_block_request &req, int nChains, int slots, const uint32_t *perChain, uint64_t *launchTotal,
                   bool flushOcclusion = false, uint64_t *flushLaunches = nullptr) {
    cudaStream_t s0 = tr->stream;
    if (nChains > 1) {
        if (cudaEventRecord(tr->evFork, s0) != cudaSuccess) return 1;
        for (int c = 1; c < nChains; c++)
            if ((perChain[c] || flushOcclusion) && cudaStreamWaitEvent(tr->chain[c].stream, tr->evFork, 0) != cudaSuccess) return 1;
    }
    for (int c = 0; c < nChains; c++) {
        for (uint32_t k = 0; k < perChain[c]; k += (uint32_t)slots) {
            const uint32_t nSlots = perChain[c] - k < (uint32_t)slots ? perChain[c] - k : (uint32_t)slots;
            uint64_t L = 0;
            record_sample(tr, tr->chain[c], req, (uint32_t)(nChains * slots), nSlots, (uint32_t)nChains, &L);
            if (launchTotal) *launchTotal += L;
        }
        if (flushOcclusion) {  // the chain's last sample left its last bounce's occlusion rays behind (see k_primary)
            Chain &ch = tr->chain[c];
            const int sorted = tr->optSort && !tr->optRefOrder ? 1 : 0;
            // its own queue head: the sample's k_primary already pulled from last_occlusion_slot (for the sample before it)
            const int flushSlot = last_occlusion_slot(req.num_bounces) + 1;
            if (tr->optCounters)
                k_occlusion<false, true><<<tr->occGrid, TRAV_BLOCK, 0, ch.stream>>>(tr->sc, ch.fb.rays[2], ch.fb.paths, ch.fb.emissiveSamples, ch.fb, 1, nullptr, (TraceCtl *)ch.ctl.p, flushSlot, sorted ? ch.fb.permOcc : nullptr);
            else
                k_occlusion<false, false><<<tr->occGrid, TRAV_BLOCK, 0, ch.stream>>>(tr->sc, ch.fb.rays[2], ch.fb.paths, ch.fb.emissiveSamples, ch.fb, 1, nullptr, (TraceCtl *)ch.ctl.p, flushSlot, sorted ? ch.fb.permOcc : nullptr);
            if (flushLaunches) (*flushLaunches)++;
        }
    }
    for (int c = 1; c < nChains; c++) {
        if (!perChain[c] && !flushOcclusion) continue;
        if (cudaEventRecord(tr->chain[c].evJoin, tr->chain[c].stream) != cudaSuccess) return 1;
        if (cudaStreamWaitEvent(s0, tr->chain[c].evJoin, 0) != cudaSuccess) return 1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// (Re)allocate the per-chain state: `want` chains, each able to hold `slots` samples of `needRays` rays (the block being
// traced).  Growing re-allocates every chain (new device addresses: the captured graph is dropped); the capacity only
// grows, with 1/8 slack, so the perfect scheduler nudging row boundaries from pass to pass does not re-allocate.
int ensure_chains(pc_tracer *tr, int want, size_t needRays, int slots = 1) {
    if (want < 1) want = 1;
    if (want > MAX_CHAINS) want = MAX_CHAINS;
    if (slots < 1) slots = 1;
    const size_t px = (size_t)tr->W * tr->H;
    if (needRays > px) needRays = px;
    if (needRays > tr->rayCap || slots > tr->raySlots) {
        CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
        drop_graph(tr);
        for (int c = 0; c < MAX_CHAINS; c++) {
            Chain &ch = tr->chain[c];
            DevBuf *all[] = {&ch.rays[0], &ch.rays[1], &ch.rays[2], &ch.paths, &ch.hitFlags, &ch.hits, &ch.emSamples, &ch.status, &ch.permOcc, &ch.permInd};
            for (DevBuf *b : all) b->release();
        }
        if (needRays > tr->rayCap) {
            size_t cap = needRays + needRays / 8;
            cap = (cap + 1023) / 1024 * 1024;
            tr->rayCap = cap > px ? px : cap;
        }
        if (slots > tr->raySlots) tr->raySlots = slots;
        tr->statusStride = tr->rayCap * tr->raySlots / SHADE_TILE + 2;  // one look-back word per k_shade tile (+ the partial one)
        want = want > tr->nChains ? want : tr->nChains;
        tr->nChains = 0;
    }
    const size_t cap = tr->rayCap * (size_t)tr->raySlots;
    for (int c = 0; c < want; c++) {
        Chain &ch = tr->chain[c];
        // the slots' accumulators: frame indexed like the trace accumulator; only block rows are ever touched
        for (int sl = 0; sl < tr->raySlots; sl++) {
            if (c == 0 && sl == 0) continue;
            if (ch.accs[sl].bytes != px * 16) {
                CU(tr, PC_ERR_ALLOC, ch.accs[sl].alloc(px * 16));
                CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(ch.accs[sl].p, 0, ch.accs[sl].bytes, tr->stream));
                ch.dirtyY0[sl] = ch.dirtyY1[sl] = 0;
            }
        }
        for (int sl = 0; sl < MAX_SLOTS; sl++)
            ch.fb.slotAcc[sl] = (float4 *)(c == 0 && sl == 0 ? tr->traceAcc.p : (sl < tr->raySlots ? ch.accs[sl].p : nullptr));
        ch.fb.traceAcc = ch.fb.slotAcc[0];
    }
    for (int c = tr->nChains; c < want; c++) {
        Chain &ch = tr->chain[c];
        if (c == 0) ch.stream = tr->stream;
        else if (!ch.stream) {
            CU(tr, PC_ERR_ALLOC, cudaStreamCreateWithFlags(&ch.stream, cudaStreamNonBlocking));
            CU(tr, PC_ERR_ALLOC, cudaEventCreateWithFlags(&ch.evJoin, cudaEventDisableTiming));
        }
        for (auto &r : ch.rays) CU(tr, PC_ERR_ALLOC, r.alloc(cap * 32));
        CU(tr, PC_ERR_ALLOC, ch.paths.alloc(cap * 32));
        CU(tr, PC_ERR_ALLOC, ch.hitFlags.alloc(cap * 4));
        CU(tr, PC_ERR_ALLOC, ch.hits.alloc(cap * 32));
        CU(tr, PC_ERR_ALLOC, ch.emSamples.alloc(cap * 16));
        CU(tr, PC_ERR_ALLOC, ch.permOcc.alloc(cap * 4));
        CU(tr, PC_ERR_ALLOC, ch.permInd.alloc(cap * 4));
        CU(tr, PC_ERR_ALLOC, ch.status.alloc(tr->statusStride * MAX_BOUNCES * 8));
        if (!ch.ctl.p) {
            CU(tr, PC_ERR_ALLOC, ch.ctl.alloc(sizeof(TraceCtl)));
            CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(ch.ctl.p, 0, sizeof(TraceCtl), tr->stream));
        }
        for (int i = 0; i < 3; i++) ch.fb.rays[i] = (Ray *)ch.rays[i].p;
        ch.fb.paths = (PathRec *)ch.paths.p;
        ch.fb.hitFlags = (uint32_t *)ch.hitFlags.p;
        ch.fb.hits = (HitRec *)ch.hits.p;
        ch.fb.emissiveSamples = (float4 *)ch.emSamples.p;
        ch.fb.permOcc = (uint32_t *)ch.permOcc.p;
        ch.fb.permInd = (uint32_t *)ch.permInd.p;
        DevBuf *zero[] = {&ch.rays[0], &ch.rays[1], &ch.rays[2], &ch.paths, &ch.hitFlags, &ch.hits, &ch.emSamples, &ch.status};
        for (DevBuf *b : zero) CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(b->p, 0, b->bytes, tr->stream));
        tr->nChains = c + 1;
    }
    return 0;
}

void release_chain_buffers(pc_tracer *tr) {
    for (int c = 0; c < MAX_CHAINS; c++) {
        Chain &ch = tr->chain[c];
        DevBuf *all[] = {&ch.rays[0], &ch.rays[1], &ch.rays[2], &ch.paths, &ch.hitFlags, &ch.hits, &ch.emSamples, &ch.status, &ch.permOcc, &ch.permInd};
        for (DevBuf *b : all) b->release();
        for (int sl = 0; sl < MAX_SLOTS; sl++) { ch.accs[sl].release(); ch.dirtyY0[sl] = ch.dirtyY1[sl] = 0; }
    }
    tr->nChains = 0;
    tr->rayCap = 0;
    tr->raySlots = 0;
}

int upload(pc_tracer *tr, DevBuf &b, const void *src, size_t bytes) {
    CU(tr, PC_ERR_ALLOC, b.alloc(bytes));
    if (bytes) CU(tr, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, tr->stream));
    return 0;
}

// clearAccumulator (accumulator.cl:5-9) over rows [y0, y1) of a frame-indexed accumulator
int clear_rows(pc_tracer *tr, void *acc, uint32_t y0, uint32_t y1, cudaStream_t s) {
    if (y1 <= y0) return 0;
    const size_t n = (size_t)tr->W * (y1 - y0);
    k_clear<<<grid_for(n, 256, tr->prop.multiProcessorCount * 8), 256, 0, s>>>((float4 *)acc + (size_t)tr->W * y0, n);
    CU(tr, PC_ERR_KERNEL, cudaGetLastError());
    return 0;
}

// ClearTraceAccumulator (tracer.go:215) without touching rows that are zero already: clear what earlier traces left
// behind and the block about to be traced, remember the new block as the dirty range.
int clear_for_block(pc_tracer *tr, uint32_t &dirtyY0, uint32_t &dirtyY1, void *acc, uint32_t y0, uint32_t y1, cudaStream_t s) {
    int rc = 0;
    if (dirtyY1 > dirtyY0 && (dirtyY1 < y0 || dirtyY0 > y1)) {  // disjoint: two clears
        if ((rc = clear_rows(tr, acc, dirtyY0, dirtyY1, s))) return rc;
        rc = clear_rows(tr, acc, y0, y1, s);
    } else {
        const uint32_t lo = dirtyY1 > dirtyY0 && dirtyY0 < y0 ? dirtyY0 : y0;
        const uint32_t hi = dirtyY1 > dirtyY0 && dirtyY1 > y1 ? dirtyY1 : y1;
        rc = clear_rows(tr, acc, lo, hi, s);
    }
    dirtyY0 = y0;
    dirtyY1 = y1;
    return rc;
}

bool first_pass(const pc_block_request *r) { return r->accumulated_samples <= r->samples_per_pixel; }

// The frame accumulator's one clear per frame.  The reference clears it inside each tracer's own Trace
// (pipeline.Reset, tracer.go:208-213), racing with the other workers' MergeOutput into the primary; here
// the first first-pass merge or the tracer's own first-pass Trace -- whichever arrives first -- does it.  Caller holds
// frameMu.  A frame normally ends at pc_sync_framebuffer; when one did not (a worker failed, the client gave up) the next
// frame is recognised by a first-pass merge of rows that were merged already, or by a second own first-pass Trace.
int open_frame(pc_tracer *dst, bool newFrame) {
    if (dst->frameOpen && !newFrame) return 0;
    int rc = clear_rows(dst, dst->frameAcc.p, 0, dst->H, dst->frameStream);
    if (rc) return rc;
    dst->frameOpen = true;
    dst->ownFirstPassTraced = false;
    dst->firstPassRows.assign(dst->H, 0);
    return 0;
}
int open_frame_for_merge(pc_tracer *dst, const pc_block_request *req) {
    if (!first_pass(req)) return 0;
    bool repeat = false;
    if (dst->frameOpen && dst->firstPassRows.size() == dst->H)
        for (uint32_t y = req->block_y; y < req->block_y + req->block_h && y < dst->H; y++) repeat = repeat || dst->firstPassRows[y];
    int rc = open_frame(dst, repeat);
    if (rc) return rc;
    for (uint32_t y = req->block_y; y < req->block_y + req->block_h && y < dst->H; y++) dst->firstPassRows[y] = 1;
    return 0;
}
int open_frame_for_trace(pc_tracer *tr) {  // Trace with AccumulatedSamples == 0
    std::lock_guard<std::mutex> g(tr->frameMu);
    int rc = open_frame(tr, tr->ownFirstPassTraced);
    tr->ownFirstPassTraced = true;
    return rc;
}

}  // namespace

extern "C" {

int pc_abi_version(void) { return PC_ABI_VERSION; }

int pc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pc_device_info(int ordinal, char *name, size_t cap, uint32_t *sm_count, uint32_t *clock_mhz, uint32_t *speed) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, ordinal) != cudaSuccess) return PC_ERR_NO_DEVICE;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ordinal);
    if (name && cap) snprintf(name, cap, "%s", p.name);
    uint32_t mhz = (uint32_t)(khz / 1000);
    if (sm_count) *sm_count = (uint32_t)p.multiProcessorCount;
    if (clock_mhz) *clock_mhz = mhz;
    if (speed) *speed = (uint32_t)p.multiProcessorCount * mhz / 1000u;  // device.go:209-222
    return 0;
}

int pc_create(int ordinal, const char *id, pc_tracer **out) {
    if (!out) return PC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = pc_device_count();
    if (ordinal < 0 || ordinal >= n) return fail(nullptr, PC_ERR_NO_DEVICE, "CUDA device %d not available (%d devices)", ordinal, n);
    auto *tr = new pc_tracer();
    tr->device = ordinal;
    tr->id = id ? id : "cuda";
    cudaError_t e = cudaSetDevice(ordinal);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&tr->prop, ordinal);
    if (e == cudaSuccess && tr->prop.major < 10) {
        fail(nullptr, PC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", ordinal, tr->prop.major, tr->prop.minor);
        delete tr;
        return PC_ERR_NO_DEVICE;
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&tr->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&tr->frameStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&tr->evStart);
    if (e == cudaSuccess) e = cudaEventCreate(&tr->evStop);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&tr->evFork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = tr->chain[0].ctl.alloc(sizeof(TraceCtl));
    if (e == cudaSuccess) e = tr->params.alloc(sizeof(TraceParams));
    if (e == cudaSuccess) e = cudaMemsetAsync(tr->chain[0].ctl.p, 0, sizeof(TraceCtl), tr->stream);
    tr->chain[0].stream = tr->stream;
    if (e != cudaSuccess) {
        fail(nullptr, PC_ERR_NO_DEVICE, "pc_create(%d): %s", ordinal, cudaGetErrorString(e));
        delete tr;
        return PC_ERR_NO_DEVICE;
    }
    // Shared-memory carve-out of the traversal kernels: just enough for the resident blocks' stacks, the rest of the SM's
    // 256 KB stays L1 (the scene is re-read by every ray).  Left to the driver the carve-out came out larger: asking for the
    // minimum measured +1.5 % on configs 2 and 3, +0.5 % on config 4 (profiles/ab_r02j.txt; 72 % like k_shade: k_trace 728 -> 763 us;
    // a 5-entry shared stack under a 32 KB carve-out: +0.6 / -0.5 / +0.1 %, ab_r02k.txt).  PC_TRAV_CARVEOUT (percent)
    // overrides, -1 keeps the driver's choice.
    {
        auto carve = [&](auto kernel, int minBlocks) {
#if defined(PC_TRAV_CARVEOUT)
            const int pct = PC_TRAV_CARVEOUT;
#else
            cudaFuncAttributes fa{};
            if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); return; }
            const size_t want = (size_t)minBlocks * (fa.sharedSizeBytes + 1024);
            int pct = (int)((want * 100 + tr->prop.sharedMemPerMultiprocessor - 1) / tr->prop.sharedMemPerMultiprocessor);
            if (pct > 100) pct = 100;
#endif
            if (pct >= 0 && cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess) cudaGetLastError();
        };
        carve(k_primary<0, false>, PC_PRIMARY_MIN_BLOCKS); carve(k_primary<0, true>, PC_PRIMARY_MIN_BLOCKS);
        carve(k_trace<false, false>, PC_TRAV_MIN_BLOCKS);  carve(k_trace<false, true>, PC_TRAV_MIN_BLOCKS);
        carve(k_trace<true, false>, PC_TRAV_MIN_BLOCKS);   carve(k_trace<true, true>, PC_TRAV_MIN_BLOCKS);
        carve(k_occlusion<false, false>, PC_OCC_MIN_BLOCKS); carve(k_occlusion<false, true>, PC_OCC_MIN_BLOCKS);
        carve(k_query<false, false>, PC_TRAV_MIN_BLOCKS);  carve(k_query<false, true>, PC_TRAV_MIN_BLOCKS);
    }
    // persistent grid: every SM full of traversal blocks (multiple of the SM count)
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_query<false, false>, TRAV_BLOCK, 0);
    if (perSM < 1) perSM = 1;
    if (perSM > 16) perSM = 16;
    tr->persistentGrid = tr->prop.multiProcessorCount * perSM;
    int occPerSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occPerSM, k_occlusion<false, false>, TRAV_BLOCK, 0);
    if (occPerSM < 1) occPerSM = 1;
    if (occPerSM > 16) occPerSM = 16;
    tr->occGrid = tr->prop.multiProcessorCount * occPerSM;
    int shadePerSM = 0;
    shade_configure(tr->prop, &shadePerSM);
    if (shadePerSM < 1) shadePerSM = 1;
    tr->shadeGrid = tr->prop.multiProcessorCount * shadePerSM;
    tr->sc.sceneDiffuseMat = -1;
    *out = tr;
    return 0;
}

void pc_destroy(pc_tracer *tr) {
    if (!tr) return;
    {
        std::lock_guard<std::mutex> g(tr->mu);
        cudaSetDevice(tr->device);
        if (tr->stream) cudaStreamSynchronize(tr->stream);
        if (tr->frameStream) cudaStreamSynchronize(tr->frameStream);
        for (void *p : tr->ipcOpened) cudaIpcCloseMemHandle(p);
        tr->ipcOpened.clear();
        for (auto &e : tr->exportBuf) e.release();
        drop_graph(tr);
        DevBuf *all[] = {&tr->bvh, &tr->inst, &tr->mats, &tr->texData, &tr->texMeta, &tr->verts, &tr->normals, &tr->uvs,
                         &tr->matIdx, &tr->emissives, &tr->node64, &tr->tri48, &tr->inst80, &tr->node128, &tr->traceAcc, &tr->frameAcc,
                         &tr->frameBuf, &tr->seedsDev, &tr->scratch, &tr->params, &tr->debugBuf};
        release_chain_buffers(tr);
        for (int c = 0; c < MAX_CHAINS; c++) {
            tr->chain[c].ctl.release();
            if (c > 0 && tr->chain[c].stream) cudaStreamDestroy(tr->chain[c].stream);
            if (tr->chain[c].evJoin) cudaEventDestroy(tr->chain[c].evJoin);
        }
        if (tr->evFork) cudaEventDestroy(tr->evFork);
        for (DevBuf *b : all) b->release();
        for (cudaEvent_t e : tr->timerEvents) cudaEventDestroy(e);
        if (tr->evStart) cudaEventDestroy(tr->evStart);
        if (tr->evStop) cudaEventDestroy(tr->evStop);
        if (tr->stream) cudaStreamDestroy(tr->stream);
        if (tr->frameStream) cudaStreamDestroy(tr->frameStream);
    }
    delete tr;
}

const char *pc_id(const pc_tracer *tr) { return tr ? tr->id.c_str() : ""; }
uint32_t pc_flags(const pc_tracer *) { return 1u; /* tracer.Local */ }
uint32_t pc_speed(const pc_tracer *tr) {
    if (!tr) return 0;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, tr->device);
    return (uint32_t)tr->prop.multiProcessorCount * (uint32_t)(khz / 1000) / 1000u;
}
const char *pc_last_error(const pc_tracer *tr) {
    if (!tr) return g_create_error.empty() ? nullptr : g_create_error.c_str();
    return tr->err.empty() ? nullptr : tr->err.c_str();
}

int pc_set_option(pc_tracer *tr, int option, int value) {
    if (!tr) return PC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> g(tr->mu);
    switch (option) {
        case PC_OPT_COUNTERS: tr->optCounters = value != 0; break;
        case PC_OPT_PRIMARY_PACKETS: tr->optPackets = value != 0; break;
        case PC_OPT_REFERENCE_ORDER: tr->optRefOrder = value != 0; break;
        case PC_OPT_USE_GRAPH: tr->optGraph = value != 0; break;
        case PC_OPT_FIX_Q4: tr->optFixQ4 = value != 0; break;
        case PC_OPT_KERNEL_TIMERS: tr->optTimers = value != 0; break;
        case PC_OPT_FUSE_TRACE: tr->optFuse = value != 0; break;
        case PC_OPT_SORT_RAYS: tr->optSort = value != 0; break;
        case PC_OPT_DEFER_OCCLUSION: tr->optDeferOcc = value != 0; break;
        case PC_OPT_SAMPLE_SLOTS:
            if (value < 0 || value > MAX_SLOTS) return fail(tr, PC_ERR_INVALID_ARGUMENT, "sample slots must be in [0, %d] (0 = by block size)", MAX_SLOTS);
            tr->optSlots = value;
            break;
        case PC_OPT_TRACE_REFILL:
            if (value < -1 || value > 1) return fail(tr, PC_ERR_INVALID_ARGUMENT, "trace refill must be -1 (by scene), 0 or 1");
            tr->optRefill = value;
            tr->refillNow = value < 0 ? tr->innerNodes >= PC_REFILL_AUTO_MIN_NODES : value != 0;
            break;
        case PC_OPT_SAMPLE_CHAINS:
            if (value < 1 || value > MAX_CHAINS) return fail(tr, PC_ERR_INVALID_ARGUMENT, "sample chains must be in [1, %d]", MAX_CHAINS);
            tr->optChains = value;
            break;
        default: return fail(tr, PC_ERR_INVALID_ARGUMENT, "unknown option %d", option);
    }
    return 0;
}

// bufferSet.Resize (buffers.go:127-175).  Ray / path state is frame sized like the reference's.
int pc_resize(pc_tracer *tr, uint32_t w, uint32_t h) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    if (w == 0 || h == 0 || (uint64_t)w * h >= (1ull << 24)) {
        // path indices travel as floats in ray.dir.w and are exact only below 2^24
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "frame %ux%u unsupported: needs 0 < pixels < 2^24", w, h);
    }
    if (w == tr->W && h == tr->H) return 0;
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->frameStream));
    drop_graph(tr);
    const size_t px = (size_t)w * h;
    release_chain_buffers(tr);
    CU(tr, PC_ERR_ALLOC, tr->traceAcc.alloc(px * 16));
    CU(tr, PC_ERR_ALLOC, tr->frameAcc.alloc(px * 16));
    CU(tr, PC_ERR_ALLOC, tr->frameBuf.alloc(px * 4));
    tr->W = w;
    tr->H = h;
    DevBuf *zero[] = {&tr->traceAcc, &tr->frameAcc, &tr->frameBuf};
    for (DevBuf *b : zero) CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(b->p, 0, b->bytes, tr->stream));
    // the per-chain ray / path / hit state is allocated by the first pc_trace, sized by its block
    {
        std::lock_guard<std::mutex> fg(tr->frameMu);
        tr->frameOpen = tr->ownFirstPassTraced = false;
        tr->firstPassRows.clear();
    }
    for (auto &e : tr->exportBuf) e.release();
    return 0;
}

// bufferSet.UploadSceneData (buffers.go:177-201) + derived traversal layout (pc_layout.hpp)
int pc_upload_scene(pc_tracer *tr, const pc_scene_view *v) {
    int rc = enter(tr);
    if (rc) return rc;
    if (!v) return fail(tr, PC_ERR_INVALID_ARGUMENT, "null scene view");
    std::lock_guard<std::mutex> g(tr->mu);
    if (v->bvh_nodes_bytes < 32 || v->bvh_nodes_bytes % 32 || v->mesh_instances_bytes % 80 || v->material_nodes_bytes % 64 ||
        v->texture_metadata_bytes % 16 || v->vertices_bytes % 48 || v->normals_bytes != v->vertices_bytes ||
        v->uvs_bytes * 2 != v->vertices_bytes || v->material_indices_bytes * 12 != v->vertices_bytes || v->emissives_bytes % 80)
        return fail(tr, PC_ERR_BAD_SCENE, "scene buffer sizes are inconsistent with the reference layouts");
    const size_t nMat = v->material_nodes_bytes / 64;
    if (v->scene_diffuse_mat_index < -1 || v->scene_diffuse_mat_index >= (int64_t)nMat)
        return fail(tr, PC_ERR_BAD_SCENE, "scene diffuse material index %d out of range (-1 = none, %zu material nodes)", v->scene_diffuse_mat_index, nMat);
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
    // The copies of the ten flat buffers start first: from page-locked host memory (pc_host_register) they run while this
    // thread derives the traversal layout and validates the scene below.  A scene that fails validation leaves the handle
    // without scene data (its buffers were overwritten).
    // The per-sample graph survives a re-upload when every device buffer keeps its address and the scene scalars are
    // unchanged (DevBuf::alloc reuses allocations of the same size; GraphKey compares the DScene by value): an e2e loop that
    // re-sends the same scene every frame does not re-instantiate the kernel nodes each time
    tr->hasScene = false;
    if ((rc = upload(tr, tr->verts, v->vertices, v->vertices_bytes))) return rc;
    if ((rc = upload(tr, tr->normals, v->normals, v->normals_bytes))) return rc;
    if ((rc = upload(tr, tr->uvs, v->uvs, v->uvs_bytes))) return rc;
    if ((rc = upload(tr, tr->texData, v->texture_data, v->texture_data_bytes))) return rc;
    if ((rc = upload(tr, tr->matIdx, v->material_indices, v->material_indices_bytes))) return rc;
    if ((rc = upload(tr, tr->bvh, v->bvh_nodes, v->bvh_nodes_bytes))) return rc;
    if ((rc = upload(tr, tr->inst, v->mesh_instances, v->mesh_instances_bytes))) return rc;
    if ((rc = upload(tr, tr->mats, v->material_nodes, v->material_nodes_bytes))) return rc;
    if ((rc = upload(tr, tr->texMeta, v->texture_metadata, v->texture_metadata_bytes))) return rc;
    if ((rc = upload(tr, tr->emissives, v->emissives, v->emissives_bytes))) return rc;
    pc_layout::Builder lb((const pc_layout::RefNode *)v->bvh_nodes, v->bvh_nodes_bytes / 32,
                          (const pc_layout::RefInstance *)v->mesh_instances, v->mesh_instances_bytes / 80,
                          (const pc_layout::Q *)v->vertices, v->vertices_bytes / 16, /*derive_tris=*/false);
    pc_layout::Layout L = lb.build();
    auto bad_scene = [&](int code, const std::string &msg) {
        cudaStreamSynchronize(tr->stream);  // the caller's buffers are no longer read when this returns
        return fail(tr, code, "%s", msg.c_str());
    };
    if (!L.error.empty()) return bad_scene(PC_ERR_BAD_SCENE, L.error);
#ifdef PC_WIDE_BVH
    pc_layout::Builder::build_wide(L);
#endif
    const int stackWords = PC_POP_CULL ? 2 : 1;  // closest-hit entries carry their entry distance under PC_POP_CULL
    if (L.stack_need * stackWords > PC_STACK_SIZE)  // the reference reserves 32 entries and never checks
        return bad_scene(PC_ERR_STACK_DEPTH, "BVH needs a " + std::to_string(L.stack_need) + "-entry traversal stack, the kernels have " + std::to_string(PC_STACK_SIZE / stackWords));
    {   // validate what the shading kernels index with
        const uint32_t *mi = (const uint32_t *)v->material_indices;
        const size_t nTri = v->material_indices_bytes / 4;
        size_t badAt = nTri;
        for (size_t i = 0; i < nTri; i++)
            if (mi[i] >= nMat) { badAt = i; break; }
        if (badAt < nTri)
            return bad_scene(PC_ERR_BAD_SCENE, "triangle " + std::to_string(badAt) + " references material node " + std::to_string(mi[badAt]) + " of " + std::to_string(nMat));
    }
    if ((rc = upload(tr, tr->node64, L.node64.data(), L.node64.size() * 16))) return rc;
#ifdef PC_WIDE_BVH
    if ((rc = upload(tr, tr->node128, L.node128.data(), L.node128.size() * 16))) return rc;
#endif
    {   // tri48 is derived on the device from the vertices and BVH nodes uploaded above
        const size_t nTris = v->vertices_bytes / 48, nNodes = v->bvh_nodes_bytes / 32;
        CU(tr, PC_ERR_ALLOC, tr->tri48.alloc(nTris * 48));
        if (nTris) {
            const int blocks = tr->prop.multiProcessorCount * 8;
            k_derive_tri48<<<blocks, 256, 0, tr->stream>>>((const float4 *)tr->verts.p, (float4 *)tr->tri48.p, nTris);
            k_leaf_counts<<<blocks, 256, 0, tr->stream>>>((const float4 *)tr->bvh.p, nNodes, (float4 *)tr->tri48.p, nTris);
            CU(tr, PC_ERR_KERNEL, cudaGetLastError());
        }
    }
    if ((rc = upload(tr, tr->inst80, L.inst80.data(), L.inst80.size() * 16))) return rc;
    CU(tr, PC_ERR_COPY_TO_DEVICE, cudaStreamSynchronize(tr->stream));  // host vectors die with this scope
    DScene &s = tr->sc;
    s.node64 = (const float4 *)tr->node64.p;
#ifdef PC_WIDE_BVH
    s.node128 = (const float4 *)tr->node128.p;
#endif
    s.tri48 = (const float4 *)tr->tri48.p;
    s.inst80 = (const float4 *)tr->inst80.p;
    s.rootRef = L.root_ref;
    s.bvhNodes = (const float4 *)tr->bvh.p;
    s.meshInstances = (const float4 *)tr->inst.p;
    s.vertices = (const float4 *)tr->verts.p;
    s.normals = (const float4 *)tr->normals.p;
    s.uvs = (const float2 *)tr->uvs.p;
    s.matIndex = (const uint32_t *)tr->matIdx.p;
    s.matNodes = (const float4 *)tr->mats.p;
    s.emissives = (const float4 *)tr->emissives.p;
    s.texMeta = (const uint4 *)tr->texMeta.p;
    s.texData = (const uint8_t *)tr->texData.p;
    s.numEmissives = (uint32_t)(v->emissives_bytes / 80);
    s.sceneDiffuseMat = v->scene_diffuse_mat_index;
    {   // top-level root box -> 2 x 2 x 2 origin cells of the traversal-order sort key
        const float *n0 = (const float *)v->bvh_nodes;  // {min.xyz, L}{max.xyz, R}
        s.worldMin = make_float3(n0[0], n0[1], n0[2]);
        auto cellScale = [](float lo, float hi) { return hi > lo ? 2.0f / (hi - lo) : 0.0f; };
        s.worldCellScale = make_float3(cellScale(n0[0], n0[4]), cellScale(n0[1], n0[5]), cellScale(n0[2], n0[6]));
    }
    tr->stackNeed = L.stack_need;
    tr->innerNodes = L.node64.size() / 4;
    tr->refillNow = tr->optRefill < 0 ? tr->innerNodes >= PC_REFILL_AUTO_MIN_NODES : tr->optRefill != 0;
    tr->hasScene = true;
    tr->sceneEpoch++;
    return 0;
}

int pc_set_camera(pc_tracer *tr, const float eye[3], const float frustum[16]) {
    if (!tr || !eye || !frustum) return PC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> g(tr->mu);
    tr->cam.eye = make_float3(eye[0], eye[1], eye[2]);
    memcpy(&tr->cam.frustrumTL, frustum, 64);
    tr->hasCamera = true;
    tr->cameraEpoch++;
    return 0;
}

// Automatic sample slots: paths per launch aimed at.  8 M on small scenes; 16 M where walks are long (the scenes that also get
// the refilling k_trace): measured +1.0 % on config 3 and +1.7 % on config 4, but -2 % on the 4K Cornell frame and +-0.4 % on
// configs 1 and 2; 16 slots instead of 8 add nothing anywhere (profiles/ab_r02n.txt).
#ifndef PC_SLOT_TARGET_PATHS
#define PC_SLOT_TARGET_PATHS (8u << 20)
#endif
#ifndef PC_SLOT_TARGET_PATHS_LONG
#define PC_SLOT_TARGET_PATHS_LONG (16u << 20)
#endif
// Tracer.Trace (tracer.go:194-247); dbg != nullptr: MonteCarloIntegrator(debugFlags) -- one chain, direct launches
static int trace_common(pc_tracer *tr, pc_block_request *req, const uint32_t *seeds, size_t n_seeds, pc_stats *stats, DebugSink *dbg) {
    int rc = enter(tr);
    if (rc) return rc;
    if (!req) return fail(tr, PC_ERR_INVALID_ARGUMENT, "null block request");
    std::lock_guard<std::mutex> g(tr->mu);
    auto t0 = std::chrono::steady_clock::now();
    if (!tr->hasScene) return fail(tr, PC_ERR_NO_SCENE_DATA, "no scene data uploaded");  // errors.go:21
    if (tr->W == 0 || req->frame_w != tr->W || req->frame_h != tr->H)
        return fail(tr, PC_ERR_NO_FRAME, "block request is for a %ux%u frame, buffers are %ux%u", req->frame_w, req->frame_h, tr->W, tr->H);
    if (req->block_y + req->block_h > req->frame_h || req->block_h == 0 || req->block_w != req->frame_w || req->block_x != 0)
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "block must be full-width rows inside the frame");
    if (req->num_bounces == 0 || req->num_bounces > MAX_BOUNCES)
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "num_bounces must be in [1, %d]", MAX_BOUNCES);
    const uint32_t spp = req->samples_per_pixel, perSample = 1 + req->num_bounces;
    const size_t need = (size_t)perSample * spp;
    if (seeds && n_seeds < need) return fail(tr, PC_ERR_INVALID_ARGUMENT, "need %zu seeds, got %zu", need, n_seeds);
    std::vector<uint32_t> own;
    if (!seeds) {  // the reference draws from Go's global math/rand (tracer.go:222, pipeline.go:146)
        own.resize(need);
        for (auto &x : own) x = next_seed(tr->seedState);
        seeds = own.data();
    }
    tr->cam.texelDims = make_float2(1.0f / (float)req->frame_w, 1.0f / (float)req->frame_h);  // resources.go:130-133
    cudaStream_t s = tr->stream;
    {   // merges enqueued without waiting (possibly by other devices) may still be reading the trace accumulator
        std::lock_guard<std::mutex> rg(tr->readMu);
        for (auto &f : tr->pendingReads) CU(tr, PC_ERR_KERNEL, cudaStreamWaitEvent(s, f->ev, 0));
        tr->pendingReads.clear();
    }
    // pipeline.Reset -> ClearFrameAccumulator when the sample counter was reset (tracer.go:208-213)
    if (req->accumulated_samples == 0 && (rc = open_frame_for_trace(tr))) return rc;
    // ClearTraceAccumulator (tracer.go:215).  With the literal Q4 behaviour (and in the debug stages) emissive hits land at
    // the block-local path index, i.e. possibly outside the block's rows: those modes clear the whole frame.
    const bool wholeFrame = !tr->optFixQ4 || dbg;
    const uint32_t rowsY0 = wholeFrame ? 0u : req->block_y, rowsY1 = wholeFrame ? tr->H : req->block_y + req->block_h;
    if (need * 4 > tr->seedsDev.bytes) CU(tr, PC_ERR_ALLOC, tr->seedsDev.alloc(need * 4 + 4096));
    if (need) CU(tr, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(tr->seedsDev.p, seeds, need * 4, cudaMemcpyHostToDevice, s));
    TraceParams hp;
    hp.cam = tr->cam;
    hp.frameW = req->frame_w; hp.blockY = req->block_y; hp.blockH = req->block_h; hp.pad = 0;
    CU(tr, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(tr->params.p, &hp, sizeof(hp), cudaMemcpyHostToDevice, s));
    // ---- sample chains: chain c traces samples c, c + nc, c + 2 nc, ...
    int nc = (tr->optTimers || dbg) ? 1 : tr->optChains;
    if ((uint32_t)nc > spp) nc = spp ? (int)spp : 1;
    if (dbg) CU(tr, PC_ERR_ALLOC, tr->debugBuf.bytes >= (size_t)tr->W * tr->H * 4 + 16 ? cudaSuccess : tr->debugBuf.alloc((size_t)tr->W * tr->H * 4 + 16));
    // ---- sample slots: a set of launches carries `slots` samples of a chain (TraceCtl in pc_kernels.cuh): sample k of the
    // request is traced by chain k % nc in slot (k / nc) % slots.  Automatic: enough slots for ~8 M paths per launch, 16 M on large scenes (profiles/ab_r02g.txt, ab_r02n.txt).
    const size_t blockRays = (size_t)req->frame_w * req->block_h;
    const size_t slotTarget = tr->innerNodes >= PC_REFILL_AUTO_MIN_NODES ? (size_t)PC_SLOT_TARGET_PATHS_LONG : (size_t)PC_SLOT_TARGET_PATHS;
    int slots = tr->optSlots > 0 ? tr->optSlots : (int)(slotTarget / (blockRays ? blockRays : 1));
    if (slots > MAX_SLOTS) slots = MAX_SLOTS;
    if (slots < 1 || dbg || tr->optPackets || tr->optRefOrder) slots = 1;
    if (slots > MAX_SLOTS) slots = MAX_SLOTS;
    while (slots > 1 && (uint32_t)(nc * slots) > spp) slots--;
    while (slots > 1 && blockRays * (size_t)slots >= (1u << 24)) slots--;  // path indices travel as floats... per slot, but ray counts stay below 2^31
    if ((rc = ensure_chains(tr, nc, blockRays, slots))) return rc;
    static const uint32_t kChainIndex[MAX_CHAINS] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int c = 0; c < nc; c++)
        for (int sl = 0; sl < slots; sl++) {
            void *acc = c == 0 && sl == 0 ? tr->traceAcc.p : tr->chain[c].accs[sl].p;
            if ((rc = clear_for_block(tr, tr->chain[c].dirtyY0[sl], tr->chain[c].dirtyY1[sl], acc, rowsY0, rowsY1, s))) return rc;
        }
    for (int c = 0; c < nc; c++) {
        // reset the persistent part of the control block, keep the three ray counters
        char *ctl = (char *)tr->chain[c].ctl.p;
        CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(ctl + offsetof(TraceCtl, nextSample), 0, sizeof(TraceCtl) - offsetof(TraceCtl, nextSample), s));
        // ... except the occlusion-ray counter: the first k_primary of this call must find no rays left over (see k_primary)
        CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(ctl + 2 * sizeof(int), 0, sizeof(int), s));
        if (c > 0) CU(tr, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(ctl + offsetof(TraceCtl, nextSample), &kChainIndex[c], 4, cudaMemcpyHostToDevice, s));
    }
    uint64_t launches = 2;
    CU(tr, PC_ERR_KERNEL, cudaEventRecord(tr->evStart, s));
    if (spp > 0) {
        tr->timerClass.clear();
        uint32_t done = 0;
        if (dbg) {
            for (; done < spp; done++) {
                dbg->capture = done + 1 == spp;
                dbg->sampleIndex = done;
                uint64_t L = 0;
                record_sample(tr, tr->chain[0], *req, 1u, 1u, 1u, &L, dbg);
                launches += L;
            }
            if (cudaGetLastError() != cudaSuccess) {
                tr->dead = true;
                return fail(tr, PC_ERR_KERNEL, "debug stage launch failed");
            }
        }
        // batches every chain contributes to one replay of the captured graph: 4 one-sample batches, fewer when a batch
        // already carries several samples (a 64-spp pass at 4 chains x 8 slots is two replays)
        const int graphBatches = slots >= 4 ? 1 : (slots >= 2 ? 2 : GRAPH_SAMPLES_PER_CHAIN);
        const uint32_t perGraph = (uint32_t)(graphBatches * nc * slots);  // samples per replay
        if (tr->optGraph && !tr->optTimers && !dbg && spp >= perGraph) {
            GraphKey key;
            key.nb = req->num_bounces; key.rr = req->min_bounces_for_rr;
            key.counters = tr->optCounters; key.packets = tr->optPackets; key.reforder = tr->optRefOrder; key.fixq4 = tr->optFixQ4;
            key.chains = nc; key.slots = slots; key.fuse = tr->optFuse; key.sort = tr->optSort; key.defer = tr->optDeferOcc; key.refill = tr->refillNow ? 1 : 0; memcpy(&key.scene, &tr->sc, sizeof(DScene)); key.seedsPtr = tr->seedsDev.p;
            if (!tr->graphExec || !(key == tr->graphKey)) {
                drop_graph(tr);
                cudaGraph_t graph = nullptr;
                uint32_t perChain[MAX_CHAINS] = {};
                for (int c = 0; c < nc; c++) perChain[c] = (uint32_t)(graphBatches * slots);
                tr->launchesPerSample = 0;  // launches of one replay
                CU(tr, PC_ERR_KERNEL, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
                int bad = enqueue_chains(tr, *req, nc, slots, perChain, &tr->launchesPerSample);
                cudaError_t ce = cudaStreamEndCapture(s, &graph);
                if (bad || ce != cudaSuccess) {
                    if (graph) cudaGraphDestroy(graph);
                    tr->dead = true;
                    return fail(tr, PC_ERR_KERNEL, "stream capture of the sample graph failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : cudaGetLastError()));
                }
                cudaError_t e = cudaGraphInstantiate(&tr->graphExec, graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) {
                    tr->graphExec = nullptr;
                    return fail(tr, PC_ERR_KERNEL, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
                }
                tr->graphKey = key;
                // the device clock of this call starts after the (host-side, one-off) capture + instantiation: a scheduler
                // that balances blocks by Stats() must not see it as render time
                CU(tr, PC_ERR_KERNEL, cudaEventRecord(tr->evStart, s));
            }
            for (; done + perGraph <= spp; done += perGraph) {
                CU(tr, PC_ERR_KERNEL, cudaGraphLaunch(tr->graphExec, s));
                launches += tr->launchesPerSample;
            }
        }
        const bool flush = defer_last_occlusion(tr, dbg != nullptr);
        if (done < spp || flush) {  // the remainder (or everything, without graphs): same chain assignment, direct launches
            // sample k -> chain k % nc, slot (k / nc) % slots: what is left is `rounds` of nc * slots samples and a tail
            uint32_t perChain[MAX_CHAINS] = {};
            for (uint32_t i = done; i < spp; i++) perChain[i % (uint32_t)nc]++;
            if (enqueue_chains(tr, *req, nc, slots, perChain, &launches, flush, &launches)) {
                tr->dead = true;
                return fail(tr, PC_ERR_KERNEL, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
        }
        // chains > 0 accumulated into their own buffers: add them in chain order, the block's rows only, one launch
        if (nc * slots > 1) {
            static_assert(MAX_CHAINS * MAX_SLOTS <= MAX_CHAINS_X_SLOTS, "ChainAccs holds one pointer per (chain, slot)");
            ChainAccs ca;
            ca.n = 0;
            for (int c = 0; c < nc; c++)
                for (int sl = 0; sl < slots; sl++)
                    if (c || sl) ca.src[ca.n++] = (const float4 *)tr->chain[c].accs[sl].p;
            const size_t off = (size_t)tr->W * rowsY0, cnt = (size_t)tr->W * (rowsY1 - rowsY0);
            k_merge_chains<<<grid_for(cnt, 256, tr->prop.multiProcessorCount * 8), 256, 0, s>>>((float4 *)tr->traceAcc.p, ca, off, cnt);
            launches++;
        }
        tr->lastChain = (int)((spp - 1) % (uint32_t)nc);
        tr->lastSlot = (int)(((spp - 1) / (uint32_t)nc) % (uint32_t)slots);
        tr->lastSlotPaths = (uint32_t)blockRays;
        tr->lastBounces = req->num_bounces;
    }
    CU(tr, PC_ERR_KERNEL, cudaEventRecord(tr->evStop, s));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(s));
    CU(tr, PC_ERR_KERNEL, cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tr->evStart, tr->evStop);
    TraceCtl hc;
    memset(&hc, 0, sizeof(hc));
    for (int c = 0; c < nc; c++) {
        TraceCtl one;
        CU(tr, PC_ERR_COPY_TO_HOST, cudaMemcpy(&one, tr->chain[c].ctl.p, sizeof(one), cudaMemcpyDeviceToHost));
        for (int k = 0; k < ST_COUNT; k++) hc.stats[k] += one.stats[k];
    }
    if (spp) req->seed = seeds[(size_t)perSample * (spp - 1)];  // the last camera seed (tracer.go:222)
    req->accumulated_samples += spp;                             // tracer.go:240
    pc_stats &st = tr->stats;
    memset(&st, 0, sizeof(st));
    st.block_w = req->block_w;
    st.block_h = req->block_h;
    st.device_time_ns = (uint64_t)((double)ms * 1e6);
    st.query_rays = hc.stats[ST_QUERY_RAYS];
    st.occlusion_rays = hc.stats[ST_OCCLUSION_RAYS];
    st.kernel_launches = launches + (dbg ? dbg->launches : 0);
    st.nodes_tested = hc.stats[ST_NODES];
    st.tris_tested = hc.stats[ST_TRIS];
    st.instances_entered = hc.stats[ST_INSTANCES];
    st.shaded_hits = hc.stats[ST_SHADED];
    st.occlusion_emitted = hc.stats[ST_OCC_EMITTED];
    st.indirect_emitted = hc.stats[ST_IND_EMITTED];
    st.unoccluded = hc.stats[ST_UNOCCLUDED];
    st.missed_query_rays = hc.stats[ST_MISSED];
    tr->timerUs.clear();
    if (tr->optTimers) {
        tr->timerUs.assign(tr->timerClass.size(), 0.f);
        for (size_t k = 0; k < tr->timerClass.size(); k++) {
            float kms = 0.f;
            if (cudaEventElapsedTime(&kms, tr->timerEvents[2 * k], tr->timerEvents[2 * k + 1]) == cudaSuccess) {
                tr->timerUs[k] = kms * 1e3f;
                st.kernel_time_ns[tr->timerClass[k]] += (uint64_t)((double)kms * 1e6);
                st.kernel_count[tr->timerClass[k]]++;
            }
        }
    }
    st.render_time_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = st;
    if (dbg && dbg->overflow) return fail(tr, PC_ERR_INVALID_ARGUMENT, "debug frames do not fit: need %u frames of %zu bytes", pc_debug_frame_count(dbg->flags, req->num_bounces), (size_t)tr->W * tr->H * 4);
    return 0;
}

int pc_trace(pc_tracer *tr, pc_block_request *req, const uint32_t *seeds, size_t n_seeds, pc_stats *stats) {
    return trace_common(tr, req, seeds, n_seeds, stats, nullptr);
}

uint32_t pc_debug_frame_count(uint32_t debug_flags, uint32_t num_bounces) {
    const uint32_t once = PC_DEBUG_PRIMARY_DEPTH | PC_DEBUG_PRIMARY_NORMALS;
    const uint32_t perBounce = PC_DEBUG_ALL_EMISSIVE | PC_DEBUG_VISIBLE_EMISSIVE | PC_DEBUG_OCCLUDED_EMISSIVE | PC_DEBUG_THROUGHPUT | PC_DEBUG_ACCUMULATOR;
    return (uint32_t)__builtin_popcount(debug_flags & once) + num_bounces * (uint32_t)__builtin_popcount(debug_flags & perBounce);
}

// MonteCarloIntegrator(debugFlags) (pipeline.go:94-213)
int pc_trace_debug(pc_tracer *tr, pc_block_request *req, const uint32_t *seeds, size_t n_seeds, uint32_t debug_flags, uint8_t *frames_out,
                   uint64_t frames_cap_bytes, pc_debug_frame *infos, uint32_t infos_cap, uint32_t *n_frames, pc_stats *stats) {
    if (n_frames) *n_frames = 0;
    if (!tr) return PC_ERR_INVALID_ARGUMENT;
    if ((!frames_out || !infos) && pc_debug_frame_count(debug_flags, req ? req->num_bounces : 0))
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "null debug frame buffer");
    DebugSink d;
    d.flags = debug_flags; d.frames = frames_out; d.cap = frames_cap_bytes; d.infos = infos; d.infosCap = infos_cap;
    int rc = trace_common(tr, req, seeds, n_seeds, stats, &d);
    if (n_frames) *n_frames = d.n;
    return rc;
}

int pc_get_stats(pc_tracer *tr, pc_stats *stats) {
    if (!tr || !stats) return PC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> g(tr->mu);
    *stats = tr->stats;
    return 0;
}

int pc_get_kernel_timings(pc_tracer *tr, uint32_t *classes, float *us, uint32_t cap, uint32_t *n) {
    if (!tr || !n) return PC_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> g(tr->mu);
    const size_t have = tr->timerUs.size();
    *n = (uint32_t)have;
    for (size_t k = 0; k < have && k < cap; k++) {
        if (classes) classes[k] = (uint32_t)tr->timerClass[k];
        if (us) us[k] = tr->timerUs[k];
    }
    return 0;
}

static int check_block(pc_tracer *tr, const pc_block_request *req) {
    if (!req) return fail(tr, PC_ERR_INVALID_ARGUMENT, "null block request");
    if (tr->W == 0 || req->frame_w != tr->W || req->frame_h != tr->H) return fail(tr, PC_ERR_NO_FRAME, "frame dimensions do not match");
    if ((uint64_t)req->frame_w * req->block_y + (uint64_t)req->block_w * req->block_h > (uint64_t)tr->W * tr->H)
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "block outside the frame");
    return 0;
}

// Record "the merge just enqueued on dst's frame stream has read src's rows" and hand the fence to src.
static int fence_source(pc_tracer *dst, pc_tracer *src) {
    if (dst->fencePool.size() < 64) {
        auto f = std::make_shared<ReadFence>();
        CU(dst, PC_ERR_ALLOC, cudaEventCreateWithFlags(&f->ev, cudaEventDisableTiming));
        dst->fencePool.push_back(f);
    }
    std::shared_ptr<ReadFence> f = dst->fencePool[dst->fenceNext++ % dst->fencePool.size()];
    CU(dst, PC_ERR_KERNEL, cudaEventRecord(f->ev, dst->frameStream));
    std::lock_guard<std::mutex> rg(src->readMu);
    if (src->pendingReads.size() >= 32) src->pendingReads.erase(src->pendingReads.begin());  // a re-recorded fence only waits longer
    src->pendingReads.push_back(f);
    return 0;
}

// Tracer.MergeOutput -> aggregateAccumulator with global offset FrameW*BlockY (resources.go:108-124).
// Every worker calls this on the primary, concurrently, from its own OS thread (renderer/default.go:188-191, SURVEY
// Q18): serialised per destination by frameMu -- NOT by the handle's mutex, which the primary's own Trace holds while it
// waits for the device -- and enqueued on the destination's frame stream without waiting (Exec1DNoWait, resources.go:119).
// The source's rows are complete (pc_trace is synchronous); when src lives on another GPU the loads cross NVLink.
int pc_merge_output(pc_tracer *dst, pc_tracer *src, const pc_block_request *req) {
    if (!src) return fail(dst, PC_ERR_UNSUPPORTED_TRACER, "merge failed: unsupported tracer instance");
    int rc = enter(dst);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(dst->frameMu);
    if ((rc = check_block(dst, req))) return rc;
    if (src->W != dst->W || src->H != dst->H) return fail(dst, PC_ERR_NO_FRAME, "source tracer has different frame dimensions");
    if (src->device != dst->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, dst->device, src->device);
        if (!can) return fail(dst, PC_ERR_PEER_ACCESS, "device %d cannot access device %d", dst->device, src->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(dst, PC_ERR_PEER_ACCESS, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    if ((rc = open_frame_for_merge(dst, req))) return rc;
    const size_t off = (size_t)req->frame_w * req->block_y, n = (size_t)req->block_w * req->block_h;
    k_merge<<<grid_for(n, 256, dst->prop.multiProcessorCount * 8), 256, 0, dst->frameStream>>>((float4 *)dst->frameAcc.p, (const float4 *)src->traceAcc.p, off, off, n);
    CU(dst, PC_ERR_KERNEL, cudaGetLastError());
    return fence_source(dst, src);
}

// The same merge for rows that live outside a pc_tracer of this process: a host buffer, a device buffer, or another
// process's export buffer mapped with pc_ipc_open (then the loads cross NVLink exactly like pc_merge_output's).
int pc_merge_rows(pc_tracer *dst, const void *rows, int is_device, const pc_block_request *req) {
    int rc = enter(dst);
    if (rc) return rc;
    if (!rows) return fail(dst, PC_ERR_INVALID_ARGUMENT, "null rows");
    std::lock_guard<std::mutex> g(dst->frameMu);
    if ((rc = check_block(dst, req))) return rc;
    if ((rc = open_frame_for_merge(dst, req))) return rc;
    const size_t off = (size_t)req->frame_w * req->block_y, n = (size_t)req->block_w * req->block_h;
    const float4 *src = (const float4 *)rows;
    if (!is_device) {
        if (dst->scratch.bytes < n * 16) CU(dst, PC_ERR_ALLOC, dst->scratch.alloc(n * 16));
        CU(dst, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(dst->scratch.p, rows, n * 16, cudaMemcpyHostToDevice, dst->frameStream));
        src = (const float4 *)dst->scratch.p;
    }
    k_merge<<<grid_for(n, 256, dst->prop.multiProcessorCount * 8), 256, 0, dst->frameStream>>>((float4 *)dst->frameAcc.p, src, off, 0, n);
    CU(dst, PC_ERR_KERNEL, cudaGetLastError());
    if (!is_device) CU(dst, PC_ERR_KERNEL, cudaStreamSynchronize(dst->frameStream));  // scratch is reused
    return 0;
}

int pc_trace_rows(pc_tracer *tr, const pc_block_request *req, void **device_ptr, uint64_t *bytes) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    if ((rc = check_block(tr, req))) return rc;
    if (device_ptr) *device_ptr = (char *)tr->traceAcc.p + (size_t)req->frame_w * req->block_y * 16;
    if (bytes) *bytes = (uint64_t)req->block_w * req->block_h * 16;
    return 0;
}

// ---- one process per GPU: the reference reaches a peer tracer's accumulator through a SHARED OpenCL context
// (device/context.go:11-28, renderer/default.go:227); across processes the equivalent is a CUDA IPC mapping.
int pc_ipc_export(pc_tracer *tr, int slot, void *handle64) {
    int rc = enter(tr);
    if (rc) return rc;
    if (slot < 0 || slot > 1 || !handle64) return fail(tr, PC_ERR_INVALID_ARGUMENT, "export slot must be 0 or 1");
    std::lock_guard<std::mutex> g(tr->mu);
    if (tr->W == 0) return fail(tr, PC_ERR_NO_FRAME, "no frame dimensions");
    DevBuf &b = tr->exportBuf[slot];
    const size_t bytes = (size_t)tr->W * tr->H * 16;
    if (b.bytes != bytes) {
        b.release();
        CU(tr, PC_ERR_ALLOC, b.alloc(bytes));
        CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(b.p, 0, bytes, tr->stream));
        CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == PC_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, b.p);
    if (e != cudaSuccess) return fail(tr, PC_ERR_PEER_ACCESS, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

int pc_ipc_publish_rows(pc_tracer *tr, const pc_block_request *req, int slot) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    if ((rc = check_block(tr, req))) return rc;
    if (slot < 0 || slot > 1 || !tr->exportBuf[slot].p) return fail(tr, PC_ERR_INVALID_ARGUMENT, "export slot %d was not exported", slot);
    const size_t off = (size_t)req->frame_w * req->block_y * 16, bytes = (size_t)req->block_w * req->block_h * 16;
    CU(tr, PC_ERR_KERNEL, cudaMemcpyAsync((char *)tr->exportBuf[slot].p + off, (const char *)tr->traceAcc.p + off, bytes, cudaMemcpyDeviceToDevice, tr->stream));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));  // visible to the importing process once this returns
    return 0;
}

int pc_ipc_open(pc_tracer *dst, const void *handle64, void **peer_ptr) {
    int rc = enter(dst);
    if (rc) return rc;
    if (!handle64 || !peer_ptr) return fail(dst, PC_ERR_INVALID_ARGUMENT, "null handle");
    std::lock_guard<std::mutex> g(dst->mu);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(dst, PC_ERR_PEER_ACCESS, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    dst->ipcOpened.push_back(p);
    *peer_ptr = p;
    return 0;
}

int pc_ipc_close(pc_tracer *dst, void *peer_ptr) {
    int rc = enter(dst);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(dst->mu);
    for (size_t i = 0; i < dst->ipcOpened.size(); i++)
        if (dst->ipcOpened[i] == peer_ptr) {
            CU(dst, PC_ERR_KERNEL, cudaStreamSynchronize(dst->frameStream));
            cudaIpcCloseMemHandle(peer_ptr);
            dst->ipcOpened.erase(dst->ipcOpened.begin() + i);
            return 0;
        }
    return fail(dst, PC_ERR_INVALID_ARGUMENT, "pointer was not opened with pc_ipc_open");
}

// Device.WaitForKernels (tracer/opencl/device/device.go, called by SyncFramebuffer at tracer.go:259): every launch this
// handle enqueued -- its own trace stream and the merges on its frame stream -- has completed when this returns.
int pc_wait_for_kernels(pc_tracer *tr) {
    int rc = enter(tr);
    if (rc) return rc;
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->frameStream));
    return 0;
}

// Tracer.SyncFramebuffer (tracer.go:250-276) + TonemapSimpleReinhard (resources.go:344-360)
int pc_sync_framebuffer(pc_tracer *tr, const pc_block_request *req, uint8_t *rgba_out) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    std::lock_guard<std::mutex> fg(tr->frameMu);
    if (!tr->hasScene) return fail(tr, PC_ERR_NO_SCENE_DATA, "no scene data uploaded");
    if ((rc = check_block(tr, req))) return rc;
    const size_t n = (size_t)req->frame_w * req->block_h;
    const float sampleWeight = 1.0f / (float)(req->accumulated_samples + req->samples_per_pixel);  // resources.go:347
    cudaStream_t fs = tr->frameStream;
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));  // WaitForKernels (tracer.go:259)
    k_tonemap<<<grid_for(n, 256, tr->prop.multiProcessorCount * 8), 256, 0, fs>>>((const float4 *)tr->frameAcc.p, (uchar4 *)tr->frameBuf.p, n, sampleWeight, req->exposure);
    CU(tr, PC_ERR_KERNEL, cudaGetLastError());
    if (rgba_out) CU(tr, PC_ERR_COPY_TO_HOST, cudaMemcpyAsync(rgba_out, tr->frameBuf.p, (size_t)tr->W * tr->H * 4, cudaMemcpyDeviceToHost, fs));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(fs));
    tr->frameOpen = tr->ownFirstPassTraced = false;
    return 0;
}

// Page-lock caller-owned host memory so that the copies of pc_upload_scene / pc_sync_framebuffer run at full PCIe / C2C
// rate and asynchronously (the reference gets the same effect from CL_MEM_USE_HOST_PTR, device/buffer.go:98-104).
int pc_host_register(void *ptr, uint64_t bytes) {
    if (!ptr || !bytes) return PC_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { cudaGetLastError(); return PC_ERR_ALLOC; }
    return 0;
}
int pc_host_unregister(void *ptr) {
    if (!ptr) return PC_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return PC_ERR_INVALID_ARGUMENT; }
    return 0;
}

int pc_read_buffer(pc_tracer *tr, int which, void *dst, uint64_t bytes) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    const void *src = nullptr;
    size_t have = 0;
    const Chain &lc = tr->chain[tr->lastChain < tr->nChains ? tr->lastChain : 0];
    // per-sample state: the buffers of the chain AND SLOT that traced the LAST sample (what the reference's single set would
    // hold).  A slot's rays start at base[k][slot] of the flat buffers (TraceCtl), its paths at slot * paths-per-slot.
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
    TraceCtl hc;
    memset(&hc, 0, sizeof(hc));
    if (lc.ctl.p) CU(tr, PC_ERR_COPY_TO_HOST, cudaMemcpy(&hc, lc.ctl.p, sizeof(hc), cudaMemcpyDeviceToHost));
    const int sl = tr->lastSlot < MAX_SLOTS ? tr->lastSlot : 0;
    const int aLast = (int)((tr->lastBounces - 1u) & 1u);  // rays[aLast] is what the last closest-hit launch traced
    int32_t counters[3] = {0, 0, 0};
    auto slice = [&](const DevBuf &b, size_t elem, size_t firstElem) {
        const size_t off = firstElem * elem;
        src = off <= b.bytes ? (const char *)b.p + off : nullptr;
        have = off <= b.bytes ? b.bytes - off : 0;
    };
    switch (which) {
        case PC_BUF_RAYS0: case PC_BUF_RAYS1: case PC_BUF_RAYS2: slice(lc.rays[which], 32, hc.base[which][sl]); break;
        case PC_BUF_PATHS: slice(lc.paths, 32, (size_t)sl * tr->lastSlotPaths); break;
        case PC_BUF_HIT_FLAGS: slice(lc.hitFlags, 4, hc.base[aLast][sl]); break;
        case PC_BUF_INTERSECTIONS: slice(lc.hits, 32, hc.base[aLast][sl]); break;
        case PC_BUF_EMISSIVE_SAMPLES: slice(lc.emSamples, 16, hc.base[2][sl]); break;
        case PC_BUF_TRACE_ACCUMULATOR: src = tr->traceAcc.p; have = tr->traceAcc.bytes; break;
        case PC_BUF_FRAME_ACCUMULATOR: src = tr->frameAcc.p; have = tr->frameAcc.bytes; break;
        case PC_BUF_FRAME_BUFFER: src = tr->frameBuf.p; have = tr->frameBuf.bytes; break;
        case PC_BUF_RAY_COUNTERS:  // the slot's share of the three counters
            for (int k = 0; k < 3; k++) counters[k] = (int32_t)(hc.base[k][sl + 1] - hc.base[k][sl]);
            if (bytes > 12) return fail(tr, PC_ERR_INVALID_ARGUMENT, "the ray counters are 12 bytes");
            memcpy(dst, counters, bytes);
            return 0;
        default: return fail(tr, PC_ERR_INVALID_ARGUMENT, "unknown buffer %d", which);
    }
    // per-sample state is sized by the largest block traced (the reference's is frame sized): what lies beyond it was never
    // written, so a frame-sized read gets zeros there
    const bool perSample = which <= PC_BUF_EMISSIVE_SAMPLES;
    const size_t unit = which == PC_BUF_HIT_FLAGS ? 4 : which == PC_BUF_EMISSIVE_SAMPLES ? 16 : 32;
    if (!src || (bytes > have && !(perSample && bytes <= (uint64_t)tr->W * tr->H * unit)))
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "buffer %d holds %zu bytes, %llu requested", which, have, (unsigned long long)bytes);
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->stream));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(tr->frameStream));
    const size_t take = bytes < have ? (size_t)bytes : have;
    CU(tr, PC_ERR_COPY_TO_HOST, cudaMemcpy(dst, src, take, cudaMemcpyDeviceToHost));
    if (bytes > take) memset((char *)dst + take, 0, bytes - take);
    return 0;
}

int pc_debug_intersect(pc_tracer *tr, const void *rays, uint32_t n, int mode, uint32_t *out_flags, void *out_hits) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    if (!tr->hasScene) return fail(tr, PC_ERR_NO_SCENE_DATA, "no scene data uploaded");
    if ((size_t)n > (size_t)tr->W * tr->H) return fail(tr, PC_ERR_INVALID_ARGUMENT, "%u rays exceed the frame's ray buffer", n);
    if ((rc = ensure_chains(tr, 1, n))) return rc;
    cudaStream_t s = tr->stream;
    Chain &c0 = tr->chain[0];
    TraceCtl *ctl = (TraceCtl *)c0.ctl.p;
    CU(tr, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(c0.rays[0].p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CU(tr, PC_ERR_KERNEL, cudaMemsetAsync(ctl, 0, sizeof(TraceCtl), s));
    {   // one slot holding all n rays, in rays[0] (closest hit) and as rays[2] (any hit)
        TraceCtl hc;
        memset(&hc, 0, sizeof(hc));
        hc.numRays[0] = (int)n; hc.numRays[2] = (int)n;
        hc.nSlots = 1; hc.slotStride = 1; hc.slotPaths = n;
        for (int k = 1; k <= MAX_SLOTS; k++) { hc.base[0][k] = n; hc.base[2][k] = n; }
        CU(tr, PC_ERR_COPY_TO_DEVICE, cudaMemcpyAsync(ctl, &hc, sizeof(hc), cudaMemcpyHostToDevice, s));
        CU(tr, PC_ERR_COPY_TO_DEVICE, cudaStreamSynchronize(s));  // hc lives on this stack frame
    }
    tr->lastChain = 0; tr->lastSlot = 0; tr->lastSlotPaths = n; tr->lastBounces = 1;
    const int pg = tr->persistentGrid;
    if (mode == 0) {
        if (tr->optRefOrder) k_query<true, false><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, c0.fb.rays[0], c0.fb.hitFlags, c0.fb.hits, ctl, 0, 0, nullptr);
        else k_query<false, false><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, c0.fb.rays[0], c0.fb.hitFlags, c0.fb.hits, ctl, 0, 0, nullptr);
    } else if (mode == 1) {
        if (tr->optRefOrder) k_occlusion<true, false><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, c0.fb.rays[0], c0.fb.paths, c0.fb.emissiveSamples, c0.fb, 0, c0.fb.hitFlags, ctl, 0, nullptr);
        else k_occlusion<false, false><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, c0.fb.rays[0], c0.fb.paths, c0.fb.emissiveSamples, c0.fb, 0, c0.fb.hitFlags, ctl, 0, nullptr);
    } else if (mode == 2) {
        k_debug_packet<false><<<pg, TRAV_BLOCK, 0, s>>>(tr->sc, c0.fb.rays[0], c0.fb.hitFlags, c0.fb.hits, ctl, n, 0);
    } else {
        return fail(tr, PC_ERR_INVALID_ARGUMENT, "unknown intersect mode %d", mode);
    }
    CU(tr, PC_ERR_KERNEL, cudaGetLastError());
    if (out_flags) CU(tr, PC_ERR_COPY_TO_HOST, cudaMemcpyAsync(out_flags, c0.hitFlags.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (out_hits && mode != 1) CU(tr, PC_ERR_COPY_TO_HOST, cudaMemcpyAsync(out_hits, c0.hits.p, (size_t)n * 32, cudaMemcpyDeviceToHost, s));
    CU(tr, PC_ERR_KERNEL, cudaStreamSynchronize(s));
    return 0;
}

int pc_debug_bxdf(pc_tracer *tr, const void *in_records, uint32_t n, void *out_records) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    if (!tr->hasScene) return fail(tr, PC_ERR_NO_SCENE_DATA, "no scene data uploaded");
    DevBuf in, out;
    CU(tr, PC_ERR_ALLOC, in.alloc((size_t)n * sizeof(BxdfIn)));
    CU(tr, PC_ERR_ALLOC, out.alloc((size_t)n * sizeof(BxdfOut)));
    cudaMemcpyAsync(in.p, in_records, (size_t)n * sizeof(BxdfIn), cudaMemcpyHostToDevice, tr->stream);
    k_debug_bxdf<<<grid_for(n, 128, 65535), 128, 0, tr->stream>>>(tr->sc, (const BxdfIn *)in.p, (BxdfOut *)out.p, n);
    cudaMemcpyAsync(out_records, out.p, (size_t)n * sizeof(BxdfOut), cudaMemcpyDeviceToHost, tr->stream);
    cudaError_t e = cudaStreamSynchronize(tr->stream);
    in.release();
    out.release();
    if (e != cudaSuccess) { tr->dead = true; return fail(tr, PC_ERR_KERNEL, "k_debug_bxdf: %s", cudaGetErrorString(e)); }
    return 0;
}

int pc_debug_rng(pc_tracer *tr, uint32_t *states_inout, uint32_t n, uint32_t draws, float *out) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    DevBuf st, o;
    CU(tr, PC_ERR_ALLOC, st.alloc((size_t)n * 8));
    CU(tr, PC_ERR_ALLOC, o.alloc((size_t)n * draws * 8));
    cudaMemcpyAsync(st.p, states_inout, (size_t)n * 8, cudaMemcpyHostToDevice, tr->stream);
    k_debug_rng<<<grid_for(n, 128, 65535), 128, 0, tr->stream>>>((uint2 *)st.p, n, draws, (float2 *)o.p);
    cudaMemcpyAsync(states_inout, st.p, (size_t)n * 8, cudaMemcpyDeviceToHost, tr->stream);
    cudaMemcpyAsync(out, o.p, (size_t)n * draws * 8, cudaMemcpyDeviceToHost, tr->stream);
    cudaError_t e = cudaStreamSynchronize(tr->stream);
    st.release();
    o.release();
    if (e != cudaSuccess) { tr->dead = true; return fail(tr, PC_ERR_KERNEL, "k_debug_rng: %s", cudaGetErrorString(e)); }
    return 0;
}

int pc_debug_tonemap(pc_tracer *tr, const float *acc, uint32_t n, float sample_weight, float exposure, uint8_t *rgba_out) {
    int rc = enter(tr);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(tr->mu);
    DevBuf a, o;
    CU(tr, PC_ERR_ALLOC, a.alloc((size_t)n * 16));
    CU(tr, PC_ERR_ALLOC, o.alloc((size_t)n * 4));
    cudaMemcpyAsync(a.p, acc, (size_t)n * 16, cudaMemcpyHostToDevice, tr->stream);
    k_tonemap<<<grid_for(n, 256, 65535), 256, 0, tr->stream>>>((const float4 *)a.p, (uchar4 *)o.p, n, sample_weight, exposure);
    cudaMemcpyAsync(rgba_out, o.p, (size_t)n * 4, cudaMemcpyDeviceToHost, tr->stream);
    cudaError_t e = cudaStreamSynchronize(tr->stream);
    a.release();
    o.release();
    if (e != cudaSuccess) { tr->dead = true; return fail(tr, PC_ERR_KERNEL, "k_tonemap: %s", cudaGetErrorString(e)); }
    return 0;
}

}  // extern "C"
