// pc_kernels.cuh -- the __global__ kernels of the CUDA tracer (sm_100a).
//
// Wavefront organisation, one kernel per stage of the reference's pipeline (pipeline.go:94-213),
// fused where a stage only post-processes its predecessor's per-ray output:
//
//   k_begin_sample  picks the sample's seeds, zeroes the per-sample work-queue heads / tickets
//   k_primary       generatePrimaryRays (camera.cl) + closest-hit traversal of the primary rays:
//                   warp-packet traversal (shared stack + votes, the Guenther et al. scheme the
//                   reference uses on GPUs, intersect.cl:353-575) or per-ray traversal
//   k_shade         shadePrimaryRayMisses / shadeIndirectRayMisses + shadeHits
//                   (pt_integrator.cl:17-275): 1024-ray CTA tiles sorted by material, shaded in 32-ray
//                   chunks that the CTA's warps pull from a shared counter, then STABLE
//                   compaction of the occlusion and indirect rays (ballot + popc per warp, offsets
//                   across the CTA, a single-pass decoupled look-back across tiles, ticket ordered),
//                   so ray order == parent ray order, which is what makes bounce >= 1 reproducible
//                   
//   k_occlusion     rayIntersectionTest + accumulateEmissiveSamples (intersect.cl:26-180,
//                   pt_integrator.cl:278-296)
//   k_query         rayIntersectionQuery (intersect.cl:184-347) for the indirect rays
//   k_trace         k_occlusion of bounce b + k_query of bounce b+1 in one persistent launch (default)
//   k_clear / k_merge / k_tonemap   accumulator.cl:5-19, hdr.cl:5-28
//
// The traversal kernels are persistent: the grid is a fixed multiple of the SM count and every
// warp pulls 32-ray work units from a global queue head (atomicAdd) until the queue, whose length
// lives in device memory, is empty -- no host round trip between stages (the reference does a
// clFinish after every launch and two blocking counter writes per bounce).
#pragma once
#include "pc_device.cuh"

namespace pc {

constexpr int MAX_BOUNCES = 32;
#ifndef PC_TRAV_BLOCK
#define PC_TRAV_BLOCK 128   // 4 warps
#endif
constexpr int TRAV_BLOCK = PC_TRAV_BLOCK;
#ifndef PC_PRIMARY_MIN_BLOCKS
#define PC_PRIMARY_MIN_BLOCKS 8   // 64 registers: +1 % on configs 3 and 4 over 6 blocks at 80 registers (profiles/ab_r01j.txt)
#endif
#ifndef PC_SHADE_BLOCK
#define PC_SHADE_BLOCK 256
#endif
constexpr int SHADE_BLOCK = PC_SHADE_BLOCK;
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 2   // 2 x 256 threads x 128 registers: no spills in shadeHit (3 blocks = 80 registers spilled 280 B)
#endif
#ifndef PC_TRAV_MIN_BLOCKS
#define PC_TRAV_MIN_BLOCKS 8
#endif
#ifndef PC_OCC_MIN_BLOCKS
#define PC_OCC_MIN_BLOCKS PC_TRAV_MIN_BLOCKS
#endif

enum StatIdx {
    ST_QUERY_RAYS = 0, ST_OCCLUSION_RAYS, ST_NODES, ST_TRIS, ST_INSTANCES, ST_SHADED, ST_OCC_EMITTED,
    ST_IND_EMITTED, ST_UNOCCLUDED, ST_MISSED, ST_COUNT = 16
};

// Device control block.  [persist] survives a whole pc_trace call, [sample] is zeroed by
// k_begin_sample.
// Sample slots.  The samples of a block request are independent, and at the BASELINE frame sizes ONE sample does not fill
// the machine (config 2's scene: 3.8 Grays/s at 1 Mpx per launch, 4.2 at 4 Mpx, profiles/size_effect_r02.txt), so a launch
// carries up to MAX_SLOTS samples at once: slot s of a batch is the sample curSample + s * slotStride.  The ray buffers
// stay FLAT -- slot 0's rays, then slot 1's, ... -- and because compaction is stable that grouping survives every bounce;
// all a kernel needs is where each slot starts in each ray buffer (base[k][s]) to recover, for ray i of rays[k], its slot,
// its slot-local index (the RNG key of pt_integrator.cl:81 -- results are bit-identical to tracing the samples one by
// one), its path record (paths[slot * slotPaths + pathIndex]) and its accumulator (FrameBufs::slotAcc[slot]).
#ifndef PC_MAX_SLOTS
#define PC_MAX_SLOTS 8
#endif
constexpr int MAX_SLOTS = PC_MAX_SLOTS;
constexpr int MAX_CHAINS_X_SLOTS = 64;
struct TraceCtl {
    int numRays[3];            // [persist] the reference's three ray counters (buffers.go:69), over all slots
    uint32_t nextSample;       // [persist]
    uint32_t curSample;        // [persist] index of slot 0's sample of the batch being traced
    uint32_t nSlots;           // slots of the batch being traced (k_begin_sample)
    uint32_t slotStride;       // sample-index distance between consecutive slots (= number of chains)
    uint32_t slotPaths;        // paths per slot = FrameW * BlockH (k_primary)
    unsigned long long stats[ST_COUNT];  // [persist]
    uint32_t base[3][MAX_SLOTS + 1];     // first ray of every slot in rays[k]; entries of unused slots hold numRays[k]
    uint32_t queueHead[2 * MAX_BOUNCES + 2];  // [sample] work-queue heads, one per traversal launch
    uint32_t ticket[MAX_BOUNCES];             // [sample] block tickets of the shade launches
};

// Per-pc_trace parameters that change from block to block or frame to frame (camera moves, the
// schedulers re-assigning rows).  They live in device memory and are refreshed by a stream-ordered
// copy, so the captured per-sample graph never has to be re-instantiated for them.
struct TraceParams {
    CameraParams cam;
    uint32_t frameW, blockY, blockH, pad;
};

struct Ray { float4 origin, dir; };                     // types.cl:4-10
struct PathRec { float4 throughput; uint4 meta; };      // types.cl:12-25 (meta = pixelIndex, flags, -, -)
struct HitRec { float4 wuvt; uint4 meta; };             // types.cl:69-83 (meta = meshInstance, triIndex, -, -)

struct FrameBufs {
    Ray *rays[3];
    PathRec *paths;
    uint32_t *hitFlags;
    HitRec *hits;
    float4 *emissiveSamples;
    float4 *traceAcc;             // == slotAcc[0]
    uint32_t *permOcc, *permInd;  // traversal order of rays[2] / of the next bounce's rays (PC_OPT_SORT_RAYS)
    float4 *slotAcc[MAX_SLOTS];   // one accumulator per sample slot: two samples of a batch may hit the same pixel at once
};

// slot of ray i given the slots' first rays b[0..MAX_SLOTS] (unused slots start at the total: never selected)
__device__ __forceinline__ uint32_t slotOf(const uint32_t *b, uint32_t i) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 1; k < MAX_SLOTS; k++) s += i >= b[k] ? 1u : 0u;
    return s;
}
struct SlotBases {
    uint32_t b[MAX_SLOTS + 1];
    __device__ __forceinline__ void load(const uint32_t *src) {
#pragma unroll
        for (int k = 0; k <= MAX_SLOTS; k++) b[k] = src[k];
    }
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Ray / path / hit state is written once and read once per bounce (a stream of 100+ MB per launch at
// the BASELINE sizes) while the scene's nodes and triangles are re-read by every ray: the state goes
// through the cache-streaming path (ld.global.cs / st.global.cs, evict-first) so it does not push the
// scene out of L1/L2.
__device__ __forceinline__ Ray ld_ray(const Ray *p) { Ray r; r.origin = __ldcs(&p->origin); r.dir = __ldcs(&p->dir); return r; }
__device__ __forceinline__ void st_ray(Ray *p, float4 origin, float4 dir) { __stcs(&p->origin, origin); __stcs(&p->dir, dir); }
__device__ __forceinline__ HitRec ld_hit(const HitRec *p) { HitRec h; h.wuvt = __ldcs(&p->wuvt); h.meta = __ldcs(&p->meta); return h; }
__device__ __forceinline__ void st_hit(HitRec *p, float4 wuvt, uint32_t inst, uint32_t tri) {
    __stcs(&p->wuvt, wuvt);
    __stcs(&p->meta, make_uint4(inst, tri, 0u, 0u));
}
__device__ __forceinline__ PathRec ld_path(const PathRec *p) { PathRec r; r.throughput = __ldcs(&p->throughput); r.meta = __ldcs(&p->meta); return r; }

__device__ __forceinline__ void warp_add_stat(TraceCtl *ctl, int idx, uint32_t v) {
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0 && v) atomicAdd(&ctl->stats[idx], (unsigned long long)v);
}

// Pull the next 32-item unit from a queue head; returns the unit's first item index.
__device__ __forceinline__ uint32_t next_unit(uint32_t *head) {
    uint32_t u = 0;
    if (lane_id() == 0) u = atomicAdd(head, 32u);
    return __shfl_sync(0xFFFFFFFFu, u, 0);
}

// Units that shrink when the queue runs dry (PC_ADAPTIVE_UNITS, OFF).  Idea: a launch ends with a tail in which a few
// warps each serialise the divergent walks of a 32-ray unit, so once fewer than 32 rays per resident warp are left,
// smaller units (and smaller k_shade tiles) would spread the same rays over more warps.  MEASURED AND REJECTED
// (profiles/ab_r01i.txt): -9 % on config 2 (3 860 -> 3 512 Mrays/s), -27 % on config 1, -6 % on config 3, -3 % on the 4K
// frame, -2 % even with a single sample chain: the tail is not idle -- the other chains' kernels run in it -- and the
// extra warp instructions of the small units cost more than the shorter tail saves.  Kept for the record.
#ifndef PC_ADAPTIVE_UNITS
#define PC_ADAPTIVE_UNITS 0
#endif
struct UnitClaim {
    uint32_t seen;  // queue position after this warp's last claim (0 before the first)
};
__device__ __forceinline__ uint32_t next_unit_adaptive(uint32_t *head, uint32_t n, UnitClaim &c, uint32_t &size) {
    uint32_t u = 0, sz = 32u;
    if (lane_id() == 0) {
#if PC_ADAPTIVE_UNITS
        const uint32_t warps = gridDim.x * (blockDim.x >> 5);
        const uint32_t left = n > c.seen ? n - c.seen : 0u;
        sz = left >= 32u * warps ? 32u : left >= 16u * warps ? 16u : left >= 8u * warps ? 8u : 4u;
#endif
        u = atomicAdd(head, sz);
    }
    size = __shfl_sync(0xFFFFFFFFu, sz, 0);
    u = __shfl_sync(0xFFFFFFFFu, u, 0);
    c.seen = u + size;
    return u;
}

#ifndef PC_SHADE_TU  // non-template kernels live in the exact-arithmetic translation unit only (pc_host.cu)
// ------------------------------------------------------------------------------------------------
__global__ void k_begin_sample(TraceCtl *ctl, unsigned long long *status, size_t statusWords, uint32_t sampleAdvance, uint32_t nSlots,
                               uint32_t slotStride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    if (i == 0) {  // a chain traces batches of nSlots samples: curSample, curSample + slotStride, ...
        ctl->curSample = ctl->nextSample;
        ctl->nextSample = ctl->nextSample + sampleAdvance;
        ctl->nSlots = nSlots;
        ctl->slotStride = slotStride;
    }
    if (i < 2 * MAX_BOUNCES + 2) ctl->queueHead[i] = 0;
    if (i < MAX_BOUNCES) ctl->ticket[i] = 0;
    for (size_t k = i; k < statusWords; k += stride) status[k] = 0ull;
}

#endif  // !PC_SHADE_TU

// ------------------------------------------------------------------------------------------------
// Persistent per-ray traversal with warp-level refilling.
//
// A warp keeps up to 32 rays in flight and every lane runs the Trav state machine of pc_device.cuh.
// Measured on the Cornell configs (ncu, profiles/) a plain "32 rays per warp, while-while" loop keeps
// only 8.7 of 32 lanes busy: lanes that found their triangle leaf wait for the slowest lane's search,
// lanes with a 2-triangle leaf wait for a 10-triangle one, finished lanes wait for the longest ray.
// The schedule below (modelled lane by lane on real bounce rays with tools/simt_model.py before it was
// written: 0.29 -> 0.62 SIMD efficiency) attacks all three:
//   * refill    -- when fewer than PC_REFILL_THRESHOLD lanes hold a ray the idle lanes take new rays
//                  from the global queue.  The queue's atomicAdd is issued one batch AHEAD and only
//                  consumed (shuffled out of lane 0) when the previous batch is used up, so its
//                  latency overlaps traversal instead of stalling the warp;
//   * phase 1   -- search: inner-node steps, instance entries and exit markers, until the lane holds
//                  a triangle leaf.  Bounded: once fewer than PC_SEARCH_MIN lanes are still searching
//                  while others already hold a leaf, the warp moves on and the stragglers resume
//                  their search in the next round;
//   * phase 2   -- every lane that holds a triangle leaf tests its triangles; finished rays are
//                  committed and their lanes become idle.
// Which lane traces which ray changes nothing: a ray's result depends on that ray alone and is written
// to the slot of its index.
//   Source::load(i, o, d, tmax)   fetch (or generate) ray i
//   Sink::store(i, hit, trav)     consume the finished traversal of ray i
// ------------------------------------------------------------------------------------------------
// MEASURED (B200, config 2, profiles/ab_r01_traversal.txt): the refilling schedule executes 23 % fewer warp
// instructions and raises lane efficiency from 8.7 to 13.6 of 32, but its extra per-lane state costs 79
// registers instead of 64 (6 instead of 8 resident blocks per SM) and the kernel is latency bound, so IPC
// falls with occupancy (2.18 -> 1.51) and k_query gets SLOWER (129 -> 148 us).  It therefore stays off by
// default (PC_TRACE_REFILL=0: fixed 32-ray units through traverse()) until the state fits 64 registers.
// ROUND 2 (profiles/ab_r02d.txt): with the stack in shared memory the refilling loop fits 64 registers.  Measured per
// kernel class: it still loses on short / coherent walks (k_primary everywhere; k_trace on the Cornell configs: 136 -> 180 us)
// and WINS on long incoherent ones (k_trace on the instancing config 1 424 -> 1 165 us, on the terrain 1 088 -> 1 023 us).
// It is therefore a run-time choice of the fused bounce kernel only (PC_OPT_TRACE_REFILL: 0 off, 1 on, -1 = by scene size).
#ifndef PC_REFILL_AUTO_MIN_NODES
#define PC_REFILL_AUTO_MIN_NODES 8192   // inner BVH nodes from which the automatic policy switches k_trace to refilling
#endif
#ifndef PC_REFILL_THRESHOLD
#define PC_REFILL_THRESHOLD 16
#endif
#ifndef PC_SEARCH_MIN
#define PC_SEARCH_MIN 8
#endif
#define PC_QUEUE_BATCH 32u

// The per-thread traversal stack of a kernel: PC_SMEM_STACK entries in shared memory (+ a local spill array for deeper
// walks), or a plain local array when PC_SMEM_STACK == 0 (see pc_device.cuh above SmemStack).
#if PC_SMEM_STACK
#define PC_TRAV_STACK(name)                                                         \
    __shared__ uint32_t name##_smem[PC_SMEM_STACK * TRAV_BLOCK];                    \
    uint32_t name##_spill[PC_STACK_SIZE > PC_SMEM_STACK ? PC_STACK_SIZE - PC_SMEM_STACK : 1]; \
    const SmemStack<PC_SMEM_STACK, TRAV_BLOCK> name{name##_smem + threadIdx.x, name##_spill}
#else
#define PC_TRAV_STACK(name)                \
    uint32_t name##_local[PC_STACK_SIZE];  \
    uint32_t *const name = name##_local
#endif

// Fixed units: a warp pulls 32 consecutive rays and every lane walks its ray with traverse().
template <bool ANY_HIT, bool COUNT, class Stack, class Source, class Sink>
__device__ __forceinline__ void trace_queue_units(const DScene &sc, Stack stack, uint32_t *head, uint32_t n, TravStats &st, Source &src, Sink &sink,
                                                  const uint32_t *perm = nullptr) {
    UnitClaim claim{0u};
    for (;;) {
        uint32_t size;
        const uint32_t unit = next_unit_adaptive(head, n, claim, size);
        if (unit >= n) break;
        const uint32_t j = unit + lane_id();
        if (lane_id() < size && j < n) {
            const uint32_t i = perm ? __ldcs(perm + j) : src.map(j);  // which ray a lane walks changes nothing: results go to slot i
            float3 o, d;
            float tmax;
            src.load(i, o, d, tmax);
            Trav t;
            const int hit = traverseWith<ANY_HIT, COUNT>(sc, stack, o, d, tmax, t.best, st);
            t.tmaxRay = tmax;
            sink.store(i, hit, t);
        }
    }
}
// Refilling schedule (see the note above)
template <bool ANY_HIT, bool COUNT, class Stack, class Source, class Sink>
__device__ __forceinline__ void trace_queue_refill(const DScene &sc, Stack stack, uint32_t *head, uint32_t n, TravStats &st, Source &src, Sink &sink,
                                                   const uint32_t *perm = nullptr) {
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = lane_id();
    const unsigned ltMask = (1u << lane) - 1u;
    Trav t;
    t.cur = REF_DONE; t.sp = 0;
    uint32_t rayIndex = 0;
    bool busy = false;          // this lane holds a live ray
    bool exhausted = false;     // warp-uniform: the queue has nothing left for this warp
    uint32_t poolNext = 0, poolEnd = 0;  // warp-uniform: ray indices already claimed from the queue
    uint32_t ahead = 0;         // lane 0: base of the batch claimed one step ahead
    if (lane == 0) ahead = atomicAdd(head, PC_QUEUE_BATCH);
    for (;;) {
        // ---- refill idle lanes
        const unsigned idle = __ballot_sync(FULL, !busy);
        if (idle && !exhausted) {
            const uint32_t want = __popc(idle), myRank = __popc(idle & ltMask);
            uint32_t assigned = 0;
            while (assigned < want) {
                if (poolNext == poolEnd) {
                    const uint32_t base = __shfl_sync(FULL, ahead, 0);
                    if (base >= n) { exhausted = true; break; }
                    poolNext = base;
                    poolEnd = min(base + PC_QUEUE_BATCH, n);
                    if (lane == 0) ahead = atomicAdd(head, PC_QUEUE_BATCH);
                }
                const uint32_t take = min(want - assigned, poolEnd - poolNext);
                if (!busy && myRank >= assigned && myRank < assigned + take) {
                    const uint32_t j = poolNext + (myRank - assigned);
                    const uint32_t i = perm ? __ldcs(perm + j) : src.map(j);
                    float3 o, d;
                    float tmax;
                    src.load(i, o, d, tmax);
                    travInit(t, sc, o, d, tmax);
                    rayIndex = i;
                    busy = true;
                }
                poolNext += take;
                assigned += take;
            }
        }
        if (!__any_sync(FULL, busy)) break;
        // ---- rounds of (search, triangles) until too few lanes are busy
        for (;;) {
            for (;;) {  // phase 1
                const bool searching = busy && !refIsTriLeaf(t.cur) && t.cur != REF_DONE;
                const unsigned sm = __ballot_sync(FULL, searching);
                if (sm == 0u) break;
                if (__popc(sm) < PC_SEARCH_MIN && __any_sync(FULL, busy && refIsTriLeaf(t.cur))) break;
                if (searching) {
                    if (!(t.cur & REF_LEAF)) {
                        travInner<ANY_HIT, COUNT>(t, sc, stack, st);
                    } else if (travOther<ANY_HIT, COUNT>(t, sc, stack, st)) {
                        t.cur = REF_DONE;
                    }
                }
            }
            if (busy && (refIsTriLeaf(t.cur) || t.cur == REF_DONE)) {  // phase 2
                const int r = t.cur == REF_DONE ? 1 : travTris<ANY_HIT, COUNT>(t, sc, stack, st);
                if (r) {
                    const int hit = ANY_HIT ? (r == 2 ? 1 : 0) : (t.best.wuvt.w < t.tmaxRay ? 1 : 0);
                    sink.store(rayIndex, hit, t);
                    busy = false;
                }
            }
            const unsigned live = __ballot_sync(FULL, busy);
            if (live == 0u) break;
            if (!exhausted && __popc(live) < PC_REFILL_THRESHOLD) break;
        }
    }
}
template <bool ANY_HIT, bool COUNT, bool REFILL = false, class Stack, class Source, class Sink>
__device__ __forceinline__ void trace_queue(const DScene &sc, Stack stack, uint32_t *head, uint32_t n, TravStats &st, Source &src, Sink &sink,
                                            const uint32_t *perm = nullptr) {
    if (REFILL) trace_queue_refill<ANY_HIT, COUNT>(sc, stack, head, n, st, src, sink, perm);
    else trace_queue_units<ANY_HIT, COUNT>(sc, stack, head, n, st, src, sink, perm);
}

struct RaySource {  // rays[i] as stored by k_primary / k_shade
    const Ray *rays;
    __device__ __forceinline__ uint32_t map(uint32_t j) const { return j; }
    __device__ __forceinline__ void load(uint32_t i, float3 &o, float3 &d, float &tmax) const {
        const Ray r = ld_ray(rays + i);
        o = xyz(r.origin); d = xyz(r.dir); tmax = r.origin.w;
    }
};
template <bool COUNT>
struct HitSink {  // hitFlag + Intersection record (intersect.cl:345-346)
    uint32_t *hitFlags;
    HitRec *hits;
    uint32_t missed;
    __device__ __forceinline__ void store(uint32_t i, int hit, const Trav &t) {
        __stcs(hitFlags + i, (uint32_t)hit);
        st_hit(hits + i, t.best.wuvt, t.best.inst, t.best.tri);
        if (COUNT && !hit) missed++;
    }
};

// ------------------------------------------------------------------------------------------------
// Warp-packet closest-hit traversal: one node sequence per warp, stack of (reference, lane mask)
// in shared memory, votes by ballot, front-to-back order by majority.  A lane only works on nodes
// its own slab test accepted (the mask), so per lane the visited set -- and therefore the hit --
// is the one the per-ray traversal finds.
// ------------------------------------------------------------------------------------------------
template <bool COUNT>
__device__ __forceinline__ int traverse_packet(const DScene &sc, uint2 *stack, bool valid, float3 o0, float3 d0,
                                               float tmaxRay, Hit &best, TravStats &st) {
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lbit = 1u << lane_id();
    int sp = 0;
    float3 o = o0, d = d0;
    float3 invDir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    uint32_t curInst = 0, curRank = 0;
    best.wuvt = make_float4(0.0f, 0.0f, 0.0f, tmaxRay);
    best.inst = 0; best.tri = 0; best.rank = 0;
    uint32_t cur = sc.rootRef;
    unsigned curMask = __ballot_sync(FULL, valid);
    if (curMask == 0) return 0;
    for (;;) {
        const bool mine = (curMask & lbit) != 0;
        if (!(cur & REF_LEAF)) {
            const float4 *np = sc.node64 + 4 * (size_t)cur;
            float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3);
            float tl = FLT_MAX, tr = FLT_MAX;
            if (mine) {
                if (COUNT) st.nodes++;
                tl = slabEntry(xyz(q0), xyz(q1), o, invDir, tmaxRay);
                tr = slabEntry(xyz(q2), xyz(q3), o, invDir, tmaxRay);
                float lim = best.wuvt.w * PC_CULL_SLACK;
                if (tl > lim) tl = FLT_MAX;
                if (tr > lim) tr = FLT_MAX;
            }
            bool wl = tl < FLT_MAX, wr = tr < FLT_MAX;
            unsigned bl = __ballot_sync(FULL, wl), br = __ballot_sync(FULL, wr);
            uint32_t lref = f2u(q0.w), rref = f2u(q1.w);
            if (bl && br) {
                unsigned prefL = __ballot_sync(FULL, wl && (!wr || tl <= tr));
                unsigned prefR = __ballot_sync(FULL, wr && (!wl || tr < tl));
                bool leftFirst = __popc(prefL) >= __popc(prefR);
                if (lane_id() == 0) stack[sp] = leftFirst ? make_uint2(rref, br) : make_uint2(lref, bl);
                sp++;
                cur = leftFirst ? lref : rref;
                curMask = leftFirst ? bl : br;
                continue;
            }
            if (bl | br) {
                cur = bl ? lref : rref;
                curMask = bl ? bl : br;
                continue;
            }
        } else if (cur == REF_POP_INSTANCE) {
            o = o0; d = d0;
            invDir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        } else if (cur & REF_TOP) {
            curInst = cur & 0x3FFFFFFFu;
            if (COUNT && mine) st.instances++;
            const float4 *ip = sc.inst80 + 5 * (size_t)curInst;
            float4 hdr = __ldg(ip);
            curRank = f2u(hdr.z);
            if (!(f2u(hdr.y) & INST_FLAG_IDENTITY)) {
                float4 m0 = __ldg(ip + 1), m1 = __ldg(ip + 2), m2 = __ldg(ip + 3), m3 = __ldg(ip + 4);
                o = mul4x1(o, m0, m1, m2, m3);
                d = mul3x1(d, m0, m1, m2);
                invDir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                if (lane_id() == 0) stack[sp] = make_uint2(REF_POP_INSTANCE, FULL);
                sp++;
            }
            cur = f2u(hdr.x);
            continue;
        } else {
            uint32_t tri = cur & 0x3FFFFFFFu;
            const float4 *tp = sc.tri48 + 3 * (size_t)tri;
            float4 a = __ldg(tp);
            uint32_t count = f2u(a.w);
            for (;;) {
                float4 b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (mine) {
                    if (COUNT) st.tris++;
                    float u, v, t;
                    if (triTest(xyz(a), xyz(b), xyz(c), o, d, u, v, t) && t > PC_EPS) {
                        float bt = best.wuvt.w;
                        bool closer = t < bt;
                        bool tie = (t == bt) && bt < tmaxRay && (curRank < best.rank || (curRank == best.rank && tri < best.tri));
                        if (closer || tie) {
                            best.wuvt = make_float4(1.0f - (u + v), u, v, t);
                            best.tri = tri; best.inst = curInst; best.rank = curRank;
                        }
                    }
                }
                if (--count == 0) break;
                tri++;
                tp += 3;
                a = __ldg(tp);
            }
        }
        if (sp == 0) break;
        --sp;
        __syncwarp();
        uint2 e = stack[sp];
        cur = e.x;
        curMask = e.y;
        __syncwarp();
    }
    return best.wuvt.w < tmaxRay ? 1 : 0;
}

// generatePrimaryRays (camera.cl:5-58) as a ray source: ray i is pixel (i % frameW, i / frameW) of the block
// PC_PRIMARY_TILES (default on): the j-th item of the work queue is not ray j but the ray of an 8 x 4 PIXEL TILE, so that the
// 32 lanes of a warp walk a compact bundle instead of a 32 x 1 strip of the image (fewer instance / leaf boundaries inside
// a warp).  A bijection on every sample slot's index range: rows [0, blockH & ~3) are tiled when frameW % 8 == 0, the ragged
// rest keeps the linear order.  Nothing else changes -- ray i, its path record and its hit record live at slot i either way.
// MEASURED (profiles/ab_r02j.txt): k_primary 4 391 -> 3 788 us on config 3, 1 800 -> 1 585 us on config 4, 456 -> 437 us on
// config 2 (frames +2.3 / +2.5 / +0.1 %).  Making 2 x 2 groups of tiles consecutive in the queue adds nothing (ab_r02k.txt).
#ifndef PC_PRIMARY_TILES
#define PC_PRIMARY_TILES 1
#endif
struct PrimarySource {
    FrameBufs fb;
    CameraParams cam;
    uint32_t frameW, blockY, slotPaths;
    const uint32_t *seeds;      // camera seed of slot s at seeds[(curSample + s * slotStride) * seedsPerSample]
    uint32_t curSample, slotStride, seedsPerSample;
    uint32_t tiledRays;         // how many of a slot's rays are tiled (see tiledCount): whole tiles only, the ragged rest stays linear
    static __device__ __forceinline__ uint32_t tiledCount(uint32_t frameW, uint32_t blockH) {
        return (frameW & 7u) == 0u ? frameW * (blockH & ~3u) : 0u;
    }
    __device__ __forceinline__ uint32_t map(uint32_t j) const {
#if PC_PRIMARY_TILES
        const uint32_t slot = j / slotPaths, local = j - slot * slotPaths;
        if (local >= tiledRays) return j;
        const uint32_t l = local & 31u;
        const uint32_t tile = local >> 5, tilesX = frameW >> 3;
        const uint32_t ty = tile / tilesX, tx = tile - ty * tilesX;
        return slot * slotPaths + (ty * 4u + (l >> 3)) * frameW + tx * 8u + (l & 7u);
#else
        return j;
#endif
    }
    __device__ __forceinline__ void load(uint32_t index, float3 &o, float3 &d, float &tmax) const {
        const uint32_t slot = index / slotPaths, local = index - slot * slotPaths;  // ray index == slot * slotPaths + path index
        const uint32_t gx = local % frameW, gy = local / frameW;
        const uint32_t randSeed = __ldg(seeds + (size_t)(curSample + slot * slotStride) * seedsPerSample);
        d = primaryRayDir(cam, gx, gy, blockY, randSeed);
        o = cam.eye;
        tmax = FLT_MAX;
        st_ray(fb.rays[0] + index, f4(cam.eye, FLT_MAX), f4(d, (float)local));  // rayNew (util/ray.cl:13-16)
        // pathNew (util/path.cl:13-17)
        __stcs(&fb.paths[index].throughput, make_float4(1.0f, 1.0f, 1.0f, 0.0f));
        __stcs(&fb.paths[index].meta, make_uint4((gy + blockY) * frameW + gx, 0u, 0u, 0u));
    }
};

// rayIntersectionTest over rays[2] + accumulateEmissiveSamples for the unoccluded ones.
// At most one occlusion ray per path and bounce, so the accumulator update needs no atomic
// (same argument as the reference, pt_integrator.cl:294-295).  hitFlags may be null.
template <bool COUNT>
struct OcclusionSink {
    const Ray *rays;
    const PathRec *paths;
    const float4 *emissiveSamples;
    float4 *const *slotAcc;   // the slots' accumulators (shared memory); null: flags only (test hook)
    uint32_t *hitFlags;
    uint32_t unocc;
    const uint32_t *base;     // slots' first rays in `rays` (shared memory)
    uint32_t slotPaths;
    __device__ __forceinline__ void store(uint32_t i, int hit, const Trav &) {
        if (hitFlags) hitFlags[i] = (uint32_t)hit;
        if (!hit && slotAcc) {
            const uint32_t slot = slotOf(base, i);
            const uint32_t pathIndex = (uint32_t)__ldcs(&rays[i].dir.w);  // rayGetPathIndex (util/ray.cl:26-28)
            const uint32_t pixel = paths[(size_t)slot * slotPaths + pathIndex].meta.x;
            const float4 s = __ldcs(emissiveSamples + i);
            float4 *acc = slotAcc[slot];
            float4 c = acc[pixel];
            c.x += s.x; c.y += s.y; c.z += s.z;
            acc[pixel] = c;
            if (COUNT) unocc++;
        }
    }
};
// the slots' first rays of rays[k] and the slots' accumulators, staged in shared memory once per CTA (dynamic indexing
// of a kernel parameter would put it in local memory); call before the first use, syncs the CTA
__device__ __forceinline__ void loadSlotBases(uint32_t *dst, float4 **accDst, const TraceCtl *ctl, int k, const FrameBufs &fb) {
    if (threadIdx.x <= MAX_SLOTS) dst[threadIdx.x] = ctl->base[k][threadIdx.x];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < MAX_SLOTS; q++) accDst[q] = fb.slotAcc[q];
    }
    __syncthreads();
}

// MODE 0: per-ray traversal, 1: warp packets over 8x4 pixel tiles, 2: reference-order per-ray.
// occSlot >= 0 (MODE 0 only): the launch ALSO drains the any-hit queue the PREVIOUS sample of this chain left behind -- its last
// bounce's occlusion test + emissive accumulation (pipeline.go:160-165), which nothing in the next sample depends on and
// which as a launch of its own is a pure tail (ncu: 9 % of the warps active, IPC 0.1).  Every warp first pulls primary
// units, then occlusion units.  The two halves touch disjoint state, with one benign overlap: pathNew rewrites
// paths[i].meta.x (the pixel index) with the value the occlusion half reads there -- it is the same for every sample of a
// block.  The last sample's queue is drained by a stand-alone k_occlusion at the end of pc_trace.
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(TRAV_BLOCK, PC_PRIMARY_MIN_BLOCKS) k_primary(DScene sc, FrameBufs fb, TraceCtl *ctl, const uint32_t *seeds,
                                                       const TraceParams *params, uint32_t seedsPerSample, int queueSlot, int occSlot, int sorted) {
    __shared__ uint2 s_stack[MODE == 1 ? (TRAV_BLOCK / 32) * PC_STACK_SIZE : 1];
    const CameraParams cam = params->cam;
    const uint32_t frameW = params->frameW, blockY = params->blockY, blockH = params->blockH;
    const uint32_t n = frameW * blockH;                       // paths per slot
    const uint32_t nSlots = MODE == 0 ? ctl->nSlots : 1u;      // packets / the literal walk trace one sample per launch
    const uint32_t total = n * nSlots;
    const uint32_t curSample = ctl->curSample, slotStride = ctl->slotStride;
    const uint32_t randSeed = seeds[(size_t)curSample * seedsPerSample];
    __shared__ uint32_t s_base2[MAX_SLOTS + 1];
    __shared__ float4 *s_acc[MAX_SLOTS];
    if (occSlot >= 0) loadSlotBases(s_base2, s_acc, ctl, 2, fb);  // the previous batch's occlusion rays
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->numRays[0] = (int)total;  // camera.cl:24-26
        ctl->slotPaths = n;
        for (uint32_t k = 0; k <= MAX_SLOTS; k++) ctl->base[0][k] = (k < nSlots ? k : nSlots) * n;
        atomicAdd(&ctl->stats[ST_QUERY_RAYS], (unsigned long long)total);
    }
    TravStats st{0, 0, 0};
    uint32_t missed = 0;
    if (MODE == 0) {
        PC_TRAV_STACK(stack);
        PrimarySource src{fb, cam, frameW, blockY, n, seeds, curSample, slotStride, seedsPerSample,
                          PrimarySource::tiledCount(frameW, blockH)};
        HitSink<COUNT> sink{fb.hitFlags, fb.hits, 0u};
        trace_queue<false, COUNT>(sc, stack, &ctl->queueHead[queueSlot], total, st, src, sink);
        missed = sink.missed;
        if (occSlot >= 0) {
            const uint32_t nO = (uint32_t)ctl->numRays[2];
            if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->stats[ST_OCCLUSION_RAYS], (unsigned long long)nO);
            OcclusionSink<COUNT> osink{fb.rays[2], fb.paths, fb.emissiveSamples, s_acc, nullptr, 0u, s_base2, n};
            RaySource osrc{fb.rays[2]};
            trace_queue<true, COUNT>(sc, stack, &ctl->queueHead[occSlot], nO, st, osrc, osink, sorted ? fb.permOcc : nullptr);
            if (COUNT) warp_add_stat(ctl, ST_UNOCCLUDED, osink.unocc);
        }
    } else {
        const uint32_t tilesX = (frameW + 7) / 8, tilesY = (blockH + 3) / 4;
        const uint32_t totalItems = MODE == 1 ? tilesX * tilesY * 32u : n;
        for (;;) {
            uint32_t unit = next_unit(&ctl->queueHead[queueSlot]);
            if (unit >= totalItems) break;
            uint32_t gx, gy;
            bool valid;
            if (MODE == 1) {
                uint32_t tile = unit / 32u;
                gx = (tile % tilesX) * 8u + (lane_id() & 7u);
                gy = (tile / tilesX) * 4u + (lane_id() >> 3);
                valid = gx < frameW && gy < blockH;
            } else {
                uint32_t i = unit + lane_id();
                valid = i < n;
                gx = i % frameW;
                gy = i / frameW;
            }
            const uint32_t index = gy * frameW + gx;
            float3 dir = f3(0.0f, 0.0f, 1.0f);
            if (valid) {
                dir = primaryRayDir(cam, gx, gy, blockY, randSeed);
                st_ray(fb.rays[0] + index, f4(cam.eye, FLT_MAX), f4(dir, (float)index));  // rayNew (util/ray.cl:13-16)
                // pathNew (util/path.cl:13-17)
                __stcs(&fb.paths[index].throughput, make_float4(1.0f, 1.0f, 1.0f, 0.0f));
                __stcs(&fb.paths[index].meta, make_uint4((gy + blockY) * frameW + gx, 0u, 0u, 0u));
            }
            Hit best;
            int hit = 0;
            if (MODE == 1) {
                hit = traverse_packet<COUNT>(sc, s_stack + (threadIdx.x / 32) * PC_STACK_SIZE, valid, cam.eye, dir, FLT_MAX, best, st);
            } else if (valid) {
                hit = traverseReference<false>(sc, cam.eye, dir, FLT_MAX, best);
            }
            if (valid) {
                __stcs(fb.hitFlags + index, (uint32_t)hit);
                st_hit(fb.hits + index, best.wuvt, best.inst, best.tri);
                if (COUNT && !hit) missed++;
            }
        }
    }
    if (COUNT) {
        warp_add_stat(ctl, ST_NODES, st.nodes);
        warp_add_stat(ctl, ST_TRIS, st.tris);
        warp_add_stat(ctl, ST_INSTANCES, st.instances);
        warp_add_stat(ctl, ST_MISSED, missed);
    }
}

// rayIntersectionQuery over rays[a][0 .. numRays[a])
template <bool REFERENCE, bool COUNT>
__global__ void __launch_bounds__(TRAV_BLOCK, PC_TRAV_MIN_BLOCKS) k_query(DScene sc, const Ray *rays, uint32_t *hitFlags, HitRec *hits,
                                                     TraceCtl *ctl, int a, int queueSlot, const uint32_t *perm) {
    const uint32_t n = (uint32_t)ctl->numRays[a];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->stats[ST_QUERY_RAYS], (unsigned long long)n);
    TravStats st{0, 0, 0};
    uint32_t missed = 0;
    if (!REFERENCE) {
        PC_TRAV_STACK(stack);
        RaySource src{rays};
        HitSink<COUNT> sink{hitFlags, hits, 0u};
        trace_queue<false, COUNT>(sc, stack, &ctl->queueHead[queueSlot], n, st, src, sink, perm);
        missed = sink.missed;
    } else {
        for (;;) {
            uint32_t unit = next_unit(&ctl->queueHead[queueSlot]);
            if (unit >= n) break;
            uint32_t i = unit + lane_id();
            if (i < n) {
                float4 ro = rays[i].origin, rd = rays[i].dir;
                Hit best;
                int hit = traverseReference<false>(sc, xyz(ro), xyz(rd), ro.w, best);
                hitFlags[i] = (uint32_t)hit;
                HitRec h;
                h.wuvt = best.wuvt;
                h.meta = make_uint4(best.inst, best.tri, 0u, 0u);
                hits[i] = h;
                if (COUNT && !hit) missed++;
            }
        }
    }
    if (COUNT) {
        warp_add_stat(ctl, ST_NODES, st.nodes);
        warp_add_stat(ctl, ST_TRIS, st.tris);
        warp_add_stat(ctl, ST_INSTANCES, st.instances);
        warp_add_stat(ctl, ST_MISSED, missed);
    }
}

template <bool REFERENCE, bool COUNT>
__global__ void __launch_bounds__(TRAV_BLOCK, PC_OCC_MIN_BLOCKS) k_occlusion(DScene sc, const Ray *rays, const PathRec *paths,
                                                         const float4 *emissiveSamples, FrameBufs fbAcc, int accumulate, uint32_t *hitFlags,
                                                         TraceCtl *ctl, int queueSlot, const uint32_t *perm) {
    const uint32_t n = (uint32_t)ctl->numRays[2];
    __shared__ uint32_t s_base2[MAX_SLOTS + 1];
    __shared__ float4 *s_acc[MAX_SLOTS];
    loadSlotBases(s_base2, s_acc, ctl, 2, fbAcc);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctl->stats[ST_OCCLUSION_RAYS], (unsigned long long)n);
    TravStats st{0, 0, 0};
    OcclusionSink<COUNT> sink{rays, paths, emissiveSamples, accumulate ? s_acc : nullptr, hitFlags, 0u, s_base2, ctl->slotPaths};
    if (!REFERENCE) {
        PC_TRAV_STACK(stack);
        RaySource src{rays};
        trace_queue<true, COUNT>(sc, stack, &ctl->queueHead[queueSlot], n, st, src, sink, perm);
    } else {
        Trav dummy;
        for (;;) {
            uint32_t unit = next_unit(&ctl->queueHead[queueSlot]);
            if (unit >= n) break;
            uint32_t i = unit + lane_id();
            if (i < n) {
                float4 ro = rays[i].origin, rd = rays[i].dir;
                Hit best;
                int hit = traverseReference<true>(sc, xyz(ro), xyz(rd), ro.w, best);
                sink.store(i, hit, dummy);
            }
        }
    }
    if (COUNT) {
        warp_add_stat(ctl, ST_NODES, st.nodes);
        warp_add_stat(ctl, ST_TRIS, st.tris);
        warp_add_stat(ctl, ST_INSTANCES, st.instances);
        warp_add_stat(ctl, ST_UNOCCLUDED, sink.unocc);
    }
}

// k_trace: the bounce's two independent traversal stages in ONE persistent launch -- the closest-hit
// query of the indirect rays (rays[a]) and the any-hit test of the occlusion rays (rays[2]) with its
// emissive accumulation.  Neither reads what the other writes (hits / hit flags vs. the trace
// accumulator), the reference only runs them back to back because its host loop is serial
// (pipeline.go:160-165 then :203-209).  Two queue heads: every warp drains the closest-hit queue first (the
// longer walks), then the any-hit queue, whose cheap rays make the launch's tail.  Saves one launch tail per bounce.
// REFILL: the warp-refilling schedule of trace_queue_refill instead of fixed 32-ray units.  Same results; which is faster
// depends on the scene (PC_OPT_TRACE_REFILL).
template <bool COUNT, bool REFILL>
__global__ void __launch_bounds__(TRAV_BLOCK, PC_TRAV_MIN_BLOCKS) k_trace(DScene sc, FrameBufs fb, TraceCtl *ctl, int a, int queueSlot, int sorted) {
    const uint32_t nQ = (uint32_t)ctl->numRays[a], nO = (uint32_t)ctl->numRays[2];
    __shared__ uint32_t s_base2[MAX_SLOTS + 1];
    __shared__ float4 *s_acc[MAX_SLOTS];
    loadSlotBases(s_base2, s_acc, ctl, 2, fb);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&ctl->stats[ST_QUERY_RAYS], (unsigned long long)nQ);
        atomicAdd(&ctl->stats[ST_OCCLUSION_RAYS], (unsigned long long)nO);
    }
    TravStats st{0, 0, 0};
    PC_TRAV_STACK(stack);
    // the closest-hit queue first (the longer walks) ...
    RaySource qsrc{fb.rays[a]};
    HitSink<COUNT> qsink{fb.hitFlags, fb.hits, 0u};
    trace_queue<false, COUNT, REFILL>(sc, stack, &ctl->queueHead[queueSlot], nQ, st, qsrc, qsink, sorted ? fb.permInd : nullptr);
    // ... then the any-hit queue, whose short walks make the launch's tail
    RaySource osrc{fb.rays[2]};
    OcclusionSink<COUNT> osink{fb.rays[2], fb.paths, fb.emissiveSamples, s_acc, nullptr, 0u, s_base2, ctl->slotPaths};
    trace_queue<true, COUNT, REFILL>(sc, stack, &ctl->queueHead[queueSlot + 1], nO, st, osrc, osink, sorted ? fb.permOcc : nullptr);
    const uint32_t missed = qsink.missed, unocc = osink.unocc;
    if (COUNT) {
        warp_add_stat(ctl, ST_NODES, st.nodes);
        warp_add_stat(ctl, ST_TRIS, st.tris);
        warp_add_stat(ctl, ST_INSTANCES, st.instances);
        warp_add_stat(ctl, ST_MISSED, missed);
        warp_add_stat(ctl, ST_UNOCCLUDED, unocc);
    }
}

// ------------------------------------------------------------------------------------------------
// Decoupled look-back over the shade TILES' (occlusion, indirect) totals.  A tile is the 32 rays a
// warp shades at once; tiles are handed out by ticket, so tile t covers rays [32t, 32t+32) and every
// predecessor of a tile is owned by a warp that is already running (or done): spinning on a
// predecessor cannot deadlock, also when the grid is not fully resident.
// status word: [63:62] 0 empty / 1 aggregate / 2 inclusive prefix, [61:31] occlusion, [30:0] indirect
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_status(unsigned long long flag, uint32_t occ, uint32_t ind) {
    return (flag << 62) | ((unsigned long long)occ << 31) | (unsigned long long)ind;
}
__device__ __forceinline__ void lookback(volatile unsigned long long *status, uint32_t ticket, uint32_t occTot,
                                         uint32_t indTot, uint32_t &occBase, uint32_t &indBase) {
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = lane_id();
    if (ticket == 0) {
        if (lane == 0) status[0] = pack_status(2ull, occTot, indTot);
        occBase = 0; indBase = 0;
        return;
    }
    if (lane == 0) status[ticket] = pack_status(1ull, occTot, indTot);
    uint32_t accO = 0, accI = 0;
    int base = (int)ticket - 1;
    for (;;) {
        int idx = base - (int)lane;
        unsigned long long v = pack_status(2ull, 0u, 0u);  // before tile 0: an inclusive prefix of zero
        if (idx >= 0) {
            do { v = status[idx]; } while ((v >> 62) == 0ull);
        }
        unsigned incMask = __ballot_sync(FULL, (v >> 62) == 2ull);
        int firstInc = incMask ? (__ffs((int)incMask) - 1) : 32;
        uint32_t o = 0, i = 0;
        if ((int)lane <= firstInc) {
            o = (uint32_t)((v >> 31) & 0x7FFFFFFFull);
            i = (uint32_t)(v & 0x7FFFFFFFull);
        }
        accO += __reduce_add_sync(FULL, o);
        accI += __reduce_add_sync(FULL, i);
        if (incMask) break;
        base -= 32;
    }
    occBase = accO; indBase = accI;
    if (lane == 0) status[ticket] = pack_status(2ull, accO + occTot, accI + indTot);
}

// ------------------------------------------------------------------------------------------------
// k_shade: shadePrimaryRayMisses / shadeIndirectRayMisses / shadeHits for rays[a][0 .. numRays[a]),
// persistent over CTA tiles of SHADE_TILE consecutive rays handed out by ticket.
//
// A warp that simply shades 32 consecutive rays meets up to six different materials (plus misses) and
// executes every BxDF / light / texture path present one after the other: measured on config 2 (ncu,
// profiles/) 13 of 32 lanes were active and 26 % of the stalls were instruction fetch (100 KB of shading
// code wanted by all warps all the time).  Sorting the tile by material first executes 38 % fewer warp
// instructions.
//   1. each thread reads the (hit flag, triangle) of its SHADE_RPT rays and takes the triangle's material
//      root as sort key; a counting sort over the CTA (warp match + shared atomics for the per-key counts,
//      a 64-entry scan) yields a permutation that groups equal materials;
//   2. the warps pull 32-ray chunks of the sorted tile from a shared counter and shade them: a chunk holds
//      (mostly) one material, and a warp with a cheap chunk takes the next one instead of idling;
//   3. results are staged in shared memory at the ray's ORIGINAL slot, compacted in original order
//      (ballot + popc per 32-slot group, offsets across the tile, ONE decoupled look-back per tile) and
//      written out -- the output order is parent ray order, whichever lane did the arithmetic for a ray.
// ------------------------------------------------------------------------------------------------
constexpr int SHADE_WARPS = SHADE_BLOCK / 32;
constexpr int SHADE_KEYS = 64;  // histogram bins: material roots modulo 61, roulette victims, misses, inactive lanes
constexpr uint32_t KEY_KILLED = SHADE_KEYS - 3, KEY_MISS = SHADE_KEYS - 2, KEY_INACTIVE = SHADE_KEYS - 1;
// PC_SORT_RR (default on): from the bounce where Russian roulette starts, a ray that the roulette is going to kill gets a sort
// key of its own.  The decision (pt_integrator.cl:113-124) depends on the path's throughput and on the third draw of the
// ray's random stream only -- not on the surface -- so the sort phase can foresee it; shadeHit still makes the decision
// itself, the key is a scheduling hint and cannot change a result.  Without it half of the lanes of every chunk return
// at the roulette and the BxDF / light-sampling code that follows -- most of shadeHit -- runs half empty (ncu, config 2:
// 11.7 of 32 lanes in the launch of bounce 3 against 24-27 before the roulette starts).
#ifndef PC_SORT_RR
#define PC_SORT_RR 1
#endif
// PC_SORT_LEAF: key a hit by the LEAF of its material tree instead of the root, as far as the leaf can be foreseen in the sort
// phase (constant mixes draw from the ray's own random stream; texture-weighted mixes and disperse nodes end the forecast).
// A layered material then no longer puts its diffuse, conductor and dielectric lanes into the same chunk.  Also only a hint.
#ifndef PC_SORT_LEAF
#define PC_SORT_LEAF 0
#endif
#ifndef PC_SHADE_PREFETCH
#define PC_SHADE_PREFETCH 0
#endif
// Rays per thread and tile.  A tile is SHADE_BLOCK * SHADE_RPT consecutive rays; after the sort the CTA's
// warps PULL 32-ray chunks of the sorted tile from a shared counter, so a warp that drew a cheap material
// (diffuse) takes the next chunk instead of waiting at the tile's barrier for the warp that drew the
// rough dielectric: with one chunk per warp (the r01c kernel) ncu showed 52 % of the warp stalls at
// barriers.  SHADE_RPT * SHADE_WARPS <= 32 (one warp scans the per-group counts).
// MEASURED (B200, config 2 at 128 spp, profiles/ab_r01de.txt), Mrays/s for (threads, rays per thread, CTAs per SM):
// (256,1,3) 3216 = the r01c kernel | (256,2,3) 3571 | (256,3,3) 3547 | (256,4,2 @80 regs) 3428 | (256,2,2) 3546 |
// (256,3,2) 3679 | (256,4,2) 3833 <- default | (512,2,1) 3705 | (128,4,4) 3497 | (128,4,5) 3268 | (128,8,3) 3224.
// ncu on the default: 28.3 of 32 lanes active (21.3 before), barrier stalls 26 % (52 %), IPC 1.00 (0.75).
#ifndef PC_SHADE_RPT
#define PC_SHADE_RPT 4
#endif
constexpr int SHADE_RPT = PC_SHADE_RPT;
constexpr int SHADE_TILE = SHADE_BLOCK * SHADE_RPT;
constexpr int SHADE_GROUPS = SHADE_TILE / 32;
static_assert(SHADE_GROUPS <= 32, "one warp scans the per-group counts");

// Traversal-order key of an emitted ray (PC_OPT_SORT_RAYS): which octant of the scene it starts in, which octant it
// points to, and its dominant axis.  k_shade writes the rays themselves in parent order (the order is part of the result:
// it fixes every ray's random stream) and, next to them, a permutation of each tile's output range sorted by
// this key; the traversal kernels pull their 32-ray units through it, so a warp walks rays that start close together and
// point the same way.  Modelled on real bounce rays of config 2 (tools/simt_model.py): -19 % warp instructions for the
// indirect rays, -25 % for the occlusion rays.
constexpr int SORT_BINS = 256;
__device__ __forceinline__ uint32_t sortKey(const DScene &sc, float ox, float oy, float oz, float dx, float dy, float dz) {
    const uint32_t cx = (ox - sc.worldMin.x) * sc.worldCellScale.x >= 1.0f ? 1u : 0u;
    const uint32_t cy = (oy - sc.worldMin.y) * sc.worldCellScale.y >= 1.0f ? 1u : 0u;
    const uint32_t cz = (oz - sc.worldMin.z) * sc.worldCellScale.z >= 1.0f ? 1u : 0u;
    const uint32_t oct = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    const uint32_t axis = ax >= ay ? (ax >= az ? 0u : 2u) : (ay >= az ? 1u : 2u);
    return ((cx | (cy << 1) | (cz << 2)) << 5) | (oct << 2) | axis;
}

struct ShadeShared {
    float occ[10][SHADE_TILE];   // origin.xyz, maxDist, dir.xyz, sample.xyz   (struct of arrays: conflict free)
    float ind[6][SHADE_TILE];    // origin.xyz, dir.xyz
    float pathIndexF[SHADE_TILE];
    uint32_t hitTri[SHADE_TILE];  // triangle of the hit, 0xFFFFFFFF for a miss
    uint16_t perm[SHADE_TILE];
    uint8_t flags[SHADE_TILE];    // bit 0 occlusion ray wanted, bit 1 indirect ray wanted
    uint32_t hist[SHADE_KEYS];    // per-key counts, then the keys' first positions in perm
    uint32_t groupOcc[SHADE_GROUPS], groupInd[SHADE_GROUPS];  // per 32-slot group: counts, then exclusive offsets
    uint32_t binOcc[SORT_BINS], binInd[SORT_BINS];  // traversal-order sort of the emitted rays: per-key counts, then first positions
    uint32_t tile, occBase, indBase, nextChunk, activeChunks;
    uint32_t baseA[MAX_SLOTS + 1];   // first ray of every sample slot in rays[a]
    uint32_t seed[MAX_SLOTS];        // the slots' shadeHits seeds of this bounce (pipeline.go:146)
    float4 *acc[MAX_SLOTS];          // the slots' accumulators
};

// k_shade is compiled in its own translation unit (pc_shade.cu) so that it can carry its own floating-point flags:
// traversal, ray generation, merge and tonemap -- everything whose results are compared BIT FOR BIT with the CPU oracle --
// stay IEEE exact (-fmad=false -prec-div=true -prec-sqrt=true); shading is compared within 1e-4 relative and may use the
// hardware's 2-ulp division / square root, like the reference's own OpenCL build does for its native_recip / native_sqrt
// / `/` in these very functions (Makefile: SHADE_FP).  The host side reaches it through these two functions.
void shade_configure(const cudaDeviceProp &prop, int *blocksPerSM);
void shade_launch(bool count, int grid, cudaStream_t s, const DScene &sc, const FrameBufs &fb, TraceCtl *ctl, const uint32_t *seeds,
                  unsigned long long *status, uint32_t seedsPerSample, uint32_t bounce, uint32_t minBouncesForRR, int a, int fixQ4,
                  int sortRays);
const char *shade_fp_mode();

#ifdef PC_SHADE_TU
template <bool COUNT>
__global__ void __launch_bounds__(SHADE_BLOCK, SHADE_MIN_BLOCKS) k_shade(DScene sc, FrameBufs fb, TraceCtl *ctl, const uint32_t *seeds,
                                                      unsigned long long *status, uint32_t seedsPerSample, uint32_t bounce,
                                                      uint32_t minBouncesForRR, int a, int fixQ4, int sortRays) {
    extern __shared__ __align__(16) unsigned char shade_smem[];
    ShadeShared &sh = *reinterpret_cast<ShadeShared *>(shade_smem);
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, tid = threadIdx.x;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint32_t n = (uint32_t)ctl->numRays[a];
    // (PC_ADAPTIVE_UNITS, off: measured slower, see next_unit_adaptive) tile size follows the launch: the largest tile that
    // still gives every CTA at least two tiles
#if PC_ADAPTIVE_UNITS
    uint32_t rpt = SHADE_RPT;
    while (rpt > 1u && n < 2u * gridDim.x * rpt * SHADE_BLOCK) rpt >>= 1;
#else
    const uint32_t rpt = SHADE_RPT;
#endif
    const uint32_t tileRays = rpt * SHADE_BLOCK;
    const uint32_t nTiles = (n + tileRays - 1u) / tileRays;
    if (n == 0 && blockIdx.x == 0 && tid == 0) {  // resources.go:230-238: both counters reset
        ctl->numRays[2] = 0;
        ctl->numRays[1 - a] = 0;
        for (int k = 0; k <= MAX_SLOTS; k++) { ctl->base[2][k] = 0u; ctl->base[1 - a][k] = 0u; }
    }
    const uint32_t slotPaths = ctl->slotPaths;
    const bool oneSlot = ctl->nSlots <= 1u;  // uniform: no slot lookup per ray
    if (tid <= MAX_SLOTS) sh.baseA[tid] = ctl->base[a][tid];
    if (tid < MAX_SLOTS) sh.seed[tid] = tid < ctl->nSlots ? seeds[(size_t)(ctl->curSample + tid * ctl->slotStride) * seedsPerSample + 1 + bounce] : 0u;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < MAX_SLOTS; q++) sh.acc[q] = fb.slotAcc[q];
    }
    uint32_t shaded = 0;
    for (;;) {
        __syncthreads();  // the previous tile's shared state is no longer in use
        if (tid == 0) { sh.tile = atomicAdd(&ctl->ticket[bounce], 1u); sh.nextChunk = 0u; }
        if (tid < SHADE_KEYS) sh.hist[tid] = 0u;
        if (sortRays && tid < SORT_BINS) { sh.binOcc[tid] = 0; sh.binInd[tid] = 0; }  // SORT_BINS <= SHADE_BLOCK
        __syncthreads();
        const uint32_t tile = sh.tile;
        if (tile >= nTiles) break;
        const uint32_t base = tile * tileRays;
        // ---- 1. sort keys of my SHADE_RPT rays (slot = r * SHADE_BLOCK + tid: coalesced), counted per key.
        //         Which thread later shades which ray does not matter (results are staged at the ray's slot),
        //         so the position inside a key's range is simply the order of arrival.
        uint32_t key[SHADE_RPT], rank[SHADE_RPT];
#pragma unroll
        for (int r = 0; r < SHADE_RPT; r++) {
            if ((uint32_t)r >= rpt) break;
            const uint32_t slot = r * SHADE_BLOCK + tid, i = base + slot;
            uint32_t k = KEY_INACTIVE, tri = 0xFFFFFFFFu;
            if (i < n) {
                k = KEY_MISS;
#if PC_SHADE_PREFETCH
                // MEASURED AND REJECTED (profiles/ab_r02o.txt): the shading phase starts every ray with a dependent pair of DRAM
                // reads -- the ray (for its path index) and then the path record, which sits wherever that index points.  Asking
                // for both here, into L2 only, so that they arrive while the tile is sorted, made k_shade SLOWER (698 -> 716 us on
                // config 2, 465 -> 477 us on the 4K frame): the extra load + slot lookup per ray in the sort phase cost more than
                // the shorter wait returns -- with two tiles per SM the other tile's shading already covers that latency.
                if (bounce < minBouncesForRR || !PC_SORT_RR) {
                    const uint32_t pi = (uint32_t)__ldcg(&fb.rays[a][i].dir.w);
                    const uint32_t ps = oneSlot ? 0u : slotOf(sh.baseA, i);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(fb.paths + (size_t)ps * slotPaths + pi));
                }
#endif
                if (__ldcs(fb.hitFlags + i)) {
                    tri = __ldcs(&fb.hits[i].meta).y;
                    uint32_t node = PC_LDG(sc.matIndex + tri);
#if PC_SORT_RR || PC_SORT_LEAF
                    const bool rr = PC_SORT_RR && bounce >= minBouncesForRR;  // uniform
                    if (PC_SORT_LEAF || rr) {
                        const uint32_t sslot = oneSlot ? 0u : slotOf(sh.baseA, i);
                        uint2 rnd = make_uint2(sh.seed[sslot], i - sh.baseA[sslot]);  // shadeHit's stream (pt_integrator.cl:81-84)
                        randomGetSample2f(rnd);
                        randomGetSample2f(rnd);
                        const float2 sample2 = randomGetSample2f(rnd);
                        bool killed = false;
                        if (rr) {
                            const uint32_t pathIndex = (uint32_t)__ldcs(&fb.rays[a][i].dir.w);
                            const float4 T = __ldcs(&fb.paths[(size_t)sslot * slotPaths + pathIndex].throughput);
                            const float rrProbability = cl_max(cl_min(0.5f, 0.2126f * T.x + 0.7152f * T.y + 0.0722f * T.z), 0.01f);
                            killed = rrProbability < sample2.x;
                        }
#if PC_SORT_LEAF
                        // matSelectNode's walk (material_sampler.cl:21-88) as far as it can be foreseen without the surface: constant
                        // mixes draw from the same stream, bump / normal maps do not draw; a mix map (texture weight) or a disperse
                        // node (path flags) ends the forecast and keys the ray by that node
                        for (int guard = 0; guard < 8 && !killed; guard++) {
                            const float4 hdr = PC_LDG(sc.matNodes + 4 * (size_t)node);
                            const uint32_t type = f2u(hdr.x);
                            if (type == OP_MIX) {
                                const float2 smp = randomGetSample2f(rnd);
                                node = smp.x < PC_LDG(sc.matNodes + 4 * (size_t)node + 1).x ? f2u(hdr.y) : f2u(hdr.z);
                            } else if (type == OP_BUMP_MAP || type == OP_NORMAL_MAP) {
                                node = f2u(hdr.y);
                            } else {
                                break;
                            }
                        }
#endif
                        if (killed) node = 0xFFFFFFFFu;
                    }
#endif
                    k = node == 0xFFFFFFFFu ? KEY_KILLED : node % (uint32_t)(SHADE_KEYS - 3);
                }
            }
            sh.hitTri[slot] = tri;
            sh.flags[slot] = 0;
            const unsigned peers = __match_any_sync(FULL, k);
            const int leader = __ffs((int)peers) - 1;
            uint32_t first = 0;
            if ((int)lane == leader) first = atomicAdd(&sh.hist[k], (uint32_t)__popc(peers));
            first = __shfl_sync(FULL, first, leader);
            key[r] = k;
            rank[r] = first + (uint32_t)__popc(peers & ltMask);
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 64 key counts: 2 per lane
            const uint32_t v0 = sh.hist[2 * lane], v1 = sh.hist[2 * lane + 1], sum = v0 + v1;
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d);
                if ((int)lane >= d) incl += o;
            }
            const uint32_t run = incl - sum;
            sh.hist[2 * lane] = run;
            sh.hist[2 * lane + 1] = run + v0;
            if (lane == 31) sh.activeChunks = (run + v0 + 31u) / 32u;  // KEY_INACTIVE sorts last and is never shaded
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SHADE_RPT; r++)
            if ((uint32_t)r < rpt) sh.perm[sh.hist[key[r]] + rank[r]] = (uint16_t)(r * SHADE_BLOCK + tid);
        __syncthreads();
        // ---- 2. warps pull 32-ray chunks of the sorted tile
        const uint32_t activeChunks = sh.activeChunks;
        for (;;) {
            uint32_t c = 0;
            if (lane == 0) c = atomicAdd(&sh.nextChunk, 1u);
            c = __shfl_sync(FULL, c, 0);
            if (c >= activeChunks) break;
            const uint32_t slot = sh.perm[c * 32u + lane];
            const uint32_t i = base + slot;
            if (i >= n) continue;
            ShadeOut so;
            so.wantOcc = false; so.wantInd = false;
            const float4 rd = __ldcs(&fb.rays[a][i].dir);
            const uint32_t pathIndex = (uint32_t)rd.w;  // rayGetDirAndPathIndex (util/ray.cl:19-23)
            // the ray's sample slot: its path records, its accumulator, its seed and its index WITHIN the sample (the
            // reference's get_global_id(0), the RNG key of pt_integrator.cl:81)
            const uint32_t sslot = oneSlot ? 0u : slotOf(sh.baseA, i);
            const uint32_t pathAt = sslot * slotPaths + pathIndex;
            const uint32_t tri = sh.hitTri[slot];
            if (tri == 0xFFFFFFFFu) {
                if (sc.sceneDiffuseMat != -1) {  // pipeline.go:134-143
                    float3 kd = shadeMiss(sc, xyz(rd));
                    PathRec p = ld_path(fb.paths + pathAt);
                    float3 add = bounce == 0 ? kd : xyz(p.throughput) * kd;
                    float4 *acc = sh.acc[sslot];
                    float4 cc = acc[p.meta.x];
                    cc.x += add.x; cc.y += add.y; cc.z += add.z;
                    acc[p.meta.x] = cc;
                }
                continue;
            }
            shaded++;
            {
                const uint32_t rayIndex = i - sh.baseA[sslot];
                const uint32_t randSeed = sh.seed[sslot];
                const float4 wuvt = __ldcs(&fb.hits[i].wuvt);
                const PathRec p = ld_path(fb.paths + pathAt);
                shadeHit(sc, xyz(rd), xyz(p.throughput), p.meta.y, wuvt, tri, rayIndex, bounce, minBouncesForRR, randSeed, so);
                if (so.accum) {
                    const uint32_t dst = fixQ4 ? p.meta.x : pathIndex;  // pt_integrator.cl:106, SURVEY Q4
                    float4 *acc = sh.acc[sslot];
                    float4 cc = acc[dst];
                    cc.x += so.accumAdd.x; cc.y += so.accumAdd.y; cc.z += so.accumAdd.z;
                    acc[dst] = cc;
                }
            }
            if (so.flagsChanged) fb.paths[pathAt].meta.y = so.pathFlags;
            if (so.wantInd) fb.paths[pathAt].throughput = f4(so.newThroughput, 0.0f);
            // ---- 3. stage at the ORIGINAL slot
            sh.flags[slot] = (uint8_t)((so.wantOcc ? 1u : 0u) | (so.wantInd ? 2u : 0u));
            sh.pathIndexF[slot] = rd.w;
            if (so.wantOcc) {
                sh.occ[0][slot] = so.occOrigin.x; sh.occ[1][slot] = so.occOrigin.y; sh.occ[2][slot] = so.occOrigin.z;
                sh.occ[3][slot] = so.occMaxDist;
                sh.occ[4][slot] = so.occDir.x; sh.occ[5][slot] = so.occDir.y; sh.occ[6][slot] = so.occDir.z;
                sh.occ[7][slot] = so.occSample.x; sh.occ[8][slot] = so.occSample.y; sh.occ[9][slot] = so.occSample.z;
            }
            if (so.wantInd) {
                sh.ind[0][slot] = so.indOrigin.x; sh.ind[1][slot] = so.indOrigin.y; sh.ind[2][slot] = so.indOrigin.z;
                sh.ind[3][slot] = so.indDir.x; sh.ind[4][slot] = so.indDir.y; sh.ind[5][slot] = so.indDir.z;
            }
        }
        __syncthreads();
        // ---- 4. stable compaction in original order (replaces pt_integrator.cl:162,176,188-197): warp w owns the
        //         32-slot groups w, w + SHADE_WARPS, ...
        unsigned occMask[SHADE_RPT], indMask[SHADE_RPT];
        uint32_t sortOcc[SHADE_RPT], sortInd[SHADE_RPT];  // key | rank inside the key's bin << 8
#pragma unroll
        for (int r = 0; r < SHADE_RPT; r++) {
            if ((uint32_t)r >= rpt) break;
            const uint32_t g = r * SHADE_WARPS + warp;
            const uint32_t f = sh.flags[g * 32u + lane];
            occMask[r] = __ballot_sync(FULL, (f & 1u) != 0u);
            indMask[r] = __ballot_sync(FULL, (f & 2u) != 0u);
            if (lane == 0) { sh.groupOcc[g] = (uint32_t)__popc(occMask[r]); sh.groupInd[g] = (uint32_t)__popc(indMask[r]); }
            sortOcc[r] = 0u; sortInd[r] = 0u;
            if (sortRays) {  // key, and the arrival rank inside the key's bin
                const uint32_t slot = g * 32u + lane;
                if (f & 1u) {
                    const uint32_t k = sortKey(sc, sh.occ[0][slot], sh.occ[1][slot], sh.occ[2][slot], sh.occ[4][slot], sh.occ[5][slot], sh.occ[6][slot]);
                    sortOcc[r] = k | (atomicAdd(&sh.binOcc[k], 1u) << 8);
                }
                if (f & 2u) {
                    const uint32_t k = sortKey(sc, sh.ind[0][slot], sh.ind[1][slot], sh.ind[2][slot], sh.ind[3][slot], sh.ind[4][slot], sh.ind[5][slot]);
                    sortInd[r] = k | (atomicAdd(&sh.binInd[k], 1u) << 8);
                }
            }
        }
        __syncthreads();
        if (sortRays && (warp == 1 || warp == 2)) {  // exclusive scans of the two 256-entry bin tables, 8 entries per lane
            uint32_t *bins = warp == 1 ? sh.binOcc : sh.binInd;
            uint32_t v[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { v[k] = bins[lane * 8 + k]; sum += v[k]; }
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d);
                if ((int)lane >= d) incl += o;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) { bins[lane * 8 + k] = run; run += v[k]; }
        }
        if (warp == 0) {
            const uint32_t groups = rpt * SHADE_WARPS;
            const uint32_t o = lane < groups ? sh.groupOcc[lane] : 0u, q = lane < groups ? sh.groupInd[lane] : 0u;
            uint32_t io = o, iq = q;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t uo = __shfl_up_sync(FULL, io, d), uq = __shfl_up_sync(FULL, iq, d);
                if ((int)lane >= d) { io += uo; iq += uq; }
            }
            if (lane < groups) { sh.groupOcc[lane] = io - o; sh.groupInd[lane] = iq - q; }
            const uint32_t occTot = __shfl_sync(FULL, io, 31), indTot = __shfl_sync(FULL, iq, 31);
            uint32_t occBase, indBase;
            lookback(status, tile, occTot, indTot, occBase, indBase);
            if (lane == 0) {
                sh.occBase = occBase; sh.indBase = indBase;
                if (tile == nTiles - 1) {  // the last tile publishes the queue lengths
                    ctl->numRays[2] = (int)(occBase + occTot);
                    ctl->numRays[1 - a] = (int)(indBase + indTot);
                    ctl->base[2][0] = 0u;
                    ctl->base[1 - a][0] = 0u;
                    for (int k = 1; k <= MAX_SLOTS; k++)  // slots that start at the end of the input (empty / unused ones)
                        if (sh.baseA[k] >= n) { ctl->base[2][k] = occBase + occTot; ctl->base[1 - a][k] = indBase + indTot; }
                    if (COUNT) {
                        atomicAdd(&ctl->stats[ST_OCC_EMITTED], (unsigned long long)(occBase + occTot));
                        atomicAdd(&ctl->stats[ST_IND_EMITTED], (unsigned long long)(indBase + indTot));
                    }
                }
            }
        }
        __syncthreads();
        // where every sample slot's rays start in the two OUTPUT buffers: stable compaction keeps the slots grouped, so a
        // slot's first output ray is the output position of its first input ray.  The warp owning that ray's 32-slot group
        // publishes it.
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < SHADE_RPT; r++) {
                if ((uint32_t)r >= rpt) break;
                const uint32_t g = r * SHADE_WARPS + warp;
                for (int k = 1; k <= MAX_SLOTS; k++) {
                    const uint32_t B = sh.baseA[k];
                    if (B >= n || B < base + g * 32u || B >= base + g * 32u + 32u) continue;
                    const unsigned below = (1u << (B - base - g * 32u)) - 1u;
                    ctl->base[2][k] = sh.occBase + sh.groupOcc[g] + (uint32_t)__popc(occMask[r] & below);
                    ctl->base[1 - a][k] = sh.indBase + sh.groupInd[g] + (uint32_t)__popc(indMask[r] & below);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < SHADE_RPT; r++) {
            if ((uint32_t)r >= rpt) break;
            const uint32_t g = r * SHADE_WARPS + warp, slot = g * 32u + lane;
            const float pif = sh.pathIndexF[slot];
            if (occMask[r] & (1u << lane)) {  // pt_integrator.cl:200-204
                const uint32_t k = sh.occBase + sh.groupOcc[g] + (uint32_t)__popc(occMask[r] & ltMask);
                __stcs(fb.emissiveSamples + k, make_float4(sh.occ[7][slot], sh.occ[8][slot], sh.occ[9][slot], 0.0f));
                st_ray(fb.rays[2] + k, make_float4(sh.occ[0][slot], sh.occ[1][slot], sh.occ[2][slot], sh.occ[3][slot]),
                       make_float4(sh.occ[4][slot], sh.occ[5][slot], sh.occ[6][slot], pif));
                if (sortRays) fb.permOcc[sh.occBase + sh.binOcc[sortOcc[r] & 0xFFu] + (sortOcc[r] >> 8)] = k;
            }
            if (indMask[r] & (1u << lane)) {  // :207-210
                const uint32_t k = sh.indBase + sh.groupInd[g] + (uint32_t)__popc(indMask[r] & ltMask);
                st_ray(fb.rays[1 - a] + k, make_float4(sh.ind[0][slot], sh.ind[1][slot], sh.ind[2][slot], FLT_MAX),
                       make_float4(sh.ind[3][slot], sh.ind[4][slot], sh.ind[5][slot], pif));
                if (sortRays) fb.permInd[sh.indBase + sh.binInd[sortInd[r] & 0xFFu] + (sortInd[r] >> 8)] = k;
            }
        }
    }
    if (COUNT) warp_add_stat(ctl, ST_SHADED, shaded);
}

#endif  // PC_SHADE_TU

#ifndef PC_SHADE_TU
// ------------------------------------------------------------------------------------------------
// tri48 on the device (pc_layout.hpp describes the record): the same float32 subtractions intersect.cl:255-256 does per
// test, from the vertices that were uploaded anyway, and the "triangles left in this leaf" lane from the BVH's mesh leaves
// ({min, -firstTriangle}{max, count}, optimized_scene.go:46-64).  Replaces a 48 B/triangle host pass + upload (0.5 GB for the
// 10 M-triangle terrain) when a scene is (re)uploaded.
// ------------------------------------------------------------------------------------------------
__global__ void k_derive_tri48(const float4 *__restrict__ vertices, float4 *__restrict__ tri48, size_t nTris) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nTris; t += stride) {
        const float4 v0 = vertices[3 * t], v1 = vertices[3 * t + 1], v2 = vertices[3 * t + 2];
        tri48[3 * t] = make_float4(v0.x, v0.y, v0.z, 0.0f);
        tri48[3 * t + 1] = make_float4(v1.x - v0.x, v1.y - v0.y, v1.z - v0.z, 0.0f);
        tri48[3 * t + 2] = make_float4(v2.x - v0.x, v2.y - v0.y, v2.z - v0.z, 0.0f);
    }
}
__global__ void k_leaf_counts(const float4 *__restrict__ bvhNodes, size_t nNodes, float4 *__restrict__ tri48, size_t nTris) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nNodes; i += stride) {
        const int l = __float_as_int(bvhNodes[2 * i].w), r = __float_as_int(bvhNodes[2 * i + 1].w);
        if (l <= 0 && r > 0) {  // mesh leaf: triangles [-l, -l + r); reachable leaves were validated on the host, unreachable
                                // nodes (meshes nobody instances) are only guarded
            const size_t first = (size_t)(-(long long)l);
            if (first + (size_t)r > nTris) continue;
            for (int j = 0; j < r; j++) tri48[3 * (first + j)].w = __uint_as_float((uint32_t)(r - j));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// accumulator.cl:5-19, hdr.cl:5-28
// ------------------------------------------------------------------------------------------------
__global__ void k_clear(float4 *acc, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
// dst[dstOff + g] += src[srcOff + g]; src may be a peer GPU's memory (loads cross NVLink)
__global__ void k_merge(float4 *__restrict__ dst, const float4 *__restrict__ src, size_t dstOff, size_t srcOff, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        float4 s = src[srcOff + g];
        float4 d = dst[dstOff + g];
        d.x += s.x; d.y += s.y; d.z += s.z;
        dst[dstOff + g] = d;
    }
}
// the sample chains' accumulators added to the trace accumulator in chain order, ((acc + c1) + c2) + ...: the same
// float32 sums as one k_merge per chain, in one pass over the block's rows
struct ChainAccs {
    const float4 *src[MAX_CHAINS_X_SLOTS];
    int n;
};
__global__ void k_merge_chains(float4 *__restrict__ dst, ChainAccs ca, size_t off, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        float4 d = dst[off + g];
        for (int c = 0; c < ca.n; c++) {
            const float4 s = __ldcs(ca.src[c] + off + g);
            d.x += s.x; d.y += s.y; d.z += s.z;
        }
        dst[off + g] = d;
    }
}
__global__ void k_tonemap(const float4 *acc, uchar4 *fb, size_t n, float sampleWeight, float exposure) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        fb[i] = tonemapReinhard(acc[i], sampleWeight, exposure);
}

// ------------------------------------------------------------------------------------------------
// kernels/debug.cl:16-156 -- the reference's debug visualisations (pc_trace_debug).  One thread per work-item,
// plain launches: a debugging aid, not a fast path.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uchar4 debugToneMap(float3 sample) {  // debugToneMapAndGammaCorrect (debug.cl:8-13)
    sample = sample * 1.0f;  // DEBUG_TONEMAP_EXPOSURE
    float3 mapped = sample / (sample + 1.0f);
    const float e = 1.0f / 2.2f;
    float3 v = f3(cl_clamp(powf(mapped.x, e), 0.0f, 1.0f), cl_clamp(powf(mapped.y, e), 0.0f, 1.0f), cl_clamp(powf(mapped.z, e), 0.0f, 1.0f)) * 255.0f;
    return make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z, 255);
}
__global__ void k_debug_clear(uchar4 *out, uint32_t n) {  // debugClearBuffer (:16-20)
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_uchar4(0, 0, 0, 255);
}
// max over the whole intersection buffer of the finite hit distances, start value 1 (resources.go:398-404); positive
// floats order like their bit patterns
__global__ void k_debug_max_depth(const HitRec *hits, uint32_t n, uint32_t *maxBits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float t = 1.0f;
    if (i < n) {
        float w = hits[i].wuvt.w;
        if (w != FLT_MAX && w > 1.0f) t = w;
    }
    for (int d = 16; d > 0; d >>= 1) t = fmaxf(t, __shfl_xor_sync(0xFFFFFFFFu, t, d));
    if (lane_id() == 0) atomicMax(maxBits, __float_as_uint(t));
}
__global__ void k_debug_depth(const TraceCtl *ctl, int a, const PathRec *paths, const uint32_t *hitFlags, const HitRec *hits,
                              const uint32_t *maxBits, uchar4 *out, uint32_t n) {  // debugRayIntersectionDepth (:23-48)
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n || (int)g >= ctl->numRays[a]) return;
    const uint32_t pixel = paths[g].meta.x;
    const float hitDist = hits[g].wuvt.w, maxDepth = __uint_as_float(*maxBits);
    if (!hitFlags[g] || hitDist == FLT_MAX) { out[pixel] = make_uchar4(0, 0, 0, 255); return; }
    const unsigned char sd = (unsigned char)(255.0f * (1.0f - hitDist / (maxDepth + 1.0f)));
    out[pixel] = make_uchar4(sd, sd, sd, 255);
}
__global__ void k_debug_normals(DScene sc, const TraceCtl *ctl, int a, const Ray *rays, PathRec *paths, const uint32_t *hitFlags,
                                const HitRec *hits, uchar4 *out, uint32_t n) {  // debugRayIntersectionNormals (:51-97)
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n || (int)g >= ctl->numRays[a]) return;
    const uint32_t pixel = paths[g].meta.x;
    const HitRec h = hits[g];
    if (!hitFlags[g] || h.wuvt.w == FLT_MAX) { out[pixel] = make_uchar4(0, 0, 0, 255); return; }
    Surface surface;
    surfaceInit(surface, h.wuvt, h.meta.y, sc);
    uint2 rnd = make_uint2(g, g);
    float3 tint = f3(1.0f, 1.0f, 1.0f);
    uint32_t flags = paths[g].meta.y;
    const uint32_t before = flags;
    matSelectNode(flags, surface, tint, sc, rnd);  // bump / normal maps perturb surface.normal; disperse sets path bits
    if (flags != before) paths[g].meta.y = flags;
    const float3 v = (surface.normal + 1.0f) * 255.0f * 0.5f;
    out[pixel] = make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z, 255);
}
__global__ void k_debug_emissive(const TraceCtl *ctl, const Ray *rays, const PathRec *paths, const uint32_t *hitFlags,
                                 const float4 *emissiveSamples, uint32_t maskOccluded, uint32_t maskNotOccluded, uchar4 *out,
                                 uint32_t n) {  // debugEmissiveSamples (:100-126)
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n || (int)g >= ctl->numRays[2]) return;
    const uint32_t pixel = paths[(uint32_t)rays[g].dir.w].meta.x;
    if ((maskOccluded && hitFlags[g]) || (maskNotOccluded && !hitFlags[g])) { out[pixel] = make_uchar4(0, 0, 0, 255); return; }
    out[pixel] = debugToneMap(xyz(emissiveSamples[g]));
}
__global__ void k_debug_throughput(const PathRec *paths, uchar4 *out, uint32_t n) {  // debugThroughput (:129-140)
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    out[paths[g].meta.x] = debugToneMap(xyz(paths[g].throughput));
}
__global__ void k_debug_accumulator(float sampleWeight, const float4 *acc, uchar4 *out, uint32_t n) {  // debugAccumulator (:143-154)
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    out[g] = debugToneMap(xyz(acc[g]) * sampleWeight);  // indexed by work-item, not by pixelIndex, like the reference
}

// ------------------------------------------------------------------------------------------------
// test hooks
// ------------------------------------------------------------------------------------------------
struct BxdfIn { float n[3]; uint32_t matNode; float in[3], p0; float out[3], p1; float rnd[2], uv[2]; };
struct BxdfOut { float sample[3], samplePdf; float dir[3], pdf; float eval[3], p; };
__global__ void k_debug_bxdf(DScene sc, const BxdfIn *in, BxdfOut *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Surface s;
    s.point = f3s(0.0f);
    s.normal = f3(in[i].n[0], in[i].n[1], in[i].n[2]);
    s.uv = make_float2(in[i].uv[0], in[i].uv[1]);
    s.matNodeIndex = in[i].matNode;
    MatNode m = loadMatNode(sc, in[i].matNode);
    float3 inDir = f3(in[i].in[0], in[i].in[1], in[i].in[2]), outDir = f3(in[i].out[0], in[i].out[1], in[i].out[2]);
    float3 dir = f3s(0.0f);
    float pdf = 1.0f;
    float3 smp = bxdfGetSample(s, m, sc, make_float2(in[i].rnd[0], in[i].rnd[1]), inDir, dir, pdf);
    float p = bxdfGetPdf(s, m, sc, inDir, outDir);
    float3 ev = bxdfEval(s, m, sc, inDir, outDir);
    BxdfOut o = {{smp.x, smp.y, smp.z}, pdf, {dir.x, dir.y, dir.z}, p, {ev.x, ev.y, ev.z}, 0.0f};
    out[i] = o;
}
__global__ void k_debug_rng(uint2 *states, uint32_t n, uint32_t draws, float2 *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 s = states[i];
    for (uint32_t d = 0; d < draws; d++) out[(size_t)i * draws + d] = randomGetSample2f(s);
    states[i] = s;
}
// primary-style packet traversal over an arbitrary ray list (32 consecutive rays per packet)
template <bool COUNT>
__global__ void __launch_bounds__(TRAV_BLOCK) k_debug_packet(DScene sc, const Ray *rays, uint32_t *hitFlags, HitRec *hits,
                                                            TraceCtl *ctl, uint32_t n, int queueSlot) {
    __shared__ uint2 s_stack[(TRAV_BLOCK / 32) * PC_STACK_SIZE];
    TravStats st{0, 0, 0};
    for (;;) {
        uint32_t unit = next_unit(&ctl->queueHead[queueSlot]);
        if (unit >= n) break;
        uint32_t i = unit + lane_id();
        bool valid = i < n;
        float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = make_float4(0.f, 0.f, 1.f, 0.f);
        if (valid) { ro = rays[i].origin; rd = rays[i].dir; }
        Hit best;
        int hit = traverse_packet<COUNT>(sc, s_stack + (threadIdx.x / 32) * PC_STACK_SIZE, valid, xyz(ro), xyz(rd), ro.w, best, st);
        if (valid) {
            hitFlags[i] = (uint32_t)hit;
            HitRec h;
            h.wuvt = best.wuvt;
            h.meta = make_uint4(best.inst, best.tri, 0u, 0u);
            hits[i] = h;
        }
    }
    if (COUNT) {
        warp_add_stat(ctl, ST_NODES, st.nodes);
        warp_add_stat(ctl, ST_TRIS, st.tris);
        warp_add_stat(ctl, ST_INSTANCES, st.instances);
    }
}

#endif  // !PC_SHADE_TU

}  // namespace pc
