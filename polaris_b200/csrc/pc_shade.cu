// pc_shade.cu -- the shading kernel's translation unit (see the note above k_shade in pc_kernels.cuh).
//
// Compiled with $(SHADE_FP) (Makefile).  Everything that must be bit-identical to the CPU oracle is in pc_host.cu.
#include <cuda_runtime.h>

#define PC_SHADE_TU 1
#include "pc_kernels.cuh"

#ifndef PC_SHADE_FP_MODE
#define PC_SHADE_FP_MODE "ieee"
#endif

namespace pc {

static_assert(SORT_BINS <= SHADE_BLOCK, "k_shade clears one bin per thread");

const char *shade_fp_mode() { return PC_SHADE_FP_MODE; }

void shade_configure(const cudaDeviceProp &prop, int *blocksPerSM) {
    // the shade tile's staging area exceeds the 48 KB static limit: opt in to the dynamic size for both instances
    cudaFuncSetAttribute(k_shade<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ShadeShared));
    cudaFuncSetAttribute(k_shade<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ShadeShared));
    // ask for just enough shared memory for SHADE_MIN_BLOCKS resident tiles; the rest of the 256 KB stays L1
    const size_t want = (size_t)SHADE_MIN_BLOCKS * (sizeof(ShadeShared) + 1024);
    int pct = (int)((want * 100 + prop.sharedMemPerMultiprocessor - 1) / prop.sharedMemPerMultiprocessor);
    if (pct > 100) pct = 100;
    cudaFuncSetAttribute(k_shade<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_shade<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_shade<false>, SHADE_BLOCK, sizeof(ShadeShared));
    if (perSM > SHADE_MIN_BLOCKS) perSM = SHADE_MIN_BLOCKS;
    if (perSM < 1) perSM = 1;
    *blocksPerSM = perSM;
}

void shade_launch(bool count, int grid, cudaStream_t s, const DScene &sc, const FrameBufs &fb, TraceCtl *ctl, const uint32_t *seeds,
                  unsigned long long *status, uint32_t seedsPerSample, uint32_t bounce, uint32_t minBouncesForRR, int a, int fixQ4,
                  int sortRays) {
    if (count)
        k_shade<true><<<grid, SHADE_BLOCK, sizeof(ShadeShared), s>>>(sc, fb, ctl, seeds, status, seedsPerSample, bounce, minBouncesForRR, a, fixQ4, sortRays);
    else
        k_shade<false><<<grid, SHADE_BLOCK, sizeof(ShadeShared), s>>>(sc, fb, ctl, seeds, status, seedsPerSample, bounce, minBouncesForRR, a, fixQ4, sortRays);
}

}  // namespace pc
