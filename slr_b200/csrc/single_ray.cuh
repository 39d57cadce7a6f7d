// One ray walked start to end by its own lane: the traversal of traverse.cuh (same steps, same arithmetic, same hit as the
// wave kernels) for kernels whose lanes each carry a whole path -- the tail kernel (tail.cu) and the bidirectional path
// tracer (bpt.cu).
#pragma once
#include "traverse.cuh"

namespace slrgpu {

template <bool ANY_HIT, bool INSTANCES, bool ALPHA, bool PREFETCH>
__device__ __noinline__ void singleRayWalkT(const DeviceScene& s, WalkState& w, uint32_t* stack, bool* overflow) {
    InstanceWalkState iw;
    iw.leaves.clear(); iw.saved.clear(); iw.curInst = SLRGPU_INVALID_ID;
    TraversalCounters cnt = {0, 0};
    bool ovf = false;
    walkBegin(w, stack);
    while (!walkStep<INSTANCES, ANY_HIT, false, ALPHA, PREFETCH>(s, w, iw, stack, cnt, ovf)) { }
    if (ovf) *overflow = true;
}
// w.r and w.time set by the caller; flat scenes take the leaner step without instance handling
template <bool ANY_HIT, bool PREFETCH = false>
__device__ __forceinline__ void singleRayWalk(const DeviceScene& s, WalkState& w, uint32_t* stack, bool* overflow) {
    if (s.hasAlpha) singleRayWalkT<ANY_HIT, true, true, PREFETCH>(s, w, stack, overflow);
    else if (s.numInstances != 0) singleRayWalkT<ANY_HIT, true, false, PREFETCH>(s, w, stack, overflow);
    else singleRayWalkT<ANY_HIT, false, false, PREFETCH>(s, w, stack, overflow);
}

}  // namespace slrgpu
