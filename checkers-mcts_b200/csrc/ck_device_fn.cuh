// ck_device_fn.cuh -- warp-level device helpers shared by the tree and glue kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ckb200.h"

namespace ck {

#define CK_FULL 0xFFFFFFFFu

// Checkers.predict (reference Checkers.py:436-437): sum of policy*mask over the 512 actions in
// numpy's float32 pairwise order (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum with
// PW_BLOCKSIZE 128 and 8 partial accumulators).  For n = 512 that order is: four blocks of
// 128; inside a block accumulator j sums a[j], a[8+j], ... in index order; the eight
// accumulators combine as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); blocks combine as
// (B0+B1)+(B2+B3).  Lane l owns accumulator j = l&7 of block l>>3, so the whole sum is 16
// sequential adds per lane and five xor-shuffles -- bit-exact with the reference.
// masked[i] receives the lane's masked inputs a[blk*128 + 8*i + j].  Returns the sum on all lanes.
__device__ __forceinline__ float masked_policy_sum(const float *__restrict__ policy, const uint32_t mask[8],
                                                   int lane, float masked[16]) {
    const int blk = lane >> 3, j = lane & 7;
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int plane = blk * 2 + (i >> 3), x = i & 7;      // action = plane*64 + x*8 + y, y = j
        const bool legal = ((x ^ j) & 1) && ((mask[plane] >> (4 * x + (j >> 1))) & 1u);
        const float v = policy[blk * 128 + i * 8 + j];
        masked[i] = legal ? v : __fmul_rn(v, 0.0f);
        r = (i == 0) ? masked[0] : __fadd_rn(r, masked[i]);
    }
    r = __fadd_rn(r, __shfl_xor_sync(CK_FULL, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(CK_FULL, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(CK_FULL, r, 4));
    r = __fadd_rn(r, __shfl_xor_sync(CK_FULL, r, 8));
    r = __fadd_rn(r, __shfl_xor_sync(CK_FULL, r, 16));
    return r;
}

}  // namespace ck
