// ldeq_gridsum.cuh -- deterministic grid-wide sum for the cooperative (batch-global error norm) LatentODE kernels.
#pragma once
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace ldeq {

// grid-wide deterministic sum: every CTA publishes its partial, all meet, and warp 0 of every CTA adds the partials in
// the same fixed order (lane-strided, then a shuffle tree).  Only one warp per CTA reads them: with every thread of
// every CTA polling the same few cache lines the L2 slice that holds them serialises ~260 k requests per reduction
// (40 us).  The partials alternate between two halves of the buffer, so one grid barrier per reduction is enough.
#define GRID_SUM_HALF 2048
static __device__ __forceinline__ double grid_sum(double part, double* partials, cg::grid_group& grid, int& parity) {
    __shared__ double s_total;
    double* buf = partials + parity * GRID_SUM_HALF;
    parity ^= 1;
    if (threadIdx.x == 0) buf[blockIdx.x] = part;
    grid.sync();
    if (threadIdx.x < 32) {
        double r = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) r += __ldcg(buf + i);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
        if (threadIdx.x == 0) s_total = r;
    }
    __syncthreads();
    const double r = s_total;
    __syncthreads();  // s_total may be rewritten by the next reduction only after everyone has read it
    return r;
}

}  // namespace ldeq
