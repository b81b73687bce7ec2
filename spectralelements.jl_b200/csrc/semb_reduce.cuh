// Deterministic block / grid reductions and PCG scalar helpers shared by all libsemb kernels.
#pragma once
#include "semb_internal.cuh"

// Fixed-order block reductions (deterministic for a fixed block size): warp shuffles, then warp 0.
// tid / nthreads are the linear thread id and block size (blocks may be 2-D).
__device__ __forceinline__ double semb_block_sum(double v, double* red /* >= 32 doubles smem */, int tid,
                                                 int nthreads) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = tid >> 5, l = tid & 31, nw = (nthreads + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  if (w == 0) {
    s = (l < nw) ? red[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  }
  return s;  // valid in thread 0
}

__device__ __forceinline__ double semb_block_max(double v, double* red, int tid, int nthreads) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  const int w = tid >> 5, l = tid & 31, nw = (nthreads + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  if (w == 0) {
    s = (l < nw) ? red[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_down_sync(0xffffffffu, s, o));
  }
  return s;
}

// Publish this block's partial (sum and, optionally, max) and let the LAST block to arrive reduce
// all partials in a fixed order (independent of which block is last) => deterministic.
// Returns true in thread 0 of the last block with *tsum / *tmax set.
__device__ __forceinline__ bool semb_last_block(double bsum, double bmax, double* psum, double* pmax,
                                                unsigned* counter, int nblocks, int bid, double* red, int tid,
                                                int nthreads, double* tsum, double* tmax) {
  __shared__ int s_last;
  if (tid == 0) {
    psum[bid] = bsum;
    if (pmax) pmax[bid] = bmax;
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    s_last = (ticket == (unsigned)(nblocks - 1));
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double v = 0.0, mx = 0.0;
  for (int i = tid; i < nblocks; i += nthreads) {
    v += ((volatile double*)psum)[i];
    if (pmax) mx = fmax(mx, ((volatile double*)pmax)[i]);
  }
  const double s = semb_block_sum(v, red, tid, nthreads);
  double m2 = 0.0;
  if (pmax) m2 = semb_block_max(mx, red, tid, nthreads);
  if (tid == 0) {
    *tsum = s;
    if (tmax) *tmax = m2;
    *counter = 0u;
    return true;
  }
  return false;
}

// PCG beta (pcg.jl:46-50): first iteration copies h, later ones use t / t_prev.
__device__ __forceinline__ double semb_pcg_beta(const SembScal* s) {
  return (s->iters == 0) ? 0.0 : s->t / s->t_prev;
}

