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


// ---- peer-memory exchange (multi-GPU, P2P mode) ----------------------------------------------------------
__device__ __forceinline__ unsigned long long semb_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void semb_st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Wait until a peer-written epoch flag reaches `ep`.  Bounded: after SembScal::spin_limit clock ticks the kernel
// gives up, records it in SembScal::err (the host turns that into SEMB_ENCCL) and goes on with whatever is there --
// a lost or diverged peer must not hang every other rank inside a kernel.
__device__ __forceinline__ void semb_wait_epoch(const unsigned long long* flag, unsigned long long ep, SembScal* me) {
  if (semb_ld_acquire_sys(flag) >= ep) return;
  const long long lim = me->spin_limit, t0 = clock64();
  while (semb_ld_acquire_sys(flag) < ep) {
    if (lim > 0 && clock64() - t0 > lim) {
      atomicExch(&me->err, 1);
      return;
    }
  }
}

// Flag-in-data exchange of a double through peer memory (the idea of NCCL's LL protocol): the value travels as one
// 16-byte store {lo32, tag, hi32, tag}; each 8-byte half is self-validating (8-byte stores are never torn), so the
// consumer needs neither a fence on the producer side nor a separate flag -- it re-reads the entry until both tags
// carry this exchange's epoch.  The tag of epoch e differs from that of e-2 (the previous use of the same buffer)
// and is never 0 (the initial contents).
__device__ __forceinline__ unsigned semb_ll_tag(unsigned long long ep) { return (unsigned)(ep & 0x7fffffffull) | 0x80000000u; }
__device__ __forceinline__ void semb_ll_store(uint4* dst, double v, unsigned tag) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"((unsigned)__double2loint(v)), "r"(tag),
               "r"((unsigned)__double2hiint(v)), "r"(tag)
               : "memory");
}
// Bounded like semb_wait_epoch: after SembScal::spin_limit ticks it records SembScal::err and returns what is there.
__device__ __forceinline__ double semb_ll_load(const uint4* src, unsigned tag, SembScal* me) {
  unsigned a, b, c, d;
  long long t0 = 0;
  for (int spin = 0;; ++spin) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(src) : "memory");
    if (b == tag && d == tag) break;
    if (spin == 0) t0 = clock64();
    else if ((spin & 63) == 0 && me->spin_limit > 0 && clock64() - t0 > me->spin_limit) {
      atomicExch(&me->err, 1);
      break;
    }
  }
  return __hiloint2double((int)c, (int)a);
}

// All-gather of up to two doubles per rank through peer memory, executed by ONE block per rank (all of
// its threads must call): thread r stores this rank's values into rank r's mailbox (NVLink st.global),
// releases a flag, then waits for rank r's values to arrive in the local mailbox.  Afterwards thread 0
// combines in rank order (v0: sum, v1: max) => bitwise identical results on every rank.  This is the
// collective fused into the producing kernel: no NCCL launch.  kind: 0 = pap, 1 = {t, rmax}, 2 = reductions.
// The epoch is a device-side counter (SembScal::ep_dev[1 + kind]) advanced here, so the calling kernel's
// arguments never change between iterations (CUDA-graph replay on several ranks).
__device__ __forceinline__ void semb_p2p_allgather(SembScal* me, int kind, double v0, double v1, double* sum0,
                                                   double* max1, int tid) {
  const int nranks = me->nranks, rank = me->rank;
  const unsigned long long ep = me->ep_dev[1 + kind] + 1ull;
  const int par = (int)(ep & 1ull);
  if (tid < nranks) {
    SembScal* dst = me->peers[tid];
    unsigned long long* fdst;
    const unsigned long long* fsrc;
    if (kind == 0) {
      dst->box_pap[par][rank] = v0;
      fdst = &dst->flag_pap[rank];
      fsrc = &me->flag_pap[tid];
    } else if (kind == 1) {
      dst->box_t[par][2 * rank] = v0;
      dst->box_t[par][2 * rank + 1] = v1;
      fdst = &dst->flag_t[rank];
      fsrc = &me->flag_t[tid];
    } else {
      dst->box_red[par][2 * rank] = v0;
      dst->box_red[par][2 * rank + 1] = v1;
      fdst = &dst->flag_red[rank];
      fsrc = &me->flag_red[tid];
    }
    __threadfence_system();
    semb_st_release_sys(fdst, ep);
    semb_wait_epoch(fsrc, ep, me);
  }
  __syncthreads();  // (every reader of ep_dev is past its load: thread 0 may advance it)
  if (tid == 0) {
    double s = 0.0, m = 0.0;
    for (int r = 0; r < nranks; ++r) {
      if (kind == 0) {
        s += ((volatile double*)me->box_pap[par])[r];
      } else {
        const volatile double* b = (kind == 1) ? me->box_t[par] : me->box_red[par];
        s += b[2 * r];
        m = fmax(m, b[2 * r + 1]);
      }
    }
    *sum0 = s;
    if (max1) *max1 = m;
    me->ep_dev[1 + kind] = ep;
  }
}

// like semb_last_block, but tells EVERY thread of the block whether it is the last one (block-uniform),
// leaving the totals in shared memory (valid after the call in all threads of the last block)
__device__ __forceinline__ bool semb_last_block_uniform(double bsum, double bmax, double* psum, double* pmax,
                                                        unsigned* counter, int nblocks, int bid, double* red,
                                                        int tid, int nthreads, double* sh_tot /* 2 shared doubles */) {
  __shared__ int s_last2;
  if (tid == 0) {
    psum[bid] = bsum;
    if (pmax) pmax[bid] = bmax;
    __threadfence();
    unsigned ticket = atomicAdd(counter, 1u);
    s_last2 = (ticket == (unsigned)(nblocks - 1));
  }
  __syncthreads();
  if (!s_last2) return false;
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
    sh_tot[0] = s;
    sh_tot[1] = m2;
    *counter = 0u;
  }
  __syncthreads();
  return true;
}
