// Fused tail of the strip kernel: everything that used to follow it in separate launches -- the x-seam kernel,
// the halo push, the y-seam kernel and (PCG) the final reduction / all-gather of sum(p .* Ap .* mult) -- is
// done by the strip kernel's own CTAs as they finish.  One apply = ONE launch.
//
// Replaces (reference /root/reference/src): the part of gatherScatter.jl:13 (QQtx * u * QQty') and mask.jl:14
// that crosses strip / chunk / rank boundaries, and the reduction of pcg.jl:52.
//
// Work is cut into TASKS, each owned by whichever CTA is the LAST to finish one of the chunks the task needs
// (an arrival counter per task; nobody waits for a CTA of the same GPU, so there is no co-residency requirement):
//   X(sx, c)  the two columns of x seam sx inside chunk c: x pairs, and for in-chunk y interfaces the y pair of the
//             two x pairs -- (a+b)+(c+d), the association of gatherScatter.jl:13 (QQtx before QQty, ABu.jl:14-33)
//   Y(k, s)   the two rows of y seam k (chunk boundary or periodic wrap) over the columns of strip s
//   C(k, sx)  the 2x2 corner where x seam sx crosses y seam k: (a+b)+(c+d) from the four raw values
// At a rank boundary the "other row" of Y and C is the neighbour's raw boundary row, which the neighbour's edge
// CTAs store straight into this rank's mailbox over NVLink from the registers that hold it (no x-summed
// intermediate, no push kernel) and publish with one epoch flag per strip segment; only then does a task wait
// (bounded, semb_wait_epoch) -- on the other GPU, never on this one.  The edge rows are computed FIRST (the top
// edge row is a one-row chunk its CTA marches through before the rest of its rows, and the edge CTA rows are
// scheduled first), so the rows are in flight for the whole kernel and the wait is normally zero.
// Every interface sum is 2-term, hence bit-identical to the dense QQ^T of the reference whichever CTA runs the task.
// PCG: each CTA and each task writes its partial of sum(p .* Ap .* mult) to its own slot; the last CTA of the
// grid adds the slots in index order (deterministic), all-gathers over the ranks and advances the epochs.
#pragma once
#include "semb_reduce.cuh"

__device__ __forceinline__ double semb_mask_at(const OpArgs& a, int x, int y, size_t idx) {
  if (a.M_arr) return a.M_arr[idx];
  return ((x == 0 && a.mx0) || (x == a.nxl - 1 && a.mx1) || (y == 0 && a.my0) || (y == a.nyl - 1 && a.my1)) ? 0.0
                                                                                                             : 1.0;
}

// What a CTA keeps (in shared memory) about each of the (up to) eight interface tasks of a chunk.  Everything but the
// outcome of the arrival is known when the chunk STARTS, so it is prepared then (semb_tail_prepare, off the critical
// path); semb_tail_announce only takes the ticket and drops the tasks this CTA did not arrive last at.
//   X: x0,x1 = the seam's two columns, ya..yb = the chunk's lines [ya, yb), flags bit0/bit1 = first/last line is a y seam
//   Y: x0..x1 = the strip's columns [x0, x1), ya,yb = the seam's two rows (rank boundary: ya only, flags bit2)
//   C: x0,x1 = the seam's two columns, ya,yb = the seam's two rows (rank boundary: ya only, flags bit2)
struct SembTailTask {
  int n;       // micro-tasks (32 items each); 0 = nothing to do (no such interface, or not the last arriver)
  int x0, x1, ya, yb, flags;
  int pbase;   // first PCG partial slot of the task
  int cnt;     // arrival counter of the task
  int target;  // arrivals it takes
};

// Threads 0..7 of the CTA describe the eight interfaces (slot & 7 = interface kind) of its main chunk cm in
// s_task[0..7], threads 8..15 those of its edge chunk ce (-1: none) in s_task[8..15].
static __device__ __noinline__ void semb_tail_prepare(const OpArgs& a, int s, int nstrips, int cm, int ce,
                                                      SembTailTask* s_task) {
  if (threadIdx.x >= 16) return;
  const int t = threadIdx.x & 7;
  const int ck = threadIdx.x < 8 ? cm : ce;
  if (ck < 0) return;
  const int r0 = a.chunk_r0[ck], r1 = a.chunk_r0[ck + 1];
  const int nxs = a.nxs, nch = a.nchunks, N = a.N;
  const SembTailLayout L(nstrips, nch, nxs, a.xmic_total);
  const int ls = s > 0 ? s - 1 : (a.perx ? nxs - 1 : -1);            // x seam on the left / right of this strip
  const int rs = s < nstrips - 1 ? s : (a.perx ? nxs - 1 : -1);
  const bool sb = a.ystart[r0] != 0, st = a.ystart[r1] != 0;         // the chunk's first / last line is on a y seam
  const int kb = ck, kt = (a.ywrap && ck == nch - 1) ? 0 : ck + 1;   // y seam ids below / above
  const bool hb = ck == 0 && a.has_lo, ht = ck == nch - 1 && a.has_hi;  // ... which is a rank boundary
  const int sx = (t & 1) ? rs : ls;                                  // the x seam of interfaces 0,1 and 4..7
  const bool top = (t == 3 || t >= 6);                               // interfaces on the chunk's upper y seam
  const int k = top ? kt : kb;
  const bool hal = top ? ht : hb, ys = top ? st : sb;
  SembTailTask T;
  T.n = 0, T.x0 = T.x1 = T.ya = T.yb = T.flags = T.pbase = T.cnt = T.target = 0;
  int xa = 0, xb = 0;
  if (sx >= 0) {
    if (sx < nstrips - 1) {
      const int e = semb_strip_e0(sx + 1, nstrips, a.Ex, N);
      xa = e * N - 1, xb = e * N;
    } else {
      xa = a.nxl - 1, xb = 0;  // periodic wrap
    }
  }
  // rows of y seam k: chunk boundary (or the local periodic wrap, k = 0); rank boundary: our edge row only
  const int ya = hal ? (k == 0 ? 0 : a.nyl - 1) : (k == 0 ? a.nyl - 1 : a.chunk_r0[k] * N - 1);
  const int yb = (k == 0) ? 0 : a.chunk_r0[k] * N;
  if (t < 2) {
    if (sx >= 0) {
      const int id = sx * nch + ck;
      T.n = a.xmic[ck + 1] - a.xmic[ck];
      T.x0 = xa, T.x1 = xb, T.ya = r0 * N, T.yb = r1 * N;
      T.flags = (sb ? 1 : 0) | (st ? 2 : 0);
      T.pbase = L.partX() + sx * a.xmic_total + a.xmic[ck];
      T.cnt = L.offX() + id, T.target = 2;
    }
  } else if (t < 4) {
    if (ys) {
      const int id = k * nstrips + s;
      T.x0 = semb_strip_e0(s, nstrips, a.Ex, N) * N + ((s > 0 || a.perx) ? 1 : 0);
      T.x1 = semb_strip_e0(s + 1, nstrips, a.Ex, N) * N - ((s < nstrips - 1 || a.perx) ? 1 : 0);
      T.n = (T.x1 - T.x0 + 31) >> 5;
      T.ya = ya, T.yb = yb, T.flags = hal ? 4 : 0;
      T.pbase = L.partY() + id * 8;
      T.cnt = L.offY() + id, T.target = hal ? 1 : 2;
      if (T.n == 0) T.target = 0;  // (a strip of seam columns only)
    }
  } else if (ys && sx >= 0) {
    const int id = k * nxs + sx;
    T.n = 1;
    T.x0 = xa, T.x1 = xb, T.ya = ya, T.yb = yb, T.flags = hal ? 4 : 0;
    T.pbase = L.partC() + id;
    T.cnt = L.offC() + id, T.target = hal ? 2 : 4;
  }
  s_task[threadIdx.x] = T;
}

// Called by ALL threads of the CTA right after it has finished the chunk (its stores fenced and a CTA barrier passed):
// threads 0..15 each announce one interface and keep the task only if this CTA arrived last.
__device__ __forceinline__ void semb_tail_announce(const OpArgs& a, SembTailTask* s_task) {
  const int t = threadIdx.x;
  if (t < 16) {
    const int target = s_task[t].target;
    if (target > 0) {
      unsigned* cnt = a.tcnt + s_task[t].cnt;
      const unsigned ticket = atomicAdd(cnt, 1u);
      if (ticket == (unsigned)(target - 1)) {
        *cnt = 0u;         // every arrival is in: ready for the next launch
        __threadfence();   // acquire side: the other chunks' stores are visible to this CTA from here on
      } else {
        s_task[t].n = 0;
      }
    }
  }
}

// One interface item in a uniform shape: up to two pairs of values, P0 = (x0,y0) + P1 = (x1,y1) and, if `four`,
// P2 = (x0,y2) + P3 = (x1,y2); the result (P0+P1) [+ (P2+P3)] goes back to every point that is local.  rem1: P1 is the
// neighbour rank's row (read at x1, not written); rem2: the second pair is the neighbour's row (read at x0, x1).
struct SembTailItem {
  int x0, y0, x1, y1, y2;
  bool valid, four, rem1, rem2;
  double v[4];
  double p[4];  // PCG: the search direction at the (up to four) points written
};

// Runs the tasks described in s_task[0 .. 16) (slot & 7 = interface kind, see semb_tail_announce), then the
// grid-wide finish.  Called once per CTA, by all of its threads, after its last chunk.  cta_acc: this thread's
// PCG partial from the strip kernel proper.
// The CTA that finishes last is typically the last arriver at all eight of its interfaces, and while it works on
// them it holds a whole strip-kernel slot (and, on a one-wave grid, the kernel's end), so the tasks are cut into
// MICRO-TASKS of 32 items dealt to all warps, several per warp at a time with every load of all of them issued before
// the first store: the tail costs two or three memory round trips instead of one per task (one warp per task made the
// 1e8-DOF apply 12 % slower than the separate seam kernels; profiles/r02_ab_tail_*.txt, r02_tail_timing_*.txt).
// PCG: a micro-task's partial of sum(p .* Ap .* mult) is reduced over its lanes in a fixed order and stored in the
// micro-task's own slot => the grid total does not depend on which CTA or warp ran it.
template <int U>
__device__ __forceinline__ void semb_tail_run(const OpArgs& a, const SembTailTask* s_task, const int* s_pre, int total,
                                              int ncorner, int ncta, const uint4* rem_lo, const uint4* rem_hi, unsigned tag,
                                              bool pcg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, N = a.N;
  const size_t pitch = (size_t)a.pitch;
  for (int mt0 = warp; mt0 < total; mt0 += U * nwarps) {
    SembTailItem it[U];
    int pslot[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // ---- decode: shared memory only
      SembTailItem& I = it[u];
      const int mt = mt0 + u * nwarps;
      I.valid = false, I.four = false, I.rem1 = false, I.rem2 = false;
      I.x0 = I.x1 = I.y0 = I.y1 = I.y2 = 0;
      pslot[u] = -1;
      if (mt >= total) continue;
      if (mt == total - 1 && ncorner) {
        // the CTA's corner tasks (up to eight, one item each) share the last micro-task: lane c <-> task slot 4..7, 12..15
        pslot[u] = -2;  // (per-lane partial slots, see below)
        const int q = 4 + (lane & 3) + 8 * (lane >> 2);
        if (lane >= 8 || s_task[q].n == 0) continue;
        const SembTailTask T = s_task[q];
        I.x0 = T.x0, I.x1 = T.x1, I.y0 = I.y1 = T.ya, I.y2 = T.yb, I.four = true, I.rem2 = (T.flags & 4) != 0;
        I.valid = true;
        continue;
      }
      int q = 0;
      while (s_pre[q + 1] <= mt) ++q;
      const SembTailTask T = s_task[q];
      const int mi = mt - s_pre[q], l = mi * 32 + lane, kind = q & 7;
      pslot[u] = T.pbase + mi;
      if (kind < 2) {  // X: line y of the chunk on the seam's two columns
        const int y = T.ya + l;
        if (y >= T.yb) continue;
        const int j = (y - T.ya) % N;
        if (j == 0 && y > T.ya) continue;                                               // lower line of an in-chunk pair
        if ((y == T.ya && (T.flags & 1)) || (y == T.yb - 1 && (T.flags & 2))) continue;  // on a y seam: corner task's
        I.x0 = T.x0, I.x1 = T.x1, I.y0 = I.y1 = y, I.y2 = y + 1;
        I.four = (j == N - 1 && y + 1 < T.yb);  // in-chunk y interface: (a+b)+(c+d)
        I.valid = true;
      } else if (kind < 4) {  // Y: column x of the seam's two rows
        const int x = T.x0 + l;
        if (x >= T.x1) continue;
        I.x0 = I.x1 = x, I.y0 = T.ya, I.y1 = T.yb, I.rem1 = (T.flags & 4) != 0;
        I.valid = true;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {  // ---- every load of every item
      SembTailItem& I = it[u];
      if (!I.valid) continue;
      const size_t i0 = (size_t)I.y0 * pitch + I.x0, i1 = (size_t)I.y1 * pitch + I.x1;
      const size_t i2 = (size_t)I.y2 * pitch + I.x0, i3 = (size_t)I.y2 * pitch + I.x1;
      I.v[0] = __ldcg(&a.out[i0]);
      if (!I.rem1) I.v[1] = __ldcg(&a.out[i1]);
      if (I.four && !I.rem2) {
        I.v[2] = __ldcg(&a.out[i2]);
        I.v[3] = __ldcg(&a.out[i3]);
      }
      if (pcg) {
        I.p[0] = __ldcg(&a.pout[i0]);
        if (!I.rem1) I.p[1] = __ldcg(&a.pout[i1]);
        if (I.four && !I.rem2) {
          I.p[2] = __ldcg(&a.pout[i2]);
          I.p[3] = __ldcg(&a.pout[i3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {  // ---- the neighbour rank's values (flag-in-data: re-read until this epoch's tag is there)
      SembTailItem& I = it[u];
      if (!I.valid || !(I.rem1 || I.rem2)) continue;
      const uint4* rem = (I.y0 == 0) ? rem_lo : rem_hi;
      if (I.rem1) I.v[1] = semb_ll_load(&rem[I.x1], tag, a.scal);
      if (I.rem2) {
        I.v[2] = semb_ll_load(&rem[I.x0], tag, a.scal);
        I.v[3] = semb_ll_load(&rem[I.x1], tag, a.scal);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {  // ---- sums, mask, stores, PCG partial of the micro-task
      const SembTailItem& I = it[u];
      double acc = 0.0;
      auto finish = [&](int x, int y, double val, double pv) {
        const size_t idx = (size_t)y * pitch + x;
        const double o = __dmul_rn(semb_mask_at(a, x, y, idx), val);
        a.out[idx] = o;
        // mult = 1 ./ gatherScatter(ones) (mesh.jl:94-96) = wx1d[x] * wy1d[y], factors in {1, 1/2} (exact)
        if (pcg) acc += __dmul_rn(__dmul_rn(pv, o), __ldg(&a.wx1d[x]) * __ldg(&a.wy1d[y]));
      };
      if (I.valid) {
        double sv = __dadd_rn(I.v[0], I.v[1]);                      // x pair (or the y pair of a Y task)
        if (I.four) sv = __dadd_rn(sv, __dadd_rn(I.v[2], I.v[3]));  // y pair of the two x pairs
        finish(I.x0, I.y0, sv, I.p[0]);
        if (!I.rem1) finish(I.x1, I.y1, sv, I.p[1]);
        if (I.four && !I.rem2) {
          finish(I.x0, I.y2, sv, I.p[2]);
          finish(I.x1, I.y2, sv, I.p[3]);
        }
      }
      if (pcg && pslot[u] >= 0) {  // (warp-uniform)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) a.tpart[ncta + pslot[u]] = acc;
      } else if (pcg && pslot[u] == -2 && I.valid) {  // packed corners: one slot per corner task
        a.tpart[ncta + s_task[4 + (lane & 3) + 8 * (lane >> 2)].pbase] = acc;
      }
    }
  }
}

static __device__ __noinline__ void semb_strip_tail(const OpArgs& a, const SembTailTask* s_task, int s, int nstrips,
                                                    unsigned long long ep, double cta_acc, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x, nch = a.nchunks;
  const size_t pitch = (size_t)a.pitch;
  const SembTailLayout L(nstrips, nch, a.nxs, a.xmic_total);
  const int ncta = gridDim.x * gridDim.y;
  const int par = (int)(ep & 1ull);
  const bool pcg = a.pcg != 0;
  const uint4* rem_lo = a.my_rows + (size_t)(2 * par + 0) * pitch;  // neighbour rows in OUR mailbox (side 0: from below)
  const uint4* rem_hi = a.my_rows + (size_t)(2 * par + 1) * pitch;
  const unsigned tag = semb_ll_tag(ep);
  __shared__ int s_pre[17];  // prefix of the micro-task counts of the 16 task slots
  __shared__ int s_fin, s_ncorner;
  __shared__ double s_tot;
  if (tid < 16) {
    const SembTailTask T = s_task[tid];
    const bool corner = (tid & 7) >= 4;
    const unsigned cmask = __ballot_sync(0xffffu, corner && T.n > 0);
    int v = corner ? 0 : T.n;  // (the corner tasks are packed into one micro-task after all the others)
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int w = __shfl_up_sync(0xffffu, v, o, 16);
      if (tid >= o) v += w;
    }
    s_pre[tid + 1] = v;
    if (tid == 0) s_pre[0] = 0;
    if (tid == 15) s_ncorner = cmask ? 1 : 0;
  }
  __syncthreads();
  semb_stamp(a, 5);
  const int ncorner = s_ncorner, total = s_pre[16] + ncorner;
  if (pcg) semb_tail_run<3>(a, s_task, s_pre, total, ncorner, ncta, rem_lo, rem_hi, tag, true);
  else semb_tail_run<5>(a, s_task, s_pre, total, ncorner, ncta, rem_lo, rem_hi, tag, false);
  __syncthreads();
  semb_stamp(a, 6);
#ifdef SEMB_TAIL_TIMING
  if (a.dbg && tid == 0) a.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + 7] = total;
#endif
  // ---- grid-wide finish (PCG mode only): reduction of sum(p .* Ap .* mult) (+ all-gather over the ranks) and the
  // device-side part of the halo epoch
  if (!pcg) return;
  const bool halo = a.has_lo || a.has_hi;
  const int bid = blockIdx.y * gridDim.x + blockIdx.x;
  {
    const double bs = semb_block_sum(cta_acc, red, tid, nt);
    if (tid == 0) a.tpart[bid] = bs;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned* cnt = a.tcnt + L.final_ticket();
    const unsigned ticket = atomicAdd(cnt, 1u);
    s_fin = (ticket == (unsigned)(ncta - 1));
    if (s_fin) {
      *cnt = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  if (!s_fin) return;
  {
    const int ntot = ncta + L.nparts();
    double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // fixed order: independent of who ran what
    for (int i = tid; i < ntot; i += 8 * nt) {
      double w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = (i + k * nt < ntot) ? __ldcg(&a.tpart[i + k * nt]) : 0.0;  // 8 loads in flight
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += w[k];
    }
    double tot = semb_block_sum(((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7])), red, tid, nt);
    if (tid == 0) s_tot = tot;
    __syncthreads();
    tot = s_tot;
    if (tid == 0) {
      a.scal->pap[0] = tot;
      a.scal->pap[1] = 0.0;
      a.scal->pap[2] = 0.0;
    }
    if (a.scal->nranks > 1) {
      double all;
      semb_p2p_allgather(a.scal, 0, tot, 0.0, &all, nullptr, tid);
      if (tid == 0) a.scal->pap_total = all;
    }
  }
  if (halo && tid == 0) a.scal->ep_dev[0] = ep - a.ep_host;  // = PCG-mode applies completed (see cur_ep)
}
