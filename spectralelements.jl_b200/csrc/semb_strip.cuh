// Fused strip kernel: out = mask(QQ^T(nu .* (D^T G D u) + k .* (B .* u)))  for one y-slab.
//
// Replaces (reference /root/reference/src): ABu.jl:9-37 (the four tensor contractions),
// lapl.jl:70-81 (laplace), mass.jl:17, hlmz.jl:15-16, the interior part of gatherScatter.jl:13
// and mask.jl:14, i.e. the body of opLHS (diffusion.jl:36-45).  In PCG mode it also performs
// pcg.jl:49 (u = hp + beta*u) on load and accumulates sum(u .* Au .* mult) (pcg.jl:52).
//
// Decomposition.  The field is the reference's global-lexicographic column-major matrix
// (x contiguous).  A CTA owns a strip of up to 32 x-consecutive elements and MARCHES through a
// chunk of element rows.  Two thread->data mappings alternate, exchanging data through shared
// memory (4 transposes per element row), so that every contraction runs out of registers with the
// D entries as constant-bank operands (no shared-memory traffic inside the N^2 FMA loops):
//   mapping B (y-lines): thread t <-> x index e0*N+t ; holds the N values of one column x of an
//                        element: global loads/stores are perfectly coalesced, Ds contractions
//   mapping A (x-lines): warp j <-> row j of the element row, lane <-> element ; Dr contractions
// x-interfaces inside the strip are summed through a small smem exchange, y-interfaces inside the
// chunk through a register carried from one element row to the next (the write of an element
// row's last line is deferred by one iteration).  Strip/chunk boundary lines ("seams") are written
// un-summed and finished by the two seam kernels (semb_vec.cu).  All interface sums are 2-term
// (commutative) and x pairs are formed before y pairs, which reproduces the reference's
// (a+b)+(c+d) association bit for bit.
#pragma once
#include "semb_internal.cuh"

template <int N>
struct StripParams {
  double Dr[N * N];  // Dr[i*N+k] = Dr(i,k)
  double Ds[N * N];
  OpArgs a;
};

template <int N>
struct StripCfg {
  static constexpr int S = N | 1;            // element stride in smem (odd => conflict-free x-line reads)
  static constexpr int PW = SEMB_BX * S;     // smem row pitch in doubles
  static constexpr int T = SEMB_BX * N;      // threads per CTA
  static constexpr int SMEM = (2 * N * PW + N * 2 * SEMB_BX) * 8;
  static constexpr bool PF = (N <= 10);      // register double-buffered prefetch of the next element row
  static constexpr int MINB = 1;
};

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

template <int N, bool PCGM, bool MASS>
__global__ void __launch_bounds__(StripCfg<N>::T, StripCfg<N>::MINB)
semb_strip_kernel(const __grid_constant__ StripParams<N> P) {
  using C = StripCfg<N>;
  constexpr int S = C::S, PW = C::PW;
  extern __shared__ double smem[];
  double* S1 = smem;
  double* S2 = smem + N * PW;
  double* S3 = smem + 2 * N * PW;  // x-interface exchange: [N][2*32]
  __shared__ double red[32];

  const OpArgs& a = P.a;
  double beta = 0.0;
  if (PCGM) {
    if (a.scal->done) return;
    beta = semb_pcg_beta(a.scal);
  }

  const int t = threadIdx.x;
  const int eB = t / N, iB = t - eB * N, colB = eB * S + iB;  // mapping B
  const int jA = t >> 5, eA = t & 31, colA = eA * S;         // mapping A
  const int strip = blockIdx.x, chunk = blockIdx.y;
  const int e0 = strip * SEMB_BX;
  const int nbe = min(SEMB_BX, a.Ex - e0);
  const bool actB = eB < nbe;
  const bool actA = eA < nbe;
  const int r0 = a.chunk_r0[chunk], r1 = a.chunk_r0[chunk + 1];
  const long long pitch = a.pitch;
  const int xg = e0 * N + t;
  const bool gs = a.gs != 0;
  // roles in the gather-scatter
  const bool xl = gs && actB && iB == N - 1 && eB < nbe - 1;  // left side of an in-strip x interface
  const bool xr = gs && actB && iB == 0 && eB > 0;            // right side
  const bool xs = gs && actB &&
                  ((iB == 0 && eB == 0 && (e0 > 0 || a.perx)) ||
                   (iB == N - 1 && eB == nbe - 1 && (e0 + nbe < a.Ex || a.perx)));  // strip seam column
  const double mcol = ((xg == 0 && a.mx0) || (xg == a.nxl - 1 && a.mx1)) ? 0.0 : 1.0;
  const bool seam_bot = a.ystart[r0] != 0;  // chunk's first line belongs to a y seam
  const bool seam_top = a.ystart[r1] != 0;

  double carry = 0.0, carry_u = 0.0;
  double acc = 0.0;  // PCG: sum p*Ap*mult over the nodes this thread finalises

  // Software pipelining in registers: the u (and p) loads of element row r+1 are issued at the top of
  // iteration r; the G loads of row r+1 are issued right after step 3 of iteration r has consumed
  // row r's factors (same registers).  One full strip row (~83 KB at N=9) is in flight per SM.
  constexpr bool PF = C::PF;
  double un[N], pn[N], g11[N], g12[N], g22[N];
  auto issue_u = [&](int r, double* uu, double* pp) {
    const size_t b = (size_t)r * N * pitch + xg;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const size_t idx = b + (size_t)j * pitch;
      uu[j] = actB ? a.u[idx] : 0.0;
      if (PCGM) pp[j] = actB ? a.pold[idx] : 0.0;
    }
  };
  auto issue_g = [&](int r) {
    const size_t b = (size_t)r * N * pitch + xg;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const size_t idx = b + (size_t)j * pitch;
      g11[j] = actB ? a.G11[idx] : 0.0;
      g12[j] = actB ? a.G12[idx] : 0.0;
      g22[j] = actB ? a.G22[idx] : 0.0;
    }
  };
  if (PF) {
    issue_u(r0, un, pn);
    issue_g(r0);
  }

  for (int r = r0; r < r1; ++r) {
    const size_t base = (size_t)r * N * pitch + xg;
    // ---- step 1 (B): load the column, form p in PCG mode, Ds contraction -----------------------
    double u[N], pl[N];
    if (PF) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        u[j] = un[j];
        if (PCGM) pl[j] = pn[j];
      }
      if (r + 1 < r1) issue_u(r + 1, un, pn);
    } else {
      issue_u(r, u, pl);
      issue_g(r);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const size_t idx = base + (size_t)j * pitch;
      double v = u[j];
      if (PCGM && actB) {
        if (a.precond) v = (v / a.B[idx]) / a.prec_b0;  // convectionDiffusion.jl:89
        v = __dadd_rn(v, __dmul_rn(beta, pl[j]));        // pcg.jl:49
        a.pout[idx] = v;
      }
      u[j] = v;
      S1[j * PW + colB] = v;
    }
    double us[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(P.Ds[j * N + k], u[k], s);
      us[j] = s;
    }
    __syncthreads();
    // ---- step 2 (A): ur = Dr * u along x ---------------------------------------------------------
    if (actA) {
      double c[N];
#pragma unroll
      for (int i = 0; i < N; ++i) c[i] = S1[jA * PW + colA + i];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) s = fma(P.Dr[m * N + i], c[i], s);
        S2[jA * PW + colA + m] = s;
      }
    }
    __syncthreads();
    // ---- step 3 (B): geometric factors, Ds^T contraction ---------------------------------------
    double aus[N];
#pragma unroll
    for (int m = 0; m < N; ++m) aus[m] = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double ur = S2[j * PW + colB];
      const double wr = fma(g11[j], ur, g12[j] * us[j]);  // lapl.jl:75
      const double ws = fma(g12[j], ur, g22[j] * us[j]);  // lapl.jl:76
      S1[j * PW + colB] = wr;
#pragma unroll
      for (int m = 0; m < N; ++m) aus[m] = fma(P.Ds[j * N + m], ws, aus[m]);
    }
    if (PF && r + 1 < r1) issue_g(r + 1);
    __syncthreads();
    // ---- step 4 (A): Dr^T contraction -----------------------------------------------------------
    if (actA) {
      double c[N];
#pragma unroll
      for (int i = 0; i < N; ++i) c[i] = S1[jA * PW + colA + i];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) s = fma(P.Dr[i * N + m], c[i], s);
        S2[jA * PW + colA + m] = s;
      }
    }
    __syncthreads();
    // ---- step 5 (B): combine, hlmz, gather-scatter, mask, store ------------------------------
    double v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const size_t idx = base + (size_t)j * pitch;
      double lap = __dadd_rn(S2[j * PW + colB], aus[j]);  // lapl.jl:78
      if (actB) {
        const double nu = a.nu_arr ? a.nu_arr[idx] : a.nu;
        lap = __dmul_rn(nu, lap);  // hlmz.jl:15
        if (MASS) {
          const double kk = a.k_arr ? a.k_arr[idx] : a.k;
          lap = __dadd_rn(lap, __dmul_rn(kk, __dmul_rn(a.B[idx], u[j])));  // hlmz.jl:16, mass.jl:17
        }
      }
      v[j] = lap;
    }
    if (!gs) {
      if (actB) {
#pragma unroll
        for (int j = 0; j < N; ++j) a.out[base + (size_t)j * pitch] = v[j];
      }
      continue;
    }
    // x pairs inside the strip
    if (xl) {
#pragma unroll
      for (int j = 0; j < N; ++j) S3[j * 2 * SEMB_BX + 2 * eB] = v[j];
    }
    if (xr) {
#pragma unroll
      for (int j = 0; j < N; ++j) S3[j * 2 * SEMB_BX + 2 * (eB - 1) + 1] = v[j];
    }
    __syncthreads();
    if (xl) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = __dadd_rn(v[j], S3[j * 2 * SEMB_BX + 2 * eB + 1]);
    }
    if (xr) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = __dadd_rn(v[j], S3[j * 2 * SEMB_BX + 2 * (eB - 1)]);
    }
    if (!actB) continue;
    if (xs) {
      // strip seam column: raw values for every line; the x-seam kernel forms x pairs, then y pairs
#pragma unroll
      for (int j = 0; j < N; ++j) a.out[base + (size_t)j * pitch] = v[j];
      continue;
    }
    // final value + mask + (PCG) dot contribution for one node
    auto finish = [&](size_t idx, double val, double pval, double mrow) {
      const double mk = a.M_arr ? a.M_arr[idx] : mcol * mrow;
      const double o = __dmul_rn(mk, val);  // mask.jl:14
      a.out[idx] = o;
      if (PCGM) acc += __dmul_rn(__dmul_rn(pval, o), a.mult[idx]);  // pcg.jl:52
    };
    // line 0
    if (r == r0) {
      if (seam_bot) {
        a.out[base] = v[0];
      } else {
        finish(base, v[0], u[0], (r == 0 && a.my0) ? 0.0 : 1.0);
      }
    } else {
      const double s = __dadd_rn(carry, v[0]);  // y pair (after the x pairs)
      finish(base, s, u[0], 1.0);
      finish(base - pitch, s, carry_u, 1.0);  // deferred last line of the previous element row
    }
#pragma unroll
    for (int j = 1; j < N - 1; ++j) finish(base + (size_t)j * pitch, v[j], u[j], 1.0);
    // line N-1
    if (r == r1 - 1) {
      const size_t idx = base + (size_t)(N - 1) * pitch;
      if (seam_top) {
        a.out[idx] = v[N - 1];
      } else {
        finish(idx, v[N - 1], u[N - 1], (r == a.ney - 1 && a.my1) ? 0.0 : 1.0);
      }
    } else {
      carry = v[N - 1];
      if (PCGM) carry_u = u[N - 1];
    }
  }

  if (PCGM) {
    const int nblocks = gridDim.x * gridDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    const double bs = semb_block_sum(acc, red, t, C::T);
    double total;
    if (semb_last_block(bs, 0.0, a.partials, nullptr, a.counters + 0, nblocks, bid, red, t, C::T, &total,
                        nullptr)) {
      a.scal->pap[0] = total;
    }
  }
}
