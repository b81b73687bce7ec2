// Even-odd contraction tables for (anti)symmetric operator matrices on symmetric node sets, rectangular or square,
// and the register-level contraction that uses them (one thread = one line, tables broadcast from shared memory).
// Used by the dealiased advection kernel (semb_advect_tile.cu) and the Stokes element kernels (semb_stokes_tile.cu).
#pragma once
#include <cuda_runtime.h>

// Even-odd tables of A (NO x NI) with A(NO-1-m, NI-1-k) = SG * A(m,k), SG = +1 (interpolation) or -1 (derivative):
//   s_m = (y_m + SG y_{NO-1-m})/2 = sum_{k<HI} P(m,k) e_k + [NI odd] A(m,c) x_c,   P = (A(m,k) + A(m,NI-1-k))/2
//   d_m = (y_m - SG y_{NO-1-m})/2 = sum_{k<HI} Q(m,k) o_k,                          Q = (A(m,k) - A(m,NI-1-k))/2
//   y_m = s_m + d_m,  y_{NO-1-m} = SG (s_m - d_m);  the middle output row (NO odd) lives in s (SG = +1) or d (SG = -1).
// Layout: P[k][m], k < HI + (NI odd), m < NS, rows padded to LS (even: LDS.128); then Q[k][m], k < HI, m < ND, rows LQ.
template <int NI, int NO, int SG>
struct EoTab {
  static constexpr int HI = NI / 2, OI = NI & 1, HO = NO / 2, OO = NO & 1;
  static constexpr int NS = HO + ((OO && SG > 0) ? 1 : 0);
  static constexpr int ND = HO + ((OO && SG < 0) ? 1 : 0);
  static constexpr int LS = (NS + 1) & ~1, LQ = (ND + 1) & ~1;
  static constexpr int KP = HI + OI;
  static constexpr int OFFQ = KP * LS;
  static constexpr int SIZE = OFFQ + HI * LQ;
  // fill cooperatively; A(m,k) = getA(m, k)
  template <typename F>
  static __device__ void fill(double* T, int tid, int nt, F getA) {
    for (int q = tid; q < KP * LS; q += nt) {
      const int k = q / LS, m = q - k * LS;
      double v = 0.0;
      if (m < NS) v = k < HI ? 0.5 * (getA(m, k) + getA(m, NI - 1 - k)) : getA(m, HI);
      T[q] = v;
    }
    for (int q = tid; q < HI * LQ; q += nt) {
      const int k = q / LQ, m = q - k * LQ;
      T[OFFQ + q] = m < ND ? 0.5 * (getA(m, k) - getA(m, NI - 1 - k)) : 0.0;
    }
  }
};

// y[v] = A x[v] for NV lines at once (NV = 1 or 2: one table load feeds all lines); T = EoTab<NI,NO,SG> in shared memory
template <int NI, int NO, int SG, int NV>
__device__ __forceinline__ void eo_contract(const double* __restrict__ T, const double (&x)[NV][NI], double (&y)[NV][NO]) {
  using E = EoTab<NI, NO, SG>;
  constexpr int HI = E::HI, OI = E::OI, HO = E::HO, LS = E::LS, LQ = E::LQ;
  constexpr int NS2 = E::LS / 2, ND2 = E::LQ / 2;
  double2 s[NV][NS2 > 0 ? NS2 : 1], d[NV][ND2 > 0 ? ND2 : 1];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
#pragma unroll
    for (int m = 0; m < NS2; ++m) s[v][m] = make_double2(0.0, 0.0);
#pragma unroll
    for (int m = 0; m < ND2; ++m) d[v][m] = make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int k = 0; k < HI; ++k) {
    double e[NV], o[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      e[v] = x[v][k] + x[v][NI - 1 - k];
      o[v] = x[v][k] - x[v][NI - 1 - k];
    }
    const double2* P = reinterpret_cast<const double2*>(T + k * LS);
#pragma unroll
    for (int m = 0; m < NS2; ++m) {
      const double2 p = P[m];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        s[v][m].x = fma(p.x, e[v], s[v][m].x);
        s[v][m].y = fma(p.y, e[v], s[v][m].y);
      }
    }
    const double2* Q = reinterpret_cast<const double2*>(T + E::OFFQ + k * LQ);
#pragma unroll
    for (int m = 0; m < ND2; ++m) {
      const double2 q = Q[m];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        d[v][m].x = fma(q.x, o[v], d[v][m].x);
        d[v][m].y = fma(q.y, o[v], d[v][m].y);
      }
    }
  }
  if (OI) {  // middle input
    const double2* P = reinterpret_cast<const double2*>(T + HI * LS);
#pragma unroll
    for (int m = 0; m < NS2; ++m) {
      const double2 p = P[m];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        s[v][m].x = fma(p.x, x[v][HI], s[v][m].x);
        s[v][m].y = fma(p.y, x[v][HI], s[v][m].y);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
#pragma unroll
    for (int m = 0; m < HO; ++m) {
      const double sm = (m & 1) ? s[v][m >> 1].y : s[v][m >> 1].x;
      const double dm = (m & 1) ? d[v][m >> 1].y : d[v][m >> 1].x;
      y[v][m] = sm + dm;
      y[v][NO - 1 - m] = SG > 0 ? sm - dm : dm - sm;
    }
    if (NO & 1) {
      if (SG > 0) y[v][HO] = (HO & 1) ? s[v][HO >> 1].y : s[v][HO >> 1].x;
      else y[v][HO] = (HO & 1) ? d[v][HO >> 1].y : d[v][HO >> 1].x;
    }
  }
}

