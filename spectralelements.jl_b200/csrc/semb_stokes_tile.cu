// Register-tiled element kernels of the Stokes split for nr == ns = NV on mshV and NP = NV - 2 on mshP
// (examples/semPS.jl:31), each replacing a chain of generic launches by ONE launch that reads every input once:
//
//   diverT (diver.jl:53-63):  pr (mshP) -> Jp = ABu(Js,Jr,pr) -> w = B .* Jp -> qx = Dr'(rx.*w) + Ds'(sx.*w),
//                                                                             qy = Dr'(ry.*w) + Ds'(sy.*w)   (mshV)
//       in: pr + B, rx, ry, sx, sy; out: qx, qy                 (56 B per velocity node + 8 B per pressure node)
//   diver  (diver.jl:17-31):  ux, uy (mshV) -> B .* (dx ux + dy uy) -> sign * ABu(Js',Jr', .)                (mshP)
//       in: ux, uy + B, rx, ry, sx, sy; out: mshP field         (56 B per velocity node + 8 B per pressure node)
//
// Same organisation as the dealiased advection kernel (semb_advect_tile.cu): a CTA of 128 threads works on a batch of
// EB = 128/NV x-consecutive elements; two thread->line mappings alternate through shared-memory tiles [row][e*S+i];
// every contraction runs fully unrolled out of registers with even-odd tables (semb_eo.cuh; the derivative matrices
// are centro-antisymmetric, the GLL interpolants centro-symmetric), and the two lines that share a matrix are paired.
// The metric terms are read straight from global memory by the column-owner threads (coalesced), one phase before they
// are needed.  Arithmetic order: the pointwise products are formed and rounded first (the reference's broadcasts);
// diverT interpolates with Jr first (as ABu, ABu.jl:14-33); diver projects with Js' first (the columns are already in
// registers; the two directions commute up to rounding, ~1e-16 relative).
#include "semb_eo.cuh"
#include "semb_vec.cuh"

namespace {

constexpr int STK_T = 128;
#ifndef STK_MINB
#define STK_MINB 3
#endif

// next batch's columns into L2 (one request per 32-byte sector): the column loads below then find their lines there
__device__ __forceinline__ void stk_prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct StokesTileArgs {
  const double *in1, *in2;   // diverT: in1 = pr (mshP) ; diver: ux, uy (mshV)
  double *out1, *out2;       // diverT: qx, qy (mshV)   ; diver: out1 (mshP)
  const double *rx, *ry, *sx, *sy, *B;
  const double *Dr, *Ds;     // row-major NV x NV
  const double *Jr, *Js;     // column-major NV x NP, interpMat(mshV.z, mshP.z)
  long long pitchV, pitchP;
  double sign;               // diver: factor of the result (stokesOp returns -Eq, diver.jl:88)
  int Ex, ney;
};

template <int NV, int NP>
struct StkCfg {
  static constexpr int EB = STK_T / NV;
  static constexpr int S = NV | 1, SP = NP | 1;
  static constexpr int PV = EB * S, PP = EB * SP;
  using TD = EoTab<NV, NV, -1>;    // D  (diver)
  using TDT = EoTab<NV, NV, -1>;   // D' (diverT): the transpose of a centro-antisymmetric matrix is centro-antisymmetric
  using TJ = EoTab<NP, NV, +1>;    // J  : pressure -> velocity nodes
  using TJT = EoTab<NV, NP, +1>;   // J' : velocity -> pressure nodes
};

// ---------------------------------------------------------------------------------------------------------------------
template <int NV, int NP>
__global__ void __launch_bounds__(STK_T, STK_MINB) semb_diverT_tile_kernel(const StokesTileArgs a) {
  using C = StkCfg<NV, NP>;
  constexpr int EB = C::EB, S = C::S, SP = C::SP, PV = C::PV, PP = C::PP;
  constexpr int OFF_JR = 0, OFF_JS = C::TJ::SIZE, OFF_DRT = 2 * C::TJ::SIZE, OFF_DST = OFF_DRT + C::TDT::SIZE,
                TAB = (OFF_DST + C::TDT::SIZE + 1) & ~1;
  __shared__ __align__(16) double sh[TAB + NP * PP + NP * PV + 2 * NV * PV];
  double* tP = sh + TAB;           // [NP][PP]  pressure tile
  double* tX = tP + NP * PP;       // [NP][PV]  x-interpolated
  double* tA = tX + NP * PV;       // [2][NV][PV]  rx.*w, ry.*w -> Dr' of them
  const int t = threadIdx.x;
  C::TJ::fill(sh + OFF_JR, t, STK_T, [&](int i, int k) { return a.Jr[i + k * NV]; });
  C::TJ::fill(sh + OFF_JS, t, STK_T, [&](int i, int k) { return a.Js[i + k * NV]; });
  C::TDT::fill(sh + OFF_DRT, t, STK_T, [&](int i, int k) { return a.Dr[k * NV + i]; });  // D'(i,k) = D(k,i)
  C::TDT::fill(sh + OFF_DST, t, STK_T, [&](int i, int k) { return a.Ds[k * NV + i]; });
  const int eC = t / NV, iC = t - eC * NV, colC = eC * S + iC;
  const int jR = t / EB, eR = t - jR * EB;
  const int nbx = (a.Ex + EB - 1) / EB;
  for (int b = blockIdx.x; b < nbx * a.ney; b += gridDim.x) {
    const int r = b / nbx, e0 = (b - r * nbx) * EB, nbe = min(EB, a.Ex - e0);
    const bool actC = t < nbe * NV;
    const size_t gV = (size_t)r * NV * a.pitchV + (size_t)e0 * NV + t;
    {
      const int bn = b + gridDim.x;
      if (bn < nbx * a.ney && (t & 3) == 0) {
        const int rn = bn / nbx, en = (bn - rn * nbx) * EB;
        if (t < min(EB, a.Ex - en) * NV) {
          const size_t gn = (size_t)rn * NV * a.pitchV + (size_t)en * NV + t;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const size_t g = gn + (size_t)j * a.pitchV;
            stk_prefetch_l2(a.B + g), stk_prefetch_l2(a.rx + g), stk_prefetch_l2(a.ry + g);
            stk_prefetch_l2(a.sx + g), stk_prefetch_l2(a.sy + g);
          }
        }
      }
    }
    __syncthreads();  // tables in place; previous batch done with the tiles
    for (int q = t; q < NP * nbe * NP; q += STK_T) {  // pressure tile, coalesced rows
      const int n = q / (nbe * NP), xx = q - n * (nbe * NP), e = xx / NP, m = xx - e * NP;
      tP[n * PP + e * SP + m] = a.in1[(size_t)(r * NP + n) * a.pitchP + (size_t)e0 * NP + xx];
    }
    // the column's metric terms: issued now, consumed after the two interpolation phases
    double bq[NV], c1[NV], c2[NV];
    if (actC) {
#pragma unroll
      for (int j = 0; j < NV; ++j) bq[j] = a.B[gV + (size_t)j * a.pitchV];
    }
    __syncthreads();
    // ---- R: x-interpolation of the NP pressure rows (Jr) -------------------------------------------------------------
    if (jR < NP && eR < nbe) {
      double xl[1][NP], y[1][NV];
#pragma unroll
      for (int m = 0; m < NP; ++m) xl[0][m] = tP[jR * PP + eR * SP + m];
      eo_contract<NP, NV, +1, 1>(sh + OFF_JR, xl, y);
#pragma unroll
      for (int i = 0; i < NV; ++i) tX[jR * PV + eR * S + i] = y[0][i];
    }
    __syncthreads();
    // ---- C: y-interpolation (Js), w = B .* Jp (mass.jl:17), the four products, Ds' of the pair (sx.*w, sy.*w) ------------
    double qs[2][NV];
    if (actC) {
      double xc[1][NP], jp[1][NV];
#pragma unroll
      for (int n = 0; n < NP; ++n) xc[0][n] = tX[n * PV + colC];
      eo_contract<NP, NV, +1, 1>(sh + OFF_JS, xc, jp);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        c1[j] = a.rx[gV + (size_t)j * a.pitchV];
        c2[j] = a.ry[gV + (size_t)j * a.pitchV];
      }
      double w[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        w[j] = __dmul_rn(bq[j], jp[0][j]);
        tA[j * PV + colC] = __dmul_rn(c1[j], w[j]);         // rx .* w
        tA[(NV + j) * PV + colC] = __dmul_rn(c2[j], w[j]);  // ry .* w
      }
      double xs[2][NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        xs[0][j] = __dmul_rn(a.sx[gV + (size_t)j * a.pitchV], w[j]);
        xs[1][j] = __dmul_rn(a.sy[gV + (size_t)j * a.pitchV], w[j]);
      }
      eo_contract<NV, NV, -1, 2>(sh + OFF_DST, xs, qs);
    }
    __syncthreads();
    // ---- R: Dr' of the pair (rx.*w, ry.*w), in place ---------------------------------------------------------------
    if (jR < NV && eR < nbe) {
      double xl[2][NV], y[2][NV];
      double* row = tA + jR * PV + eR * S;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xl[0][i] = row[i];
        xl[1][i] = row[NV * PV + i];
      }
      eo_contract<NV, NV, -1, 2>(sh + OFF_DRT, xl, y);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        row[i] = y[0][i];
        row[NV * PV + i] = y[1][i];
      }
    }
    __syncthreads();
    // ---- C: qx = Dr'(rx.*w) + Ds'(sx.*w), qy likewise (grad.jl:56-60 with the directions un-swapped); coalesced stores ----
    if (actC) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        a.out1[gV + (size_t)j * a.pitchV] = __dadd_rn(tA[j * PV + colC], qs[0][j]);
        a.out2[gV + (size_t)j * a.pitchV] = __dadd_rn(tA[(NV + j) * PV + colC], qs[1][j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NV, int NP>
__global__ void __launch_bounds__(STK_T, STK_MINB) semb_diver_tile_kernel(const StokesTileArgs a) {
  using C = StkCfg<NV, NP>;
  constexpr int EB = C::EB, S = C::S, SP = C::SP, PV = C::PV, PP = C::PP;
  constexpr int OFF_DR = 0, OFF_DS = C::TD::SIZE, OFF_JRT = 2 * C::TD::SIZE, OFF_JST = OFF_JRT + C::TJT::SIZE,
                TAB = (OFF_JST + C::TJT::SIZE + 1) & ~1;
  __shared__ __align__(16) double sh[TAB + 2 * NV * PV + NP * PV];
  double* tU = sh + TAB;            // [2][NV][PV]  ux, uy -> ur of them
  double* tY = tU + 2 * NV * PV;    // [NP][PV]     y-projected B.*div
  const int t = threadIdx.x;
  C::TD::fill(sh + OFF_DR, t, STK_T, [&](int i, int k) { return a.Dr[i * NV + k]; });
  C::TD::fill(sh + OFF_DS, t, STK_T, [&](int i, int k) { return a.Ds[i * NV + k]; });
  C::TJT::fill(sh + OFF_JRT, t, STK_T, [&](int m, int i) { return a.Jr[i + m * NV]; });  // J'(m,i) = J(i,m)
  C::TJT::fill(sh + OFF_JST, t, STK_T, [&](int m, int i) { return a.Js[i + m * NV]; });
  const int eC = t / NV, iC = t - eC * NV, colC = eC * S + iC;
  const int jR = t / EB, eR = t - jR * EB;
  const int nbx = (a.Ex + EB - 1) / EB;
  for (int b = blockIdx.x; b < nbx * a.ney; b += gridDim.x) {
    const int r = b / nbx, e0 = (b - r * nbx) * EB, nbe = min(EB, a.Ex - e0);
    const bool actC = t < nbe * NV;
    const size_t gV = (size_t)r * NV * a.pitchV + (size_t)e0 * NV + t;
    {
      const int bn = b + gridDim.x;
      if (bn < nbx * a.ney && (t & 3) == 0) {
        const int rn = bn / nbx, en = (bn - rn * nbx) * EB;
        if (t < min(EB, a.Ex - en) * NV) {
          const size_t gn = (size_t)rn * NV * a.pitchV + (size_t)en * NV + t;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const size_t g = gn + (size_t)j * a.pitchV;
            stk_prefetch_l2(a.in1 + g), stk_prefetch_l2(a.in2 + g), stk_prefetch_l2(a.B + g), stk_prefetch_l2(a.rx + g);
            stk_prefetch_l2(a.ry + g), stk_prefetch_l2(a.sx + g), stk_prefetch_l2(a.sy + g);
          }
        }
      }
    }
    __syncthreads();
    // ---- C: columns of ux, uy -> registers and tiles; us = Ds * (ux, uy) ---------------------------------------------
    double us[2][NV], c1[NV], c2[NV];
    if (actC) {
      double uc[2][NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        uc[0][j] = a.in1[gV + (size_t)j * a.pitchV];
        uc[1][j] = a.in2[gV + (size_t)j * a.pitchV];
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        c1[j] = a.sx[gV + (size_t)j * a.pitchV];   // consumed two phases later
        c2[j] = a.sy[gV + (size_t)j * a.pitchV];
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        tU[j * PV + colC] = uc[0][j];
        tU[(NV + j) * PV + colC] = uc[1][j];
      }
      eo_contract<NV, NV, -1, 2>(sh + OFF_DS, uc, us);
    }
    __syncthreads();
    // ---- R: ur = Dr * (ux, uy) rows, in place ------------------------------------------------------------------------
    if (jR < NV && eR < nbe) {
      double xl[2][NV], y[2][NV];
      double* row = tU + jR * PV + eR * S;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xl[0][i] = row[i];
        xl[1][i] = row[NV * PV + i];
      }
      eo_contract<NV, NV, -1, 2>(sh + OFF_DR, xl, y);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        row[i] = y[0][i];
        row[NV * PV + i] = y[1][i];
      }
    }
    __syncthreads();
    // ---- C: B .* (uxdx + uydy) (grad.jl:30-31, diver.jl:25-27), y-projection Js' -----------------------------------------
    if (actC) {
      double w[1][NV], y[1][NP];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const size_t g = gV + (size_t)j * a.pitchV;
        const double uxdx = __dadd_rn(__dmul_rn(a.rx[g], tU[j * PV + colC]), __dmul_rn(c1[j], us[0][j]));
        const double uydy = __dadd_rn(__dmul_rn(a.ry[g], tU[(NV + j) * PV + colC]), __dmul_rn(c2[j], us[1][j]));
        w[0][j] = __dmul_rn(a.B[g], __dadd_rn(uxdx, uydy));
      }
      eo_contract<NV, NP, +1, 1>(sh + OFF_JST, w, y);
#pragma unroll
      for (int n = 0; n < NP; ++n) tY[n * PV + colC] = y[0][n];
    }
    __syncthreads();
    // ---- R: x-projection Jr' of the NP rows, store to the pressure mesh (a warp covers contiguous row segments) -----------
    if (jR < NP && eR < nbe) {
      double xl[1][NV], y[1][NP];
#pragma unroll
      for (int i = 0; i < NV; ++i) xl[0][i] = tY[jR * PV + eR * S + i];
      eo_contract<NV, NP, +1, 1>(sh + OFF_JRT, xl, y);
      double* dst = a.out1 + (size_t)(r * NP + jR) * a.pitchP + (size_t)(e0 + eR) * NP;
#pragma unroll
      for (int m = 0; m < NP; ++m) dst[m] = __dmul_rn(a.sign, y[0][m]);
    }
  }
  (void)SP;
  (void)PP;
}

template <int NV, int NP>
int launch_stokes_tile(semb_ctx* ctx, const StokesTileArgs& a, bool transpose) {
  const int nbatch = ((a.Ex + StkCfg<NV, NP>::EB - 1) / StkCfg<NV, NP>::EB) * a.ney;
  int grid = ctx->sm_count * STK_MINB;   // persistent: the resident CTAs walk through the batches
  if (grid > nbatch) grid = nbatch;
  if (grid < 1) grid = 1;
  if (transpose)
    semb_diverT_tile_kernel<NV, NP><<<grid, STK_T, 0, ctx->stream>>>(a);
  else
    semb_diver_tile_kernel<NV, NP><<<grid, STK_T, 0, ctx->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

}  // namespace

// velocity orders served (pressure order NV - 2)
#define SEMB_STK_SIZES(X) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)

// transpose = 1: diverT (in1 = pr on P; out1, out2 = qx, qy on V); 0: diver (in1, in2 = ux, uy; out1 on P, times sign).
// *done = 0 if the sizes are not served or the matrices fail the symmetry test (the caller then runs the generic chain).
int semb_launch_stokes_tile(semb_ctx* ctx, semb_mesh* V, semb_mesh* P, int transpose, const double* in1, const double* in2,
                            double* out1, double* out2, const double* dJr, const double* dJs, double sign, int* done) {
  *done = 0;
  if (V->nr != V->ns || P->nr != P->ns || P->nr != V->nr - 2 || !V->eo) return SEMB_OK;
  StokesTileArgs a;
  a.in1 = in1;
  a.in2 = in2;
  a.out1 = out1;
  a.out2 = out2;
  a.rx = V->arr[SEMB_RX];
  a.ry = V->arr[SEMB_RY];
  a.sx = V->arr[SEMB_SX];
  a.sy = V->arr[SEMB_SY];
  a.B = V->arr[SEMB_B];
  a.Dr = V->dDr;
  a.Ds = V->dDs;
  a.Jr = dJr;
  a.Js = dJs;
  a.pitchV = V->pitch;
  a.pitchP = P->pitch;
  a.sign = sign;
  a.Ex = V->Ex;
  a.ney = V->ney;
#define SEMB_STK_CASE(n)                                              \
  if (V->nr == n) {                                                   \
    SEMB_TRY((launch_stokes_tile<n, n - 2>(ctx, a, transpose != 0))); \
    *done = 1;                                                        \
    return SEMB_OK;                                                   \
  }
  SEMB_STK_SIZES(SEMB_STK_CASE)
#undef SEMB_STK_CASE
  return SEMB_OK;
}
