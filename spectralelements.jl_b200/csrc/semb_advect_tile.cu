// Register-tiled fused dealiased advection, advect.jl:45-64, for nr == ns = N on mshV and nrd == nsd = M on mshD,
// templated on (N, M) so that every contraction runs fully unrolled out of registers (one thread = one line):
//
//   Tx,Ty = grad(T,mshV)                       grad.jl:15-34
//   JTx,JTy,Jux,Juy = ABu(Js,Jr,.)             advect.jl:54-57   (V grid -> dealiasing grid)
//   JCu = (Jux.*JTx + Juy.*JTy) .* mshD.B      advect.jl:59-60
//   Cu  = ABu(Js',Jr',JCu)                     advect.jl:61      (projection back to the V grid)
//
// One CTA of 128 threads works on a batch of EB = 128/M x-consecutive elements of one element row; nothing between
// the loads of T,ux,uy,rx,ry,sx,sy,B_D and the store of Cu leaves the SM.  Two thread->data mappings alternate
// (as in the strip kernel, semb_strip.cuh), exchanging lines through shared memory tiles [row][e*S + i], S odd:
//   C (y-lines): thread c <-> column (e,i) of the batch: coalesced global loads/stores, Ds / Js / Js' contractions
//   R (x-lines): thread p <-> (row, e):                                               Dr / Jr / Jr' contractions
// Phases per batch (one __syncthreads between them):
//   1 C: load T column, us = Ds*T, T -> tile            2 R: ur = Dr*T in place
//   3 C: Tx,Ty from ur,us and the metric terms; y-interpolation (Js) of Tx,Ty,ux,uy -> 4 tiles of M rows
//   4 R: x-interpolation (Jr) of the 4 rows, pointwise product with B_D, x-projection (Jr') -> tile (in place)
//   5 C: y-projection (Js'), coalesced store of Cu
// The two interpolation directions commute exactly in exact arithmetic; the reference applies Jr first (ABu.jl:14-33),
// here Js is applied first (the column mapping already holds the y-lines): results agree to rounding (~1e-16
// relative), well inside the 1e-12 contract.  The projection is applied in the reference's order (Jr' then Js').
//
// Several T's that share the advecting velocity (makeRHS!: exH[i] = -advect(uh[i],vx,vy,...) for i = 1..k,
// convectionDiffusion.jl:100-105) are processed in ONE launch: ux,uy,B_D are loaded and Jux,Juy interpolated once
// per batch (kept in registers by the R mapping), each T then costs T + Cu of HBM traffic.
#include "semb_vec.cuh"

namespace {

constexpr int ADV_T = 128;
constexpr int ADV_MAXT = 4;

struct AdvTileArgs {
  const double* T[ADV_MAXT];
  double* out[ADV_MAXT];
  const double *ux, *uy, *rx, *ry, *sx, *sy, *BD;
  const double *Dr, *Ds;  // row-major N x N (semb_mesh::dDr)
  const double *Jr, *Js;  // column-major M x N, interpMat(mshD.z, mshV.z)
  long long pitchV, pitchD;
  int nT, Ex, ney;
};

template <int N, int M>
struct AdvCfg {
  static constexpr int EB = ADV_T / M;   // elements per batch
  static constexpr int S = N | 1;        // element stride inside a tile row (odd: both mappings conflict-free)
  static constexpr int SD = M | 1;       // same for the B_D tile
  static constexpr int PV = EB * S, PD = EB * SD;
  static constexpr int NP = N + (N & 1), MP = M + (M & 1);  // table rows padded to even (LDS.128)
  // tables [k][o]: Dr, Ds (N x NP), Jr, Js (N x MP), Jr', Js' (M x NP)
  static constexpr int OFF_DR = 0, OFF_DS = N * NP, OFF_JR = 2 * N * NP, OFF_JS = OFF_JR + N * MP,
                       OFF_JRT = OFF_JS + N * MP, OFF_JST = OFF_JRT + M * NP, TAB = OFF_JST + M * NP;
  static constexpr int OFF_ST = TAB;                   // [N][PV]    T -> ur
  static constexpr int OFF_TF = OFF_ST + N * PV;       // [2][M][PV] y-interpolated Tx, Ty; [0] reused for Jr' JCu
  static constexpr int OFF_TU = OFF_TF + 2 * M * PV;   // [2][M][PD] y-interpolated ux, uy (N per element), then Jux, Juy (M per
                                                       // element) written in place by the row's owner thread
  static constexpr int OFF_BD = OFF_TU + 2 * M * PD;   // [M][PD]
  static constexpr int SMEM_DOUBLES = OFF_BD + M * PD;
  static constexpr int SMEM = SMEM_DOUBLES * 8;
  static constexpr int OCC0 = (227 * 1024) / (SMEM + 1024);
  static constexpr int OCC = OCC0 < 1 ? 1 : (OCC0 > 4 ? 4 : OCC0);
};

// y[o] = sum_k T[k*LD + o] * x[k], o < NO; T in shared memory, rows 16-byte aligned (LD even): broadcast LDS.128
template <int NI, int NO, int LD>
__device__ __forceinline__ void adv_contract(const double* __restrict__ T, const double (&x)[NI], double (&y)[NO]) {
  constexpr int NO2 = (NO + 1) / 2;
  double2 acc[NO2];
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const double2* row = reinterpret_cast<const double2*>(T + k * LD);
#pragma unroll
    for (int o = 0; o < NO2; ++o) {
      const double2 t = row[o];
      if (k == 0) {
        acc[o].x = t.x * x[0];
        acc[o].y = t.y * x[0];
      } else {
        acc[o].x = fma(t.x, x[k], acc[o].x);
        acc[o].y = fma(t.y, x[k], acc[o].y);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) y[o] = (o & 1) ? acc[o >> 1].y : acc[o >> 1].x;
}

// y[o] = sum_k T[k*LD + O0 + o] * x[k], o < NO (O0 even)
template <int NI, int O0, int NO, int LD>
__device__ __forceinline__ void adv_contract_part(const double* __restrict__ T, const double (&x)[NI], double (&y)[NO]) {
  constexpr int NO2 = (NO + 1) / 2;
  double2 acc[NO2];
#pragma unroll
  for (int k = 0; k < NI; ++k) {
    const double2* row = reinterpret_cast<const double2*>(T + k * LD + O0);
#pragma unroll
    for (int o = 0; o < NO2; ++o) {
      const double2 t = row[o];
      if (k == 0) {
        acc[o].x = t.x * x[0];
        acc[o].y = t.y * x[0];
      } else {
        acc[o].x = fma(t.x, x[k], acc[o].x);
        acc[o].y = fma(t.y, x[k], acc[o].y);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < NO; ++o) y[o] = (o & 1) ? acc[o >> 1].y : acc[o >> 1].x;
}

// outputs m in [O0, O0+NO) of one dealiasing-grid row: interpolate, multiply, and add their share of the Jr' projection
template <int N, int M, int O0, int NO>
__device__ __forceinline__ void adv_phase4_part(const double* sh, const double* rowF, const double* rowU, const double* rowB,
                                                double (&pr)[N]) {
  using C = AdvCfg<N, M>;
  constexpr int PV = C::PV, PD = C::PD, MP = C::MP, NP = C::NP;
  double xl[N], jt[NO], cu[NO];
#pragma unroll
  for (int i = 0; i < N; ++i) xl[i] = rowF[i];
  adv_contract_part<N, O0, NO, MP>(sh + C::OFF_JR, xl, jt);
#pragma unroll
  for (int o = 0; o < NO; ++o) cu[o] = __dmul_rn(rowU[O0 + o], jt[o]);
#pragma unroll
  for (int i = 0; i < N; ++i) xl[i] = rowF[M * PV + i];
  adv_contract_part<N, O0, NO, MP>(sh + C::OFF_JR, xl, jt);
#pragma unroll
  for (int o = 0; o < NO; ++o)
    cu[o] = __dmul_rn(__dadd_rn(cu[o], __dmul_rn(rowU[M * PD + O0 + o], jt[o])), rowB[O0 + o]);
  // pr[i] += sum_o Jr(O0+o, i) * cu[o]
#pragma unroll
  for (int o = 0; o < NO; ++o) {
    const double2* row = reinterpret_cast<const double2*>(sh + C::OFF_JRT + (O0 + o) * NP);
#pragma unroll
    for (int i2 = 0; i2 < (N + 1) / 2; ++i2) {
      const double2 t = row[i2];
      pr[2 * i2] = fma(t.x, cu[o], pr[2 * i2]);
      if (2 * i2 + 1 < N) pr[2 * i2 + 1] = fma(t.y, cu[o], pr[2 * i2 + 1]);
    }
  }
}

// x-interpolation of one velocity row in place: N values -> M values in the same SD-wide slot
template <int N, int M>
__device__ __forceinline__ void adv_interp_row_inplace(const double* sh, double* row) {
  using C = AdvCfg<N, M>;
  constexpr int MP = C::MP;
  constexpr int H0 = ((M + 1) / 2 + 1) & ~1;
  constexpr int NA = H0 < M ? H0 : M;
  double xl[N];
#pragma unroll
  for (int i = 0; i < N; ++i) xl[i] = row[i];
  {
    double y[NA];
    adv_contract_part<N, 0, NA, MP>(sh + C::OFF_JR, xl, y);
#pragma unroll
    for (int o = 0; o < NA; ++o) row[o] = y[o];
  }
  if constexpr (H0 < M) {
    double y[M - H0];
    adv_contract_part<N, H0, M - H0, MP>(sh + C::OFF_JR, xl, y);
#pragma unroll
    for (int o = 0; o < M - H0; ++o) row[H0 + o] = y[o];
  }
}

template <int N, int M>
__global__ void __launch_bounds__(ADV_T, AdvCfg<N, M>::OCC) semb_advect_tile_kernel(const AdvTileArgs a) {
  using C = AdvCfg<N, M>;
  constexpr int EB = C::EB, S = C::S, SD = C::SD, PV = C::PV, PD = C::PD, NP = C::NP, MP = C::MP;
  extern __shared__ __align__(16) double sh[];
  double* sT = sh + C::OFF_ST;
  double* tF = sh + C::OFF_TF;
  double* tU = sh + C::OFF_TU;
  double* sBD = sh + C::OFF_BD;
  const int t = threadIdx.x;
  // tables
  for (int q = t; q < N * NP; q += ADV_T) {
    const int k = q / NP, i = q - k * NP;
    sh[C::OFF_DR + q] = i < N ? a.Dr[i * N + k] : 0.0;
    sh[C::OFF_DS + q] = i < N ? a.Ds[i * N + k] : 0.0;
  }
  for (int q = t; q < N * MP; q += ADV_T) {
    const int k = q / MP, m = q - k * MP;
    sh[C::OFF_JR + q] = m < M ? a.Jr[m + k * M] : 0.0;
    sh[C::OFF_JS + q] = m < M ? a.Js[m + k * M] : 0.0;
  }
  for (int q = t; q < M * NP; q += ADV_T) {
    const int m = q / NP, i = q - m * NP;
    sh[C::OFF_JRT + q] = i < N ? a.Jr[m + i * M] : 0.0;
    sh[C::OFF_JST + q] = i < N ? a.Js[m + i * M] : 0.0;
  }
  // mapping C: column (eC, iC); mapping R on the V rows: (jR, eR); on the D rows: (nR, eD)
  const int eC = t / N, iC = t - eC * N, colC = eC * S + iC;
  const int jR = t / EB, eR = t - jR * EB;  // jR < N valid for phase 2, jR < M for phase 4 (same split, EB*M <= 128)
  const int nbx = (a.Ex + EB - 1) / EB;
  const int nbatch = nbx * a.ney;
  for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
    const int r = b / nbx, e0 = (b - r * nbx) * EB;
    const int nbe = min(EB, a.Ex - e0);
    const bool actC = t < nbe * N;
    const bool actR2 = jR < N && eR < nbe;
    const bool actR4 = jR < M && eR < nbe;
    const size_t gV = (size_t)r * N * a.pitchV + (size_t)e0 * N + t;  // + j*pitchV: this thread's column on mshV
    __syncthreads();  // previous batch done with every tile (and the tables are in place)
    // L2 prefetch of the next batch's inputs (128-byte lines): the phases below then wait on L2, not on HBM
    if (b + (int)gridDim.x < nbatch) {
      const int b2 = b + gridDim.x, r2 = b2 / nbx, f0 = (b2 - r2 * nbx) * EB, nb2 = min(EB, a.Ex - f0);
      const int lv = (nb2 * N * 8 + 127) / 128 + 1, ld = (nb2 * M * 8 + 127) / 128 + 1;  // lines per row (+1: misalignment)
      const int nV = (6 + a.nT) * N * lv, nD = M * ld;
      for (int q = t; q < nV + nD; q += ADV_T) {
        const char* p;
        bool ok;  // stay inside the (padded) row
        if (q < nV) {
          const int f = q / (N * lv), rem = q - f * (N * lv), j = rem / lv, l = rem - j * lv;
          const double* base = f == 0 ? a.ux : f == 1 ? a.uy : f == 2 ? a.rx : f == 3 ? a.ry : f == 4 ? a.sx : f == 5 ? a.sy
                                                                                                             : a.T[f - 6];
          p = (const char*)(base + (size_t)(r2 * N + j) * a.pitchV + (size_t)f0 * N) + l * 128;
          ok = (long long)f0 * N * 8 + l * 128 < a.pitchV * 8;
        } else {
          const int rem = q - nV, n = rem / ld, l = rem - n * ld;
          p = (const char*)(a.BD + (size_t)(r2 * M + n) * a.pitchD + (size_t)f0 * M) + l * 128;
          ok = (long long)f0 * M * 8 + l * 128 < a.pitchD * 8;
        }
        if (ok) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      }
    }
    // B_D tile of the batch: rows r*M+n, columns e0*M .. e0*M + nbe*M (coalesced), consumed in phase 4
    for (int q = t; q < M * nbe * M; q += ADV_T) {
      const int n = q / (nbe * M), xx = q - n * (nbe * M), e = xx / M, m = xx - e * M;
      sBD[n * PD + e * SD + m] = a.BD[(size_t)(r * M + n) * a.pitchD + (size_t)e0 * M + xx];
    }
#pragma unroll 1
    for (int it = 0; it < a.nT; ++it) {
      const double* __restrict__ Tg = a.T[it];
      double us[N];
      // ---- phase 1 (C): T column -> registers and tile; us = Ds * T -------------------------------------------------
      if (actC) {
        double tc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) tc[j] = Tg[gV + (size_t)j * a.pitchV];
#pragma unroll
        for (int j = 0; j < N; ++j) sT[j * PV + colC] = tc[j];
        adv_contract<N, N, NP>(sh + C::OFF_DS, tc, us);
      }
      __syncthreads();
      // ---- phase 2 (R): ur = Dr * T along x, in place ---------------------------------------------------------------
      if (actR2) {
        double xl[N], ur[N];
        double* row = sT + jR * PV + eR * S;
#pragma unroll
        for (int i = 0; i < N; ++i) xl[i] = row[i];
        adv_contract<N, N, NP>(sh + C::OFF_DR, xl, ur);
#pragma unroll
        for (int i = 0; i < N; ++i) row[i] = ur[i];
      }
      __syncthreads();
      // ---- phase 3 (C): Tx, Ty (grad.jl:30-31), y-interpolation of Tx, Ty (and ux, uy for the first T) ----------------
      if (actC) {
        // one field per trip (not unrolled: bounds the loads in flight and the live registers)
        const int nf = it == 0 ? 4 : 2;
#pragma unroll 1
        for (int f = 0; f < nf; ++f) {
          double xc[N], o[M];
          if (f < 2) {
            const double* __restrict__ ca = f == 0 ? a.rx : a.ry;
            const double* __restrict__ cb = f == 0 ? a.sx : a.sy;
#pragma unroll
            for (int j = 0; j < N; ++j) {
              const size_t g = gV + (size_t)j * a.pitchV;
              xc[j] = __dadd_rn(__dmul_rn(ca[g], sT[j * PV + colC]), __dmul_rn(cb[g], us[j]));
            }
          } else {
            const double* __restrict__ src = f == 2 ? a.ux : a.uy;
#pragma unroll
            for (int j = 0; j < N; ++j) xc[j] = src[gV + (size_t)j * a.pitchV];
          }
          adv_contract<N, M, MP>(sh + C::OFF_JS, xc, o);
          if (f < 2) {
            double* dst = tF + f * M * PV + colC;
#pragma unroll
            for (int n = 0; n < M; ++n) dst[n * PV] = o[n];
          } else {
            double* dst = tU + (f - 2) * M * PD + eC * SD + iC;
#pragma unroll
            for (int n = 0; n < M; ++n) dst[n * PD] = o[n];
          }
        }
      }
      __syncthreads();
      // ---- phase 4 (R): x-interpolation, JCu = (Jux.*JTx + Juy.*JTy).*B_D (advect.jl:59-60), x-projection Jr' -----------
      // done in two halves of the M outputs to bound the live registers (a double is two registers)
      if (actR4) {
        double pr[N];
#pragma unroll
        for (int i = 0; i < N; ++i) pr[i] = 0.0;
        const double* rowF = tF + jR * PV + eR * S;
        double* rowU = tU + jR * PD + eR * SD;
        const double* rowB = sBD + jR * PD + eR * SD;
        if (it == 0) {  // Jux, Juy of the batch, shared by all T's
          adv_interp_row_inplace<N, M>(sh, rowU);
          adv_interp_row_inplace<N, M>(sh, rowU + M * PD);
        }
        constexpr int H0 = ((M + 1) / 2 + 1) & ~1;  // even split point (LDS.128 alignment of the table rows)
        adv_phase4_part<N, M, 0, (H0 < M ? H0 : M)>(sh, rowF, rowU, rowB, pr);
        if constexpr (H0 < M) adv_phase4_part<N, M, H0, M - H0>(sh, rowF, rowU, rowB, pr);
        double* rowO = tF + jR * PV + eR * S;
#pragma unroll
        for (int i = 0; i < N; ++i) rowO[i] = pr[i];  // only this thread touches the row
      }
      __syncthreads();
      // ---- phase 5 (C): y-projection Js', store Cu ---------------------------------------------------------------------
      if (actC) {
        double cl[M], cu[N];
#pragma unroll
        for (int n = 0; n < M; ++n) cl[n] = tF[n * PV + colC];
        adv_contract<M, N, NP>(sh + C::OFF_JST, cl, cu);
        double* __restrict__ og = a.out[it];
#pragma unroll
        for (int j = 0; j < N; ++j) og[gV + (size_t)j * a.pitchV] = cu[j];
      }
      // the next T's phase 1 writes sT (last read in phase 3) and its phase 3 writes tF after two more barriers
    }
  }
}

template <int N, int M>
int launch_tile(semb_ctx* ctx, const AdvTileArgs& a) {
  using C = AdvCfg<N, M>;
  auto kern = semb_advect_tile_kernel<N, M>;
  static bool attr_done[64] = {false};
  const int dev = ctx->device & 63;
  if (!attr_done[dev]) {
    SEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done[dev] = true;
  }
  const int nbatch = ((a.Ex + C::EB - 1) / C::EB) * a.ney;
  int grid = ctx->sm_count * C::OCC;
  if (grid > nbatch) grid = nbatch;
  if (grid < 1) grid = 1;
  kern<<<grid, ADV_T, C::SMEM, ctx->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

}  // namespace

// (N, M) pairs served: M = ceil(1.5 N) (examples/semPS.jl:31; cd2d.jl:54 is 8 -> 12) plus M = floor(1.5 N) for odd N
#ifdef SEMB_ADV_ONLY_9_14
#define SEMB_ADV_PAIRS(X) X(9, 14)
#else
#define SEMB_ADV_PAIRS(X) \
  X(3, 5) X(4, 6) X(5, 7) X(5, 8) X(6, 9) X(7, 10) X(7, 11) X(8, 12) X(9, 13) X(9, 14) X(10, 15) X(11, 16) X(11, 17) X(12, 18)
#endif

// returns SEMB_OK and *done = 1 if the tiled kernel ran (nT <= 4 fields T[i] -> out[i]), *done = 0 if (N, M) is not served
int semb_launch_advect_tile(semb_ctx* ctx, semb_mesh* V, semb_mesh* D, int nT, const double* const* T, const double* ux,
                            const double* uy, const double* dJr, const double* dJs, double* const* out, int* done) {
  *done = 0;
  if (V->nr != V->ns || D->nr != D->ns || nT < 1 || nT > ADV_MAXT) return SEMB_OK;
  AdvTileArgs a;
  for (int i = 0; i < ADV_MAXT; ++i) {
    a.T[i] = i < nT ? T[i] : nullptr;
    a.out[i] = i < nT ? out[i] : nullptr;
  }
  a.ux = ux;
  a.uy = uy;
  a.rx = V->arr[SEMB_RX];
  a.ry = V->arr[SEMB_RY];
  a.sx = V->arr[SEMB_SX];
  a.sy = V->arr[SEMB_SY];
  a.BD = D->arr[SEMB_B];
  a.Dr = V->dDr;
  a.Ds = V->dDs;
  a.Jr = dJr;
  a.Js = dJs;
  a.pitchV = V->pitch;
  a.pitchD = D->pitch;
  a.nT = nT;
  a.Ex = V->Ex;
  a.ney = V->ney;
#define SEMB_ADV_CASE(n, m)                       \
  if (V->nr == n && D->nr == m) {                 \
    SEMB_TRY((launch_tile<n, m>(ctx, a)));        \
    *done = 1;                                    \
    return SEMB_OK;                               \
  }
  SEMB_ADV_PAIRS(SEMB_ADV_CASE)
#undef SEMB_ADV_CASE
  return SEMB_OK;
}
