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
// All inputs are staged into shared memory with cp.async one pass / one batch AHEAD (issued right after the barrier that
// retires the last read of their buffer), so HBM latency overlaps the contractions of the current pass.
// Phases per batch (one __syncthreads between them):
//   1 C: T column from the staged tile, us = Ds*T       2 R: ur = Dr*T in place
//   3 C: Tx,Ty from ur,us and the metric terms; y-interpolation (Js) of Tx,Ty,ux,uy -> 4 tiles of M rows
//   4 R: x-interpolation (Jr) of the 4 rows, pointwise product with B_D, x-projection (Jr') -> tile (in place)
//   5 C: y-projection (Js'), coalesced store of Cu
//
// The kernel is bound by shared-memory wavefronts, not by FP64 issue or HBM (profiles/r01_advtile_r1h.txt: a broadcast
// table load per FMA), so the contractions are organised to need as few table loads as possible:
//   * even-odd: GLL interpolation matrices are centro-symmetric (J(M-1-m,N-1-k) = J(m,k)) and derivative matrices
//     centro-antisymmetric, so y = A x splits into two half-size products on e_k = x_k + x_{N-1-k}, o_k = x_k - x_{N-1-k}
//     (half the FMAs and half the table loads; same idea as StripTab in semb_strip.cuh, here for rectangular A);
//   * pairing: two lines that take the same matrix (Tx,Ty / ux,uy) are contracted together, one table load feeds both.
// The host launches this kernel only when the matrices pass the symmetry test (mesh `eo` flag; the J's are built by the
// library from GLL nodes).
//
// The two interpolation directions commute exactly in exact arithmetic; the reference applies Jr first (ABu.jl:14-33),
// here Js is applied first (the column mapping already holds the y-lines): results agree to rounding (~1e-16
// relative), well inside the 1e-12 contract.  The projection is applied in the reference's order (Jr' then Js').
//
// Several T's that share the advecting velocity (makeRHS!: exH[i] = -advect(uh[i],vx,vy,...) for i = 1..k,
// convectionDiffusion.jl:100-105) are processed in ONE launch: ux,uy,B_D and the metric terms are loaded and Jux,Juy
// interpolated once per batch, each T then costs T + Cu of HBM traffic.
#include "semb_eo.cuh"
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
  static constexpr int SD = M | 1;       // same for the D-grid tiles
  static constexpr int PV = EB * S, PD = EB * SD;
  using TD = EoTab<N, N, -1>;   // Dr, Ds
  using TJ = EoTab<N, M, +1>;   // Jr, Js
  using TP = EoTab<M, N, +1>;   // Jr', Js'
  static constexpr int OFF_DR = 0, OFF_DS = TD::SIZE, OFF_JR = 2 * TD::SIZE, OFF_JS = OFF_JR + TJ::SIZE,
                       OFF_JRT = OFF_JS + TJ::SIZE, OFF_JST = OFF_JRT + TP::SIZE, TAB = (OFF_JST + TP::SIZE + 1) & ~1;
  static constexpr int OFF_ST = TAB;                   // [N][PV]    T (staged by cp.async) -> ur
  static constexpr int OFF_IN = OFF_ST + N * PV;       // [6][N][PV] staged ux, uy, rx, ry, sx, sy of the batch
  static constexpr int OFF_TF = OFF_IN + 6 * N * PV;   // [2][M][PV] y-interpolated Tx, Ty; [0] reused for Jr' JCu
  static constexpr int OFF_TU = OFF_TF + 2 * M * PV;   // [2][M][PD] y-interpolated ux, uy (N per element), then Jux, Juy (M per
                                                       // element) written in place by the row's owner thread
  static constexpr int OFF_BD = OFF_TU + 2 * M * PD;   // [M][PD]
  static constexpr int SMEM_DOUBLES = OFF_BD + M * PD;
  static constexpr int SMEM = SMEM_DOUBLES * 8;
  static constexpr int OCC0 = (227 * 1024) / (SMEM + 1024);
  static constexpr int OCC = OCC0 < 1 ? 1 : (OCC0 > 4 ? 4 : OCC0);
};

template <int N, int M>
__global__ void __launch_bounds__(ADV_T, AdvCfg<N, M>::OCC) semb_advect_tile_kernel(const AdvTileArgs a) {
  using C = AdvCfg<N, M>;
  constexpr int EB = C::EB, S = C::S, SD = C::SD, PV = C::PV, PD = C::PD;
  extern __shared__ __align__(16) double sh[];
  double* sT = sh + C::OFF_ST;
  double* sIn = sh + C::OFF_IN;
  double* tF = sh + C::OFF_TF;
  double* tU = sh + C::OFF_TU;
  double* sBD = sh + C::OFF_BD;
  const int t = threadIdx.x;
  // even-odd tables (a.Dr: row-major D(i,k) = Dr[i*N+k]; a.Jr: column-major J(m,i) = Jr[m + i*M])
  C::TD::fill(sh + C::OFF_DR, t, ADV_T, [&](int i, int k) { return a.Dr[i * N + k]; });
  C::TD::fill(sh + C::OFF_DS, t, ADV_T, [&](int i, int k) { return a.Ds[i * N + k]; });
  C::TJ::fill(sh + C::OFF_JR, t, ADV_T, [&](int m, int i) { return a.Jr[m + i * M]; });
  C::TJ::fill(sh + C::OFF_JS, t, ADV_T, [&](int m, int i) { return a.Js[m + i * M]; });
  C::TP::fill(sh + C::OFF_JRT, t, ADV_T, [&](int i, int m) { return a.Jr[m + i * M]; });
  C::TP::fill(sh + C::OFF_JST, t, ADV_T, [&](int i, int m) { return a.Js[m + i * M]; });
  // mapping C: column (eC, iC); mapping R on the V rows: (jR, eR); on the D rows: (nR, eD)
  const int eC = t / N, iC = t - eC * N, colC = eC * S + iC;
  const int jR = t / EB, eR = t - jR * EB;  // jR < N valid for phase 2, jR < M for phase 4 (same split, EB*M <= 128)
  const int nbx = (a.Ex + EB - 1) / EB;
  const int nbatch = nbx * a.ney;
  // ---- asynchronous staging (cp.async, 8-byte granules: any strip alignment): the inputs of the NEXT pass / batch are
  // issued as soon as the barrier that retires the last read of their buffer has passed, and land while the current
  // pass computes; each column-owner thread copies its own column, so the global side is coalesced ----------------------
  auto cp8 = [](double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
  };
  auto batch_of = [&](int b, int* r, int* e0, int* nbe) {
    *r = b / nbx;
    *e0 = (b - *r * nbx) * EB;
    *nbe = min(EB, a.Ex - *e0);
  };
  auto issue_T = [&](int b, int it) {
    int r, e0, nbe;
    batch_of(b, &r, &e0, &nbe);
    if (t < nbe * N) {
      const double* src = a.T[it] + (size_t)r * N * a.pitchV + (size_t)e0 * N + t;
#pragma unroll
      for (int j = 0; j < N; ++j) cp8(sT + j * PV + colC, src + (size_t)j * a.pitchV);
    }
  };
  auto issue_inputs = [&](int b) {
    int r, e0, nbe;
    batch_of(b, &r, &e0, &nbe);
    if (t < nbe * N) {
      const size_t g = (size_t)r * N * a.pitchV + (size_t)e0 * N + t;
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        const double* src = (f == 0 ? a.ux : f == 1 ? a.uy : f == 2 ? a.rx : f == 3 ? a.ry : f == 4 ? a.sx : a.sy) + g;
#pragma unroll
        for (int j = 0; j < N; ++j) cp8(sIn + (f * N + j) * PV + colC, src + (size_t)j * a.pitchV);
      }
    }
  };
  auto issue_BD = [&](int b) {
    int r, e0, nbe;
    batch_of(b, &r, &e0, &nbe);
    for (int q = t; q < M * nbe * M; q += ADV_T) {
      const int n = q / (nbe * M), xx = q - n * (nbe * M), e = xx / M, m = xx - e * M;
      cp8(sBD + n * PD + e * SD + m, a.BD + (size_t)(r * M + n) * a.pitchD + (size_t)e0 * M + xx);
    }
  };
  if ((int)blockIdx.x < nbatch) {
    issue_T(blockIdx.x, 0);
    issue_inputs(blockIdx.x);
    issue_BD(blockIdx.x);
  }
  for (int b = blockIdx.x; b < nbatch; b += gridDim.x) {
    int r, e0, nbe;
    batch_of(b, &r, &e0, &nbe);
    const bool actC = t < nbe * N;
    const bool actR2 = jR < N && eR < nbe;
    const bool actR4 = jR < M && eR < nbe;
    const size_t gV = (size_t)r * N * a.pitchV + (size_t)e0 * N + t;  // + j*pitchV: this thread's column on mshV
    const int bnext = b + gridDim.x;
#pragma unroll 1
    for (int it = 0; it < a.nT; ++it) {
      const bool last = it + 1 == a.nT;
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();  // staged T (and, for the first T, the batch inputs) visible; the tables are in place;
                        // the previous pass is done with tF
      double us[1][N];
      // ---- phase 1 (C): T column -> registers; us = Ds * T --------------------------------------------------------------
      if (actC) {
        double tc[1][N];
#pragma unroll
        for (int j = 0; j < N; ++j) tc[0][j] = sT[j * PV + colC];
        eo_contract<N, N, -1, 1>(sh + C::OFF_DS, tc, us);
      }
      __syncthreads();
      // ---- phase 2 (R): ur = Dr * T along x, in place ---------------------------------------------------------------
      if (actR2) {
        double xl[1][N], ur[1][N];
        double* row = sT + jR * PV + eR * S;
#pragma unroll
        for (int i = 0; i < N; ++i) xl[0][i] = row[i];
        eo_contract<N, N, -1, 1>(sh + C::OFF_DR, xl, ur);
#pragma unroll
        for (int i = 0; i < N; ++i) row[i] = ur[0][i];
      }
      __syncthreads();
      // ---- phase 3 (C): Tx, Ty (grad.jl:30-31), y-interpolation of the pair Tx, Ty (and ux, uy for the first T) -----------
      if (actC) {
        {
          double xc[2][N], o[2][M];
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const int q = j * PV + colC;
            const double ur = sT[q];
            xc[0][j] = __dadd_rn(__dmul_rn(sIn[2 * N * PV + q], ur), __dmul_rn(sIn[4 * N * PV + q], us[0][j]));  // rx, sx
            xc[1][j] = __dadd_rn(__dmul_rn(sIn[3 * N * PV + q], ur), __dmul_rn(sIn[5 * N * PV + q], us[0][j]));  // ry, sy
          }
          eo_contract<N, M, +1, 2>(sh + C::OFF_JS, xc, o);
#pragma unroll
          for (int n = 0; n < M; ++n) {
            tF[n * PV + colC] = o[0][n];
            tF[(M + n) * PV + colC] = o[1][n];
          }
        }
        if (it == 0) {
          double xc[2][N], o[2][M];
#pragma unroll
          for (int j = 0; j < N; ++j) {
            xc[0][j] = sIn[j * PV + colC];        // ux
            xc[1][j] = sIn[(N + j) * PV + colC];  // uy
          }
          eo_contract<N, M, +1, 2>(sh + C::OFF_JS, xc, o);
          double* dst = tU + eC * SD + iC;
#pragma unroll
          for (int n = 0; n < M; ++n) {
            dst[n * PD] = o[0][n];
            dst[(M + n) * PD] = o[1][n];
          }
        }
      }
      __syncthreads();
      // sT is free (and, after the last T, the staged inputs): start the copies of the next pass
      if (!last) {
        issue_T(b, it + 1);
      } else if (bnext < nbatch) {
        issue_T(bnext, 0);
        issue_inputs(bnext);
      }
      // ---- phase 4 (R): x-interpolation, JCu = (Jux.*JTx + Juy.*JTy).*B_D (advect.jl:59-60), x-projection Jr' -----------
      if (actR4) {
        double* rowF = tF + jR * PV + eR * S;
        double* rowU = tU + jR * PD + eR * SD;
        const double* rowB = sBD + jR * PD + eR * SD;
        if (it == 0) {  // Jux, Juy of the batch, shared by all T's: N values -> M values in the same SD-wide slot
          double xl[2][N], ju[2][M];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            xl[0][i] = rowU[i];
            xl[1][i] = rowU[M * PD + i];
          }
          eo_contract<N, M, +1, 2>(sh + C::OFF_JR, xl, ju);
#pragma unroll
          for (int m = 0; m < M; ++m) {
            rowU[m] = ju[0][m];
            rowU[M * PD + m] = ju[1][m];
          }
        }
        double cu[1][M], pr[1][N];
        {
          double xl[2][N], jt[2][M];
#pragma unroll
          for (int i = 0; i < N; ++i) {
            xl[0][i] = rowF[i];
            xl[1][i] = rowF[M * PV + i];
          }
          eo_contract<N, M, +1, 2>(sh + C::OFF_JR, xl, jt);
#pragma unroll
          for (int m = 0; m < M; ++m)
            cu[0][m] = __dmul_rn(__dadd_rn(__dmul_rn(rowU[m], jt[0][m]), __dmul_rn(rowU[M * PD + m], jt[1][m])), rowB[m]);
        }
        eo_contract<M, N, +1, 1>(sh + C::OFF_JRT, cu, pr);
#pragma unroll
        for (int i = 0; i < N; ++i) rowF[i] = pr[0][i];  // only this thread touches the row
      }
      __syncthreads();
      if (last && bnext < nbatch) issue_BD(bnext);  // the B_D tile is free
      // ---- phase 5 (C): y-projection Js', store Cu ---------------------------------------------------------------------
      if (actC) {
        double cl[1][M], cu[1][N];
#pragma unroll
        for (int n = 0; n < M; ++n) cl[0][n] = tF[n * PV + colC];
        eo_contract<M, N, +1, 1>(sh + C::OFF_JST, cl, cu);
        double* __restrict__ og = a.out[it];
#pragma unroll
        for (int j = 0; j < N; ++j) og[gV + (size_t)j * a.pitchV] = cu[0][j];
      }
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
// or the mesh's derivative matrices are not centro-antisymmetric (V->eo; the J's are GLL interpolants, symmetric)
int semb_launch_advect_tile(semb_ctx* ctx, semb_mesh* V, semb_mesh* D, int nT, const double* const* T, const double* ux,
                            const double* uy, const double* dJr, const double* dJs, double* const* out, int* done) {
  *done = 0;
  if (V->nr != V->ns || D->nr != D->ns || nT < 1 || nT > ADV_MAXT || !V->eo) return SEMB_OK;
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
