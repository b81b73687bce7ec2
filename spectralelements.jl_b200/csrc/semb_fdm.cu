// Fast-diagonalisation (FDM) Laplacian / Helmholtz preconditioner as the opM of pcg (pcg.jl:37) -- SURVEY 8f-3.
//
// The reference holds it only as commented-out sketches: the element-wise tensor solve
//     lapl_fdm(b,Bi,Sx,Sy,Sxi,Syi,Di):  u = b.*Bi; u = ABu(Syi,Sxi,u); u = u.*Di; u = ABu(Sy,Sx,u)     (lapl.jl:105-119)
// and its construction from eigen(Ax,Bx), eigen(Ay,By) of the 1-D stiffness / mass matrices with the null mode cut off
// at |1/lambda| > 1e8 (examples/p2d_explicit.jl:109-141).  As written (element Neumann problems, no overlap) it makes pcg
// slower (tests/tools/fdm_prototype.py); what works -- and what is built here, restated by the CPU checker as
// fdm_schwarz -- is the same solve on every element EXTENDED BY ONE NODE into its neighbours, combined symmetrically:
//     opM(r) = mask( W .* gs( sum_e R_e' A_e^-1 R_e (W .* r) ) ),   A_e^-1 = (Sy (x) Sx) Di (Sy (x) Sx)'
// with W = 1/sqrt(number of subdomains holding the node), S' B S = I (so Si = S' B and the Bi of the sketch is folded
// in), Di = 1/(nu*(lx + ly) + k).  Element half-lengths hx, hy come from the element-averaged metric; an element's
// extension is taken with its own half-length, so S and lambda depend on the element only through S/sqrt(h), lambda/h^2
// and three reference decompositions per direction (first / interior / last element) serve the whole mesh
// (semb_fdm_tables, semb_host.cpp).  5-8x fewer PCG iterations than no preconditioner on the BASELINE meshes.
//
// Two kernels per application:
//   semb_fdm_solve_kernel<N>   one CTA = a batch of x-consecutive elements of one element row: gathers the (N+2)^2 tiles of
//                              W.*r, applies the four contractions out of registers (the strip kernel's two thread<->line
//                              mappings), writes the tiles to a tile-major ("fat") buffer
//   semb_fdm_combine_kernel    per node: sums the (up to four) tile entries that land on it and on its duplicates in the
//                              fixed (x pairs, then y pairs) association of gatherScatter.jl:13, applies W and the mask,
//                              and -- inside pcg -- accumulates sum(r .* h .* mult) (pcg.jl:45) deterministically
#include "semb_reduce.cuh"
#include "semb_vec.cuh"

namespace {

struct FdmArgs {
  const double* r;      // residual (continuous)
  double* out;          // h = opM(r)
  double* fat;          // tile buffer: row (ey*N2 + jj) * fpitch + ex*N2 + ii
  const double* tab;    // [dir 2][class 4][N2*N2 + N2]: S (column-major: S[ii + c*N2]) then lambda
  const double* hx;     // [ney][Ex] element half-lengths
  const double* hy;
  const double* wx;     // W = wx[x] * wy[y]
  const double* wy;
  const double* mult_x; // mult(x,y) = mult_x[x] * mult_y[y]  (PCG reduction)
  const double* mult_y;
  long long pitch, fpitch;
  int N, Ex, Ey, ey0, ney, nxl, nyl, perx, pery;
  int mx0, mx1, my0, my1;
  double nu, k;
  SembScal* scal;
  double* partials;
  unsigned* counter;
  int pcg;      // 1: inside pcg (early exit on done; reduction + advance), 2: same, first call (pcg.jl:25-33 state)
};

// class of an element in a direction: 0 interior (neighbours on both sides), 1 first, 2 last, 3 single
__device__ __forceinline__ int fdm_class(int e, int E, int per) {
  if (per) return 0;
  return (e == 0 ? 1 : 0) | (e == E - 1 ? 2 : 0);
}

template <int N>
struct FdmCfg {
  static constexpr int N2 = N + 2;
  static constexpr int T = 256;
  static constexpr int BX = T / N2;        // elements per CTA
  static constexpr int S = N2 | 1;         // element stride in the tile buffer (odd: conflict-free in both mappings)
  static constexpr int PW = BX * S;        // row pitch of the tile buffer
  static constexpr int TSZ = N2 * N2 + N2; // one table: S then lambda
  static constexpr int SMEM = (N2 * PW + 5 * TSZ) * 8;
};

template <int N>
__global__ void __launch_bounds__(256) semb_fdm_solve_kernel(const FdmArgs a) {
  using C = FdmCfg<N>;
  constexpr int N2 = C::N2, BX = C::BX, S = C::S, PW = C::PW, TSZ = C::TSZ;
  extern __shared__ __align__(16) double sm[];
  double* S1 = sm;                 // [N2][PW] tiles
  double* sTx = S1 + N2 * PW;      // [4][TSZ] x tables, all classes
  double* sTy = sTx + 4 * TSZ;     // [TSZ] y table of this element row
  if (a.pcg && a.scal->done) return;
  const int t = threadIdx.x;
  const int ey = blockIdx.y, eyg = a.ey0 + ey;
  const int e0 = blockIdx.x * BX;
  const int nbe = min(BX, a.Ex - e0);
  const int cy = fdm_class(eyg, a.Ey, a.pery);
  for (int q = t; q < 4 * TSZ; q += C::T) sTx[q] = a.tab[q];
  for (int q = t; q < TSZ; q += C::T) sTy[q] = a.tab[4 * TSZ + cy * TSZ + q];
  // ---- mapping B: thread <-> (element eB, tile column iB): load the column of W.*r ---------------------------------
  const int eB = t / N2, iB = t - eB * N2;
  const bool actB = eB < nbe;
  const int ex = e0 + eB;
  double col[N2];
  {
    // tile column -> global column: [left neighbour's node N-2, own 0..N-1, right neighbour's node 1]
    int x = -1;
    if (actB) {
      if (iB == 0) x = (ex > 0 || a.perx) ? ((ex + a.Ex - 1) % a.Ex) * N + N - 2 : -1;
      else if (iB == N + 1) x = (ex < a.Ex - 1 || a.perx) ? ((ex + 1) % a.Ex) * N + 1 : -1;
      else x = ex * N + iB - 1;
    }
    const double wxv = x >= 0 ? a.wx[x] : 0.0;
#pragma unroll
    for (int jj = 0; jj < N2; ++jj) {
      int y;  // tile row -> local row (this rank's slab; the periodic wrap is local on one rank)
      if (jj == 0) y = (ey > 0) ? (ey - 1) * N + N - 2 : ((a.pery && a.ney == a.Ey) ? (a.ney - 1) * N + N - 2 : -1);
      else if (jj == N + 1) y = (ey < a.ney - 1) ? (ey + 1) * N + 1 : ((a.pery && a.ney == a.Ey) ? 1 : -1);
      else y = ey * N + jj - 1;
      col[jj] = (x >= 0 && y >= 0) ? __dmul_rn(__dmul_rn(wxv, a.wy[y]), a.r[(size_t)y * a.pitch + x]) : 0.0;
    }
  }
  __syncthreads();  // tables are in shared memory
  // t1 = Sy' * col (along y): t1[c] = sum_jj Sy[jj][c] col[jj]
  {
    double o[N2];
#pragma unroll
    for (int c = 0; c < N2; ++c) o[c] = 0.0;
#pragma unroll
    for (int jj = 0; jj < N2; ++jj) {
#pragma unroll
      for (int c = 0; c < N2; ++c) o[c] = fma(sTy[jj + c * N2], col[jj], o[c]);
    }
    if (t < BX * N2) {
#pragma unroll
      for (int c = 0; c < N2; ++c) S1[c * PW + eB * S + iB] = o[c];
    }
  }
  __syncthreads();
  // ---- mapping A: thread <-> (row jA = y-mode, element eA): Sx' along x, scale by Di, Sx back ---------------------------
  {
    const int jA = t / BX, eA = t - jA * BX;
    if (jA < N2 && eA < nbe) {
      const int exA = e0 + eA;
      const double* Tx = sTx + fdm_class(exA, a.Ex, a.perx) * TSZ;
      const double hx = a.hx[(size_t)ey * a.Ex + exA], hy = a.hy[(size_t)ey * a.Ex + exA];
      const double ly = sTy[N2 * N2 + jA] / (hy * hy);
      double c[N2], o[N2];
#pragma unroll
      for (int i = 0; i < N2; ++i) c[i] = S1[jA * PW + eA * S + i];
#pragma unroll
      for (int m = 0; m < N2; ++m) o[m] = 0.0;
#pragma unroll
      for (int i = 0; i < N2; ++i) {
#pragma unroll
        for (int m = 0; m < N2; ++m) o[m] = fma(Tx[i + m * N2], c[i], o[m]);
      }
      // Di = 1/(nu*(lx+ly)+k) (p2d_explicit.jl:131-134), times the 1/(hx*hy) of the two S/sqrt(h) pairs
      const double sc = 1.0 / (hx * hy);
#pragma unroll
      for (int m = 0; m < N2; ++m) {
        const double lam = a.nu * (Tx[N2 * N2 + m] / (hx * hx) + ly) + a.k;
        double d = 1.0 / lam;
        if (!(fabs(d) <= 1e8)) d = 0.0;  // null mode of an all-free subdomain; padding modes (lambda = inf) give 0 anyway
        o[m] *= d * sc;
      }
#pragma unroll
      for (int i = 0; i < N2; ++i) c[i] = 0.0;
#pragma unroll
      for (int m = 0; m < N2; ++m) {
#pragma unroll
        for (int i = 0; i < N2; ++i) c[i] = fma(Tx[i + m * N2], o[m], c[i]);
      }
#pragma unroll
      for (int i = 0; i < N2; ++i) S1[jA * PW + eA * S + i] = c[i];
    }
  }
  __syncthreads();
  // ---- mapping B: Sy back along y, store the tile column -----------------------------------------------------------
  if (actB) {
    double c[N2], o[N2];
#pragma unroll
    for (int m = 0; m < N2; ++m) c[m] = S1[m * PW + eB * S + iB];
#pragma unroll
    for (int jj = 0; jj < N2; ++jj) o[jj] = 0.0;
#pragma unroll
    for (int m = 0; m < N2; ++m) {
#pragma unroll
      for (int jj = 0; jj < N2; ++jj) o[jj] = fma(sTy[jj + m * N2], c[m], o[jj]);
    }
    double* dst = a.fat + (size_t)(ey * N2) * a.fpitch + (size_t)ex * N2 + iB;
#pragma unroll
    for (int jj = 0; jj < N2; ++jj) dst[(size_t)jj * a.fpitch] = o[jj];
  }
}

// value the tiles leave on the LOCAL copy (ex,i,ey,j) of a node: own tile entry + the extension entries of the
// neighbouring tiles that land on it (fixed order of additions)
__device__ __forceinline__ double fdm_zloc(const FdmArgs& a, int ex, int i, int ey, int j) {
  const int N = a.N, N2 = N + 2;
  auto fat = [&](int fx, int ii, int fy, int jj) {
    return a.fat[(size_t)(fy * N2 + jj) * a.fpitch + (size_t)fx * N2 + ii];
  };
  const bool wrapy = a.pery && a.ney == a.Ey;
  const bool hasL = ex > 0 || a.perx, hasR = ex < a.Ex - 1 || a.perx;
  const bool hasB = ey > 0 || wrapy, hasT = ey < a.ney - 1 || wrapy;
  const int exL = (ex + a.Ex - 1) % a.Ex, exR = (ex + 1) % a.Ex;
  const int eyB = (ey + a.ney - 1) % a.ney, eyT = (ey + 1) % a.ney;
  const bool l = i == 1 && hasL, r = i == N - 2 && hasR, b = j == 1 && hasB, tt = j == N - 2 && hasT;
  double z = fat(ex, i + 1, ey, j + 1);
  if (l) z = __dadd_rn(z, fat(exL, N + 1, ey, j + 1));
  if (r) z = __dadd_rn(z, fat(exR, 0, ey, j + 1));
  if (b) z = __dadd_rn(z, fat(ex, i + 1, eyB, N + 1));
  if (tt) z = __dadd_rn(z, fat(ex, i + 1, eyT, 0));
  if (l && b) z = __dadd_rn(z, fat(exL, N + 1, eyB, N + 1));
  if (r && b) z = __dadd_rn(z, fat(exR, 0, eyB, N + 1));
  if (l && tt) z = __dadd_rn(z, fat(exL, N + 1, eyT, 0));
  if (r && tt) z = __dadd_rn(z, fat(exR, 0, eyT, 0));
  return z;
}

__global__ void __launch_bounds__(256) semb_fdm_combine_kernel(const FdmArgs a) {
  __shared__ double red[32];
  __shared__ double sh_tot[2];
  if (a.pcg && a.scal->done) return;
  const int N = a.N;
  const bool wrapy = a.pery && a.ney == a.Ey;
  double acc = 0.0;
  for (int row = blockIdx.y * blockDim.y + threadIdx.y; row < a.nyl; row += gridDim.y * blockDim.y) {
    const int ey = row / N, j = row - ey * N;
    // partner copy in y (gatherScatter.jl:13: duplicates of a node on an element interface)
    int eyp = -1, jp = 0;
    if (j == N - 1 && (ey < a.ney - 1 || wrapy)) eyp = (ey + 1) % a.ney, jp = 0;
    else if (j == 0 && (ey > 0 || wrapy)) eyp = (ey + a.ney - 1) % a.ney, jp = N - 1;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < a.nxl; x += gridDim.x * blockDim.x) {
      const int ex = x / N, i = x - ex * N;
      int exp_ = -1, ip = 0;
      if (i == N - 1 && (ex < a.Ex - 1 || a.perx)) exp_ = (ex + 1) % a.Ex, ip = 0;
      else if (i == 0 && (ex > 0 || a.perx)) exp_ = (ex + a.Ex - 1) % a.Ex, ip = N - 1;
      double g = fdm_zloc(a, ex, i, ey, j);
      if (exp_ >= 0) g = __dadd_rn(g, fdm_zloc(a, exp_, ip, ey, j));      // x pair first ...
      if (eyp >= 0) {
        double h = fdm_zloc(a, ex, i, eyp, jp);
        if (exp_ >= 0) h = __dadd_rn(h, fdm_zloc(a, exp_, ip, eyp, jp));
        g = __dadd_rn(g, h);                                             // ... then the y pair of the x pairs
      }
      const bool z = (x == 0 && a.mx0) || (x == a.nxl - 1 && a.mx1) || (row == 0 && a.my0) || (row == a.nyl - 1 && a.my1);
      const double h = __dmul_rn(z ? 0.0 : 1.0, __dmul_rn(__dmul_rn(a.wx[x], a.wy[row]), g));
      const size_t idx = (size_t)row * a.pitch + x;
      a.out[idx] = h;
      if (a.pcg) acc += __dmul_rn(__dmul_rn(a.r[idx], h), a.mult_x[x] * a.mult_y[row]);  // pcg.jl:45
    }
  }
  if (a.pcg) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x, nb = gridDim.x * gridDim.y;
    const double bs = semb_block_sum(acc, red, tid, nt);
    if (semb_last_block_uniform(bs, 0.0, a.partials, nullptr, a.counter, nb, bid, red, tid, nt, sh_tot)) {
      if (tid == 0) {
        // t = sum(r .* h .* mult); norm(r,Inf) was left in red[2] by the init / update kernel: advance the PCG state
        SembScal* s = a.scal;
        const double tnew = sh_tot[0], rmax = s->red[2];
        if (a.pcg == 2) {
          s->t = tnew, s->t_prev = 0.0, s->iters = 0, s->warned = 0;
        } else {
          s->t_prev = s->t, s->t = tnew, s->iters += 1;
        }
        s->rmax = rmax;
        int done = !(rmax > s->tol);                                        // pcg.jl:36
        if (!done && s->iters >= s->maxiter) { done = 1; s->warned = 1; }   // pcg.jl:39
        s->done = done;
      }
    }
  }
}

// element half-lengths: hx = mean over the element of Jac*sqrt(sx^2+sy^2), hy = mean of Jac*sqrt(rx^2+ry^2), written with
// the arrays every mesh holds: Jac = B/(wr_i*ws_j), G22 = B*(sx^2+sy^2), G11 = B*(rx^2+ry^2) (mesh.jl:117-123)
__global__ void semb_fdm_lengths_kernel(const double* __restrict__ B, const double* __restrict__ G11,
                                        const double* __restrict__ G22, const double* __restrict__ wr,
                                        const double* __restrict__ ws, long long pitch, int N, int Ex, int ney, double* hx,
                                        double* hy) {
  __shared__ double red[32];
  const int e = blockIdx.x;
  if (e >= Ex * ney) return;
  const int ey = e / Ex, ex = e - ey * Ex;
  double sx = 0.0, sy = 0.0;
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) {
    const int j = q / N, i = q - j * N;
    const size_t idx = (size_t)(ey * N + j) * pitch + (size_t)ex * N + i;
    const double b = B[idx], jw = b / (wr[i] * ws[j]);
    sx += jw * sqrt(G22[idx] / b);
    sy += jw * sqrt(G11[idx] / b);
  }
  const double tx = semb_block_sum(sx, red, threadIdx.x, blockDim.x);
  const double ty = semb_block_sum(sy, red, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) {
    hx[e] = tx / (double)(N * N);
    hy[e] = ty / (double)(N * N);
  }
}

template <int N>
int launch_solve(semb_ctx* ctx, const FdmArgs& a) {
  using C = FdmCfg<N>;
  auto kern = semb_fdm_solve_kernel<N>;
  static bool attr_done[64] = {false};
  const int dev = ctx->device & 63;
  if (!attr_done[dev]) {
    SEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done[dev] = true;
  }
  kern<<<dim3((a.Ex + C::BX - 1) / C::BX, a.ney), 256, C::SMEM, ctx->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

}  // namespace

struct semb_fdm {
  semb_mesh* m = nullptr;
  double nu = 1.0, k = 0.0;
  int mx0 = 0, mx1 = 0, my0 = 0, my1 = 0;
  double *d_hx = nullptr, *d_hy = nullptr, *d_tab = nullptr, *d_fat = nullptr, *d_wx = nullptr, *d_wy = nullptr;
  long long fpitch = 0;
};

int semb_fdm_free_impl(semb_fdm* f) {
  if (!f) return SEMB_OK;
  cudaFree(f->d_hx);
  cudaFree(f->d_hy);
  cudaFree(f->d_tab);
  cudaFree(f->d_fat);
  cudaFree(f->d_wx);
  cudaFree(f->d_wy);
  delete f;
  return SEMB_OK;
}

// bcflags: Dirichlet flags of the GLOBAL boundary lines (x0, x1, y0, y1) after the periodic override (parse_bc);
// bcglob: the same for the whole domain (this rank's slab may not touch the y boundaries)
int semb_fdm_create_impl(semb_mesh* m, double nu, double k, int mx0, int mx1, int my0, int my1, int gy0, int gy1,
                         semb_fdm** out) {
  semb_ctx* c = m->ctx;
  const int N = m->nr, N2 = N + 2, TSZ = N2 * N2 + N2;
  semb_fdm* f = new semb_fdm();
  *out = f;
  f->m = m;
  f->nu = nu;
  f->k = k;
  f->mx0 = mx0, f->mx1 = mx1, f->my0 = my0, f->my1 = my1;
  // reference decompositions: classes 0 interior, 1 first, 2 last, 3 single; kinds 0 neighbour, 1 Dirichlet, 2 free
  std::vector<double> tab((size_t)2 * 4 * TSZ, 0.0);
  for (int dir = 0; dir < 2; ++dir) {
    const std::vector<double>& D = dir == 0 ? m->hDr : m->hDs;
    const std::vector<double>& w = dir == 0 ? m->hwr : m->hws;
    const int klo = dir == 0 ? (mx0 ? 1 : 2) : (gy0 ? 1 : 2), khi = dir == 0 ? (mx1 ? 1 : 2) : (gy1 ? 1 : 2);
    for (int cls = 0; cls < 4; ++cls) {
      double* T = tab.data() + ((size_t)dir * 4 + cls) * TSZ;
      SEMB_TRY(semb_fdm_tables(N, D.data(), w.data(), (cls & 1) ? klo : 0, (cls & 2) ? khi : 0, T, T + N2 * N2));
    }
  }
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_tab, tab.size() * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(f->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  // counting weights W = 1/sqrt(cx*cy), separable: c = 1 + [neighbour below && i <= 1] + [neighbour above && i >= N-2]
  std::vector<double> wx((size_t)m->pitch, 0.0), wy((size_t)m->nyl, 0.0);
  for (int x = 0; x < m->nxl; ++x) {
    const int e = x / N, i = x % N;
    const int cx = 1 + (((e > 0 || m->perx) && i <= 1) ? 1 : 0) + (((e < m->Ex - 1 || m->perx) && i >= N - 2) ? 1 : 0);
    wx[x] = 1.0 / std::sqrt((double)cx);
  }
  for (int y = 0; y < m->nyl; ++y) {
    const int eg = m->ey0 + y / N, j = y % N;
    const int cy = 1 + (((eg > 0 || m->pery) && j <= 1) ? 1 : 0) + (((eg < m->Ey - 1 || m->pery) && j >= N - 2) ? 1 : 0);
    wy[y] = 1.0 / std::sqrt((double)cy);
  }
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_wx, wx.size() * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_wy, wy.size() * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(f->d_wx, wx.data(), wx.size() * sizeof(double), cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMemcpy(f->d_wy, wy.data(), wy.size() * sizeof(double), cudaMemcpyHostToDevice));
  // element half-lengths
  const size_t ne = (size_t)m->Ex * m->ney;
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_hx, ne * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_hy, ne * sizeof(double)));
  double *d_wr = nullptr, *d_ws = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d_wr, N * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&d_ws, N * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(d_wr, m->hwr.data(), N * sizeof(double), cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMemcpy(d_ws, m->hws.data(), N * sizeof(double), cudaMemcpyHostToDevice));
  semb_fdm_lengths_kernel<<<(unsigned)ne, 64, 0, c->stream>>>(m->arr[SEMB_B], m->arr[SEMB_G11], m->arr[SEMB_G22], d_wr, d_ws,
                                                              m->pitch, N, m->Ex, m->ney, f->d_hx, f->d_hy);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_wr);
  cudaFree(d_ws);
  SEMB_CHECK_CUDA(e);
  c->launches++;
  f->fpitch = ((long long)m->Ex * N2 + 15) / 16 * 16;
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_fat, (size_t)f->fpitch * m->ney * N2 * sizeof(double)));
  return SEMB_OK;
}

// h = opM(r); pcg: 0 stand-alone, 1 inside pcg (reduction + advance), 2 first call of a solve
int semb_fdm_apply_impl(semb_fdm* f, const double* r, double* out, int pcg) {
  semb_mesh* m = f->m;
  semb_ctx* c = m->ctx;
  FdmArgs a;
  a.r = r;
  a.out = out;
  a.fat = f->d_fat;
  a.tab = f->d_tab;
  a.hx = f->d_hx;
  a.hy = f->d_hy;
  a.wx = f->d_wx;
  a.wy = f->d_wy;
  a.mult_x = m->d_wx1d;
  a.mult_y = m->d_wy1d;
  a.pitch = m->pitch;
  a.fpitch = f->fpitch;
  a.N = m->nr;
  a.Ex = m->Ex;
  a.Ey = m->Ey;
  a.ey0 = m->ey0;
  a.ney = m->ney;
  a.nxl = m->nxl;
  a.nyl = m->nyl;
  a.perx = m->perx;
  a.pery = m->pery;
  a.mx0 = f->mx0, a.mx1 = f->mx1, a.my0 = f->my0, a.my1 = f->my1;
  a.nu = f->nu;
  a.k = f->k;
  a.scal = m->d_scal;
  a.partials = m->d_partials;
  a.counter = m->d_counters + 6;
  a.pcg = pcg;
  switch (m->nr) {
#define SEMB_CASE(n) \
  case n:            \
    SEMB_TRY(launch_solve<n>(c, a)); \
    break;
    SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9) SEMB_CASE(10) SEMB_CASE(11)
    SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16) SEMB_CASE(17)
#undef SEMB_CASE
    default:
      semb_set_error("fdm: no kernel for nr = %d (3..17)", m->nr);
      return SEMB_EINVAL;
  }
  int bx = 32;
  while (bx < 256 && bx < m->nxl) bx <<= 1;
  const int by = 256 / bx;
  int gx = (m->nxl + bx - 1) / bx;
  if (gx > 64) gx = 64;
  int gy = (m->nyl + by - 1) / by;
  const int cap = m->npartials / gx;
  if (gy > cap) gy = cap;
  if (gy > c->sm_count * 16) gy = c->sm_count * 16;
  if (gy < 1) gy = 1;
  semb_fdm_combine_kernel<<<dim3(gx, gy), dim3(bx, by), 0, c->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  c->launches++;
  return SEMB_OK;
}
