// Element-local kernels of the Stokes pressure/velocity split: grad^T (grad.jl:44-63), the velocity-grid
// part of diver (diver.jl:17-31) and of diver^T (diver.jl:53-63), and the pointwise middle of approxHlmzInv
// (diver.jl:92-104).  The reference's stokes.jl / diver.jl are not executable as shipped; these kernels
// follow the docstring math; the deviations from the literal code are listed in include/semb.h (Stokes section).
//
// One CTA per batch of EB x-consecutive elements of an element row; the element tiles live in shared memory, so every
// input array is read from HBM once and every output written once:
//   gradT : 1 input (+B) + 4 metric arrays in, 2 out      (56 or 64 B/DOF)
//   diver : 2 inputs + 4 metric arrays + B in, 1 out      (64 B/DOF)
// Arithmetic order: products with the metric terms / B are formed and rounded first (the reference's broadcasts),
// the contractions then accumulate with FMAs (the reference's BLAS calls: order unpinned).
#include "semb_vec.cuh"

namespace {

struct StokesLocalArgs {
  const double *a = nullptr, *b = nullptr;  // gradT: a = u ; diver: a = ux, b = uy
  const double* W = nullptr;                // gradT: optional pointwise weight applied to u first (mass: B)
  const double *rx, *ry, *sx, *sy, *B;
  const double *Dr, *Ds;                    // row-major D[i*n+k] = D(i,k)
  double *o1 = nullptr, *o2 = nullptr;
  long long pitch;
  int nr, ns, Ex, ney, EB;
};

// shared layout: Dr [nr*nr], Ds [ns*ns], then 4 tiles [EB][ns][nr|1]
__device__ __forceinline__ int tile_idx(int e, int j, int i, int ns, int S) { return (e * ns + j) * S + i; }

// out_x = Dr'_r (rx .* w) + Ds'_s (sx .* w),  out_y = Dr'_r (ry .* w) + Ds'_s (sy .* w),  w = W .* u  (or u)
__global__ void __launch_bounds__(256) semb_gradT_kernel(const StokesLocalArgs a) {
  extern __shared__ double sh[];
  const int nr = a.nr, ns = a.ns, S = nr | 1, EB = a.EB, tile = EB * ns * S;
  double* sDr = sh;
  double* sDs = sDr + nr * nr;
  double* t0 = sDs + ns * ns;  // rx .* w
  double* t1 = t0 + tile;      // sx .* w
  double* t2 = t1 + tile;      // ry .* w
  double* t3 = t2 + tile;      // sy .* w
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = tid; q < nr * nr; q += nt) sDr[q] = a.Dr[q];
  for (int q = tid; q < ns * ns; q += nt) sDs[q] = a.Ds[q];
  const int nbx = (a.Ex + EB - 1) / EB;
  for (int bidx = blockIdx.x; bidx < nbx * a.ney; bidx += gridDim.x) {
    const int r = bidx / nbx, e0 = (bidx - r * nbx) * EB, nbe = min(EB, a.Ex - e0);
    const int rowlen = nbe * nr;
    __syncthreads();
    for (int q = tid; q < ns * rowlen; q += nt) {
      const int j = q / rowlen, xx = q - j * rowlen, e = xx / nr, i = xx - e * nr;
      const size_t g = (size_t)(r * ns + j) * a.pitch + (size_t)e0 * nr + xx;
      double w = a.a[g];
      if (a.W) w = __dmul_rn(a.W[g], w);  // mass(Jp), mass.jl:17
      const int o = tile_idx(e, j, i, ns, S);
      t0[o] = __dmul_rn(a.rx[g], w);
      t1[o] = __dmul_rn(a.sx[g], w);
      t2[o] = __dmul_rn(a.ry[g], w);
      t3[o] = __dmul_rn(a.sy[g], w);
    }
    __syncthreads();
    for (int q = tid; q < ns * rowlen; q += nt) {
      const int j = q / rowlen, xx = q - j * rowlen, e = xx / nr, i = xx - e * nr;
      double ax = 0.0, bx = 0.0, ay = 0.0, by = 0.0;
      for (int k = 0; k < nr; ++k) {  // (Dr' along r): sum_k Dr(k,i) v(k,j)
        const double d = sDr[k * nr + i];
        const int o = tile_idx(e, j, k, ns, S);
        ax = fma(d, t0[o], ax);
        ay = fma(d, t2[o], ay);
      }
      for (int k = 0; k < ns; ++k) {  // (Ds' along s): sum_k Ds(k,j) v(i,k)
        const double d = sDs[k * ns + j];
        const int o = tile_idx(e, k, i, ns, S);
        bx = fma(d, t1[o], bx);
        by = fma(d, t3[o], by);
      }
      const size_t g = (size_t)(r * ns + j) * a.pitch + (size_t)e0 * nr + xx;
      a.o1[g] = __dadd_rn(ax, bx);
      a.o2[g] = __dadd_rn(ay, by);
    }
  }
}

// out = B .* (dx(ux) + dy(uy)),  dx(u) = rx .* ur + sx .* us,  dy(u) = ry .* ur + sy .* us  (grad.jl:27-31, diver.jl:22-27)
__global__ void __launch_bounds__(256) semb_diver_local_kernel(const StokesLocalArgs a) {
  extern __shared__ double sh[];
  const int nr = a.nr, ns = a.ns, S = nr | 1, EB = a.EB, tile = EB * ns * S;
  double* sDr = sh;
  double* sDs = sDr + nr * nr;
  double* t0 = sDs + ns * ns;  // ux
  double* t1 = t0 + tile;      // uy
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = tid; q < nr * nr; q += nt) sDr[q] = a.Dr[q];
  for (int q = tid; q < ns * ns; q += nt) sDs[q] = a.Ds[q];
  const int nbx = (a.Ex + EB - 1) / EB;
  for (int bidx = blockIdx.x; bidx < nbx * a.ney; bidx += gridDim.x) {
    const int r = bidx / nbx, e0 = (bidx - r * nbx) * EB, nbe = min(EB, a.Ex - e0);
    const int rowlen = nbe * nr;
    __syncthreads();
    for (int q = tid; q < ns * rowlen; q += nt) {
      const int j = q / rowlen, xx = q - j * rowlen, e = xx / nr, i = xx - e * nr;
      const size_t g = (size_t)(r * ns + j) * a.pitch + (size_t)e0 * nr + xx;
      const int o = tile_idx(e, j, i, ns, S);
      t0[o] = a.a[g];
      t1[o] = a.b[g];
    }
    __syncthreads();
    for (int q = tid; q < ns * rowlen; q += nt) {
      const int j = q / rowlen, xx = q - j * rowlen, e = xx / nr, i = xx - e * nr;
      double urx = 0.0, ury = 0.0, usx = 0.0, usy = 0.0;
      for (int k = 0; k < nr; ++k) {
        const double d = sDr[i * nr + k];
        const int o = tile_idx(e, j, k, ns, S);
        urx = fma(d, t0[o], urx);
        ury = fma(d, t1[o], ury);
      }
      for (int k = 0; k < ns; ++k) {
        const double d = sDs[j * ns + k];
        const int o = tile_idx(e, k, i, ns, S);
        usx = fma(d, t0[o], usx);
        usy = fma(d, t1[o], usy);
      }
      const size_t g = (size_t)(r * ns + j) * a.pitch + (size_t)e0 * nr + xx;
      const double uxdx = __dadd_rn(__dmul_rn(a.rx[g], urx), __dmul_rn(a.sx[g], usx));
      const double uydy = __dadd_rn(__dmul_rn(a.ry[g], ury), __dmul_rn(a.sy[g], usy));
      a.o1[g] = __dmul_rn(a.B[g], __dadd_rn(uxdx, uydy));  // mass(div), diver.jl:25-27
    }
  }
}

// middle of approxHlmzInv (diver.jl:97-99): out = (M .* g) .* Bi ./ b0, M from the Dirichlet flags of this slab
__global__ void semb_hinv_mid_kernel(const double* __restrict__ g, const double* __restrict__ Bi, double b0,
                                     long long pitch, int nxl, int nyl, int mx0, int mx1, int my0, int my1, double* out) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nxl; x += gridDim.x * blockDim.x) {
      const size_t idx = (size_t)row * pitch + x;
      const bool z = (x == 0 && mx0) || (x == nxl - 1 && mx1) || (row == 0 && my0) || (row == nyl - 1 && my1);
      const double v = __dmul_rn(z ? 0.0 : 1.0, g[idx]);
      out[idx] = __ddiv_rn(__dmul_rn(v, Bi[idx]), b0);
    }
}

// One-pass gatherScatter (gatherScatter.jl:8-21) of a single-rank field with a pointwise epilogue: every node adds its
// x partner, then the (x-summed) y partner -- the reference's (a+b)+(c+d) association, bitwise equal to the two-pass
// gs_x + seam_y form -- and applies  mode 0: nothing | 1: (M.*g).*Bi./b0 (diver.jl:96-98) | 2: M.*g (mask.jl:14).
struct GsFusedArgs {
  const double* u;
  double* out;
  const double* Bi;
  double b0;
  long long pitch;
  int nr, ns, Ex, ney, nxl, nyl, perx, pery, mode, mx0, mx1, my0, my1;
  int raw_lo, raw_hi;  // multi-rank: row 0 / nyl-1 still wait for the neighbour's row: no epilogue yet
};
// A thread owns a pair of columns (16-byte loads / stores) and walks down the rows of its CTA row; everything that only
// depends on the column -- element index, x partners, Dirichlet columns -- is computed once per thread.  (The first
// version, one node per thread with the index arithmetic per node, ran at 27 % of the HBM roof and was 47 % of the
// Stokes Schur apply: profiles/r02_stokes_launches_r3j.txt.)
__global__ void __launch_bounds__(256) semb_gs_fused_kernel(const GsFusedArgs a) {
  const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x >= a.nxl) return;
  const bool vb = x + 1 < a.nxl;  // (nxl may be odd: the pad column keeps its zero)
  // x partners of the two nodes: -1 none, else the column (the pair's other node, or a neighbouring pair's)
  auto partner = [&](int xx) {
    const int e = xx / a.nr, i = xx - e * a.nr;
    if (i == a.nr - 1) return e < a.Ex - 1 ? xx + 1 : (a.perx ? 0 : -1);
    if (i == 0) return e > 0 ? xx - 1 : (a.perx ? a.nxl - 1 : -1);
    return -1;
  };
  const int pa = partner(x), pb = vb ? partner(x + 1) : -1;
  const bool za = (x == 0 && a.mx0) || (x == a.nxl - 1 && a.mx1), zb = vb && (x + 1 == a.nxl - 1 && a.mx1);
  const size_t pitch = (size_t)a.pitch;
  const double* __restrict__ u = a.u;      // (out never aliases u or Bi: lets the loads of the next row pass this row's store)
  const double* __restrict__ Bi = a.Bi;
  double* __restrict__ out = a.out;
  auto row_sum = [&](const double* __restrict__ r, double& ga, double& gb) {   // x pairs of one row
    const double2 v = *reinterpret_cast<const double2*>(r + x);
    ga = v.x, gb = v.y;
    if (pa >= 0) ga = __dadd_rn(ga, pa == x + 1 ? v.y : r[pa]);
    if (pb >= 0) gb = __dadd_rn(gb, pb == x ? v.x : r[pb]);
  };
#pragma unroll 2
  for (int row = blockIdx.y; row < a.nyl; row += gridDim.y) {
    const int rl = row / a.ns, j = row - rl * a.ns;
    int yp = -1;
    if (j == a.ns - 1) yp = rl < a.ney - 1 ? row + 1 : (a.pery ? 0 : -1);
    else if (j == 0) yp = rl > 0 ? row - 1 : (a.pery ? a.nyl - 1 : -1);
    double ga, gb;
    row_sum(u + (size_t)row * pitch, ga, gb);
    if (yp >= 0) {   // the y pair of the two x pairs: (a+b)+(c+d), gatherScatter.jl:13
      double ha, hb;
      row_sum(u + (size_t)yp * pitch, ha, hb);
      ga = __dadd_rn(ga, ha), gb = __dadd_rn(gb, hb);
    }
    if (a.mode && !(row == 0 && a.raw_lo) && !(row == a.nyl - 1 && a.raw_hi)) {
      const bool zr = (row == 0 && a.my0) || (row == a.nyl - 1 && a.my1);
      ga = __dmul_rn((za || zr) ? 0.0 : 1.0, ga);
      gb = __dmul_rn((zb || zr) ? 0.0 : 1.0, gb);
      if (a.mode == 1) {
        const double2 bi = *reinterpret_cast<const double2*>(Bi + (size_t)row * pitch + x);
        ga = __dmul_rn(ga, bi.x), gb = __dmul_rn(gb, bi.y);
        if (a.b0 != 1.0) ga = __ddiv_rn(ga, a.b0), gb = __ddiv_rn(gb, a.b0);   // (x / 1.0 == x: same bits, no division)
      }
    }
    *reinterpret_cast<double2*>(out + (size_t)row * pitch + x) = make_double2(ga, vb ? gb : 0.0);
  }
}

// the epilogue of semb_gs_fused_kernel on the slab's first / last row, once the neighbour's row has been added
__global__ void semb_gs_rows_epilogue_kernel(const GsFusedArgs a) {
  for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < 2 * a.nxl; id += gridDim.x * blockDim.x) {
    const int which = id / a.nxl, x = id - which * a.nxl;
    if (which == 0 ? !a.raw_lo : !a.raw_hi) continue;
    const int row = which == 0 ? 0 : a.nyl - 1;
    const size_t idx = (size_t)row * a.pitch + x;
    const bool z = (x == 0 && a.mx0) || (x == a.nxl - 1 && a.mx1) || (row == 0 && a.my0) || (row == a.nyl - 1 && a.my1);
    double g = __dmul_rn(z ? 0.0 : 1.0, a.out[idx]);
    if (a.mode == 1) g = __ddiv_rn(__dmul_rn(g, a.Bi[idx]), a.b0);
    a.out[idx] = g;
  }
}

int local_launch_cfg(semb_ctx* ctx, semb_mesh* m, int ntiles, StokesLocalArgs* a, size_t* smem, int* grid) {
  const int S = m->nr | 1;
  int EB = 256 / (m->nr * m->ns);  // about one node per thread
  if (EB < 1) EB = 1;
  if (EB > m->Ex) EB = m->Ex;
  auto bytes = [&](int eb) { return (size_t)(m->nr * m->nr + m->ns * m->ns + ntiles * eb * m->ns * S) * 8; };
  while (EB > 1 && bytes(EB) > 48 * 1024) --EB;
  SEMB_REQUIRE(bytes(EB) <= 48 * 1024, "stokes: element tile of %d x %d nodes does not fit shared memory", m->nr, m->ns);
  a->EB = EB;
  a->nr = m->nr;
  a->ns = m->ns;
  a->Ex = m->Ex;
  a->ney = m->ney;
  a->pitch = m->pitch;
  a->rx = m->arr[SEMB_RX];
  a->ry = m->arr[SEMB_RY];
  a->sx = m->arr[SEMB_SX];
  a->sy = m->arr[SEMB_SY];
  a->B = m->arr[SEMB_B];
  a->Dr = m->dDr;
  a->Ds = m->dDs;
  *smem = bytes(EB);
  const long long nb = (long long)((m->Ex + EB - 1) / EB) * m->ney;
  long long g = (long long)ctx->sm_count * 8;
  *grid = (int)(nb < g ? nb : g);
  if (*grid < 1) *grid = 1;
  return SEMB_OK;
}

}  // namespace

int semb_launch_gradT(semb_ctx* ctx, semb_mesh* m, const double* u, const double* W, double* ox, double* oy) {
  StokesLocalArgs a;
  size_t smem;
  int grid;
  SEMB_TRY(local_launch_cfg(ctx, m, 4, &a, &smem, &grid));
  a.a = u;
  a.W = W;
  a.o1 = ox;
  a.o2 = oy;
  semb_gradT_kernel<<<grid, 256, smem, ctx->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

int semb_launch_diver_local(semb_ctx* ctx, semb_mesh* m, const double* ux, const double* uy, double* out) {
  StokesLocalArgs a;
  size_t smem;
  int grid;
  SEMB_TRY(local_launch_cfg(ctx, m, 2, &a, &smem, &grid));
  a.a = ux;
  a.b = uy;
  a.o1 = out;
  semb_diver_local_kernel<<<grid, 256, smem, ctx->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

int semb_launch_hinv_mid(semb_ctx* ctx, semb_mesh* m, const double* g, double b0, int mx0, int mx1, int my0, int my1,
                         double* out) {
  int gx = (m->nxl + 255) / 256;
  if (gx > 1024) gx = 1024;
  const int gy = m->nyl > 32768 ? 32768 : m->nyl;
  semb_hinv_mid_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(g, m->arr[SEMB_BI], b0, m->pitch, m->nxl, m->nyl, mx0, mx1,
                                                             my0, my1, out);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

// fused gatherScatter + epilogue (mode: 0 none, 1 (M.*g).*Bi./b0, 2 M.*g); out must not alias u.
// stage 0: the one-pass kernel over the slab (rows with a neighbour rank are left as raw local sums);
// stage 1: the epilogue on those rows, to be launched after the halo rows have been added (multi-rank only).
int semb_launch_gs_fused(semb_ctx* ctx, semb_mesh* m, const double* u, double* out, int mode, double b0, int mx0, int mx1,
                         int my0, int my1, int stage) {
  GsFusedArgs a;
  a.u = u;
  a.out = out;
  a.Bi = m->arr[SEMB_BI];
  a.b0 = b0;
  a.pitch = m->pitch;
  a.nr = m->nr;
  a.ns = m->ns;
  a.Ex = m->Ex;
  a.ney = m->ney;
  a.nxl = m->nxl;
  a.nyl = m->nyl;
  a.perx = m->perx;
  a.pery = (m->pery && ctx->nranks == 1) ? 1 : 0;  // with several ranks the periodic wrap is a halo
  a.mode = mode;
  a.mx0 = mx0;
  a.mx1 = mx1;
  a.my0 = my0;
  a.my1 = my1;
  a.raw_lo = m->halo_lo;
  a.raw_hi = m->halo_hi;
  if (stage == 0) {
    const int gx = ((m->nxl + 1) / 2 + 255) / 256;
    int gy = (ctx->sm_count * 8 + gx - 1) / gx;   // eight resident CTAs per SM, each walking down its share of the rows
    if (gy > m->nyl) gy = m->nyl;
    semb_gs_fused_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(a);
  } else {
    if (mode == 0 || (!a.raw_lo && !a.raw_hi)) return SEMB_OK;
    int blocks = (2 * m->nxl + 255) / 256;
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    semb_gs_rows_epilogue_kernel<<<blocks, 256, 0, ctx->stream>>>(a);
  }
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}
