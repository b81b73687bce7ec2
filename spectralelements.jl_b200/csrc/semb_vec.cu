// Streaming kernels of libsemb: seam completion of the fused operator, stand-alone gather-scatter and
// mask, PCG vector updates with fused deterministic reductions, geometry set-up, generic fallbacks.
// Reference citations are to /root/reference/src.
#include "semb_reduce.cuh"
#include "semb_vec.cuh"

#define SEMB_PI 3.14159265358979323846

namespace {

__device__ __forceinline__ double mask_at(const OpArgs& a, int x, int y, size_t idx) {
  if (a.M_arr) return a.M_arr[idx];
  return ((x == 0 && a.mx0) || (x == a.nxl - 1 && a.mx1) || (y == 0 && a.my0) || (y == a.nyl - 1 && a.my1)) ? 0.0
                                                                                                             : 1.0;
}

// ------------------------------------------------------------------------------------------------
// x seams: column pairs (xa, xb) at strip boundaries (+ the periodic wrap).  Forms the x pair for every
// line, and, for lines that are an in-chunk y interface, the y pair of the two x pairs:
// (a+b)+(c+d), the association of gatherScatter.jl:13 (QQtx before QQty, ABu.jl:14-33).
// Lines on a y seam are left x-summed but unmasked for the y-seam kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) semb_seam_x_kernel(const OpArgs a) {
  __shared__ double red[32];
  if (a.pcg && a.scal->done) return;
  const int N = a.N;
  double acc = 0.0;
  const int yb = a.y_begin, ny = (a.y_end > a.y_begin ? a.y_end : a.nyl) - yb;
  const long long total = (long long)a.nxseam * ny;
  auto finish = [&](int x, int y, double val) {
    const size_t idx = (size_t)y * a.pitch + x;
    const double o = __dmul_rn(mask_at(a, x, y, idx), val);
    a.out[idx] = o;
    if (a.pcg) acc += __dmul_rn(__dmul_rn(a.pout[idx], o), a.mult[idx]);
  };
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total;
       id += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(id / ny), y = yb + (int)(id - (long long)s * ny);
    const int xa = a.xseam[2 * s], xb = a.xseam[2 * s + 1];
    const int r = y / N, j = y - r * N;
    if (j == 0 && r > 0 && !a.ystart[r]) continue;  // lower line of an in-chunk pair: done by line y-1
    const size_t ia = (size_t)y * a.pitch + xa, ib = (size_t)y * a.pitch + xb;
    const double s0 = __dadd_rn(a.out[ia], a.out[ib]);
    if (j == N - 1 && r + 1 < a.ney && !a.ystart[r + 1]) {
      const double s1 = __dadd_rn(a.out[ia + a.pitch], a.out[ib + a.pitch]);
      const double tot = __dadd_rn(s0, s1);
      finish(xa, y, tot);
      finish(xb, y, tot);
      finish(xa, y + 1, tot);
      finish(xb, y + 1, tot);
    } else if ((j == 0 && a.ystart[r]) || (j == N - 1 && a.ystart[r + 1])) {
      a.out[ia] = s0;
      a.out[ib] = s0;
    } else {
      finish(xa, y, s0);
      finish(xb, y, s0);
    }
  }
  if (a.pcg) {
    const double bs = semb_block_sum(acc, red, threadIdx.x, blockDim.x);
    double tot;
    if (semb_last_block(bs, 0.0, a.partials, nullptr, a.counters, gridDim.x, blockIdx.x, red, threadIdx.x,
                        blockDim.x, &tot, nullptr))
      a.scal->pap[1] = tot;
  }
}

// ------------------------------------------------------------------------------------------------
// y seams: row pairs (ya, yb) (chunk boundaries, periodic wrap), then the halo rows received from the
// neighbouring ranks.  Rows are contiguous => coalesced.  Applies the mask (mask.jl:14) last.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) semb_seam_y_kernel(const OpArgs a, int nhalo_lo, int nhalo_hi, int domask,
                                                          const P2PArgs x) {
  __shared__ double red[32];
  __shared__ double sh_tot[2];
  if (a.pcg && a.scal->done) return;
  if (x.on && (nhalo_lo || nhalo_hi)) {
    // P2P halo: the neighbours' boundary rows arrive in the local mailbox over NVLink (semb_halo_push_kernel
    // on the neighbour); wait for this apply's epoch before touching them
    if (threadIdx.x == 0) {
      if (nhalo_lo) semb_wait_epoch(&a.scal->flag_halo[0], x.epoch, a.scal);
      if (nhalo_hi) semb_wait_epoch(&a.scal->flag_halo[1], x.epoch, a.scal);
    }
    __syncthreads();
  }
  double acc = 0.0;
  const int nq = a.nyseam + nhalo_lo + nhalo_hi;
  const long long total = (long long)nq * a.nxl;
  auto finish = [&](int x_, int y, double val) {
    const size_t idx = (size_t)y * a.pitch + x_;
    const double o = domask ? __dmul_rn(mask_at(a, x_, y, idx), val) : val;
    a.out[idx] = o;
    if (a.pcg) acc += __dmul_rn(__dmul_rn(a.pout[idx], o), a.mult[idx]);
  };
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total;
       id += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(id / a.nxl), xx = (int)(id - (long long)q * a.nxl);
    if (q < a.nyseam) {
      const int ya = a.yseam[2 * q], yb = a.yseam[2 * q + 1];
      const double s = __dadd_rn(a.out[(size_t)ya * a.pitch + xx], a.out[(size_t)yb * a.pitch + xx]);
      finish(xx, ya, s);
      finish(xx, yb, s);
    } else {
      const bool lo = (q == a.nyseam) && nhalo_lo;
      const int y = lo ? 0 : a.nyl - 1;
      const double h = lo ? a.halo_lo[xx] : a.halo_hi[xx];
      finish(xx, y, __dadd_rn(a.out[(size_t)y * a.pitch + xx], h));
    }
  }
  if (a.pcg) {
    const double bs = semb_block_sum(acc, red, threadIdx.x, blockDim.x);
    if (semb_last_block_uniform(bs, 0.0, a.partials, nullptr, a.counters, gridDim.x, blockIdx.x, red, threadIdx.x,
                                blockDim.x, sh_tot)) {
      if (threadIdx.x == 0) a.scal->pap[2] = sh_tot[0];
      if (x.on) {
        // sum(p .* Ap .* mult) over all ranks, fused here: this kernel is the last contributor
        const double mine = __dadd_rn(__dadd_rn(a.scal->pap[0], a.scal->pap[1]), sh_tot[0]);
        double tot;
        semb_p2p_allgather(a.scal, 0, mine, 0.0, &tot, nullptr, threadIdx.x);
        if (threadIdx.x == 0) a.scal->pap_total = tot;
      }
    }
  }
}

// P2P halo push: copy this slab's first / last row (x-summed by now) into the neighbours' mailboxes with
// plain stores to mapped peer memory (NVLink), then release the epoch flag.  Replaces ncclSend/ncclRecv.
__global__ void __launch_bounds__(256) semb_halo_push_kernel(const double* __restrict__ out, long long pitch, int nxl,
                                                             int nyl, double* dst_lo, double* dst_hi,
                                                             unsigned long long* flag_lo, unsigned long long* flag_hi,
                                                             unsigned long long epoch, unsigned* counter,
                                                             const SembScal* scal, int pcg) {
  __shared__ int s_last;
  if (pcg && scal->done) return;
  const double* first = out;
  const double* last = out + (size_t)(nyl - 1) * pitch;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nxl; i += gridDim.x * blockDim.x) {
    if (dst_hi) dst_hi[i] = last[i];   // my last row is the upper neighbour's "row from below"
    if (dst_lo) dst_lo[i] = first[i];  // my first row is the lower neighbour's "row from above"
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(counter, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence_system();
    if (flag_hi) semb_st_release_sys(flag_hi, epoch);
    if (flag_lo) semb_st_release_sys(flag_lo, epoch);
    *counter = 0u;
  }
}

// ---- 2-D streaming iteration over a pitched field ------------------------------------------------
struct Grid2D {
  dim3 grid, block;
};
Grid2D grid2d(long long pitch, int nyl, int sm_count, int max_blocks = 2048) {
  const int p2 = (int)(pitch / 2);
  int bx = 32;
  while (bx < 256 && bx < p2) bx <<= 1;
  int by = 256 / bx;
  int gx = (p2 + bx - 1) / bx;
  if (gx > 64) gx = 64;
  int rows = (nyl + by - 1) / by;
  int gy = (sm_count * 8 + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  while ((long long)gx * gy > max_blocks) gy--;
  Grid2D g;
  g.grid = dim3(gx, gy);
  g.block = dim3(bx, by);
  return g;
}

#define SEMB_FOR_2D(p2, nyl)                                                                     \
  for (int row = blockIdx.y * blockDim.y + threadIdx.y; row < (nyl); row += gridDim.y * blockDim.y) \
    for (int c2 = blockIdx.x * blockDim.x + threadIdx.x; c2 < (p2); c2 += gridDim.x * blockDim.x)

#define SEMB_TID (threadIdx.y * blockDim.x + threadIdx.x)
#define SEMB_NT (blockDim.x * blockDim.y)
#define SEMB_BID (blockIdx.y * gridDim.x + blockIdx.x)
#define SEMB_NB (gridDim.x * gridDim.y)

__global__ void semb_mask_kernel(const double2* __restrict__ u, const double2* __restrict__ M, double2* out,
                                 size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = u[i];
    if (M) {
      const double2 m = M[i];
      v.x = __dmul_rn(m.x, v.x);  // mask.jl:14
      v.y = __dmul_rn(m.y, v.y);
    }
    out[i] = v;  // mask.jl:13 (copy) when M is empty
  }
}

__global__ void semb_axpby_kernel(double a, const double2* __restrict__ x, double b, double2* y, size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const double2 xv = x[i];
    double2 yv = y[i];
    yv.x = __dadd_rn(__dmul_rn(a, xv.x), __dmul_rn(b, yv.x));
    yv.y = __dadd_rn(__dmul_rn(a, xv.y), __dmul_rn(b, yv.y));
    y[i] = yv;
  }
}

__global__ void semb_fill_kernel(double* x, double v, long long pitch, int nxl, int nyl) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < pitch; c += gridDim.x * blockDim.x)
      x[(size_t)row * pitch + c] = (c < nxl) ? v : 0.0;
}

// splitmix64 uniform(-1,1), indexed by the GLOBAL column-major linear index
__global__ void semb_fill_random_kernel(double* x, long long pitch, int nxl, int nyl, long long gnxl,
                                        long long gy0, uint64_t seed) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxl; c += gridDim.x * blockDim.x) {
      const uint64_t lin = (uint64_t)((gy0 + row) * gnxl + c) + 1ull;
      uint64_t z = seed + lin * 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z = z ^ (z >> 31);
      const double u01 = (double)(z >> 11) * (1.0 / 9007199254740992.0);
      x[(size_t)row * pitch + c] = __dadd_rn(__dmul_rn(2.0, u01), -1.0);
    }
}

// x part of QQ^T (QQtx, mesh.jl:81): each interface node gets the sum of its x duplicate
__global__ void semb_gs_x_kernel(const double* __restrict__ u, double* __restrict__ out, long long pitch, int N,
                                 int Ex, int nxl, int nyl, int perx) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nxl; x += gridDim.x * blockDim.x) {
      const int e = x / N, i = x - e * N;
      int xp = -1;
      if (i == N - 1 && e < Ex - 1) xp = x + 1;
      else if (i == 0 && e > 0) xp = x - 1;
      else if (perx && x == nxl - 1) xp = 0;
      else if (perx && x == 0) xp = nxl - 1;
      const size_t b = (size_t)row * pitch;
      double v = u[b + x];
      if (xp >= 0) v = __dadd_rn(v, u[b + xp]);
      out[b + x] = v;
    }
}

// mult = 1 ./ gatherScatter(ones), mesh.jl:94-96 -- exact values 1, 1/2, 1/4
__global__ void semb_mult_kernel(double* mult, long long pitch, int nr, int ns, int Ex, int Ey, int ey0, int ney,
                                 int perx, int pery) {
  const int nxl = nr * Ex, nyl = ns * ney;
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nxl; x += gridDim.x * blockDim.x) {
      const int e = x / nr, i = x - e * nr;
      const int rl = row / ns, j = row - rl * ns, rg = ey0 + rl;
      const bool dx = (i == nr - 1 && e < Ex - 1) || (i == 0 && e > 0) ||
                      (perx && (x == 0 || x == nxl - 1));
      const bool dy = (j == ns - 1 && rg < Ey - 1) || (j == 0 && rg > 0) ||
                      (pery && ((rg == 0 && j == 0) || (rg == Ey - 1 && j == ns - 1)));
      const double cx = dx ? 2.0 : 1.0, cy = dy ? 2.0 : 1.0;
      mult[(size_t)row * pitch + x] = 1.0 / (cx * cy);
    }
}

// generateMask, mesh.jl:149-175 (flags already resolved for periodicity / slab position)
__global__ void semb_mask_gen_kernel(double* M, long long pitch, int nxl, int nyl, int mx0, int mx1, int my0,
                                     int my1) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nxl; x += gridDim.x * blockDim.x) {
      const bool z = (x == 0 && mx0) || (x == nxl - 1 && mx1) || (row == 0 && my0) || (row == nyl - 1 && my1);
      M[(size_t)row * pitch + x] = z ? 0.0 : 1.0;
    }
}

// semmesh.jl:9-27 + ndgrid.jl:8-13 + a built-in deform (mesh.jl:98-108), generated on device.
// z = dz*z0 + ze[e], ze = linspace(-1,1,E+1) (same op order as the NumPy oracle).
__device__ __forceinline__ double semb_coord(int g, int n, int E, const double* z0) {
  const int e = g / n, i = g - e * n;
  const double step = 2.0 / (double)E;
  const double ze0 = (e == E) ? 1.0 : __dadd_rn(__dmul_rn((double)e, step), -1.0);
  const double ze1 = (e + 1 == E) ? 1.0 : __dadd_rn(__dmul_rn((double)(e + 1), step), -1.0);
  const double dz = __dadd_rn(ze1, -ze0);
  return __dadd_rn(__dmul_rn(dz, z0[i]), ze0);
}

__global__ void semb_grid_kernel(double* x, double* y, long long pitch, int nr, int ns, int Ex, int Ey, int ey0,
                                 int ney, const double* __restrict__ z0r, const double* __restrict__ z0s,
                                 int kind, double p0, double p1, double p2) {
  const int nxl = nr * Ex, nyl = ns * ney;
  for (int row = blockIdx.y; row < nyl; row += gridDim.y) {
    const double s = semb_coord(ey0 * ns + row, ns, Ey, z0s);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxl; c += gridDim.x * blockDim.x) {
      const double r = semb_coord(c, nr, Ex, z0r);
      double xo = r, yo = s;
      if (kind == SEMB_DEFORM_ANNULUS) {  // geom.jl:40-49
        const double R = __dadd_rn(__dmul_rn((p1 - p0) / 2.0, __dadd_rn(r, 1.0)), p0);
        const double th = __dadd_rn(__dmul_rn(p2 / 2.0, __dadd_rn(s, 1.0)), 0.0);
        xo = __dmul_rn(R, cos(th));
        yo = __dmul_rn(R, sin(th));
      } else if (kind == SEMB_DEFORM_WAVY) {
        const double d = __dmul_rn(__dmul_rn(p0, sin(__dmul_rn(SEMB_PI, r))), sin(__dmul_rn(SEMB_PI, s)));
        xo = __dadd_rn(r, d);
        yo = __dadd_rn(s, d);
      }
      x[(size_t)row * pitch + c] = xo;
      y[(size_t)row * pitch + c] = yo;
    }
  }
}

// jac.jl:24-40 + mesh.jl:114-123.  One thread per node; D row-major in global (L1-resident).
// Pointwise products are kept un-fused (the reference's broadcasts do not contract).
__global__ void semb_geom_kernel(const double* __restrict__ x, const double* __restrict__ y, long long pitch,
                                 int nr, int ns, int Ex, int ney, const double* __restrict__ Dr,
                                 const double* __restrict__ Ds, const double* __restrict__ wr,
                                 const double* __restrict__ ws, double* J, double* Ji, double* rx, double* ry,
                                 double* sx, double* sy, double* B, double* Bi, double* G11, double* G12,
                                 double* G22) {
  const int nxl = nr * Ex, nyl = ns * ney;
  for (int row = blockIdx.y; row < nyl; row += gridDim.y) {
    const int rl = row / ns, j = row - rl * ns;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxl; c += gridDim.x * blockDim.x) {
      const int e = c / nr, i = c - e * nr;
      const size_t rowb = (size_t)row * pitch + (size_t)e * nr;       // start of the x-line
      const size_t colb = (size_t)rl * ns * pitch + c;                 // start of the y-line
      double xr = 0, yr = 0, xs = 0, ys = 0;
      for (int k = 0; k < nr; ++k) {
        const double d = Dr[i * nr + k];
        xr = fma(d, x[rowb + k], xr);
        yr = fma(d, y[rowb + k], yr);
      }
      for (int k = 0; k < ns; ++k) {
        const double d = Ds[j * ns + k];
        xs = fma(d, x[colb + (size_t)k * pitch], xs);
        ys = fma(d, y[colb + (size_t)k * pitch], ys);
      }
      const size_t idx = (size_t)row * pitch + c;
      const double Jv = __dadd_rn(__dmul_rn(xr, ys), -__dmul_rn(xs, yr));  // jac.jl:31
      const double Jiv = 1.0 / Jv;
      const double rxv = __dmul_rn(Jiv, ys), ryv = __dmul_rn(-Jiv, xs);
      const double sxv = __dmul_rn(-Jiv, yr), syv = __dmul_rn(Jiv, xr);
      const double Bv = __dmul_rn(Jv, __dmul_rn(wr[i], ws[j]));  // mesh.jl:117
      if (J) J[idx] = Jv;
      if (Ji) Ji[idx] = Jiv;
      if (rx) rx[idx] = rxv;
      if (ry) ry[idx] = ryv;
      if (sx) sx[idx] = sxv;
      if (sy) sy[idx] = syv;
      if (B) B[idx] = Bv;
      if (Bi) Bi[idx] = 1.0 / Bv;
      if (G11) G11[idx] = __dmul_rn(Bv, __dadd_rn(__dmul_rn(rxv, rxv), __dmul_rn(ryv, ryv)));  // mesh.jl:121-123
      if (G12) G12[idx] = __dmul_rn(Bv, __dadd_rn(__dmul_rn(rxv, sxv), __dmul_rn(ryv, syv)));
      if (G22) G22[idx] = __dmul_rn(Bv, __dadd_rn(__dmul_rn(sxv, sxv), __dmul_rn(syv, syv)));
    }
  }
}

// ---- generic local operator (any nr, ns), two passes through wr/ws (lapl.jl:70-81) -----------------
__global__ void semb_generic_pass1(const OpArgs a, int nr, int ns, const double* __restrict__ Dr,
                                   const double* __restrict__ Ds, double* wrt, double* wst) {
  const int nxl = a.nxl, nyl = a.nyl;
  for (int row = blockIdx.y; row < nyl; row += gridDim.y) {
    const int rl = row / ns, j = row - rl * ns;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxl; c += gridDim.x * blockDim.x) {
      const int e = c / nr, i = c - e * nr;
      const size_t rowb = (size_t)row * a.pitch + (size_t)e * nr, colb = (size_t)rl * ns * a.pitch + c;
      double ur = 0, us = 0;
      for (int k = 0; k < nr; ++k) ur = fma(Dr[i * nr + k], a.u[rowb + k], ur);
      for (int k = 0; k < ns; ++k) us = fma(Ds[j * ns + k], a.u[colb + (size_t)k * a.pitch], us);
      const size_t idx = (size_t)row * a.pitch + c;
      wrt[idx] = fma(a.G11[idx], ur, a.G12[idx] * us);
      wst[idx] = fma(a.G12[idx], ur, a.G22[idx] * us);
    }
  }
}

__global__ void semb_generic_pass2(const OpArgs a, int nr, int ns, const double* __restrict__ Dr,
                                   const double* __restrict__ Ds, const double* __restrict__ wrt,
                                   const double* __restrict__ wst, int massterm) {
  const int nxl = a.nxl, nyl = a.nyl;
  for (int row = blockIdx.y; row < nyl; row += gridDim.y) {
    const int rl = row / ns, j = row - rl * ns;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxl; c += gridDim.x * blockDim.x) {
      const int e = c / nr, i = c - e * nr;
      const size_t rowb = (size_t)row * a.pitch + (size_t)e * nr, colb = (size_t)rl * ns * a.pitch + c;
      double ar = 0, as = 0;
      for (int k = 0; k < nr; ++k) ar = fma(Dr[k * nr + i], wrt[rowb + k], ar);
      for (int k = 0; k < ns; ++k) as = fma(Ds[k * ns + j], wst[colb + (size_t)k * a.pitch], as);
      const size_t idx = (size_t)row * a.pitch + c;
      double lap = __dadd_rn(ar, as);
      const double nu = a.nu_arr ? a.nu_arr[idx] : a.nu;
      lap = __dmul_rn(nu, lap);
      if (massterm) {
        const double kk = a.k_arr ? a.k_arr[idx] : a.k;
        lap = __dadd_rn(lap, __dmul_rn(kk, __dmul_rn(a.B[idx], a.u[idx])));
      }
      a.out[idx] = lap;
    }
  }
}

// mult = 1 ./ gatherScatter(ones) (mesh.jl:94-96) is separable: mult(x,y) = wx(x) * wy(y) with weights in
// {1, 1/2} (exact products), so the streaming kernels read two tiny 1-D tables instead of an n-sized array
__device__ __forceinline__ double2 semb_mult2(const double2* __restrict__ wx1d, const double* __restrict__ wy1d,
                                              int c2, int row) {
  const double2 w = wx1d[c2];
  const double wy = wy1d[row];
  return make_double2(w.x * wy, w.y * wy);
}

// ---- deterministic reductions ------------------------------------------------------------------------
// which = 0: sum(a .* b .* mult) (pcg.jl:45,52);  which = 1: norm(a, Inf) (pcg.jl:36);
// which = 2: max |a - ref| over the valid nodes (is an array coefficient really a constant?)
__global__ void __launch_bounds__(256) semb_reduce_kernel(int which, const double2* __restrict__ av,
                                                          const double2* __restrict__ bv,
                                                          const double2* __restrict__ wx1d,
                                                          const double* __restrict__ wy1d, int p2, int nyl,
                                                          double* partials, unsigned* counter, SembScal* scal,
                                                          const P2PArgs x, double ref, int nxl) {
  __shared__ double red[32];
  double acc = 0.0;
  SEMB_FOR_2D(p2, nyl) {
    const size_t i = (size_t)row * p2 + c2;
    const double2 a2 = av[i];
    if (which == 0) {
      const double2 b2 = bv[i];
      const double2 m2 = semb_mult2(wx1d, wy1d, c2, row);
      acc += __dmul_rn(__dmul_rn(a2.x, b2.x), m2.x);
      acc += __dmul_rn(__dmul_rn(a2.y, b2.y), m2.y);
    } else if (which == 1) {
      acc = fmax(acc, fmax(fabs(a2.x), fabs(a2.y)));
    } else {
      if (2 * c2 < nxl) acc = fmax(acc, fabs(a2.x - ref));
      if (2 * c2 + 1 < nxl) acc = fmax(acc, fabs(a2.y - ref));
    }
  }
  __shared__ double sh_tot[2];
  bool last;
  if (which == 0) {
    const double bs = semb_block_sum(acc, red, SEMB_TID, SEMB_NT);
    last = semb_last_block_uniform(bs, 0.0, partials, nullptr, counter, SEMB_NB, SEMB_BID, red, SEMB_TID, SEMB_NT,
                                   sh_tot);
  } else {
    const double bs = semb_block_max(acc, red, SEMB_TID, SEMB_NT);
    last = semb_last_block_uniform(0.0, bs, partials, partials + SEMB_NB, counter, SEMB_NB, SEMB_BID, red, SEMB_TID,
                                   SEMB_NT, sh_tot);
  }
  if (last) {
    const double v = (which == 0) ? sh_tot[0] : sh_tot[1];
    if (x.on) {
      double s0, m1;
      semb_p2p_allgather(scal, 2, which == 0 ? v : 0.0, which == 0 ? 0.0 : v, &s0, &m1, SEMB_TID);
      if (SEMB_TID == 0) scal->red[which] = (which == 0) ? s0 : m1;
    } else if (SEMB_TID == 0) {
      scal->red[which] = v;
      scal->xchg_red[scal->rank] = v;
    }
  }
}

// combine the per-rank values gathered into xchg_red (fixed rank order => identical on all ranks)
__global__ void semb_reduce_finalize_kernel(SembScal* scal, int which) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double v = 0.0;
  for (int r = 0; r < scal->nranks; ++r) v = (which == 0) ? v + scal->xchg_red[r] : fmax(v, scal->xchg_red[r]);
  scal->red[which] = v;
}

// ---- PCG ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double semb_prec(double r, double B, int precond, double b0) {
  return precond ? (r / B) / b0 : r;  // opPrecond: identity (diffusion.jl:47-49) or u ./ B ./ b0 (convectionDiffusion.jl:87-91)
}

// shared tail of init/update: publish t = sum(r.*h.*mult), rmax = norm(r,Inf); advance the state
__device__ __forceinline__ void semb_pcg_advance(SembScal* s, double tnew, double rmax, bool first) {
  if (first) {
    s->t = tnew;
    s->t_prev = 0.0;
    s->iters = 0;
    s->warned = 0;
  } else {
    s->t_prev = s->t;
    s->t = tnew;
    s->iters += 1;
  }
  s->rmax = rmax;
  int done = !(rmax > s->tol);                                   // pcg.jl:36
  if (!done && s->iters >= s->maxiter) { done = 1; s->warned = 1; }  // pcg.jl:39
  s->done = done;
}

// x = 0, r = b, p = 0 (pcg.jl:25-33); t0, rmax0
__global__ void __launch_bounds__(256) semb_pcg_init_kernel(const double2* __restrict__ b, double2* x, double2* r,
                                                            double2* p, double2* hout, const double2* __restrict__ Bm,
                                                            const double2* __restrict__ wx1d,
                                                            const double* __restrict__ wy1d, int p2, int nxl,
                                                            int nyl, int precond, double b0, double tol,
                                                            long long maxiter, double* partials,
                                                            unsigned* counter, SembScal* scal, const P2PArgs xp) {
  __shared__ double red[32];
  double acc = 0.0, amx = 0.0;
  SEMB_FOR_2D(p2, nyl) {
    const size_t i = (size_t)row * p2 + c2;
    const double2 bv = b[i];
    const double2 z = make_double2(0.0, 0.0);
    x[i] = z;
    p[i] = z;
    r[i] = bv;
    const double2 m2 = semb_mult2(wx1d, wy1d, c2, row);
    double hx = bv.x, hy = bv.y;
    if (precond == 1) {
      const double2 B2 = Bm[i];
      hx = (2 * c2 < nxl) ? semb_prec(bv.x, B2.x, 1, b0) : 0.0;
      hy = (2 * c2 + 1 < nxl) ? semb_prec(bv.y, B2.y, 1, b0) : 0.0;
      if (hout) hout[i] = make_double2(hx, hy);  // h = opM(r), kept for the strip kernel (staged instead of r)
    }
    acc += __dmul_rn(__dmul_rn(bv.x, hx), m2.x);
    acc += __dmul_rn(__dmul_rn(bv.y, hy), m2.y);
    amx = fmax(amx, fmax(fabs(bv.x), fabs(bv.y)));
  }
  __shared__ double sh_tot[2];
  const double bs = semb_block_sum(acc, red, SEMB_TID, SEMB_NT);
  const double bm = semb_block_max(amx, red, SEMB_TID, SEMB_NT);
  if (semb_last_block_uniform(bs, bm, partials, partials + SEMB_NB, counter, SEMB_NB, SEMB_BID, red, SEMB_TID,
                              SEMB_NT, sh_tot)) {
    double tot = sh_tot[0], tmx = sh_tot[1];
    if (SEMB_TID == 0) {
      scal->tol = tol;
      scal->maxiter = maxiter;
      scal->xchg_t[2 * scal->rank] = tot;
      scal->xchg_t[2 * scal->rank + 1] = tmx;
    }
    if (precond == 2) {  // opM is a kernel of its own (FDM): it forms t = sum(r.*h.*mult) and advances the state
      if (SEMB_TID == 0) scal->red[2] = tmx;
      return;
    }
    if (xp.on) semb_p2p_allgather(scal, 1, tot, tmx, &tot, &tmx, SEMB_TID);
    if (SEMB_TID == 0 && (scal->nranks == 1 || xp.on)) semb_pcg_advance(scal, tot, tmx, true);
  }
}

// x += a*p ; r -= a*Ap (pcg.jl:53-54) with a = t / sum(p.*Ap.*mult) (pcg.jl:52); new t and norm(r,Inf)
__global__ void __launch_bounds__(256) semb_pcg_update_kernel(double2* x, double2* r, const double2* __restrict__ p,
                                                              const double2* __restrict__ Ap, double2* hout,
                                                              const double2* __restrict__ Bm,
                                                              const double2* __restrict__ wx1d,
                                                              const double* __restrict__ wy1d, int p2, int nxl,
                                                              int nyl, int precond, double b0, double* partials,
                                                              unsigned* counter, SembScal* scal, const P2PArgs xp) {
  __shared__ double red[32];
  if (scal->done) return;
  double pap;
  if (scal->nranks == 1) {
    pap = __dadd_rn(__dadd_rn(scal->pap[0], scal->pap[1]), scal->pap[2]);
  } else {
    pap = scal->pap_total;  // combined over ranks by the y-seam kernel (P2P) or semb_pcg_combine_pap_kernel (NCCL)
  }
  const double alpha = scal->t / pap;
  double acc = 0.0, amx = 0.0;
  SEMB_FOR_2D(p2, nyl) {
    const size_t i = (size_t)row * p2 + c2;
    const double2 pv = p[i], av = Ap[i];
    // B travels with the other four loads (inside the `if (precond)` below it was issued only after they had
    // arrived: two exposed memory latencies per trip, long-scoreboard stalls 78 %, profiles/r01_update_prec_r1n.txt)
    const double2 B2 = precond == 1 ? Bm[i] : make_double2(1.0, 1.0);
    const double2 m2 = semb_mult2(wx1d, wy1d, c2, row);
    double2 xv = x[i], rv = r[i];
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
    rv.x = __dadd_rn(rv.x, -__dmul_rn(alpha, av.x));
    rv.y = __dadd_rn(rv.y, -__dmul_rn(alpha, av.y));
    x[i] = xv;
    r[i] = rv;
    double hx = rv.x, hy = rv.y;
    if (precond == 1) {
      hx = (2 * c2 < nxl) ? semb_prec(rv.x, B2.x, 1, b0) : 0.0;
      hy = (2 * c2 + 1 < nxl) ? semb_prec(rv.y, B2.y, 1, b0) : 0.0;
      if (hout) hout[i] = make_double2(hx, hy);
    }
    acc += __dmul_rn(__dmul_rn(rv.x, hx), m2.x);
    acc += __dmul_rn(__dmul_rn(rv.y, hy), m2.y);
    amx = fmax(amx, fmax(fabs(rv.x), fabs(rv.y)));
  }
  __shared__ double sh_tot[2];
  const double bs = semb_block_sum(acc, red, SEMB_TID, SEMB_NT);
  const double bm = semb_block_max(amx, red, SEMB_TID, SEMB_NT);
  if (semb_last_block_uniform(bs, bm, partials, partials + SEMB_NB, counter, SEMB_NB, SEMB_BID, red, SEMB_TID,
                              SEMB_NT, sh_tot)) {
    double tot = sh_tot[0], tmx = sh_tot[1];
    if (SEMB_TID == 0) {
      scal->xchg_t[2 * scal->rank] = tot;
      scal->xchg_t[2 * scal->rank + 1] = tmx;
    }
    if (precond == 2) {  // FDM: the preconditioner kernel that follows forms t and advances the state
      if (SEMB_TID == 0) scal->red[2] = tmx;
      return;
    }
    if (xp.on) semb_p2p_allgather(scal, 1, tot, tmx, &tot, &tmx, SEMB_TID);  // fused all-gather over NVLink
    if (SEMB_TID == 0 && (scal->nranks == 1 || xp.on)) semb_pcg_advance(scal, tot, tmx, false);
  }
}

// multi-GPU: after the all-gather of {t_local, rmax_local}, combine in rank order and advance
__global__ void semb_pcg_finalize_kernel(SembScal* scal, int first) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (!first && scal->done) return;
  double t = 0.0, mx = 0.0;
  for (int q = 0; q < scal->nranks; ++q) {
    t += scal->xchg_t[2 * q];
    mx = fmax(mx, scal->xchg_t[2 * q + 1]);
  }
  semb_pcg_advance(scal, t, mx, first != 0);
}

__global__ void semb_pcg_pack_pap_kernel(SembScal* scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  scal->xchg_pap[scal->rank] = __dadd_rn(__dadd_rn(scal->pap[0], scal->pap[1]), scal->pap[2]);
}

// NCCL path: after the all-gather of xchg_pap, combine in rank order
__global__ void semb_pcg_combine_pap_kernel(SembScal* scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double pap = 0.0;
  for (int q = 0; q < scal->nranks; ++q) pap += scal->xchg_pap[q];
  scal->pap_total = pap;
}

// generic path: p = h + beta*p (pcg.jl:46-50)
__global__ void semb_pcg_dir_kernel(const double2* __restrict__ r, double2* p, const double2* __restrict__ Bm,
                                    int p2, int nxl, int nyl, int precond, double b0, SembScal* scal) {
  if (scal->done) return;
  const double beta = semb_pcg_beta(scal);
  SEMB_FOR_2D(p2, nyl) {
    const size_t i = (size_t)row * p2 + c2;
    const double2 rv = r[i];
    double2 pv = p[i];
    double hx = rv.x, hy = rv.y;
    if (precond) {
      const double2 B2 = Bm[i];
      hx = (2 * c2 < nxl) ? semb_prec(rv.x, B2.x, 1, b0) : 0.0;
      hy = (2 * c2 + 1 < nxl) ? semb_prec(rv.y, B2.y, 1, b0) : 0.0;
    }
    pv.x = __dadd_rn(hx, __dmul_rn(beta, pv.x));
    pv.y = __dadd_rn(hy, __dmul_rn(beta, pv.y));
    p[i] = pv;
  }
}

// generic path: out = M .* out (mask.jl:14) and sum(p .* out .* mult) (pcg.jl:52) -> scal->pap[0]
__global__ void __launch_bounds__(256) semb_mask_dot_kernel(const OpArgs a) {
  __shared__ double red[32];
  if (a.pcg && a.scal->done) return;
  double acc = 0.0;
  for (int row = blockIdx.y * blockDim.y + threadIdx.y; row < a.nyl; row += gridDim.y * blockDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < a.nxl; x += gridDim.x * blockDim.x) {
      const size_t idx = (size_t)row * a.pitch + x;
      const double o = __dmul_rn(mask_at(a, x, row, idx), a.out[idx]);
      a.out[idx] = o;
      if (a.pcg) acc += __dmul_rn(__dmul_rn(a.pout[idx], o), a.mult[idx]);
    }
  if (a.pcg) {
    const double bs = semb_block_sum(acc, red, SEMB_TID, SEMB_NT);
    double tot;
    if (semb_last_block(bs, 0.0, a.partials, nullptr, a.counters, SEMB_NB, SEMB_BID, red, SEMB_TID, SEMB_NT, &tot,
                        nullptr)) {
      a.scal->pap[0] = tot;
      a.scal->pap[1] = 0.0;
      a.scal->pap[2] = 0.0;
    }
  }
}

// ---- generic ABu (ABu.jl:9-37) --------------------------------------------------------------------------
// out (m*mb/nb x n) = (I (x) Br) u : Br (mb x nb, column-major) acts on each nb-row chunk of every column
__global__ void semb_abu_r_kernel(const double* __restrict__ Br, int mb, int nb, const double* __restrict__ u,
                                  int m, int n, long long ldu, double* out, long long ldo) {
  const int mo = m / nb * mb;
  for (int col = blockIdx.y; col < n; col += gridDim.y)
    for (int ro = blockIdx.x * blockDim.x + threadIdx.x; ro < mo; ro += gridDim.x * blockDim.x) {
      const int blk = ro / mb, i = ro - blk * mb;
      const double* uc = u + (size_t)col * ldu + (size_t)blk * nb;
      double s = 0.0;
      for (int k = 0; k < nb; ++k) s = fma(Br[i + (size_t)k * mb], uc[k], s);
      out[(size_t)col * ldo + ro] = s;
    }
}
// out (m x n*ma/na) : out[:, ii] = u[:, jj] * As'  (As ma x na)
__global__ void semb_abu_s_kernel(const double* __restrict__ As, int ma, int na, const double* __restrict__ u,
                                  int m, int n, long long ldu, double* out, long long ldo) {
  const int no = n / na * ma;
  for (int col = blockIdx.y; col < no; col += gridDim.y) {
    const int blk = col / ma, j = col - blk * ma;
    for (int ro = blockIdx.x * blockDim.x + threadIdx.x; ro < m; ro += gridDim.x * blockDim.x) {
      double s = 0.0;
      for (int k = 0; k < na; ++k) s = fma(As[j + (size_t)k * ma], u[(size_t)(blk * na + k) * ldu + ro], s);
      out[(size_t)col * ldo + ro] = s;
    }
  }
}

// grad(u,msh), grad.jl:15-34: ux = rx.*ur + sx.*us, uy = ry.*ur + sy.*us with ur = (I (x) Dr) u, us = (Ds (x) I) u
__global__ void semb_grad_kernel(const double* __restrict__ u, long long pitch, int nr, int ns, int Ex, int ney,
                                 const double* __restrict__ Dr, const double* __restrict__ Ds,
                                 const double* __restrict__ rx, const double* __restrict__ ry,
                                 const double* __restrict__ sx, const double* __restrict__ sy, double* ux, double* uy) {
  const int nxl = nr * Ex, nyl = ns * ney;
  for (int row = blockIdx.y; row < nyl; row += gridDim.y) {
    const int rl = row / ns, j = row - rl * ns;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nxl; c += gridDim.x * blockDim.x) {
      const int e = c / nr, i = c - e * nr;
      const size_t rowb = (size_t)row * pitch + (size_t)e * nr, colb = (size_t)rl * ns * pitch + c;
      double ur = 0, us = 0;
      for (int k = 0; k < nr; ++k) ur = fma(Dr[i * nr + k], u[rowb + k], ur);
      for (int k = 0; k < ns; ++k) us = fma(Ds[j * ns + k], u[colb + (size_t)k * pitch], us);
      const size_t idx = (size_t)row * pitch + c;
      ux[idx] = __dadd_rn(__dmul_rn(rx[idx], ur), __dmul_rn(sx[idx], us));  // grad.jl:30
      uy[idx] = __dadd_rn(__dmul_rn(ry[idx], ur), __dmul_rn(sy[idx], us));  // grad.jl:31
    }
  }
}

// advect.jl:59-60 (dealiased) / :36-37 (plain): Cu = (ux.*Tx + uy.*Ty) .* B ; sign = -1 stores exH = -advect
__global__ void semb_advect_pointwise_kernel(const double* __restrict__ jux, const double* __restrict__ jtx,
                                             const double* __restrict__ juy, const double* __restrict__ jty,
                                             const double* __restrict__ B, long long pitch, int nxl, int nyl,
                                             double* out) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nxl; x += gridDim.x * blockDim.x) {
      const size_t idx = (size_t)row * pitch + x;
      const double c = __dadd_rn(__dmul_rn(jux[idx], jtx[idx]), __dmul_rn(juy[idx], jty[idx]));
      out[idx] = __dmul_rn(c, B[idx]);
    }
}

struct RhsArgs {
  const double* uh[4];
  const double* adv[4];  // advect(uh[i],...) (convectionDiffusion.jl:102); exH[i] = -adv[i]
  double b[4], a[4];
  int k;
};

// makeRHS! (diffusion.jl:55-62), un-fused like the reference's broadcasts:
//   rhs = B.*f ; rhs -= nu.*lapl(ub) ; rhs -= bdfB[1+i] .* (B.*uh[i]) ; rhs = M.*rhs
__global__ void semb_rhs_kernel(const double* __restrict__ f, const double* __restrict__ nu,
                                const double* __restrict__ lub, const double* __restrict__ B, RhsArgs h, long long pitch,
                                int nxl, int nyl, int mx0, int mx1, int my0, int my1, double* rhs) {
  for (int row = blockIdx.y; row < nyl; row += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nxl; x += gridDim.x * blockDim.x) {
      const size_t idx = (size_t)row * pitch + x;
      const double Bv = B[idx];
      double r = __dmul_rn(Bv, f[idx]);                               // mass(f), diffusion.jl:55
      r = __dadd_rn(r, -__dmul_rn(nu[idx], lub[idx]));                // - nu .* lapl(ub), :56
      for (int i = 0; i < h.k; ++i) {
        r = __dadd_rn(r, -__dmul_rn(h.b[i], __dmul_rn(Bv, h.uh[i][idx])));  // :58-60 / convectionDiffusion.jl:103
        if (h.adv[i]) r = __dadd_rn(r, __dmul_rn(h.a[i], -h.adv[i][idx]));   // rhs .+= bdfA[i] .* exH[i], :104
      }
      const bool z = (x == 0 && mx0) || (x == nxl - 1 && mx1) || (row == 0 && my0) || (row == nyl - 1 && my1);
      rhs[idx] = __dmul_rn(z ? 0.0 : 1.0, r);                         // mask, :62
    }
}

// ------------------------------------------------------------------------------------------------
// Fused dealiased advection, advect.jl:45-64, for nr == ns = N on mshV and nrd == nsd = M on mshD:
//   Tx,Ty = grad(T) ; J* = ABu(Js,Jr,*) for Tx,Ty,ux,uy ; JCu = (Jux.*JTx + Juy.*JTy).*B_D ; Cu = ABu(Js',Jr',JCu)
// One CTA works on EB x-consecutive elements of an element row at a time, everything between the loads
// of T,ux,uy,rx,ry,sx,sy,B_D and the store of Cu stays in shared memory (no intermediate fields in HBM:
// ~8 V-sized + 1 D-sized array of traffic instead of ~40).  Contraction order as the reference: the x
// matrix (Br) before the y matrix (As) in both ABu calls (ABu.jl:14-33).
// ------------------------------------------------------------------------------------------------
struct AdvectFusedArgs {
  const double *T, *ux, *uy, *rx, *ry, *sx, *sy, *BD;
  const double *Dr, *Ds;  // row-major N x N
  const double *Jr, *Js;  // column-major M x N (interpMat(mshD.z, mshV.z))
  double* out;
  long long pitchV, pitchD;
  int N, M, Ex, ney, EB;
};

__global__ void __launch_bounds__(256) semb_advect_fused_kernel(const AdvectFusedArgs a) {
  extern __shared__ double sh[];
  const int N = a.N, M = a.M, EB = a.EB, NN = N * N, MN = M * N, MM = M * M;
  double* sDr = sh;            // [N][N]   Dr(i,k)
  double* sDs = sDr + NN;      // [N][N]   Ds(j,k)
  double* sJr = sDs + NN;      // [M][N]   Jr(m,i)
  double* sJs = sJr + MN;      // [M][N]   Js(n,j)
  double* tT = sJs + MN;       // [EB][N][N] T, later unused
  double* tF = tT + EB * NN;   // [4][EB][N][N]  Tx, Ty, ux, uy   (element tile: [j][i], i = x contiguous)
  double* tB = tF + 4 * EB * NN;  // [4][EB][N][M] x-interpolated; later [EB][M][N] x-projected JCu
  double* tD = tB + 4 * EB * MN;  // [4][EB][M][M] on the dealiasing grid; tD[0] becomes JCu
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int q = tid; q < NN; q += nt) {
    sDr[q] = a.Dr[q];
    sDs[q] = a.Ds[q];
  }
  for (int q = tid; q < MN; q += nt) {
    const int m = q / N, i = q - m * N;
    sJr[q] = a.Jr[m + (size_t)i * M];
    sJs[q] = a.Js[m + (size_t)i * M];
  }
  const int e0 = blockIdx.x * EB;
  const int nbe = min(EB, a.Ex - e0);
  for (int r = blockIdx.y; r < a.ney; r += gridDim.y) {
    __syncthreads();
    // load T, ux, uy tiles (coalesced along x)
    for (int q = tid; q < N * nbe * N; q += nt) {
      const int j = q / (nbe * N), xx = q - j * (nbe * N), e = xx / N, i = xx - e * N;
      const size_t g = (size_t)(r * N + j) * a.pitchV + (size_t)e0 * N + xx;
      tT[e * NN + j * N + i] = a.T[g];
      tF[(2 * EB + e) * NN + j * N + i] = a.ux[g];
      tF[(3 * EB + e) * NN + j * N + i] = a.uy[g];
    }
    __syncthreads();
    // grad: Tx = rx.*ur + sx.*us, Ty = ry.*ur + sy.*us (grad.jl:27-31)
    for (int q = tid; q < N * nbe * N; q += nt) {
      const int j = q / (nbe * N), xx = q - j * (nbe * N), e = xx / N, i = xx - e * N;
      const double* t = tT + e * NN;
      double ur = 0.0, us = 0.0;
      for (int k = 0; k < N; ++k) ur = fma(sDr[i * N + k], t[j * N + k], ur);
      for (int k = 0; k < N; ++k) us = fma(sDs[j * N + k], t[k * N + i], us);
      const size_t g = (size_t)(r * N + j) * a.pitchV + (size_t)e0 * N + xx;
      tF[(0 * EB + e) * NN + j * N + i] = __dadd_rn(__dmul_rn(a.rx[g], ur), __dmul_rn(a.sx[g], us));
      tF[(1 * EB + e) * NN + j * N + i] = __dadd_rn(__dmul_rn(a.ry[g], ur), __dmul_rn(a.sy[g], us));
    }
    __syncthreads();
    // interpolate along x: tB[f][e][j][m] = sum_i Jr(m,i) f[e][j][i]
    for (int q = tid; q < 4 * nbe * N * M; q += nt) {
      const int m = q % M, rest = q / M, j = rest % N, fe = rest / N, e = fe % nbe, f = fe / nbe;
      const double* src = tF + (f * EB + e) * NN + j * N;
      double s_ = 0.0;
      for (int i = 0; i < N; ++i) s_ = fma(sJr[m * N + i], src[i], s_);
      tB[(f * EB + e) * MN + j * M + m] = s_;
    }
    __syncthreads();
    // interpolate along y: tD[f][e][n][m] = sum_j Js(n,j) tB[f][e][j][m]
    for (int q = tid; q < 4 * nbe * M * M; q += nt) {
      const int m = q % M, rest = q / M, n = rest % M, fe = rest / M, e = fe % nbe, f = fe / nbe;
      const double* src = tB + (f * EB + e) * MN + m;
      double s_ = 0.0;
      for (int j = 0; j < N; ++j) s_ = fma(sJs[n * N + j], src[j * M], s_);
      tD[(f * EB + e) * MM + n * M + m] = s_;
    }
    __syncthreads();
    // JCu = (Jux.*JTx + Juy.*JTy) .* B_D (advect.jl:59-60), in place in tD[0]
    for (int q = tid; q < M * nbe * M; q += nt) {
      const int n = q / (nbe * M), xx = q - n * (nbe * M), e = xx / M, m = xx - e * M;
      const int o = e * MM + n * M + m;
      const double c = __dadd_rn(__dmul_rn(tD[2 * EB * MM + o], tD[0 * EB * MM + o]),
                                 __dmul_rn(tD[3 * EB * MM + o], tD[1 * EB * MM + o]));
      const size_t g = (size_t)(r * M + n) * a.pitchD + (size_t)e0 * M + xx;
      tD[o] = __dmul_rn(c, a.BD[g]);
    }
    __syncthreads();
    // project back along x (Br = Jr'): tB[e][n][i] = sum_m Jr(m,i) JCu[e][n][m]
    for (int q = tid; q < nbe * M * N; q += nt) {
      const int i = q % N, rest = q / N, n = rest % M, e = rest / M;
      const double* src = tD + e * MM + n * M;
      double s_ = 0.0;
      for (int m = 0; m < M; ++m) s_ = fma(sJr[m * N + i], src[m], s_);
      tB[e * MN + n * N + i] = s_;
    }
    __syncthreads();
    // project back along y (As = Js'): Cu[e][j][i] = sum_n Js(n,j) tB[e][n][i] ; coalesced store
    for (int q = tid; q < N * nbe * N; q += nt) {
      const int j = q / (nbe * N), xx = q - j * (nbe * N), e = xx / N, i = xx - e * N;
      const double* src = tB + e * MN + i;
      double s_ = 0.0;
      for (int n = 0; n < M; ++n) s_ = fma(sJs[n * N + j], src[n * N], s_);
      a.out[(size_t)(r * N + j) * a.pitchV + (size_t)e0 * N + xx] = s_;
    }
  }
}

dim3 rows_grid(int ncols, int nrows, int threads) {
  int gx = (ncols + threads - 1) / threads;
  if (gx < 1) gx = 1;
  if (gx > 1024) gx = 1024;
  int gy = nrows < 1 ? 1 : (nrows > 32768 ? 32768 : nrows);
  return dim3(gx, gy);
}

int flat_blocks(size_t n, int sm_count) {
  size_t b = (n + 255) / 256;
  size_t cap = (size_t)sm_count * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

#define SEMB_POST_LAUNCH(ctx)              \
  SEMB_CHECK_CUDA(cudaGetLastError());     \
  (ctx)->launches++;                       \
  return SEMB_OK

int semb_launch_seam_x(semb_ctx* ctx, const OpArgs& a) {
  if (a.nxseam == 0) return SEMB_OK;
  const long long total = (long long)a.nxseam * ((a.y_end > a.y_begin ? a.y_end : a.nyl) - a.y_begin);
  int blocks = (int)((total + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  OpArgs b = a;
  b.partials = a.partials;  // caller passes the slot for this kernel
  semb_seam_x_kernel<<<blocks, 256, 0, ctx->stream>>>(b);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_seam_y(semb_ctx* ctx, const OpArgs& a, int nhalo_lo, int nhalo_hi, bool domask, const P2PArgs& x) {
  const int nq = a.nyseam + nhalo_lo + nhalo_hi;
  if (nq == 0) return SEMB_OK;
  const long long total = (long long)nq * a.nxl;
  int blocks = (int)((total + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  semb_seam_y_kernel<<<blocks, 256, 0, ctx->stream>>>(a, nhalo_lo, nhalo_hi, domask ? 1 : 0, x);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_halo_push(semb_ctx* ctx, const double* out, long long pitch, int nxl, int nyl, double* dst_lo,
                          double* dst_hi, unsigned long long* flag_lo, unsigned long long* flag_hi,
                          unsigned long long epoch, unsigned* counter, const SembScal* scal, int pcg) {
  int blocks = (nxl + 1023) / 1024;
  if (blocks < 1) blocks = 1;
  if (blocks > 32) blocks = 32;
  semb_halo_push_kernel<<<blocks, 256, 0, ctx->stream>>>(out, pitch, nxl, nyl, dst_lo, dst_hi, flag_lo, flag_hi, epoch,
                                                         counter, scal, pcg);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_gs_x(semb_ctx* ctx, const double* u, double* out, long long pitch, int N, int Ex, int nxl, int nyl,
                     int perx) {
  semb_gs_x_kernel<<<rows_grid(nxl, nyl, 256), 256, 0, ctx->stream>>>(u, out, pitch, N, Ex, nxl, nyl, perx);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_mask(semb_ctx* ctx, const double* u, const double* M, double* out, size_t n) {
  semb_mask_kernel<<<flat_blocks(n / 2, ctx->sm_count), 256, 0, ctx->stream>>>(
      (const double2*)u, (const double2*)M, (double2*)out, n / 2);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_axpby(semb_ctx* ctx, double a, const double* x, double b, double* y, size_t n) {
  semb_axpby_kernel<<<flat_blocks(n / 2, ctx->sm_count), 256, 0, ctx->stream>>>(a, (const double2*)x, b,
                                                                                 (double2*)y, n / 2);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_fill(semb_ctx* ctx, double* x, double v, long long pitch, int nxl, int nyl) {
  semb_fill_kernel<<<rows_grid((int)pitch, nyl, 256), 256, 0, ctx->stream>>>(x, v, pitch, nxl, nyl);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_fill_random(semb_ctx* ctx, double* x, long long pitch, int nxl, int nyl, long long gnxl,
                            long long gy0, uint64_t seed) {
  semb_fill_random_kernel<<<rows_grid(nxl, nyl, 256), 256, 0, ctx->stream>>>(x, pitch, nxl, nyl, gnxl, gy0, seed);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_mult(semb_ctx* ctx, double* mult, long long pitch, int nr, int ns, int Ex, int Ey, int ey0,
                     int ney, int perx, int pery) {
  semb_mult_kernel<<<rows_grid(nr * Ex, ns * ney, 256), 256, 0, ctx->stream>>>(mult, pitch, nr, ns, Ex, Ey, ey0,
                                                                               ney, perx, pery);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_mask_gen(semb_ctx* ctx, double* M, long long pitch, int nxl, int nyl, int mx0, int mx1, int my0,
                         int my1) {
  semb_mask_gen_kernel<<<rows_grid(nxl, nyl, 256), 256, 0, ctx->stream>>>(M, pitch, nxl, nyl, mx0, mx1, my0, my1);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_grid(semb_ctx* ctx, double* x, double* y, long long pitch, int nr, int ns, int Ex, int Ey, int ey0,
                     int ney, const double* d_z0r, const double* d_z0s, int kind, const double* params) {
  semb_grid_kernel<<<rows_grid(nr * Ex, ns * ney, 256), 256, 0, ctx->stream>>>(
      x, y, pitch, nr, ns, Ex, Ey, ey0, ney, d_z0r, d_z0s, kind, params[0], params[1], params[2]);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_geom(semb_ctx* ctx, const double* x, const double* y, long long pitch, int nr, int ns, int Ex,
                     int ney, const double* dDr, const double* dDs, const double* d_wr, const double* d_ws,
                     double* J, double* Ji, double* rx, double* ry, double* sx, double* sy, double* B, double* Bi,
                     double* G11, double* G12, double* G22) {
  semb_geom_kernel<<<rows_grid(nr * Ex, ns * ney, 256), 256, 0, ctx->stream>>>(
      x, y, pitch, nr, ns, Ex, ney, dDr, dDs, d_wr, d_ws, J, Ji, rx, ry, sx, sy, B, Bi, G11, G12, G22);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_generic_local(semb_ctx* ctx, const OpArgs& a, int nr, int ns, const double* dDr, const double* dDs,
                              double* tmp_wr, double* tmp_ws, bool massterm) {
  dim3 g = rows_grid(a.nxl, a.nyl, 256);
  semb_generic_pass1<<<g, 256, 0, ctx->stream>>>(a, nr, ns, dDr, dDs, tmp_wr, tmp_ws);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  semb_generic_pass2<<<g, 256, 0, ctx->stream>>>(a, nr, ns, dDr, dDs, tmp_wr, tmp_ws, massterm ? 1 : 0);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_reduce(semb_ctx* ctx, semb_mesh* m, int which, const double* a, const double* b, const P2PArgs& x,
                       double ref) {
  Grid2D g = grid2d(m->pitch, m->nyl, ctx->sm_count, m->npartials / 2);
  semb_reduce_kernel<<<g.grid, g.block, 0, ctx->stream>>>(which, (const double2*)a, (const double2*)b,
                                                          (const double2*)m->d_wx1d, m->d_wy1d, (int)(m->pitch / 2),
                                                          m->nyl, m->d_partials, m->d_counters + 4, m->d_scal, x, ref, m->nxl);
  SEMB_POST_LAUNCH(ctx);
}

// custom-operator PCG: the reduction kernel left sum(p.*Ap.*mult) (all ranks) in red[0]; hand it to the update kernel
__global__ void semb_pcg_set_pap_kernel(SembScal* scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double v = scal->red[0];
  scal->pap[0] = v;
  scal->pap[1] = 0.0;
  scal->pap[2] = 0.0;
  scal->pap_total = v;
}
int semb_launch_pcg_set_pap(semb_ctx* ctx, semb_mesh* m) {
  semb_pcg_set_pap_kernel<<<1, 32, 0, ctx->stream>>>(m->d_scal);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_reduce_finalize(semb_ctx* ctx, semb_mesh* m, int which) {
  semb_reduce_finalize_kernel<<<1, 32, 0, ctx->stream>>>(m->d_scal, which);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_pcg_init(semb_ctx* ctx, semb_mesh* m, const double* b, double* x, double* r, double* p, double* hout,
                         int precond, double prec_b0, double tol, long long maxiter, const P2PArgs& xa) {
  Grid2D g = grid2d(m->pitch, m->nyl, ctx->sm_count, m->npartials / 2);
  semb_pcg_init_kernel<<<g.grid, g.block, 0, ctx->stream>>>(
      (const double2*)b, (double2*)x, (double2*)r, (double2*)p, (double2*)hout, (const double2*)m->arr[SEMB_B],
      (const double2*)m->d_wx1d, m->d_wy1d, (int)(m->pitch / 2), m->nxl, m->nyl, precond, prec_b0, tol, maxiter,
      m->d_partials, m->d_counters + 3, m->d_scal, xa);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_pcg_update(semb_ctx* ctx, semb_mesh* m, double* x, double* r, const double* p, const double* Ap,
                           double* hout, int precond, double prec_b0, const P2PArgs& xa) {
  Grid2D g = grid2d(m->pitch, m->nyl, ctx->sm_count, m->npartials / 2);
  semb_pcg_update_kernel<<<g.grid, g.block, 0, ctx->stream>>>(
      (double2*)x, (double2*)r, (const double2*)p, (const double2*)Ap, (double2*)hout, (const double2*)m->arr[SEMB_B],
      (const double2*)m->d_wx1d, m->d_wy1d, (int)(m->pitch / 2), m->nxl, m->nyl, precond, prec_b0, m->d_partials,
      m->d_counters + 3, m->d_scal, xa);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_pcg_dir(semb_ctx* ctx, semb_mesh* m, const double* r, double* p, int precond, double prec_b0) {
  Grid2D g = grid2d(m->pitch, m->nyl, ctx->sm_count);
  semb_pcg_dir_kernel<<<g.grid, g.block, 0, ctx->stream>>>((const double2*)r, (double2*)p,
                                                           (const double2*)m->arr[SEMB_B], (int)(m->pitch / 2),
                                                           m->nxl, m->nyl, precond, prec_b0, m->d_scal);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_mask_dot(semb_ctx* ctx, semb_mesh* m, const OpArgs& a) {
  int bx = 32;
  while (bx < 256 && bx < a.nxl) bx <<= 1;
  const int by = 256 / bx;
  int gx = (a.nxl + bx - 1) / bx;
  if (gx > 64) gx = 64;
  int gy = (a.nyl + by - 1) / by;
  const int cap = m->npartials / gx;
  if (gy > cap) gy = cap;
  if (gy > ctx->sm_count * 8) gy = ctx->sm_count * 8;
  semb_mask_dot_kernel<<<dim3(gx, gy), dim3(bx, by), 0, ctx->stream>>>(a);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_pcg_pack_pap(semb_ctx* ctx, semb_mesh* m) {
  semb_pcg_pack_pap_kernel<<<1, 32, 0, ctx->stream>>>(m->d_scal);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_pcg_combine_pap(semb_ctx* ctx, semb_mesh* m) {
  semb_pcg_combine_pap_kernel<<<1, 32, 0, ctx->stream>>>(m->d_scal);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_pcg_finalize(semb_ctx* ctx, semb_mesh* m, int first) {
  semb_pcg_finalize_kernel<<<1, 32, 0, ctx->stream>>>(m->d_scal, first);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_abu_r(semb_ctx* ctx, const double* Br, int mb, int nb, const double* u, int m, int n, long long ldu,
                      double* out, long long ldo) {
  semb_abu_r_kernel<<<rows_grid(m / nb * mb, n, 128), 128, 0, ctx->stream>>>(Br, mb, nb, u, m, n, ldu, out, ldo);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_abu_s(semb_ctx* ctx, const double* As, int ma, int na, const double* u, int m, int n, long long ldu,
                      double* out, long long ldo) {
  semb_abu_s_kernel<<<rows_grid(m, n / na * ma, 128), 128, 0, ctx->stream>>>(As, ma, na, u, m, n, ldu, out, ldo);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_grad(semb_ctx* ctx, semb_mesh* m, const double* u, double* ux, double* uy) {
  semb_grad_kernel<<<rows_grid(m->nxl, m->nyl, 256), 256, 0, ctx->stream>>>(
      u, m->pitch, m->nr, m->ns, m->Ex, m->ney, m->dDr, m->dDs, m->arr[SEMB_RX], m->arr[SEMB_RY], m->arr[SEMB_SX],
      m->arr[SEMB_SY], ux, uy);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_advect_pointwise(semb_ctx* ctx, semb_mesh* m, const double* jux, const double* jtx, const double* juy,
                                 const double* jty, double* out) {
  semb_advect_pointwise_kernel<<<rows_grid(m->nxl, m->nyl, 256), 256, 0, ctx->stream>>>(
      jux, jtx, juy, jty, m->arr[SEMB_B], m->pitch, m->nxl, m->nyl, out);
  SEMB_POST_LAUNCH(ctx);
}

int semb_launch_rhs(semb_ctx* ctx, semb_mesh* m, const double* f, const double* nu, const double* lub, int k,
                    const double* const* uh, const double* b, const double* const* adv, const double* a, int mx0,
                    int mx1, int my0, int my1, double* rhs) {
  RhsArgs h;
  h.k = k;
  for (int i = 0; i < 4; ++i) {
    h.uh[i] = i < k ? uh[i] : nullptr;
    h.b[i] = i < k ? b[i] : 0.0;
    h.adv[i] = (adv && i < k) ? adv[i] : nullptr;
    h.a[i] = (a && i < k) ? a[i] : 0.0;
  }
  semb_rhs_kernel<<<rows_grid(m->nxl, m->nyl, 256), 256, 0, ctx->stream>>>(f, nu, lub, m->arr[SEMB_B], h, m->pitch,
                                                                          m->nxl, m->nyl, mx0, mx1, my0, my1, rhs);
  SEMB_POST_LAUNCH(ctx);
}

// returns SEMB_OK and *done = 1 if the fused kernel ran, *done = 0 if the sizes do not fit it
int semb_launch_advect_fused(semb_ctx* ctx, semb_mesh* V, semb_mesh* D, const double* T, const double* ux,
                             const double* uy, const double* dJr, const double* dJs, double* out, int* done) {
  *done = 0;
  if (V->nr != V->ns || D->nr != D->ns || D->nr < V->nr || D->nr > 32) return SEMB_OK;
  const int N = V->nr, M = D->nr;
  int EB = 4;
  auto bytes = [&](int eb) { return (size_t)(2 * N * N + 2 * M * N + eb * (5 * N * N + 4 * M * N + 4 * M * M)) * 8; };
  while (EB > 1 && bytes(EB) > 72 * 1024) EB >>= 1;
  if (bytes(EB) > 200 * 1024) return SEMB_OK;
  static size_t attr[64] = {0};  // per device
  const int dev = ctx->device & 63;
  if (bytes(EB) > attr[dev]) {
    SEMB_CHECK_CUDA(cudaFuncSetAttribute(semb_advect_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bytes(EB)));
    attr[dev] = bytes(EB);
  }
  AdvectFusedArgs a;
  a.T = T;
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
  a.out = out;
  a.pitchV = V->pitch;
  a.pitchD = D->pitch;
  a.N = N;
  a.M = M;
  a.Ex = V->Ex;
  a.ney = V->ney;
  a.EB = EB;
  const int gx = (V->Ex + EB - 1) / EB;
  int gy = (ctx->sm_count * 12 + gx - 1) / gx;
  if (gy > V->ney) gy = V->ney;
  if (gy < 1) gy = 1;
  semb_advect_fused_kernel<<<dim3(gx, gy), 256, bytes(EB), ctx->stream>>>(a);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  *done = 1;
  return SEMB_OK;
}
