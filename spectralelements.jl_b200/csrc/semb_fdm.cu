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
// (semb_fdm_tables, semb_host.cpp).  6-8x fewer PCG iterations than no preconditioner on the BASELINE meshes.
//
// One kernel per application (semb_fdm_kernel<N>): gathers the (N+2)^2 tiles of r, applies the four contractions out of
// registers (the strip kernel's two thread<->line mappings; W is folded into the tables), sums the tile entries that land
// on a node and on its duplicates -- x overlaps, x pairs, y overlaps, y pairs, the association of gatherScatter.jl:13 --
// applies the mask and, inside pcg, accumulates sum(r .* h .* mult) (pcg.jl:45) deterministically.  16 bytes per node
// of HBM traffic; on several ranks the neighbour slabs' boundary rows of r arrive through peer memory (below).
#include "semb_reduce.cuh"
#include "semb_vec.cuh"

namespace {

struct FdmArgs {
  const double* r;      // residual (continuous)
  double* out;          // h = opM(r)
  const double* tab;    // [dir 2][class 4][FdmTab<N+2>::SIZE] contraction tables (layouts: FdmTab)
  const double* el;     // [3][nel]: 1/hx^2, 1/hy^2, 1/(hx*hy) of the elements (hx, hy: half-lengths)
  long long nel;
  const double* wx;     // W = wx[x] * wy[y]: only read for N = 3 (N >= 4: folded into the tables)
  const double* wy;
  const double* mult_x; // mult(x,y) = mult_x[x] * mult_y[y]  (PCG reduction)
  const double* mult_y;
  long long pitch;
  int N, Ex, Ey, ey0, ney, nxl, nyl, perx, pery;
  int mx0, mx1, my0, my1;
  double nu, k;
  SembScal* scal;
  double* partials;
  unsigned* counter;
  int pcg;      // 1: inside pcg (early exit on done; reduction + advance), 2: same, first call (pcg.jl:25-33 state)
  // ---- several ranks (y-slabs): the N rows of r next to a slab boundary that the neighbour's halo tiles need travel
  // through peer memory as flag-in-data entries (semb_ll_store), pushed by the first CTA row at kernel start
  int has_lo, has_hi;
  const uint4* ghost;   // local ghost rows [parity][side][N][pitch]; side 0: rows Y = -N-2, -N..-2 below the slab,
                        // side 1: rows Y = nyl+1..nyl+N-1, nyl+N+1 above it (Y = -1 and Y = nyl are the shared lines)
  uint4* peer_lo;       // the lower neighbour's ghost rows (we fill its side 1), the upper neighbour's (its side 0)
  uint4* peer_hi;
  const double* el_lo;  // [3][Ex] scalings of the neighbours' adjoining element rows
  const double* el_hi;
  unsigned long long ep_host;   // epoch = ep_host (stand-alone applications) + *ep_dev (applications inside pcg) + 1
  unsigned long long* ep_dev;
};

// class of an element in a direction: 0 interior (neighbours on both sides), 1 first, 2 last, 3 single
__device__ __forceinline__ int fdm_class(int e, int E, int per) {
  if (per) return 0;
  return (e == 0 ? 1 : 0) | (e == E - 1 ? 2 : 0);
}

// ---- contraction tables -------------------------------------------------------------------------------------------------
// One slot per (direction, element class): [eo flag][lambda (N2)][F][G] with F the nodes -> modes matrix (rows = node,
// padded to an even length so that a row is read with LDS.128) and G the modes -> nodes matrix (rows = mode).
// The extended 1-D operator of an element with the same kind of neighbour on both sides is symmetric under the
// reflection i <-> N2-1-i, so its eigenvectors are even or odd: with the even modes first, y = S'x and x = S y split
// into two half-size products on e_k = x_k + x_{N2-1-k}, o_k = x_k - x_{N2-1-k} (half the FMAs and table loads) -- the
// even-odd trick of the strip kernel, here for a symmetric eigenbasis.  Boundary classes use the full products.
template <int N2>
struct FdmTab {
  static constexpr int NP = (N2 + 1) & ~1;
  static constexpr int H = N2 / 2, ODD = N2 & 1, NE = H + ODD, NO = H;
  static constexpr int NEP = (NE + 1) & ~1, NOP = (NO + 1) & ~1;
  static constexpr int OFF_LAM = 1, OFF_F = (1 + N2 + 1) & ~1, OFF_G = OFF_F + N2 * NP;
  static constexpr int SIZE = OFF_G + N2 * NP;
  // S: (N2 x N2) column-major (S[i + m*N2]), lam[N2]; returns the slot (host)
  static void fill(const double* S, const double* lam, double* slot, bool allow_eo) {
    for (int q = 0; q < SIZE; ++q) slot[q] = 0.0;
    // parity of every mode (padding modes: zero vectors, lambda = inf, count as even)
    int par[N2];
    bool eo = allow_eo;
    for (int m = 0; m < N2; ++m) {
      double se = 0.0, so = 0.0, nn = 0.0;
      for (int i = 0; i < N2; ++i) {
        const double a = S[i + (size_t)m * N2], b = S[(N2 - 1 - i) + (size_t)m * N2];
        se += (a - b) * (a - b);
        so += (a + b) * (a + b);
        nn += a * a;
      }
      if (se <= 1e-20 * nn) par[m] = 0;
      else if (so <= 1e-20 * nn) par[m] = 1;
      else eo = false, par[m] = 0;
    }
    int ne = 0, no = 0;
    for (int m = 0; m < N2; ++m) (par[m] ? no : ne)++;
    if (ne != NE || no != NO) eo = false;
    if (!eo) {
      slot[0] = 0.0;
      for (int m = 0; m < N2; ++m) slot[OFF_LAM + m] = lam[m];
      for (int i = 0; i < N2; ++i)
        for (int m = 0; m < N2; ++m) {
          slot[OFF_F + i * NP + m] = S[i + (size_t)m * N2];
          slot[OFF_G + m * NP + i] = S[i + (size_t)m * N2];
        }
      return;
    }
    slot[0] = 1.0;
    int order[N2], c = 0;  // even modes first, then the odd ones (each group in ascending eigenvalue order)
    for (int m = 0; m < N2; ++m)
      if (!par[m]) order[c++] = m;
    for (int m = 0; m < N2; ++m)
      if (par[m]) order[c++] = m;
    for (int m = 0; m < N2; ++m) slot[OFF_LAM + m] = lam[order[m]];
    double* FE = slot + OFF_F;            // [k < NE][m < NE]: S(k, even m)   (k = H is the middle node when N2 is odd)
    double* FO = FE + NE * NEP;           // [k < H][m < NO]:  S(k, odd m)
    double* GE = slot + OFF_G;            // [m < NE][i < NE]: S(i, even m)
    double* GO = GE + NE * NEP;           // [m < NO][i < H]:  S(i, odd m)
    for (int k = 0; k < NE; ++k)
      for (int m = 0; m < NE; ++m) {
        FE[k * NEP + m] = S[k + (size_t)order[m] * N2];
        GE[m * NEP + k] = S[k + (size_t)order[m] * N2];
      }
    for (int k = 0; k < H; ++k)
      for (int m = 0; m < NO; ++m) {
        FO[k * NOP + m] = S[k + (size_t)order[NE + m] * N2];
        GO[m * NOP + k] = S[k + (size_t)order[NE + m] * N2];
      }
  }
};

// y = S' x (nodes -> modes) out of registers; slot = FdmTab layout in shared memory
template <int N2, bool EO>
__device__ __forceinline__ void fdm_fwd(const double* __restrict__ slot, const double (&x)[N2], double (&y)[N2]) {
  using TB = FdmTab<N2>;
  if constexpr (!EO) {
    const double* F = slot + TB::OFF_F;
#pragma unroll
    for (int m = 0; m < N2; ++m) y[m] = 0.0;
#pragma unroll
    for (int k = 0; k < N2; ++k) {
#pragma unroll
      for (int m = 0; m < N2; ++m) y[m] = fma(F[k * TB::NP + m], x[k], y[m]);
    }
  } else {
    constexpr int H = TB::H, NE = TB::NE, NO = TB::NO;
    const double* FE = slot + TB::OFF_F;
    const double* FO = FE + NE * TB::NEP;
#pragma unroll
    for (int m = 0; m < N2; ++m) y[m] = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const double e = x[k] + x[N2 - 1 - k], o = x[k] - x[N2 - 1 - k];
#pragma unroll
      for (int m = 0; m < NE; ++m) y[m] = fma(FE[k * TB::NEP + m], e, y[m]);
#pragma unroll
      for (int m = 0; m < NO; ++m) y[NE + m] = fma(FO[k * TB::NOP + m], o, y[NE + m]);
    }
    if (TB::ODD) {
#pragma unroll
      for (int m = 0; m < NE; ++m) y[m] = fma(FE[H * TB::NEP + m], x[H], y[m]);
    }
  }
}

// x = S y (modes -> nodes)
template <int N2, bool EO>
__device__ __forceinline__ void fdm_bwd(const double* __restrict__ slot, const double (&y)[N2], double (&x)[N2]) {
  using TB = FdmTab<N2>;
  if constexpr (!EO) {
    const double* G = slot + TB::OFF_G;
#pragma unroll
    for (int i = 0; i < N2; ++i) x[i] = 0.0;
#pragma unroll
    for (int m = 0; m < N2; ++m) {
#pragma unroll
      for (int i = 0; i < N2; ++i) x[i] = fma(G[m * TB::NP + i], y[m], x[i]);
    }
  } else {
    constexpr int H = TB::H, NE = TB::NE, NO = TB::NO;
    const double* GE = slot + TB::OFF_G;
    const double* GO = GE + NE * TB::NEP;
    double sv[NE], dv[H > 0 ? H : 1];
#pragma unroll
    for (int i = 0; i < NE; ++i) sv[i] = 0.0;
#pragma unroll
    for (int i = 0; i < H; ++i) dv[i] = 0.0;
#pragma unroll
    for (int m = 0; m < NE; ++m) {
#pragma unroll
      for (int i = 0; i < NE; ++i) sv[i] = fma(GE[m * TB::NEP + i], y[m], sv[i]);
    }
#pragma unroll
    for (int m = 0; m < NO; ++m) {
#pragma unroll
      for (int i = 0; i < H; ++i) dv[i] = fma(GO[m * TB::NOP + i], y[NE + m], dv[i]);
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      x[i] = sv[i] + dv[i];
      x[N2 - 1 - i] = sv[i] - dv[i];
    }
    if (TB::ODD) x[H] = sv[H];
  }
}

template <int N>
struct FdmCfg {
  static constexpr int N2 = N + 2;
  static constexpr int T = 256;
  static constexpr int BX = T / N2;        // tiles per CTA: BX - 2 output elements + one halo tile on either side
  static constexpr int S = N2 | 1;         // element stride in the tile buffer (odd: conflict-free in both mappings)
  static constexpr int PW = BX * S;        // row pitch of a tile buffer
  static constexpr int TSZ = FdmTab<N2>::SIZE;
  static constexpr int SMEM = (2 * N2 * PW + 4 * TSZ) * 8;   // two tile buffers, 3 x tables + the y table in use
  static constexpr int MAXB = (227 * 1024) / (SMEM + 1024);
  // CTAs per SM: three (80 registers) only while that does not spill; from N = 7 on two CTAs at 128 registers are faster
  // (N = 9: 136 bytes of spills and 34 % long-scoreboard stalls at 80 registers, 1.19 -> 0.96 ms per application at 128)
  static constexpr int WANTB = N <= 6 ? 3 : 2;
  static constexpr int MINB = MAXB >= WANTB ? WANTB : (MAXB >= 2 ? 2 : 1);
};

__device__ __forceinline__ void fdm_cp_async8(uint32_t dst_smem, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void fdm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fdm_cp_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// 1/x without the special-case branch of the IEEE division: MUFU.RCP64H and two Newton steps (<= 1 ulp); x = 0, +-inf
// and NaN come out as NaN, which the caller's |d| <= 1e8 test turns into 0 like the reference's cut-off
__device__ __forceinline__ double fdm_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// One launch per application.  A CTA owns a strip of up to BX-2 x-consecutive elements plus one halo element on either
// side, and MARCHES through a chunk of element rows plus one halo row below and above (the halo tiles are computed twice,
// by the two CTAs that need them: no exchange between CTAs, no second pass over the field).  Per element row:
//   P1 (thread <-> tile column)   the (N+2) x (N+2) tiles of r -- prefetched one element row ahead with cp.async into
//                                 the other tile buffer -- Sy' along y out of registers, back into the buffer in place
//   P2 (thread <-> tile row)      Sx' along x, Di, Sx back along x, in place
//   P3 (thread <-> tile column)   Sy back along y; the columns next to an element interface go back to the buffer
//   P4 (same threads)             x sums (tile overlaps, then the x pair of gatherScatter.jl:13), y sums against the three
//                                 rows carried in registers from the previous element row (overlaps, then the y pair),
//                                 mask, store; inside pcg the sum(r.*h.*mult) partial (pcg.jl:45)
// The counting weights W are folded into the rows of the eigenvector tables (N >= 4: the weight of a tile node only
// depends on the class of the element; N = 3 multiplies explicitly).  Every sum has two terms (three for N = 3), in a
// fixed order: the result does not depend on the launch geometry.
// Kernel parameters: the arguments plus the INTERIOR class's even-odd tables (x, y).  Read from the parameter constant bank
// they reach the DFMAs through LDCU.128 / uniform registers instead of broadcast LDS.128 (the strip kernel's trick for
// N >= 10, semb_strip.cuh): the shared-memory pipe, 62 % busy at N = 9 with half of its wavefronts being table loads,
// keeps the tile traffic only.  Boundary classes (no even-odd split) keep their tables in shared memory.
template <int N>
struct FdmParams {
  FdmArgs a;
  int cok;  // the interior tables are even-odd (symmetric nodes): use them; otherwise every class goes through shared memory
  alignas(16) double tc[2][FdmTab<N + 2>::SIZE];
};

template <int N>
__global__ void __launch_bounds__(256, FdmCfg<N>::MINB) semb_fdm_kernel(const __grid_constant__ FdmParams<N> P) {
  using C = FdmCfg<N>;
  using TB = FdmTab<C::N2>;
  const FdmArgs& a = P.a;
  constexpr int N2 = C::N2, BX = C::BX, S = C::S, PW = C::PW, TSZ = C::TSZ;
  constexpr bool FOLDW = N >= 4;
  extern __shared__ __align__(16) double sm[];
  double* sTx = sm;                // [3][TSZ] x tables: slot = class (class 3, a single element, only occurs alone: slot 0)
  double* sTy = sTx + 3 * TSZ;     // [TSZ] y table of the current element row's class
  double* bufs = sTy + TSZ;        // [2][N2][PW] tile buffers
  __shared__ double red[32];
  __shared__ double sh_tot[2];
  if (a.pcg && a.scal->done) return;
  const int t = threadIdx.x;
  const size_t pitch = (size_t)a.pitch;
  // strip: output elements [o0, o1), tile slot e <-> element o0 - 1 + e
  const int o0 = (int)(((long long)blockIdx.x * a.Ex) / gridDim.x), o1 = (int)(((long long)(blockIdx.x + 1) * a.Ex) / gridDim.x);
  const int nout = o1 - o0;
  // chunk: output element rows [r0, r1), tile rows r0-1 .. r1
  const int r0 = (int)(((long long)blockIdx.y * a.ney) / gridDim.y), r1 = (int)(((long long)(blockIdx.y + 1) * a.ney) / gridDim.y);
  const bool wrapy = a.pery && a.ney == a.Ey;
  const bool lo_any = wrapy || a.has_lo, hi_any = wrapy || a.has_hi;   // an element row below / above the slab exists
  const bool ranks = a.has_lo || a.has_hi;
  // (re-read where it is needed instead of kept in registers: *ep_dev only moves when the last CTA is done)
  auto cur_epoch = [&]() { return a.ep_host + *(volatile unsigned long long*)a.ep_dev + 1ull; };
  if (ranks && blockIdx.y == 0) {
    const unsigned long long epoch = cur_epoch();
    const int par = (int)(epoch & 1ull);
    const unsigned tag = semb_ll_tag(epoch);
    // the rows the neighbours' halo tiles read, for this strip's columns, straight into their memory over NVLink
    const int xs0 = o0 * N, ncols = nout * N;
    for (int i = t; i < N * ncols; i += C::T) {
      const int sl = i / ncols, x = xs0 + i - sl * ncols;
      if (a.has_lo) {
        const int row = (sl < N - 1) ? 1 + sl : N + 1;
        semb_ll_store(a.peer_lo + ((size_t)(2 * par + 1) * N + sl) * pitch + x, a.r[(size_t)row * pitch + x], tag);
      }
      if (a.has_hi) {
        const int row = (sl == 0) ? a.nyl - N - 2 : a.nyl - N - 1 + sl;
        semb_ll_store(a.peer_hi + ((size_t)(2 * par) * N + sl) * pitch + x, a.r[(size_t)row * pitch + x], tag);
      }
    }
  }
  {
    const int sl1 = (a.Ex == 1 && !a.perx) ? 3 : 0;  // slot 0 holds class 0, or class 3 when that is the only one
    for (int q = t; q < 3 * TSZ; q += C::T) {
      const int sl = q / TSZ;
      sTx[q] = a.tab[(size_t)(sl == 0 ? sl1 : sl) * TSZ + (q - sl * TSZ)];
    }
  }
  // ---- mapping B: thread <-> (tile slot eB, tile column iB) --------------------------------------------------------
  const int eB = t / N2, iB = t - eB * N2;
  const bool inB = t < BX * N2;
  int exB = o0 - 1 + eB;
  const bool validB = inB && eB < nout + 2 && (a.perx || (exB >= 0 && exB < a.Ex));
  exB = (exB + a.Ex) % a.Ex;
  const bool hasL = exB > 0 || a.perx, hasR = exB < a.Ex - 1 || a.perx;
  int xB = -1;  // tile column -> global column: [left neighbour's node N-2, own 0..N-1, right neighbour's node 1]
  if (validB) {
    if (iB == 0) xB = hasL ? ((exB + a.Ex - 1) % a.Ex) * N + N - 2 : -1;
    else if (iB == N + 1) xB = hasR ? ((exB + 1) % a.Ex) * N + 1 : -1;
    else xB = exB * N + iB - 1;
  }
  const double wxB = (!FOLDW && xB >= 0) ? a.wx[xB] : 0.0;
  const bool outB = validB && eB >= 1 && eB <= nout && iB >= 1 && iB <= N;   // this thread finishes node column xB
  const int colB = eB * S + iB;
  const bool mzx = outB && ((xB == 0 && a.mx0) || (xB == a.nxl - 1 && a.mx1));
  const double mxB = (a.pcg && outB) ? a.mult_x[xB] : 0.0, mxBh = mxB * 0.5;
  // the tile column whose entries land on this node column as well (overlap with / x pair of the neighbour element)
  int srcCol = -1;
  if (outB) {
    const int i = iB - 1;
    if (i == 1 && hasL) srcCol = (eB - 1) * S + N + 1;
    else if (i == N - 2 && hasR) srcCol = (eB + 1) * S;
    else if (i == 0 && hasL) srcCol = (eB - 1) * S + N;
    else if (i == N - 1 && hasR) srcCol = (eB + 1) * S + 1;
  }
  const bool src3 = N == 3 && outB && iB == 2 && hasL && hasR;  // N = 3: node 1 lies in both neighbours' extensions
  const ptrdiff_t rd = a.r - a.out;
  // ---- mapping A: thread <-> (tile row jA, tile slot eA) -------------------------------------------------------------
  const int jA = t / BX, eA = t - jA * BX;
  int exA = o0 - 1 + eA;
  const bool validA = jA < N2 && eA < nout + 2 && (a.perx || (exA >= 0 && exA < a.Ex));
  exA = (exA + a.Ex) % a.Ex;
  const int clsA = fdm_class(exA, a.Ex, a.perx);
  const double* Tx = sTx + (clsA == 3 ? 0 : clsA) * TSZ;
  const int colA = jA * PW + eA * S;

  // one tile entry from outside the rows [N, nyl-N): local row, periodic image, neighbour rank's row (ghost) or nothing
  auto fetch_edge = [&](int Y, double* dst) {
    const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst);
    if (Y >= 0 && Y < a.nyl) fdm_cp_async8(d32, a.r + (size_t)Y * pitch + xB);
    else if (wrapy) fdm_cp_async8(d32, a.r + (size_t)((Y + a.nyl) % a.nyl) * pitch + xB);
    else if (Y < 0 && a.has_lo) {
      if (Y == -1) fdm_cp_async8(d32, a.r + xB);   // the shared line: the neighbour's last row = our row 0
      else {
        const unsigned long long epoch = cur_epoch();
        *dst = semb_ll_load(a.ghost + ((size_t)(2 * (int)(epoch & 1ull)) * N + (Y == -N - 2 ? 0 : Y + N + 1)) * pitch + xB,
                            semb_ll_tag(epoch), a.scal);
      }
    } else if (Y >= a.nyl && a.has_hi) {
      const int d = Y - a.nyl;
      if (d == 0) fdm_cp_async8(d32, a.r + (size_t)(a.nyl - 1) * pitch + xB);
      else {
        const unsigned long long epoch = cur_epoch();
        *dst = semb_ll_load(a.ghost + ((size_t)(2 * (int)(epoch & 1ull) + 1) * N + (d == N + 1 ? N - 1 : d - 1)) * pitch + xB,
                            semb_ll_tag(epoch), a.scal);
      }
    } else *dst = 0.0;
  };
  // an element row's tile columns: rows rr*N + {-2, 0..N-1, N+1} of this thread's global column
  auto prefetch = [&](int rr, double* buf) {
    if (inB) {
      double* dst = buf + colB;
      if (xB < 0) {
#pragma unroll
        for (int jj = 0; jj < N2; ++jj) dst[jj * PW] = 0.0;
      } else if (rr >= 1 && rr <= a.ney - 2) {
        const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst);
        const double* p = a.r + ((size_t)rr * N) * pitch + xB;
        fdm_cp_async8(d32, p - 2 * pitch);
#pragma unroll
        for (int jj = 1; jj <= N; ++jj) {
          fdm_cp_async8(d32 + jj * PW * 8, p);
          p += pitch;
        }
        fdm_cp_async8(d32 + (N + 1) * PW * 8, p + pitch);
      } else {
#pragma unroll
        for (int jj = 0; jj < N2; ++jj) fetch_edge(rr * N + (jj == 0 ? -2 : (jj == N + 1 ? N + 1 : jj - 1)), dst + jj * PW);
      }
    }
    fdm_cp_async_commit();
  };
  // tile rows of this CTA: q = 0 .. nq-1 <-> rr = ra + q (halo rows outside a non-periodic slab do not exist)
  const int ra = (r0 > 0 || lo_any) ? r0 - 1 : r0, rb = (r1 < a.ney || hi_any) ? r1 : r1 - 1;
  const int nq = rb - ra + 1;
  prefetch(ra, bufs);
  double el_n[3] = {1.0, 1.0, 0.0};
  auto load_el = [&](int rr) {
    if (validA) {
      const double* pe;
      long long st = a.nel;
      if (rr >= 0 && rr < a.ney) pe = a.el + (size_t)rr * a.Ex + exA;
      else if (wrapy) pe = a.el + (size_t)((rr + a.ney) % a.ney) * a.Ex + exA;
      else pe = (rr < 0 ? a.el_lo : a.el_hi) + exA, st = a.Ex;
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(el_n[0]) : "l"(pe));
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(el_n[1]) : "l"(pe + st));
      asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(el_n[2]) : "l"(pe + 2 * st));
    }
  };
  load_el(ra);
  __syncthreads();
  const bool cxA = P.cok && clsA == 0;   // interior element: even-odd tables from the constant bank
  int cy_loaded = -1;
  double pend0 = 0.0, pend1 = 0.0, prevN1 = 0.0;
  double acc = 0.0;

  for (int q = 0; q < nq; ++q) {
    const int rr = ra + q;
    double* S1 = bufs + (q & 1) * (N2 * PW);
    const int cy = fdm_class((a.ey0 + rr + a.Ey) % a.Ey, a.Ey, a.pery);
    if (cy != cy_loaded) {  // (block-uniform; at most three times per CTA)
      __syncthreads();
      for (int i = t; i < TSZ; i += C::T) sTy[i] = a.tab[(size_t)(4 + cy) * TSZ + i];
      cy_loaded = cy;
      __syncthreads();
    }
    const bool cyc = P.cok && cy == 0;   // interior element row (block-uniform)
    // element scalings 1/hx^2, 1/hy^2, 1/(hx*hy): loaded one element row ahead (volatile asm: the loads stay here)
    const double ihx2 = el_n[0], ihy2 = el_n[1], sc = el_n[2];
    if (q + 1 < nq) load_el(rr + 1);
    // ---- P1: (W .*) r, Sy' along y ----------------------------------------------------------------------------------
    fdm_cp_async_wait();
    if (inB) {
      double col[N2], o[N2];
#pragma unroll
      for (int jj = 0; jj < N2; ++jj) col[jj] = S1[jj * PW + colB];
      if (!FOLDW) {
#pragma unroll
        for (int jj = 0; jj < N2; ++jj) {
          int y = rr * N + (jj == 0 ? -2 : (jj == N + 1 ? N + 1 : jj - 1));   // (N = 3: single rank only)
          if (wrapy) y = (y + a.nyl) % a.nyl;
          col[jj] = (y >= 0 && y < a.nyl) ? __dmul_rn(__dmul_rn(wxB, a.wy[y]), col[jj]) : 0.0;
        }
      }
      if (cyc) fdm_fwd<N2, true>(P.tc[1], col, o);
      else fdm_fwd<N2, false>(sTy, col, o);
#pragma unroll
      for (int c = 0; c < N2; ++c) S1[c * PW + colB] = o[c];
    }
    __syncthreads();
    // the other buffer is free (its last readers were the P4 of the previous element row): next element row's tiles
    if (q + 1 < nq) prefetch(rr + 1, bufs + ((q + 1) & 1) * (N2 * PW));
    // ---- P2: Sx' along x, Di, Sx back ----------------------------------------------------------------------------------
    if (validA) {
      const double ly = sTy[TB::OFF_LAM + jA] * ihy2;
      double c[N2], o[N2];
#pragma unroll
      for (int i = 0; i < N2; ++i) c[i] = S1[colA + i];
      // Di = 1/(nu*(lx+ly)+k) (p2d_explicit.jl:131-134), times the 1/(hx*hy) of the two S/sqrt(h) pairs
      auto scale = [&](const double* T) {
#pragma unroll
        for (int m = 0; m < N2; ++m) {
          double d = fdm_rcp(fma(a.nu, fma(T[TB::OFF_LAM + m], ihx2, ly), a.k));
          if (!(fabs(d) <= 1e8)) d = 0.0;  // null mode of an all-free subdomain; padding modes (lambda = inf) give 0 too
          o[m] *= d * sc;
        }
      };
      if (cxA) {
        fdm_fwd<N2, true>(P.tc[0], c, o);
        scale(P.tc[0]);
        fdm_bwd<N2, true>(P.tc[0], o, c);
      } else {
        fdm_fwd<N2, false>(Tx, c, o);
        scale(Tx);
        fdm_bwd<N2, false>(Tx, o, c);
      }
#pragma unroll
      for (int i = 0; i < N2; ++i) S1[colA + i] = c[i];
    }
    __syncthreads();
    // ---- P3: Sy back along y; interface columns go back to the buffer ------------------------------------------------
    double g[N2];
    if (inB) {
      double c[N2];
#pragma unroll
      for (int m = 0; m < N2; ++m) c[m] = S1[m * PW + colB];
      if (cyc) fdm_bwd<N2, true>(P.tc[1], c, g);
      else fdm_bwd<N2, false>(sTy, c, g);
      if (iB <= 1 || iB >= N) {
#pragma unroll
        for (int jj = 0; jj < N2; ++jj) S1[jj * PW + colB] = g[jj];
      }
    }
    __syncthreads();
    // ---- P4: sums over the tiles that hold a node, mask, store ------------------------------------------------------
    if (!outB) continue;
    if (srcCol >= 0) {  // x: the neighbour tile's extension column, or the x pair
#pragma unroll
      for (int jj = 0; jj < N2; ++jj) g[jj] = __dadd_rn(g[jj], S1[jj * PW + srcCol]);
    }
    if (N == 3 && src3) {
#pragma unroll
      for (int jj = 0; jj < N2; ++jj) g[jj] = __dadd_rn(g[jj], S1[jj * PW + (eB + 1) * S]);
    }
    // h = mask .* (W .*) sum, and the pcg.jl:45 partial; my = the node's mult along y (structural: 1 or 1/2)
    auto write = [&](double* po, int row, double gv, bool zrow, bool half) {
      double h = gv;
      if (!FOLDW) h = __dmul_rn(__dmul_rn(wxB, a.wy[row]), gv);
      h = __dmul_rn((mzx || zrow) ? 0.0 : 1.0, h);
      *po = h;
      if (a.pcg) acc += __dmul_rn(__dmul_rn(po[rd], h), half ? mxBh : mxB);
    };
    if (rr < r0) {  // halo row below the chunk: only what the next element row needs of it
      pend0 = g[N - 1], pend1 = g[N], prevN1 = g[N + 1];
      continue;
    }
    const bool hb = rr > 0 || lo_any, ht = rr < a.ney - 1 || hi_any;
    const int ey = rr, ep = rr - 1;  // element rows of this and of the previous tile row (both inside the slab here)
    double* pp = a.out + ((size_t)ep * N + N - 2) * pitch + xB;
    if (rr >= r1) {  // halo row above the chunk: finishes the last two lines of the chunk
      write(pp, ep * N + N - 2, __dadd_rn(pend0, g[0]), false, false);
      write(pp + pitch, ep * N + N - 1, __dadd_rn(pend1, g[1]), false, true);
      continue;
    }
    double* pc = a.out + ((size_t)ey * N) * pitch + xB;
    const double val1 = hb ? __dadd_rn(g[2], prevN1) : g[2];  // line 1: own tile + the tile below's extension row
    if (hb) {
      const double s = __dadd_rn(pend1, g[1]);                // y pair (after the x pairs)
      if (rr > r0) {
        write(pp, ep * N + N - 2, __dadd_rn(pend0, g[0]), false, false);   // + this tile's extension row
        write(pp + pitch, ep * N + N - 1, s, false, true);
      }
      write(pc, ey * N, s, false, true);
    } else {
      write(pc, ey * N, g[1], a.my0 != 0, false);
    }
#pragma unroll
    for (int j = 1; j <= N - 3; ++j) {
      pc += pitch;
      write(pc, ey * N + j, j == 1 ? val1 : g[j + 1], false, false);
    }
    pend0 = (N == 3) ? val1 : g[N - 1], pend1 = g[N], prevN1 = g[N + 1];
    if (!ht) {
      write(pc + pitch, ey * N + N - 2, pend0, false, false);
      write(pc + 2 * pitch, ey * N + N - 1, pend1, a.my1 != 0, false);
    }
  }

  if (a.pcg) {
    const int bid = blockIdx.y * gridDim.x + blockIdx.x, nb = gridDim.x * gridDim.y;
    const double bs = semb_block_sum(acc, red, t, C::T);
    if (semb_last_block_uniform(bs, 0.0, a.partials, nullptr, a.counter, nb, bid, red, t, C::T, sh_tot)) {
      // t = sum(r .* h .* mult); norm(r,Inf) was left in red[2] by the init / update kernel: advance the PCG state
      SembScal* s = a.scal;
      double tnew = sh_tot[0], rmax = s->red[2];
      if (s->nranks > 1) semb_p2p_allgather(s, 1, tnew, rmax, &tnew, &rmax, t);  // combined in rank order over NVLink
      if (t == 0) {
        if (ranks) *a.ep_dev += 1ull;   // (every CTA read the epoch long ago)
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
                                        const double* __restrict__ ws, long long pitch, int N, int Ex, int ney, double* el) {
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
    const double hx = tx / (double)(N * N), hy = ty / (double)(N * N);
    const size_t nel = (size_t)Ex * ney;
    el[e] = 1.0 / (hx * hx);
    el[nel + e] = 1.0 / (hy * hy);
    el[2 * nel + e] = 1.0 / (hx * hy);
  }
}

// Launch geometry: strips of at most BX-2 output elements; the number of chunks minimises (waves of CTAs) x (element
// rows a CTA marches through, its two halo rows included)
template <int N>
int launch_fdm(semb_ctx* ctx, const FdmArgs& a, int npartials, const double* htab, int cok) {
  using C = FdmCfg<N>;
  auto kern = semb_fdm_kernel<N>;
  FdmParams<N> P;
  P.a = a;
  P.cok = cok;
  memcpy(P.tc[0], htab, sizeof(double) * C::TSZ);                          // x, interior class
  memcpy(P.tc[1], htab + (size_t)4 * C::TSZ, sizeof(double) * C::TSZ);     // y, interior class
  static bool attr_done[64] = {false};
  static int occ[64] = {0};
  const int dev = ctx->device & 63;
  if (!attr_done[dev]) {
    SEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    SEMB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[dev], kern, 256, C::SMEM));
    if (occ[dev] < 1) occ[dev] = 1;
    attr_done[dev] = true;
  }
  const int nstrips = (a.Ex + C::BX - 3) / (C::BX - 2);
  const long long slots = (long long)occ[dev] * ctx->sm_count;
  int best = 1;
  long long best_cost = -1;
  for (int nch = 1; nch <= a.ney && (long long)nch * nstrips <= npartials; ++nch) {
    const long long waves = ((long long)nstrips * nch + slots - 1) / slots;
    const long long cost = waves * ((a.ney + nch - 1) / nch + 2);
    if (best_cost < 0 || cost < best_cost) best_cost = cost, best = nch;
  }
  kern<<<dim3(nstrips, best), 256, C::SMEM, ctx->stream>>>(P);
  SEMB_CHECK_CUDA(cudaGetLastError());
  ctx->launches++;
  return SEMB_OK;
}

int fdm_slot_size(int N) {
  switch (N) {
#define SEMB_CASE(n) \
  case n:            \
    return FdmTab<n + 2>::SIZE;
    SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9) SEMB_CASE(10) SEMB_CASE(11)
    SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16) SEMB_CASE(17)
#undef SEMB_CASE
  }
  return 0;
}
void fdm_slot_fill(int N, const double* S, const double* lam, double* slot, bool allow_eo) {
  switch (N) {
#define SEMB_CASE(n)                                 \
  case n:                                            \
    FdmTab<n + 2>::fill(S, lam, slot, allow_eo);     \
    break;
    SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9) SEMB_CASE(10) SEMB_CASE(11)
    SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16) SEMB_CASE(17)
#undef SEMB_CASE
  }
}

}  // namespace

struct semb_fdm {
  semb_mesh* m = nullptr;
  double nu = 1.0, k = 0.0;
  int mx0 = 0, mx1 = 0, my0 = 0, my1 = 0;
  double *d_el = nullptr, *d_tab = nullptr, *d_wx = nullptr, *d_wy = nullptr;
  std::vector<double> h_tab;   // host copy of the tables (the interior class's travel as kernel parameters)
  int cok = 0;                 // the interior tables are even-odd in both directions
  // several ranks: ghost rows of r (IPC-exported, written by the neighbours), the neighbours' element scalings, epochs
  uint4* d_ghost = nullptr;
  void *peer_lo = nullptr, *peer_hi = nullptr;   // the neighbours' ghost rows as mapped into this process
  double *d_el_lo = nullptr, *d_el_hi = nullptr;
  unsigned long long* d_ep = nullptr;
  unsigned long long ep_host = 0;
  bool ranks_ready = false;   // the collective set-up completed (on every rank)
};

int semb_fdm_free_impl(semb_fdm* f) {
  if (!f) return SEMB_OK;
  cudaFree(f->d_el);
  cudaFree(f->d_tab);
  cudaFree(f->d_wx);
  cudaFree(f->d_wy);
  if (f->d_ghost) {
    // no neighbour may still be writing into (or mapping) the ghost rows (a set-up that failed did so on every rank, before
    // anyone could push: no barrier then)
    if (f->ranks_ready) semb_comm_barrier(f->m->ctx);
    if (f->peer_lo) cudaIpcCloseMemHandle(f->peer_lo);
    if (f->peer_hi && f->peer_hi != f->peer_lo) cudaIpcCloseMemHandle(f->peer_hi);
    cudaFree(f->d_ghost);
  }
  cudaFree(f->d_el_lo);
  cudaFree(f->d_el_hi);
  cudaFree(f->d_ep);
  delete f;
  return SEMB_OK;
}

// Several ranks: the halo tiles of a slab's first / last element row read N rows of r of the neighbour rank.  They arrive
// in `d_ghost`, an allocation exported through CUDA IPC like the mesh's mailbox, as flag-in-data entries; the scalings of
// the neighbours' adjoining element rows are exchanged once, here.  Collective over the communicator.
static int fdm_setup_ranks(semb_fdm* f) {
  semb_mesh* m = f->m;
  semb_ctx* c = m->ctx;
  const int N = m->nr, P = c->nranks, rk = c->rank;
  SEMB_REQUIRE(m->p2p, "fdm: on several ranks the preconditioner needs the peer-memory transport (CUDA IPC between the GPUs)");
  SEMB_REQUIRE(N >= 4, "fdm: on several ranks it needs nr >= 4");
  double small = m->ney < 2 ? 1.0 : 0.0;
  SEMB_TRY(semb_comm_allreduce_max(c, &small, 1));
  SEMB_REQUIRE(small == 0.0, "fdm: on several ranks every slab needs at least 2 element rows");
  const size_t gbytes = (size_t)4 * N * m->pitch * sizeof(uint4);
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_ghost, gbytes));
  SEMB_CHECK_CUDA(cudaMemset(f->d_ghost, 0, gbytes));
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_ep, sizeof(unsigned long long)));
  SEMB_CHECK_CUDA(cudaMemset(f->d_ep, 0, sizeof(unsigned long long)));
  // IPC handles of every rank's ghost rows, all-gathered through the communicator; only the neighbours' are mapped
  cudaIpcMemHandle_t mine;
  SEMB_CHECK_CUDA(cudaIpcGetMemHandle(&mine, f->d_ghost));
  char* d_h = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d_h, (size_t)P * sizeof(cudaIpcMemHandle_t)));
  SEMB_CHECK_CUDA(cudaMemcpy(d_h + (size_t)rk * sizeof(cudaIpcMemHandle_t), &mine, sizeof(mine), cudaMemcpyHostToDevice));
  SEMB_CHECK_NCCL(ncclAllGather(d_h + (size_t)rk * sizeof(cudaIpcMemHandle_t), d_h, sizeof(cudaIpcMemHandle_t), ncclChar, c->comm,
                                c->stream));
  std::vector<cudaIpcMemHandle_t> all(P);
  SEMB_CHECK_CUDA(cudaMemcpyAsync(all.data(), d_h, (size_t)P * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_h);
  cudaError_t eo = cudaSuccess;
  if (m->halo_lo) eo = cudaIpcOpenMemHandle(&f->peer_lo, all[m->rank_lo], cudaIpcMemLazyEnablePeerAccess);
  if (eo == cudaSuccess && m->halo_hi) {
    if (m->halo_lo && m->rank_hi == m->rank_lo) f->peer_hi = f->peer_lo;   // two ranks, periodic: one mapping
    else eo = cudaIpcOpenMemHandle(&f->peer_hi, all[m->rank_hi], cudaIpcMemLazyEnablePeerAccess);
  }
  if (eo != cudaSuccess) cudaGetLastError();
  // every rank must take the same path from here on (the exchanges below and the barrier in the destructor are collective)
  double bad = eo == cudaSuccess ? 0.0 : 1.0;
  SEMB_TRY(semb_comm_allreduce_max(c, &bad, 1));
  SEMB_REQUIRE(bad == 0.0, "fdm: mapping a neighbour rank's ghost rows (CUDA IPC) failed on some rank: %s", cudaGetErrorString(eo));
  // element scalings of the adjoining rows: our first row goes down, our last row up (receives in the opposite
  // order, so that two ranks that are each other's neighbour on both sides pair the messages correctly)
  const size_t Ex = (size_t)m->Ex, nel = Ex * m->ney;
  double* d_send = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d_send, 6 * Ex * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_el_lo, 3 * Ex * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_el_hi, 3 * Ex * sizeof(double)));
  for (int q = 0; q < 3; ++q) {
    SEMB_CHECK_CUDA(cudaMemcpyAsync(d_send + q * Ex, f->d_el + q * nel, Ex * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    SEMB_CHECK_CUDA(cudaMemcpyAsync(d_send + (3 + q) * Ex, f->d_el + q * nel + (size_t)(m->ney - 1) * Ex, Ex * sizeof(double),
                                    cudaMemcpyDeviceToDevice, c->stream));
  }
  SEMB_CHECK_NCCL(ncclGroupStart());
  if (m->halo_lo) SEMB_CHECK_NCCL(ncclSend(d_send, 3 * Ex, ncclDouble, m->rank_lo, c->comm, c->stream));
  if (m->halo_hi) SEMB_CHECK_NCCL(ncclSend(d_send + 3 * Ex, 3 * Ex, ncclDouble, m->rank_hi, c->comm, c->stream));
  if (m->halo_hi) SEMB_CHECK_NCCL(ncclRecv(f->d_el_hi, 3 * Ex, ncclDouble, m->rank_hi, c->comm, c->stream));
  if (m->halo_lo) SEMB_CHECK_NCCL(ncclRecv(f->d_el_lo, 3 * Ex, ncclDouble, m->rank_lo, c->comm, c->stream));
  SEMB_CHECK_NCCL(ncclGroupEnd());
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_send);
  SEMB_TRY(semb_comm_barrier(c));  // every rank's ghost rows exist and are mapped before anyone pushes
  f->ranks_ready = true;
  return SEMB_OK;
}

// bcflags: Dirichlet flags of the GLOBAL boundary lines (x0, x1, y0, y1) after the periodic override (parse_bc);
// bcglob: the same for the whole domain (this rank's slab may not touch the y boundaries)
int semb_fdm_create_impl(semb_mesh* m, double nu, double k, int mx0, int mx1, int my0, int my1, int gy0, int gy1,
                         semb_fdm** out) {
  semb_ctx* c = m->ctx;
  const int N = m->nr, N2 = N + 2;
  semb_fdm* f = new semb_fdm();
  *out = f;
  f->m = m;
  f->nu = nu;
  f->k = k;
  f->mx0 = mx0, f->mx1 = mx1, f->my0 = my0, f->my1 = my1;
  // reference decompositions: classes 0 interior, 1 first, 2 last, 3 single; kinds 0 neighbour, 1 Dirichlet, 2 free
  const int TSZ = fdm_slot_size(N);
  SEMB_REQUIRE(TSZ > 0, "fdm: no kernel for nr = %d (3..17)", N);
  std::vector<double> tab((size_t)2 * 4 * TSZ, 0.0), Sm((size_t)N2 * N2), lam(N2);
  for (int pass = 0; pass < 2; ++pass) {
  for (int dir = 0; dir < 2; ++dir) {
    const std::vector<double>& D = dir == 0 ? m->hDr : m->hDs;
    const std::vector<double>& w = dir == 0 ? m->hwr : m->hws;
    const int klo = dir == 0 ? (mx0 ? 1 : 2) : (gy0 ? 1 : 2), khi = dir == 0 ? (mx1 ? 1 : 2) : (gy1 ? 1 : 2);
    for (int cls = 0; cls < 4; ++cls) {
      SEMB_TRY(semb_fdm_tables(N, D.data(), w.data(), (cls & 1) ? klo : 0, (cls & 2) ? khi : 0, Sm.data(), lam.data()));
      if (N >= 4) {
        // counting weights W = 1/sqrt(number of subdomains holding the node) folded into the rows of S (both the
        // nodes -> modes and the modes -> nodes products carry one W): the two nodes on either side of an interface
        // lie in two subdomains along this direction (for N >= 4 that only depends on the class)
        const bool lo = !(cls & 1), hi = !(cls & 2);
        for (int i = 0; i < N2; ++i) {
          const bool two = (lo && i <= 2) || (hi && i >= N - 1);
          if (two)
            for (int c = 0; c < N2; ++c) Sm[i + (size_t)c * N2] *= std::sqrt(0.5);
        }
      }
      // the even-odd layout only for the interior class (read from the constant bank); decided below for both directions
      fdm_slot_fill(N, Sm.data(), lam.data(), tab.data() + ((size_t)dir * 4 + cls) * TSZ, cls == 0 && pass == 0);
    }
  }
  const bool cok = tab[0] != 0.0 && tab[(size_t)4 * TSZ] != 0.0;
  if (cok || pass == 1) {
    f->cok = cok ? 1 : 0;
    break;
  }
  }  // (second pass: an asymmetric D -- no even-odd split anywhere)
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_tab, tab.size() * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(f->d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  f->h_tab = tab;
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
  SEMB_CHECK_CUDA(cudaMalloc(&f->d_el, 3 * ne * sizeof(double)));
  double *d_wr = nullptr, *d_ws = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d_wr, N * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&d_ws, N * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(d_wr, m->hwr.data(), N * sizeof(double), cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMemcpy(d_ws, m->hws.data(), N * sizeof(double), cudaMemcpyHostToDevice));
  semb_fdm_lengths_kernel<<<(unsigned)ne, 64, 0, c->stream>>>(m->arr[SEMB_B], m->arr[SEMB_G11], m->arr[SEMB_G22], d_wr, d_ws,
                                                              m->pitch, N, m->Ex, m->ney, f->d_el);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_wr);
  cudaFree(d_ws);
  SEMB_CHECK_CUDA(e);
  c->launches++;
  if (c->nranks > 1) SEMB_TRY(fdm_setup_ranks(f));
  return SEMB_OK;
}

// the tables do not depend on the coefficients of nu*lapl + k*mass: only Di = 1/(nu*(lx+ly)+k) does, inside the kernel
int semb_fdm_set_coeffs_impl(semb_fdm* f, double nu, double k) {
  SEMB_REQUIRE(f && nu > 0.0 && k >= 0.0, "fdm: needs nu > 0, k >= 0");
  f->nu = nu;
  f->k = k;
  return SEMB_OK;
}

// h = opM(r); pcg: 0 stand-alone, 1 inside pcg (reduction + advance), 2 first call of a solve
int semb_fdm_apply_impl(semb_fdm* f, const double* r, double* out, int pcg) {
  semb_mesh* m = f->m;
  semb_ctx* c = m->ctx;
  FdmArgs a;
  a.r = r;
  a.out = out;
  a.tab = f->d_tab;
  a.el = f->d_el;
  a.nel = (long long)m->Ex * m->ney;
  a.wx = f->d_wx;
  a.wy = f->d_wy;
  a.mult_x = m->d_wx1d;
  a.mult_y = m->d_wy1d;
  a.pitch = m->pitch;
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
  a.has_lo = f->d_ghost ? m->halo_lo : 0;
  a.has_hi = f->d_ghost ? m->halo_hi : 0;
  a.ghost = f->d_ghost;
  a.peer_lo = (uint4*)f->peer_lo;
  a.peer_hi = (uint4*)f->peer_hi;
  a.el_lo = f->d_el_lo;
  a.el_hi = f->d_el_hi;
  a.ep_host = f->ep_host;
  a.ep_dev = f->d_ep;
  if (!pcg && f->d_ghost) f->ep_host++;   // (inside pcg the kernel's last CTA advances *ep_dev instead: graph replay)
  switch (m->nr) {
#define SEMB_CASE(n) \
  case n:            \
    SEMB_TRY(launch_fdm<n>(c, a, m->npartials, f->h_tab.data(), f->cok)); \
    break;
    SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9) SEMB_CASE(10) SEMB_CASE(11)
    SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16) SEMB_CASE(17)
#undef SEMB_CASE
    default:
      semb_set_error("fdm: no kernel for nr = %d (3..17)", m->nr);
      return SEMB_EINVAL;
  }
  return SEMB_OK;
}
