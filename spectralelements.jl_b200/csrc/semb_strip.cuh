// Fused strip kernel: out = mask(QQ^T(nu .* (D^T G D u) + k .* (B .* u)))  for one y-slab.
//
// Replaces (reference /root/reference/src): ABu.jl:9-37 (the four tensor contractions),
// lapl.jl:70-81 (laplace), mass.jl:17, hlmz.jl:15-16, the interior part of gatherScatter.jl:13
// and mask.jl:14, i.e. the body of opLHS (diffusion.jl:36-45).  In PCG mode it also performs
// pcg.jl:49 (u = hp + beta*u) on load and accumulates sum(u .* Au .* mult) (pcg.jl:52).
//
// Decomposition.  The field is the reference's global-lexicographic column-major matrix
// (x contiguous).  A CTA (256 threads) owns a strip of BX = 256/N x-consecutive elements and MARCHES
// through a chunk of element rows.  Two thread->data mappings alternate, exchanging data through
// shared memory (4 transposes per element row), so that every contraction runs out of registers with
// the D entries as constant-bank (uniform-register) operands -- no shared-memory traffic inside the
// N^2 FMA loops:
//   mapping B (y-lines): thread t <-> column x = e0*N+t of the strip; holds the N values of that
//                        column inside the current element row: coalesced global loads/stores,
//                        Ds contractions
//   mapping A (x-lines): thread p <-> (row j = p/BX, element e = p%BX); holds the N values of one
//                        x-line of one element: Dr contractions.  Shared memory is laid out [j][e*S+i]
//                        with S = N|1 odd, so both mappings are bank-conflict free.
// x-interfaces inside the strip are summed through a small smem exchange, y-interfaces inside the
// chunk through a register carried from one element row to the next (the write of an element
// row's last line is deferred by one iteration).  Strip/chunk boundary lines ("seams") are written
// un-summed and finished by the two seam kernels (semb_vec.cu).  All interface sums are 2-term
// (commutative) and x pairs are formed before y pairs, which reproduces the reference's
// (a+b)+(c+d) association bit for bit.
#pragma once
#include "semb_reduce.cuh"
#include "semb_tail.cuh"

// Contraction tables.  Four N x N matrices are applied per element row: A1 = Ds (y-lines of u),
// A2 = Dr (x-lines of u), A3 = Ds^T (y-lines of ws), A4 = Dr^T (x-lines of wr).  Each is shipped to the
// kernel as a table in the exact layout the contraction reads with broadcast LDS.128:
//   plain:    T[k][i] = A(i,k), rows padded to NP = N + (N&1)
//   even-odd: for centro-antisymmetric A (A(N-1-i,N-1-k) = -A(i,k), true of derivative matrices on
//             symmetric nodes) y = A x splits into two half-size products on e_k = x_k + x_{N-1-k} and
//             o_k = x_k - x_{N-1-k}:  y_i = Se_i + So_i, y_{N-1-i} = So_i - Se_i  (about half the FMAs
//             and half the table loads):  TP[k][i] = (A(i,k)+A(i,N-1-k))/2, TM[k][i] = (A(i,k)-A(i,N-1-k))/2
template <int N>
struct StripTab {
  static constexpr int NP = N + (N & 1);
  static constexpr int H = N / 2, ODD = N & 1;
  static constexpr int RP = (H + 1) & ~1;          // padded row of TP (H entries)
  static constexpr int RM = (H + ODD + 1) & ~1;    // padded row of TM (H + ODD entries)
  static constexpr int EOSIZE = (H + ODD) * RP + H * RM;
  static constexpr int SIZE = N * NP;              // >= EOSIZE
  // A is row-major A[i*N+k] = A(i,k)
  static void fill(const double* A, bool eo, double* T) {
    for (int q = 0; q < SIZE; ++q) T[q] = 0.0;
    if (!eo) {
      for (int k = 0; k < N; ++k)
        for (int i = 0; i < N; ++i) T[k * NP + i] = A[i * N + k];
      return;
    }
    double* TP = T;
    double* TM = T + (H + ODD) * RP;
    for (int k = 0; k < H; ++k)
      for (int i = 0; i < H; ++i) {
        TP[k * RP + i] = 0.5 * (A[i * N + k] + A[i * N + (N - 1 - k)]);
        TM[k * RM + i] = 0.5 * (A[i * N + k] - A[i * N + (N - 1 - k)]);
      }
    if (ODD) {
      for (int i = 0; i < H; ++i) TP[H * RP + i] = 0.5 * (A[i * N + H] - A[(N - 1 - i) * N + H]);  // mid column
      for (int k = 0; k < H; ++k) TM[k * RM + H] = 0.5 * (A[H * N + k] - A[H * N + (N - 1 - k)]);  // mid row
    }
  }
  // max_ik |A(i,k) + A(N-1-i,N-1-k)| / max|A|
  static double antisymmetry_defect(const double* A) {
    double d = 0.0, mx = 0.0;
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < N; ++k) {
        const double s = A[i * N + k] + A[(N - 1 - i) * N + (N - 1 - k)];
        d = s < 0 ? (-s > d ? -s : d) : (s > d ? s : d);
        const double v = A[i * N + k] < 0 ? -A[i * N + k] : A[i * N + k];
        mx = v > mx ? v : mx;
      }
    return mx > 0 ? d / mx : 0.0;
  }
};

template <int N>
struct StripParams {
  double tab[4][StripTab<N>::SIZE];  // A1 = Ds, A2 = Dr, A3 = Ds^T, A4 = Dr^T (StripTab layouts)
  OpArgs a;
};

// y = A x out of registers; T is the shared-memory table of A (StripTab layout), N independent accumulators
template <int N, bool EO>
__device__ __forceinline__ void semb_contract(const double* __restrict__ T, const double (&x)[N], double (&y)[N]) {
  using TB = StripTab<N>;
  if constexpr (!EO) {
    constexpr int NP = TB::NP;
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = (k == 0) ? T[i] * x[0] : fma(T[k * NP + i], x[k], y[i]);
    }
  } else {
    constexpr int H = TB::H, ODD = TB::ODD, RP = TB::RP, RM = TB::RM;
    const double* TP = T;
    const double* TM = T + (H + ODD) * RP;
    double e[H > 0 ? H : 1], o[H > 0 ? H : 1], se[H > 0 ? H : 1], so[H + ODD];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      e[k] = x[k] + x[N - 1 - k];
      o[k] = x[k] - x[N - 1 - k];
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
#pragma unroll
      for (int i = 0; i < H; ++i) se[i] = (k == 0) ? TP[i] * e[0] : fma(TP[k * RP + i], e[k], se[i]);
#pragma unroll
      for (int i = 0; i < H + ODD; ++i) so[i] = (k == 0) ? TM[i] * o[0] : fma(TM[k * RM + i], o[k], so[i]);
    }
    if (ODD) {
#pragma unroll
      for (int i = 0; i < H; ++i) se[i] = fma(TP[H * RP + i], x[H], se[i]);
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      y[i] = so[i] + se[i];
      y[N - 1 - i] = so[i] - se[i];
    }
    if (ODD) y[H] = so[H];
  }
}

template <int N>
struct StripCfg {
  static constexpr int T = semb_strip_threads(N);  // threads per CTA
  static constexpr int BX = semb_strip_bx(N);    // elements per strip (BX*N even: 16-byte bulk-copy rows)
  static constexpr int S = N | 1;                // element stride in smem (odd)
  static constexpr int PW = BX * S;              // row pitch of the transposition buffers S1/S2 (doubles)
  static constexpr int PWS = BX * N;             // row pitch of the TMA staging buffers (dense in x)
  static constexpr int NSTAGE = 4;               // u, G11, G12, G22
  static constexpr int SMEM_DOUBLES = N * PW + NSTAGE * N * PWS + N * 2 * BX + 4 * StripTab<N>::SIZE;
  static constexpr int SMEM = SMEM_DOUBLES * 8 + 64;
  static constexpr int MAXB_SMEM = (228 * 1024) / (SMEM + 1024);              // CTAs/SM the shared memory allows
  static constexpr int MAXB_REGS = 65536 / (T * 128) > 0 ? 65536 / (T * 128) : 1;  // ... at 128 registers per thread
  static constexpr int MINB0 = MAXB_SMEM < MAXB_REGS ? MAXB_SMEM : MAXB_REGS;
  static constexpr int MINB = (N <= 6 && 3 * SMEM + 3072 <= 228 * 1024 && T == 256) ? 3 : (MINB0 < 1 ? 1 : MINB0);
};

// ---- mbarrier / bulk-copy (TMA 1-D) PTX wrappers ------------------------------------------------------
__device__ __forceinline__ uint32_t semb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void semb_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(semb_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void semb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(semb_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void semb_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(semb_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (SASS: UBLKCP), completion signalled on an mbarrier as transaction bytes
__device__ __forceinline__ void semb_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   semb_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(semb_smem_u32(bar))
               : "memory");
}
// L2 residency hints.  The coefficient / input rows are read exactly once (evict_first: they should not displace
// anything), while the raw interface values a CTA leaves for the fused tail are re-read by another CTA up to a chunk
// later (evict_last: without it they come back from DRAM as scattered 32-byte sectors at the very end of the kernel)
__device__ __forceinline__ uint64_t semb_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t semb_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void semb_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                   uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          semb_smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(semb_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void semb_st_keep(double* p, double v) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(semb_policy_evict_last()) : "memory");
}
// global -> L2 bulk prefetch (no destination, no registers): rows a later plain load will hit in L2
__device__ __forceinline__ void semb_bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void semb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef SEMB_LATE_B
#define SEMB_LATE_B 1
#endif
// Where the contractions read their tables.  N <= 9: shared memory, broadcast LDS.128 (two entries per load).  N >= 10: the
// kernel-parameter constant bank -- LDCU.128 into uniform registers, DFMA with a UR operand: the table loads leave the
// shared-memory pipe, which at these sizes is the busiest unit (73 % of its cycles at N = 13, half of the wavefronts
// being table loads: profiles/r01_strip13_r1l.txt).  Measured (profiles/r02_sweep_tables_constant_bank_r2x.txt):
// N = 11: 70 -> 81 % of the HBM roof, N = 13: 63 -> 73 %, N = 9: 78 -> 79 % plain but 92 -> 85 % in the PCG variant.
// Helmholtz (plain apply with a mass term): the B rows of the NEXT element row go to L2 with the G rows (bulk prefetch), so
// that the early column load at the top of the row is an L2 hit.  Measured per N (profiles/r02_sweep_b_prefetch_r3p.txt,
// % of the 48-B roof): N = 6 76 -> 81, 8 76 -> 81, 9 72 -> 80, 10 76 -> 79; N = 7, 11, 12, 13 lose 1-3 points and keep the
// plain form.  -DSEMB_B_PREFETCH_ALL=1 forces it everywhere (the A/B build).
#ifndef SEMB_B_PREFETCH_ALL
#define SEMB_B_PREFETCH_ALL 0
#endif
#ifndef SEMB_TABC_MIN_N
#define SEMB_TABC_MIN_N 10
#endif
#define SEMB_TAB(q) (TABC ? (const double*)P.tab[q] : (const double*)(sT + (q) * TSZ))

template <int N, bool PCGM, bool MASS, bool EO>
__global__ void __launch_bounds__(StripCfg<N>::T, StripCfg<N>::MINB)
semb_strip_kernel(const __grid_constant__ StripParams<N> P) {
  using C = StripCfg<N>;
  constexpr int S = C::S, PW = C::PW, PWS = C::PWS, BX = C::BX, TSZ = StripTab<N>::SIZE;
  // where the B column of the mass term is loaded: at the top of the row (plain apply: 20 bytes of spills, latency
  // hidden behind steps 1-2) or in step 5 after an L2 prefetch (PCG variant, which also carries the p prefetch:
  // 184 -> 56 bytes of spills, 0.618 -> 0.590 ms per preconditioned iteration at 512x512 order 8; the plain
  // apply measured 4 % slower in the late form, so it keeps the early one)
  constexpr bool LATE_B = SEMB_LATE_B && PCGM;
  constexpr bool BPF = SEMB_B_PREFETCH_ALL || N == 6 || (N >= 8 && N <= 10);
  // (below 10 the A/B is mixed -- profiles/r02_sweep_tables_constant_bank_low_orders_r3d.txt: N = 6 gains in every
  // variant (79 -> 84 % plain), N = 7 and 9 only in the plain apply (79 -> 85 %, 78 -> 79 %), N = 3-5 and 8 lose)
  constexpr bool TABC = N >= SEMB_TABC_MIN_N || N == 6 || ((N == 7 || N == 9) && !PCGM && !MASS);
  extern __shared__ __align__(128) double smem[];
  // [N][PW] transposition buffer, updated IN PLACE by the alternating mappings (each phase touches
  // every location from exactly one thread): u -> Dr u -> wr -> Dr^T wr
  double* S1 = smem;
  double* SU = S1 + N * PW;             // [N][PWS] staged u rows (bulk copies)
  double* SG = SU + N * PWS;            // [3][N][PWS] staged G11, G12, G22 rows
  double* S3 = SG + 3 * N * PWS;        // [N][2*BX] x-interface exchange
  // contraction tables (StripTab layouts), fetched with broadcast LDS.128: kernel-parameter constants
  // end up as LDC + R2UR pairs, one pair per FMA, on sm_100 (profiles/r01_strip9_r1b.txt)
  double* sT = S3 + N * 2 * BX;         // [4][TSZ]
  uint64_t* bars = (uint64_t*)(sT + 4 * TSZ);  // bars[0]: u stage full, bars[1]: G stage full
  __shared__ double red[32];

  const OpArgs& a = P.a;
  double beta = 0.0;
  if (PCGM) {
    if (a.scal->done) return;
    beta = semb_pcg_beta(a.scal);
  }

  const int t = threadIdx.x;
  const int eB = t / N, iB = t - eB * N;        // mapping B: element / x-node of this thread's column
  const int colB = eB * S + iB;
  const int jA = t / BX, eA = t - jA * BX;      // mapping A: row / element of this thread's x-line
  const int colA = jA * PW + eA * S;
  const int e0 = semb_strip_e0(blockIdx.x, gridDim.x, a.Ex, N);   // balanced strips (plan: mesh_build_plan)
  const int nbe = semb_strip_e0(blockIdx.x + 1, gridDim.x, a.Ex, N) - e0;
  const bool actB = eB < nbe;                   // (eB < BX is implied: nbe <= BX)
  const bool inB = t < BX * N;                  // threads BX*N..T-1 own no column (T need not divide by N)
  const bool actA = (jA < N) && (eA < nbe);
  const int pitch = (int)a.pitch;
  const int x0 = e0 * N;
  const int xg = x0 + t;
  const bool gs = a.gs != 0;
  // roles in the gather-scatter
  const bool xl = gs && actB && iB == N - 1 && eB < nbe - 1;  // left side of an in-strip x interface
  const bool xr = gs && actB && iB == 0 && eB > 0;            // right side
  const bool xs = gs && actB &&
                  ((iB == 0 && eB == 0 && (e0 > 0 || a.perx)) ||
                   (iB == N - 1 && eB == nbe - 1 && (e0 + nbe < a.Ex || a.perx)));  // strip seam column
  const bool mzero = (xg == 0 && a.mx0) || (xg == a.nxl - 1 && a.mx1);  // Dirichlet column
  // mult = 1 ./ gatherScatter(ones) (mesh.jl:94-96) is structural: 1/(cx*cy)
  const double wx = (xl || xr || xs) ? 0.5 : 1.0;
  // Fused tail (semb_tail.cuh): this CTA row marches through one chunk -- or two: with a neighbour rank above, the
  // slab's LAST element row is a chunk of its own that the top CTA row computes FIRST (so that the boundary line is on
  // its way over NVLink for the whole kernel), followed without a pipeline bubble by the rows of its main chunk.
  // Afterwards the CTA announces its chunk(s) and runs the interface tasks it arrived last at.
  const bool tail = a.tail != 0;
  __shared__ SembTailTask s_task[16];
  const bool two = tail && a.grp[2 * blockIdx.y + 1] >= 0;   // edge row first, then the main chunk
  // epoch of this apply's halo exchange = applies issued by the host + PCG-mode applies completed on the device (the
  // latter count lives in device memory so that a captured batch of PCG iterations replays with unchanged arguments;
  // only the grid's very last CTA of a PCG-mode launch advances it)
  auto cur_ep = [&]() { return a.ep_host + *(volatile unsigned long long*)&a.scal->ep_dev[0] + 1ull; };
  if (tail && t < 16) s_task[t].n = 0, s_task[t].target = 0;

  // ---- bulk-copy producer (warp 0): one row = nbe*N doubles (rounded up to 16 bytes; the pad double
  // lies inside the padded row pitch), N rows per array, completion counted on an mbarrier ------------
  const uint32_t row_bytes = (uint32_t)(((nbe * N + 1) & ~1) * 8);
  auto issue_rows = [&](int r, int first_arr, int narr, double* stage, uint64_t* bar) {
    // called by all lanes of warp 0 (after a CTA barrier that retired every read of `stage`)
    if (t == 0) {
      semb_fence_proxy_async();
      semb_mbar_expect_tx(bar, row_bytes * (uint32_t)(narr * N));
    }
    __syncwarp();
#ifndef SEMB_NO_L2_HINTS
    const uint64_t pol = semb_policy_evict_first();
#endif
    for (int c = t; c < narr * N; c += 32) {
      const int q = c / N, j = c - q * N;
      const double* src = (first_arr + q == 0) ? a.u : (first_arr + q == 1) ? a.G11 : (first_arr + q == 2) ? a.G12 : a.G22;
#ifndef SEMB_NO_L2_HINTS
      semb_bulk_g2s_hint(stage + (q * N + j) * PWS, src + (size_t)(r * N + j) * pitch + x0, row_bytes, bar, pol);
#else
      semb_bulk_g2s(stage + (q * N + j) * PWS, src + (size_t)(r * N + j) * pitch + x0, row_bytes, bar);
#endif
    }
    // general coefficients: B is read late (step 5) with plain loads; its rows travel to L2 alongside the G rows
    if (MASS && (LATE_B || BPF) && first_arr == 1 && a.B && t < N)
      semb_bulk_prefetch_l2(a.B + (size_t)(r * N + t) * pitch + x0, row_bytes);
  };
  if (!TABC) {
    for (int q = t; q < 4 * TSZ; q += C::T) sT[q] = (&P.tab[0][0])[q];
  }
  if (t == 0) {
    semb_mbar_init(&bars[0], 1);
    semb_mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double carry = 0.0, carry_u = 0.0;
  double acc = 0.0;  // PCG: sum p*Ap*mult over the nodes this thread finalises
  double pn[N];      // PCG: register prefetch of the previous search direction, one element row ahead
  auto issue_p = [&](int r) {
    const int b = r * N * pitch + xg;
#pragma unroll
    for (int j = 0; j < N; ++j) pn[j] = actB ? a.pold[b + j * pitch] : 0.0;
  };
  semb_stamp(a, 0);
  const int cm = tail ? a.grp[2 * blockIdx.y + (two ? 1 : 0)] : a.chunk0 + (int)blockIdx.y;  // the (main) chunk
  const int r0 = a.chunk_r0[cm], r1 = a.chunk_r0[cm + 1];
  const int nrows = (r1 - r0) + (two ? 1 : 0);
  auto row_of = [&](int i) { return two ? (i == 0 ? a.ney - 1 : r0 + i - 1) : r0 + i; };
  const bool seam_bot_m = a.ystart[r0] != 0;  // main chunk's first / last line belongs to a y seam
  const bool seam_top_m = a.ystart[r1] != 0;
  // rank boundary: the raw boundary line also goes straight into the neighbour's mailbox over NVLink, from the
  // registers that hold it, as flag-in-data entries (semb_ll_store: no fence, no flag, nobody stalls); rows
  // [parity][side][pitch]: our first line is the lower neighbour's "row from above" (side 1), our last line the upper
  // neighbour's "row from below" (side 0)
  auto push_row = [&](bool lo, double val) {
    const unsigned long long ep = cur_ep();
    const int par = (int)(ep & 1ull);
    uint4* row = lo ? a.peer_rows_lo + (size_t)(2 * par + 1) * pitch : a.peer_rows_hi + (size_t)(2 * par) * pitch;
    semb_ll_store(row + xg, val, semb_ll_tag(ep));
  };
  if (t < 32 && nrows > 0) {
    issue_rows(row_of(0), 0, 1, SU, &bars[0]);
    issue_rows(row_of(0), 1, 3, SG, &bars[1]);
  }
  if (tail) semb_tail_prepare(a, (int)blockIdx.x, (int)gridDim.x, cm, two ? a.grp[2 * blockIdx.y] : -1, s_task);
  if (PCGM && nrows > 0) issue_p(row_of(0));

  // inactive threads read through clamped indices (no selects in the inner loops); they never store
  const int tr = inB ? t : 0;  // (the staging buffers are only written by the async proxy, between barriers)
  // partner column of an in-strip x interface (element stride S, N nodes per element)
  const int slotW = xl ? 2 * eB : 2 * (eB - 1) + 1, slotR = xl ? 2 * eB + 1 : 2 * (eB - 1);
  const bool xi = xl || xr;
  // nu .* (Dr^T wr + Ds^T ws) + k .* (B .* u)  (lapl.jl:78, hlmz.jl:15-16), un-fused like the reference
  auto combine = [&](double aur, double ausj, double nuj, double mj) {
    double lap = __dmul_rn(nuj, __dadd_rn(aur, ausj));
    if (MASS) lap = __dadd_rn(lap, mj);
    return lap;
  };

  for (int i = 0; i < nrows; ++i) {
    const int r = row_of(i);
    const int base = r * N * pitch + xg;
    const uint32_t parity = (uint32_t)(i & 1);  // both mbarriers complete one phase per element row
    const bool edge = two && i == 0;            // the one-row chunk [ney-1, ney): y seams on both sides
    const bool first = edge || r == r0, last = edge || r == r1 - 1;   // first / last row of its chunk
    const bool seam_bot = edge || seam_bot_m, seam_top = edge || seam_top_m;
    const bool plo = tail && a.has_lo && r == 0, phi = tail && a.has_hi && r == a.ney - 1;
    // coefficient columns that are not staged: issued now, consumed in step 3
    double bq[N];
    if (MASS && !LATE_B) {  // MASS = "general coefficients": k != 0, array k, or (non-constant) array nu
#pragma unroll
      for (int j = 0; j < N; ++j) bq[j] = (a.B && actB) ? a.B[base + j * pitch] : 0.0;
    }
    // ---- step 1 (B): the column of u (p in PCG mode), Ds contraction --------------------------------
    double u[N];
    semb_mbar_wait(&bars[0], parity);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v = SU[j * PWS + tr];
      if (PCGM && actB) {
        const int idx = base + j * pitch;
        if (a.precond) v = (v / ((MASS && !LATE_B) ? bq[j] : a.B[idx])) / a.prec_b0;  // convectionDiffusion.jl:89
        v = __dadd_rn(v, __dmul_rn(beta, pn[j]));        // pcg.jl:49
        a.pout[idx] = v;
      }
      u[j] = v;
      if (inB) S1[j * PW + colB] = v;
    }
    if (PCGM && i + 1 < nrows) issue_p(row_of(i + 1));
    double us[N];  // us = Ds * u along y: us[j] = sum_k Ds(j,k) u[k]
    semb_contract<N, EO>(SEMB_TAB(0), u, us);
    __syncthreads();
    if (t < 32 && i + 1 < nrows) issue_rows(row_of(i + 1), 0, 1, SU, &bars[0]);  // u stage is free: prefetch the next row
    // ---- step 2 (A): ur = Dr * u along x -------------------------------------------------------------
    if (actA) {
      double c[N], o[N];  // ur[m] = sum_i Dr(m,i) u[i]
#pragma unroll
      for (int i = 0; i < N; ++i) c[i] = S1[colA + i];
      semb_contract<N, EO>(SEMB_TAB(1), c, o);
#pragma unroll
      for (int m = 0; m < N; ++m) S1[colA + m] = o[m];
    }
    __syncthreads();
    // ---- step 3 (B): geometric factors, Ds^T contraction -------------------------------------------
    double aus[N], ws[N];
    semb_mbar_wait(&bars[1], parity);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double ur = inB ? S1[j * PW + colB] : 0.0;  // (threads BX*N..T-1 own no column)
      const double g11 = SG[(0 * N + j) * PWS + tr], g12 = SG[(1 * N + j) * PWS + tr], g22 = SG[(2 * N + j) * PWS + tr];
      const double wr = fma(g11, ur, g12 * us[j]);  // lapl.jl:75
      ws[j] = fma(g12, ur, g22 * us[j]);            // lapl.jl:76
      if (inB) S1[j * PW + colB] = wr;
    }
    semb_contract<N, EO>(SEMB_TAB(2), ws, aus);  // (Ds^T ws)[m] = sum_j Ds(j,m) ws[j]
    double mt[N];  // k .* (B .* u), hlmz.jl:16 / mass.jl:17
    if (MASS && !LATE_B) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double kk = (a.k_arr && actB) ? a.k_arr[base + j * pitch] : a.k;
        mt[j] = __dmul_rn(kk, __dmul_rn(bq[j], u[j]));
      }
    }
    __syncthreads();
    if (t < 32 && i + 1 < nrows) issue_rows(row_of(i + 1), 1, 3, SG, &bars[1]);  // G stage is free: prefetch the next row
    // ---- step 4 (A): Dr^T contraction ---------------------------------------------------------------
    if (actA) {
      double c[N], o[N];  // (Dr^T wr)[m] = sum_i Dr(i,m) wr[i]
#pragma unroll
      for (int i = 0; i < N; ++i) c[i] = S1[colA + i];
      semb_contract<N, EO>(SEMB_TAB(3), c, o);
#pragma unroll
      for (int m = 0; m < N; ++m) S1[colA + m] = o[m];
    }
    __syncthreads();
    // ---- step 5 (B): combine, hlmz, gather-scatter, mask, store -----------------------------------
    if (MASS && LATE_B) {
      // the B column is fetched here (L2 hits: prefetched with the G rows) instead of at the top of the row, where it
      // stayed live across steps 1-3 next to u, us, ws and the p prefetch (184 bytes of spills in the PCG variant)
#pragma unroll
      for (int j = 0; j < N; ++j) bq[j] = (a.B && actB) ? a.B[base + j * pitch] : 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double kk = (a.k_arr && actB) ? a.k_arr[base + j * pitch] : a.k;
        mt[j] = __dmul_rn(kk, __dmul_rn(bq[j], u[j]));
      }
    }
    double v[N];
#pragma unroll
    for (int j = 0; j < N; ++j)
      v[j] = combine(inB ? S1[j * PW + colB] : 0.0, aus[j],
                     (MASS && a.nu_arr && actB) ? a.nu_arr[base + j * pitch] : a.nu,  // rare: loaded in place
                     MASS ? mt[j] : 0.0);
    if (!gs) {
      if (actB) {
#pragma unroll
        for (int j = 0; j < N; ++j) a.out[base + j * pitch] = v[j];
      }
      continue;
    }
    // x pairs inside the strip: exchange the finished local values through S3
    if (xi) {
#pragma unroll
      for (int j = 0; j < N; ++j) S3[j * 2 * BX + slotW] = v[j];
    }
    __syncthreads();
    if (xi) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = __dadd_rn(v[j], S3[j * 2 * BX + slotR]);
    }
    if (!actB) continue;
    if (xs) {
      // strip seam column: raw values for every line; the x-seam kernel forms x pairs, then y pairs
#pragma unroll
      for (int j = 0; j < N; ++j) semb_st_keep(&a.out[base + j * pitch], v[j]);
      if (plo) push_row(true, v[0]);
      if (phi) push_row(false, v[N - 1]);
      continue;
    }
    // final value (+ mask, mask.jl:14) and the PCG dot contribution (pcg.jl:52) of one node;
    // multiplying by a mask of 1.0 or a mult of 1.0 is the identity, so it is skipped
    auto finish = [&](int idx, double val, double pval, bool zero_line, double wy) {
      double o = val;
      if (a.M_arr) o = __dmul_rn(a.M_arr[idx], val);
      else if (mzero || zero_line) o = __dmul_rn(0.0, val);
      a.out[idx] = o;
      if (PCGM) acc += __dmul_rn(__dmul_rn(pval, o), wx * wy);
    };
    // line 0
    if (first) {
      if (seam_bot) {
        semb_st_keep(&a.out[base], v[0]);
        if (plo) push_row(true, v[0]);
      } else finish(base, v[0], u[0], r == 0 && a.my0, 1.0);
    } else {
      const double s = __dadd_rn(carry, v[0]);  // y pair (after the x pairs)
      finish(base, s, u[0], false, 0.5);
      finish(base - pitch, s, carry_u, false, 0.5);  // deferred last line of the previous element row
    }
#pragma unroll
    for (int j = 1; j < N - 1; ++j) finish(base + j * pitch, v[j], u[j], false, 1.0);
    // line N-1
    if (last) {
      const int idx = base + (N - 1) * pitch;
      if (seam_top) {
        semb_st_keep(&a.out[idx], v[N - 1]);
        if (phi) push_row(false, v[N - 1]);
      } else finish(idx, v[N - 1], u[N - 1], r == a.ney - 1 && a.my1, 1.0);
    } else {
      carry = v[N - 1];
      if (PCGM) carry_u = u[N - 1];
    }
  }

  if (tail) {  // rows finished: fence this thread's stores, announce the chunk(s), run the tasks this CTA now owns
    semb_stamp(a, 1);
    __threadfence();
    __syncthreads();
    semb_tail_announce(a, s_task);
    semb_stamp(a, 2);
    __syncthreads();
    semb_strip_tail(a, s_task, (int)blockIdx.x, (int)gridDim.x, (a.has_lo || a.has_hi) ? cur_ep() : 0ull, acc, red);
    return;
  }
  if (PCGM) {
    const int nblocks = gridDim.x * gridDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    const double bs = semb_block_sum(acc, red, t, C::T);
    double total;
    if (semb_last_block(bs, 0.0, a.partials, nullptr, a.counters, nblocks, bid, red, t, C::T, &total, nullptr)) {
      a.scal->pap[0] = total;
    }
  }
}
