// Internal definitions of libsemb (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <functional>
#include <mutex>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/semb.h"

#define SEMB_MAXN 17       // largest nr == ns served by the templated strip kernel
// CTA size of the strip kernel: 256 threads (8 warps, 2 per SM sub-partition); 192 for N = 12..14 so that two
// CTAs still fit the 227 KB of shared memory (the staging buffers grow like N^2 per element)
#ifndef SEMB_T9
#define SEMB_T9 256
#endif
#ifndef SEMB_T12
#define SEMB_T12 192
#endif
#ifndef SEMB_T15
#define SEMB_T15 256
#endif
constexpr int semb_strip_threads(int n) { return (n >= 12 && n <= 14) ? SEMB_T12 : (n >= 15 ? SEMB_T15 : (n == 9 ? SEMB_T9 : 256)); }
// elements per strip: as many as fit the CTA, with BX*N even so that every staged row is a multiple of 16 bytes
constexpr int semb_strip_bx(int n) {
  return ((semb_strip_threads(n) / n) * n) % 2 == 0 ? semb_strip_threads(n) / n : semb_strip_threads(n) / n - 1;
}
// First element of strip s when Ex elements are dealt to nstrips strips of (nearly) equal width: every CTA of a wave
// then streams the same number of bytes (a fixed width of BX leaves a narrow last strip and BX-wide critical paths).
// For odd N the start is rounded down to an even element so that x0 = e0*N doubles stays 16-byte aligned (bulk copies).
__host__ __device__ inline int semb_strip_e0(int s, int nstrips, int Ex, int N) {
  if (s >= nstrips) return Ex;
  int e = (int)((long long)s * Ex / nstrips);
  if (N & 1) e &= ~1;
  return e;
}
#define SEMB_MAX_RANKS 16

void semb_set_error(const char* fmt, ...);

#define SEMB_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      semb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                              \
      return SEMB_ECUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define SEMB_CHECK_NCCL(expr)                                                         \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != ncclSuccess) {                                                          \
      semb_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, ncclGetErrorString(_r)); \
      return SEMB_ENCCL;                                                              \
    }                                                                                 \
  } while (0)

#define SEMB_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      semb_set_error(__VA_ARGS__);   \
      return SEMB_EINVAL;            \
    }                                \
  } while (0)

#define SEMB_TRY(expr)          \
  do {                          \
    int _rc = (expr);           \
    if (_rc < 0) return _rc;    \
  } while (0)

// Device-resident PCG / reduction scalars.  One per mesh.
struct SembScal {
  double t;        // sum(r .* h .* mult) of the current residual   (pcg.jl:45)
  double t_prev;   // previous t (the reference recomputes it, pcg.jl:49; same bits)
  double rmax;     // norm(r, Inf)                                    (pcg.jl:36)
  double pap[3];   // partial sum(p .* Ap .* mult): strip kernel, x-seam kernel, y-seam kernel
  double red[4];   // generic reduction results (dot / norm)
  double tol;
  long long iters;
  long long maxiter;
  int done;        // 1 once norm(r,Inf) <= tol or iters == maxiter
  int warned;      // 1 if stopped by maxiter (pcg.jl:39)
  // multi-GPU gathered scalars: slot r = rank r's local contribution
  double xchg_pap[SEMB_MAX_RANKS];
  double xchg_t[2 * SEMB_MAX_RANKS];  // {t_local, rmax_local} per rank
  double xchg_red[2 * SEMB_MAX_RANKS];
  int nranks, rank;
  double pap_total;  // multi-rank: combined sum(p .* Ap .* mult) of the current iteration
  // ---- peer-memory mailbox (multi-GPU, P2P mode): written by the OTHER ranks over NVLink -----------
  // flag_*[r] = epoch of the last value rank r pushed; data slots are double-buffered by epoch parity
  unsigned long long flag_halo[2];  // [0]: row from the lower neighbour, [1]: from the upper neighbour
  unsigned long long flag_pap[SEMB_MAX_RANKS];
  unsigned long long flag_t[SEMB_MAX_RANKS];
  unsigned long long flag_red[SEMB_MAX_RANKS];
  double box_pap[2][SEMB_MAX_RANKS];
  double box_t[2][2 * SEMB_MAX_RANKS];
  double box_red[2][2 * SEMB_MAX_RANKS];
  // ---- persistent device-side state (never overwritten by the host after mesh creation) ------------------
  // Epochs of the peer-memory exchanges live HERE, not in kernel arguments, so that a batch of PCG iterations
  // can be replayed as a CUDA graph on several ranks: [0] fused-tail halo, [1] pap, [2] {t, rmax}, [3] reductions
  unsigned long long ep_dev[4];
  SembScal* peers[SEMB_MAX_RANKS];  // peers[r] = rank r's mailbox as mapped into THIS process (r == rank: local)
  long long spin_limit;             // clock64 ticks a kernel may wait for a peer before it gives up (0 = forever)
  int err;                          // 1 once a peer wait timed out (surfaced as SEMB_ENCCL by the host)
};

// Switch + halo epoch handed to the kernels that exchange data through mapped peer memory (CUDA IPC); the peer
// pointers themselves and the epochs of the scalar all-gathers live in device memory (SembScal::peers, ep_dev).
struct P2PArgs {
  int on = 0, nranks = 1, rank = 0;
  unsigned long long epoch = 0;    // epoch of the stand-alone halo exchange this kernel consumes (host-side counter)
};

struct semb_ctx {
  std::recursive_mutex mutex;  // one in-flight call per context
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t in_stream = nullptr, out_stream = nullptr;  // H2D / D2H streams of the pipelined host twin
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long launches = 0;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double* flush_buf = nullptr;
  size_t flush_bytes = 0;
  double* d_sync = nullptr;  // 1-double scratch for barrier
  // optional per-kernel timing of the strip kernel (bench.py's roofline leg)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
  size_t prof_used = 0;
};

struct semb_field {
  semb_mesh* mesh = nullptr;
  double* d = nullptr;
};

struct semb_mesh {
  semb_ctx* ctx = nullptr;
  int nr = 0, ns = 0, Ex = 0, Ey = 0;  // Ey global
  int ey0 = 0, ney = 0;                // this rank's slab of element rows
  int perx = 0, pery = 0;
  int nxl = 0, nyl = 0;                // local node counts
  long long pitch = 0;                 // row pitch in doubles (multiple of 16)
  size_t nalloc = 0;                   // doubles per field allocation (pitch * nyl)
  std::vector<double> hDr, hDs, hwr, hws;  // column-major host copies
  double* dDr = nullptr;               // device copies (row-major: D[i*n+k] = D(i,k)) for generic kernels
  double* dDs = nullptr;
  double* arr[SEMB_MESH_ARRAY_COUNT] = {nullptr};
  double* d_wx1d = nullptr;            // separable mult: mult(x,y) = wx1d[x] * wy1d[y] (pitch / nyl entries)
  double* d_wy1d = nullptr;
  bool fast = false;                   // nr == ns in [2, SEMB_MAXN]: templated strip kernel
  bool eo = false;                     // Dr, Ds centro-antisymmetric: even-odd contraction variant
  // launch plan of the strip kernel
  int bx = 0;                          // elements per strip (semb_strip_bx(nr))
  int nstrips = 0, nchunks = 0;
  std::vector<int> h_chunk_r0;         // nchunks+1 element-row offsets
  int* d_chunk_r0 = nullptr;
  unsigned char* d_ystart = nullptr;   // ney+1 flags: element row r starts a y-seam (chunk / periodic / halo)
  std::vector<unsigned char> h_ystart;
  int nxseam = 0;                      // x seams (strip boundaries + periodic wrap): column pairs
  int* d_xseam = nullptr;              // 2*nxseam ints (xa, xb)
  int nyseam = 0;                      // local y seams (row pairs); halo seams handled separately
  int* d_yseam = nullptr;              // 2*nyseam ints (ya, yb)
  int halo_lo = 0, halo_hi = 0;        // 1 if the slab has a neighbour rank below / above
  int rank_lo = -1, rank_hi = -1;
  double* d_halo_lo = nullptr;         // nxl doubles each (received neighbour rows), NCCL path
  double* d_halo_hi = nullptr;
  // P2P path: one IPC-exported allocation per mesh = [SembScal][halo rows: 2 parities x 2 sides x pitch]
  bool p2p = false;
  void* d_mailbox = nullptr;
  double* d_mail_halo = nullptr;       // local halo rows inside the mailbox
  void* peer_mailbox[SEMB_MAX_RANKS] = {nullptr};
  unsigned long long ep_tail_host = 0; // fused-tail applies outside PCG issued so far (the PCG-mode count lives in ep_dev[0])
  unsigned long long ep_halo = 0;      // epoch of the stand-alone halo exchange (host-side; the others live in SembScal::ep_dev)
  // fused tail (semb_tail.cuh)
  bool tail = false;                   // interface completion fused into the strip kernel
  bool tail_plain = false;             // ... also for applies outside PCG (several ranks; one rank: the seam kernels are faster)
  int ngroups = 0;                     // grid.y of the strip kernel (CTA rows); a CTA row marches through 1 or 2 chunks
  int* d_grp = nullptr;                // 2*ngroups chunk ids
  unsigned* d_tcnt = nullptr;
  double* d_tpart = nullptr;
  int ntcnt = 0, ntpart = 0, xmic_total = 0;
  long long* d_dbg = nullptr;          // -DSEMB_TAIL_TIMING builds only
  // reductions
  int npartials = 0;
  double* d_partials = nullptr;        // 3 * npartials doubles
  unsigned* d_counters = nullptr;      // 8 tickets
  SembScal* d_scal = nullptr;
  SembScal* h_scal = nullptr;          // pinned mirror
  // pcg state
  semb_field* w_r = nullptr;
  semb_field* w_p = nullptr;
  semb_field* w_Ap = nullptr;
  semb_field* w_h = nullptr;           // h = opM(r) of the preconditioned PCG (written by init / update, staged by the strip kernel)
  semb_field* w_tmp = nullptr;
  semb_field* w_t1 = nullptr;
  semb_field* w_t2 = nullptr;
  semb_field* pcg_x = nullptr;
  semb_pcg_opts pcg_opts;
  bool pcg_active = false;
  bool host_pipe = false;              // semb_oplhs_host may take the pipelined path (agreed on by all ranks)
  bool pcg_keep_h = false;             // preconditioned PCG on the fused path: h kept in w_h
  // operator hook of the device-resident PCG: when set, an iteration is p = h + beta*p, w_Ap = pcg_custom(w_p),
  // sum(p.*Ap.*mult) by the reduction kernel, update -- instead of the fused strip kernel (Stokes Schur operator)
  std::function<int()> pcg_custom;
  struct semb_fdm* fdm = nullptr;      // FDM preconditioner registered on this mesh (semb_fdm_create), used by precond = 2
  char fdm_bc[4] = {0, 0, 0, 0};       // the boundary flags it was built for
  std::vector<semb_field*> fields;     // live fields (for leak-free destroy)
  std::vector<semb_field*> host_tmp;   // cached device fields of the *_host twins
};

// Arguments shared by the operator kernels (strip kernel, seam kernels, generic kernels).
struct OpArgs {
  const double* u = nullptr;      // input field (PCG mode: the residual r)
  const double* pold = nullptr;   // PCG mode: previous search direction
  double* pout = nullptr;         // PCG mode: new search direction p = h + beta*pold
  const double* G11 = nullptr;
  const double* G12 = nullptr;
  const double* G22 = nullptr;
  const double* B = nullptr;
  const double* nu_arr = nullptr;
  const double* k_arr = nullptr;
  const double* M_arr = nullptr;
  const double* mult = nullptr;
  double* out = nullptr;
  const double* halo_lo = nullptr;
  const double* halo_hi = nullptr;
  double nu = 1.0, k = 0.0, prec_b0 = 1.0;
  long long pitch = 0;
  int N = 0, Ex = 0, ney = 0, nxl = 0, nyl = 0;
  int perx = 0;
  int gs = 0;          // 1: fused QQ^T + mask; 0: local operator only
  int precond = 0;     // PCG mode: h = (r ./ B) ./ b0
  int mx0 = 0, mx1 = 0, my0 = 0, my1 = 0;  // Dirichlet flags that apply to THIS slab's boundary lines
  const int* chunk_r0 = nullptr;
  int nchunks = 0;
  int chunk0 = 0;               // first chunk of this launch (pipelined host twin: one slab of chunks per launch)
  int y_begin = 0, y_end = 0;   // row range of the x-seam kernel (0, 0 = all rows)
  const unsigned char* ystart = nullptr;
  const int* xseam = nullptr;
  int nxseam = 0;
  const int* yseam = nullptr;
  int nyseam = 0;
  SembScal* scal = nullptr;
  double* partials = nullptr;
  unsigned* counters = nullptr;
  int pcg = 0;         // 1: PCG mode (read scal, fuse p update and dot)
  // ---- fused tail (semb_tail.cuh): interface sums, halo exchange and the PCG dot finished INSIDE the strip kernel
  int tail = 0;                  // 1: seams are completed by the last CTA to arrive at each of them (no seam kernels)
  const int* grp = nullptr;      // 2 ints per blockIdx.y: the chunks that CTA row marches through, in order (-1: none)
  const int* xmic = nullptr;     // nchunks+1 prefix of ceil(lines of chunk / 32): micro-tasks of an x seam
  int xmic_total = 0;            // = xmic[nchunks]
  unsigned* tcnt = nullptr;      // arrival counters [x tasks][y tasks][corner tasks][final ticket]
  double* tpart = nullptr;       // PCG partial sums [CTA][x tasks][y tasks][corner tasks]
  int nxs = 0;                   // x seams (strip boundaries + the periodic wrap)
  int ywrap = 0;                 // periodic y closed inside this rank: boundary nchunks == boundary 0
  int has_lo = 0, has_hi = 0;    // neighbour ranks below / above
  unsigned long long ep_host = 0;  // halo exchanges of non-PCG applies issued so far (host-side part of the epoch)
  // boundary rows travel as 16-byte {lo32, tag, hi32, tag} entries (flag-in-data: no fence, no separate flag)
  uint4* peer_rows_lo = nullptr;   // neighbour-below's tail rows  [parity][side][pitch] (we write side 1)
  uint4* peer_rows_hi = nullptr;   // neighbour-above's tail rows  (we write side 0)
  const uint4* my_rows = nullptr;  // rows the neighbours wrote into OUR mailbox
  const double* wx1d = nullptr;  // mult(x,y) = wx1d[x] * wy1d[y] (factors 1 or 1/2)
  const double* wy1d = nullptr;
  long long* dbg = nullptr;      // -DSEMB_TAIL_TIMING builds: 8 globaltimer stamps per CTA (tools/tail_timing.py)
};

#ifdef SEMB_TAIL_TIMING
__device__ __forceinline__ void semb_stamp(const OpArgs& a, int i) {
  if (a.dbg && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + i] = t;
  }
}
#else
#define semb_stamp(a, i) ((void)0)
#endif

// Layout of the arrival counters / PCG partial slots of the fused tail (host: mesh_set_groups; device: semb_tail.cuh)
struct SembTailLayout {
  int nX, nY, nC;  // task counts: nxs*nchunks, (nchunks+1)*nstrips, (nchunks+1)*nxs
  int pX, pY;      // partial slots (one per MICRO-task of 32 items): nxs * xmic, nY * 8
  __host__ __device__ SembTailLayout(int nstrips, int nchunks, int nxs, int xmic)
      : nX(nxs * nchunks), nY((nchunks + 1) * nstrips), nC((nchunks + 1) * nxs), pX(nxs * xmic), pY(nY * 8) {}
  __host__ __device__ int offX() const { return 0; }
  __host__ __device__ int offY() const { return nX; }
  __host__ __device__ int offC() const { return nX + nY; }
  __host__ __device__ int ntasks() const { return nX + nY + nC; }
  __host__ __device__ int final_ticket() const { return nX + nY + nC; }
  __host__ __device__ int partX() const { return 0; }
  __host__ __device__ int partY() const { return pX; }
  __host__ __device__ int partC() const { return pX + pY; }
  __host__ __device__ int nparts() const { return pX + pY + nC; }
};

// FDM preconditioner (semb_fdm.cu)
struct semb_fdm;
int semb_fdm_create_impl(semb_mesh* m, double nu, double k, int mx0, int mx1, int my0, int my1, int gy0, int gy1,
                         semb_fdm** out);
int semb_fdm_free_impl(semb_fdm* f);
int semb_fdm_apply_impl(semb_fdm* f, const double* r, double* out, int pcg);
int semb_fdm_set_coeffs_impl(semb_fdm* f, double nu, double k);

// launchers implemented in the .cu files
int semb_launch_strip(semb_ctx* ctx, const OpArgs& a, const double* hDr, const double* hDs, int nstrips,
                      int nchunks, bool pcg, bool massterm, bool eo);
int semb_strip_regs(int N, bool pcg, bool massterm, int* regs, int* smem, int* occ);

// helpers
static inline long long semb_pitch_for(int nxl) { return ((long long)nxl + 15) / 16 * 16; }
