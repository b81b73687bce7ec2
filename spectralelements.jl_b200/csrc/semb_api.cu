// C-ABI entry points of libsemb.so (see include/semb.h).  Host-side orchestration only: contexts,
// meshes, fields, the launch plan of the fused operator, the device-resident PCG loop, NCCL plumbing.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <mutex>

#include <stddef.h>

#include "semb_internal.cuh"
#include "semb_vec.cuh"

#define SEMB_SCAL_PTR(m, member) ((double*)((char*)(m)->d_scal + offsetof(SembScal, member)))

// ---- error handling ---------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void semb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* semb_last_error(void) { return g_err; }
extern "C" int semb_version(void) { return 100; }

#define SEMB_NPARTIALS 4096
// bytes of SembScal the host may overwrite: everything before the peer-written mailbox part
#define SEMB_SCAL_HOST_BYTES offsetof(SembScal, flag_halo)

// ---- strip-kernel dispatch -----------------------------------------------------------------------------
#define SEMB_DECL_STRIP(n)                                                                                       \
  int semb_launch_strip_n##n(semb_ctx*, const OpArgs&, const double*, const double*, int, int, bool, bool, bool); \
  double semb_strip_defect_n##n(const double*, const double*);                                                    \
  int semb_strip_attr_n##n(bool, bool, int*, int*, int*);
SEMB_DECL_STRIP(2) SEMB_DECL_STRIP(3) SEMB_DECL_STRIP(4) SEMB_DECL_STRIP(5) SEMB_DECL_STRIP(6) SEMB_DECL_STRIP(7)
SEMB_DECL_STRIP(8) SEMB_DECL_STRIP(9) SEMB_DECL_STRIP(10) SEMB_DECL_STRIP(11) SEMB_DECL_STRIP(12)
SEMB_DECL_STRIP(13) SEMB_DECL_STRIP(14) SEMB_DECL_STRIP(15) SEMB_DECL_STRIP(16) SEMB_DECL_STRIP(17)

int semb_launch_strip(semb_ctx* ctx, const OpArgs& a, const double* hDr, const double* hDs, int nstrips,
                      int nchunks, bool pcg, bool massterm, bool eo) {
  switch (a.N) {
#define SEMB_CASE(n) \
  case n:            \
    return semb_launch_strip_n##n(ctx, a, hDr, hDs, nstrips, nchunks, pcg, massterm, eo);
    SEMB_CASE(2) SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9)
    SEMB_CASE(10) SEMB_CASE(11) SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16)
    SEMB_CASE(17)
#undef SEMB_CASE
  }
  semb_set_error("no strip kernel for N=%d", a.N);
  return SEMB_EINVAL;
}

double semb_strip_defect(int N, const double* hDr, const double* hDs) {
  switch (N) {
#define SEMB_CASE(n) \
  case n:            \
    return semb_strip_defect_n##n(hDr, hDs);
    SEMB_CASE(2) SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9)
    SEMB_CASE(10) SEMB_CASE(11) SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16)
    SEMB_CASE(17)
#undef SEMB_CASE
  }
  return 1.0;
}

int semb_strip_regs(int N, bool pcg, bool massterm, int* regs, int* smem, int* occ) {
  switch (N) {
#define SEMB_CASE(n) \
  case n:            \
    return semb_strip_attr_n##n(pcg, massterm, regs, smem, occ);
    SEMB_CASE(2) SEMB_CASE(3) SEMB_CASE(4) SEMB_CASE(5) SEMB_CASE(6) SEMB_CASE(7) SEMB_CASE(8) SEMB_CASE(9)
    SEMB_CASE(10) SEMB_CASE(11) SEMB_CASE(12) SEMB_CASE(13) SEMB_CASE(14) SEMB_CASE(15) SEMB_CASE(16)
    SEMB_CASE(17)
#undef SEMB_CASE
  }
  return SEMB_EINVAL;
}

extern "C" int semb_strip_kernel_info(int N, int pcg, int massterm, int* regs, int* smem, int* occ) {
  return semb_strip_regs(N, pcg != 0, massterm != 0, regs, smem, occ);
}

// ---- context ---------------------------------------------------------------------------------------------
extern "C" int semb_init(int device, semb_ctx** out) {
  SEMB_REQUIRE(out, "semb_init: null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    semb_set_error("semb_init: no CUDA device available (%s); libsemb has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    cudaGetLastError();
    return SEMB_ECUDA;
  }
  SEMB_REQUIRE(device >= 0 && device < ndev, "semb_init: device %d out of range (have %d)", device, ndev);
  SEMB_CHECK_CUDA(cudaSetDevice(device));
  semb_ctx* c = new semb_ctx();
  c->device = device;
  cudaDeviceProp prop;
  SEMB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  SEMB_CHECK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  SEMB_CHECK_CUDA(cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking));
  SEMB_CHECK_CUDA(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
  SEMB_CHECK_CUDA(cudaEventCreate(&c->ev0));
  SEMB_CHECK_CUDA(cudaEventCreate(&c->ev1));
  SEMB_CHECK_CUDA(cudaMalloc(&c->d_sync, 64));
  SEMB_CHECK_CUDA(cudaMemset(c->d_sync, 0, 64));
  *out = c;
  return SEMB_OK;
}

static int ctx_enter(semb_ctx* c) {
  SEMB_REQUIRE(c, "null context");
  SEMB_CHECK_CUDA(cudaSetDevice(c->device));
  return SEMB_OK;
}

// A context is not re-entrant: one in-flight call per ctx (callable from any host thread).  The lock is
// recursive because entry points are composed of each other.
struct CtxGuard {
  std::unique_lock<std::recursive_mutex> lock;
  explicit CtxGuard(semb_ctx* c) {
    if (c) lock = std::unique_lock<std::recursive_mutex>(c->mutex);
  }
};
#define SEMB_ENTER(c)       \
  CtxGuard _semb_guard(c);  \
  SEMB_TRY(ctx_enter(c))

extern "C" int semb_finalize(semb_ctx* c) {
  if (!c) return SEMB_OK;
  SEMB_TRY(ctx_enter(c));  // no guard: the context (and its mutex) is destroyed here
  cudaStreamSynchronize(c->stream);
  if (c->comm) ncclCommDestroy(c->comm);
  if (c->flush_buf) cudaFree(c->flush_buf);
  if (c->d_sync) cudaFree(c->d_sync);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->in_stream);
  cudaStreamDestroy(c->out_stream);
  delete c;
  return SEMB_OK;
}

extern "C" int semb_sync(semb_ctx* c) {
  SEMB_ENTER(c);
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return SEMB_OK;
}

extern "C" int semb_stream(semb_ctx* c, void** s) {
  SEMB_REQUIRE(c && s, "semb_stream: null argument");
  *s = (void*)c->stream;
  return SEMB_OK;
}

extern "C" int semb_timer_start(semb_ctx* c) {
  SEMB_ENTER(c);
  SEMB_CHECK_CUDA(cudaEventRecord(c->ev0, c->stream));
  return SEMB_OK;
}

extern "C" int semb_timer_stop(semb_ctx* c, double* ms) {
  SEMB_ENTER(c);
  SEMB_CHECK_CUDA(cudaEventRecord(c->ev1, c->stream));
  SEMB_CHECK_CUDA(cudaEventSynchronize(c->ev1));
  float f = 0.f;
  SEMB_CHECK_CUDA(cudaEventElapsedTime(&f, c->ev0, c->ev1));
  if (ms) *ms = (double)f;
  return SEMB_OK;
}

extern "C" int semb_launch_count(semb_ctx* c, long long* n) {
  SEMB_REQUIRE(c && n, "semb_launch_count: null argument");
  *n = c->launches;
  return SEMB_OK;
}

extern "C" int semb_flush_l2(semb_ctx* c) {
  SEMB_ENTER(c);
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    SEMB_CHECK_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  SEMB_CHECK_CUDA(cudaMemsetAsync(c->flush_buf, 0, c->flush_bytes, c->stream));
  return SEMB_OK;
}

extern "C" int semb_profile_enable(semb_ctx* c, int max_launches) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(max_launches >= 0 && max_launches <= 65536, "semb_profile_enable: 0..65536 launches");
  while (c->prof_ev.size() < (size_t)2 * max_launches) {
    cudaEvent_t e;
    SEMB_CHECK_CUDA(cudaEventCreate(&e));
    c->prof_ev.push_back(e);
  }
  c->prof_used = 0;
  c->profile = max_launches > 0;
  return SEMB_OK;
}

extern "C" int semb_profile_read(semb_ctx* c, double* total_ms, int* launches) {
  SEMB_ENTER(c);
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
    float f = 0.f;
    SEMB_CHECK_CUDA(cudaEventElapsedTime(&f, c->prof_ev[i], c->prof_ev[i + 1]));
    tot += f;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int)(c->prof_used / 2);
  c->prof_used = 0;
  return SEMB_OK;
}

extern "C" int semb_alloc_pinned(size_t bytes, void** p) {
  SEMB_REQUIRE(p, "semb_alloc_pinned: null");
  SEMB_CHECK_CUDA(cudaMallocHost(p, bytes));
  return SEMB_OK;
}

extern "C" int semb_free_pinned(void* p) {
  if (p) SEMB_CHECK_CUDA(cudaFreeHost(p));
  return SEMB_OK;
}

// ---- NCCL plumbing ---------------------------------------------------------------------------------------
extern "C" int semb_comm_unique_id(char id[128]) {
  SEMB_REQUIRE(id, "semb_comm_unique_id: null");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  SEMB_CHECK_NCCL(ncclGetUniqueId(&u));
  memcpy(id, &u, 128);
  return SEMB_OK;
}

extern "C" int semb_comm_init(semb_ctx* c, int nranks, int rank, const char id[128]) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(nranks >= 1 && nranks <= SEMB_MAX_RANKS && rank >= 0 && rank < nranks,
               "semb_comm_init: bad nranks/rank %d/%d", nranks, rank);
  SEMB_REQUIRE(!c->comm, "semb_comm_init: communicator already initialised");
  if (nranks > 1) {
    ncclUniqueId u;
    memcpy(&u, id, 128);
    SEMB_CHECK_NCCL(ncclCommInitRank(&c->comm, nranks, u, rank));
  }
  c->nranks = nranks;
  c->rank = rank;
  return SEMB_OK;
}

extern "C" int semb_comm_info(semb_ctx* c, int* nranks, int* rank) {
  SEMB_REQUIRE(c, "null context");
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  return SEMB_OK;
}

extern "C" int semb_comm_barrier(semb_ctx* c) {
  SEMB_ENTER(c);
  if (c->comm) SEMB_CHECK_NCCL(ncclAllReduce(c->d_sync, c->d_sync, 1, ncclDouble, ncclSum, c->comm, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return SEMB_OK;
}

extern "C" int semb_comm_allreduce_max(semb_ctx* c, double* v, int n) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(v && n >= 1 && n <= 7, "semb_comm_allreduce_max: 1..7 values");
  if (!c->comm) return SEMB_OK;
  SEMB_CHECK_CUDA(cudaMemcpyAsync(c->d_sync + 1, v, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SEMB_CHECK_NCCL(ncclAllReduce(c->d_sync + 1, c->d_sync + 1, n, ncclDouble, ncclMax, c->comm, c->stream));
  SEMB_CHECK_CUDA(cudaMemcpyAsync(v, c->d_sync + 1, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return SEMB_OK;
}

// ---- mesh --------------------------------------------------------------------------------------------------
static int mesh_alloc_array(semb_mesh* m, int which) {
  if (m->arr[which]) return SEMB_OK;
  SEMB_CHECK_CUDA(cudaMalloc(&m->arr[which], m->nalloc * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemsetAsync(m->arr[which], 0, m->nalloc * sizeof(double), m->ctx->stream));
  return SEMB_OK;
}

static int upload_pitched(semb_mesh* m, double* dst, const double* host) {
  SEMB_CHECK_CUDA(cudaMemcpy2DAsync(dst, m->pitch * sizeof(double), host, (size_t)m->nxl * sizeof(double),
                                    (size_t)m->nxl * sizeof(double), m->nyl, cudaMemcpyHostToDevice,
                                    m->ctx->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));  // never retain a host pointer after return
  return SEMB_OK;
}

static int download_pitched(semb_mesh* m, const double* src, double* host) {
  SEMB_CHECK_CUDA(cudaMemcpy2DAsync(host, (size_t)m->nxl * sizeof(double), src, m->pitch * sizeof(double),
                                    (size_t)m->nxl * sizeof(double), m->nyl, cudaMemcpyDeviceToHost,
                                    m->ctx->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return SEMB_OK;
}

// layout of a mesh's mailbox allocation (exported through CUDA IPC):
//   [SembScal][stand-alone halo rows: 2 parities x 2 sides x pitch doubles]
//   [fused-tail rows: 2 parities x 2 sides x pitch 16-byte flag-in-data entries (semb_ll_store)]
static size_t mail_scal_bytes() { return (sizeof(SembScal) + 255) / 256 * 256; }
static size_t mail_bytes(semb_mesh* m) {
  return mail_scal_bytes() + 4 * (size_t)m->pitch * sizeof(double) + 4 * (size_t)m->pitch * sizeof(uint4);
}
static uint4* mail_tail_rows(semb_mesh* m, void* mailbox) {
  return (uint4*)((char*)mailbox + mail_scal_bytes() + 4 * (size_t)m->pitch * sizeof(double));
}

// Chunk tables of the strip kernel for `groups` CTA rows: chunk offsets, y-seam flags, seam lists, and the tables
// of the fused tail (semb_tail.cuh).  With a neighbour rank above, the slab's last element row becomes a chunk of
// its own that the top CTA row marches through FIRST, and the two edge CTA rows are scheduled first, so that both
// boundary rows are on their way to the neighbours while the rest of the slab is computed.
static int mesh_set_groups(semb_mesh* m, int groups) {
  semb_ctx* c = m->ctx;
  const int N = m->ns, P = c->nranks;
  const bool wrap_local = m->pery && P == 1;
  if (groups > m->ney) groups = m->ney;
  if (groups < 1) groups = 1;
  std::vector<int> g0(groups + 1);
  // (rounded up: the LAST group is the short one -- with a neighbour above it is the CTA row with the extra edge chunk)
  for (int k = 0; k <= groups; ++k) g0[k] = (int)(((long long)k * m->ney + groups - 1) / groups);
  const bool split = m->tail && m->halo_hi && g0[groups] - g0[groups - 1] >= 2 && !getenv("SEMB_NO_EDGE_FIRST");
  m->ngroups = groups;
  m->nchunks = groups + (split ? 1 : 0);
  m->h_chunk_r0.assign(g0.begin(), g0.end());
  if (split) m->h_chunk_r0.insert(m->h_chunk_r0.end() - 1, m->ney - 1);
  std::vector<int> grp(2 * (size_t)groups, -1);
  {
    std::vector<int> order(groups);  // CTA row -> group: the top group first when it feeds a neighbour
    for (int g = 0; g < groups; ++g) order[g] = g;
    if (m->tail && m->halo_hi && groups > 1) {
      order.pop_back();
      order.insert(order.begin(), groups - 1);
    }
    for (int by = 0; by < groups; ++by) {
      const int g = order[by];
      if (split && g == groups - 1) {
        grp[2 * by] = groups;  // the one-row chunk [ney-1, ney)
        grp[2 * by + 1] = groups - 1;
      } else {
        grp[2 * by] = g;
      }
    }
  }
  // y seam flags per element row
  m->h_ystart.assign(m->ney + 1, 0);
  for (int k = 1; k < m->nchunks; ++k) m->h_ystart[m->h_chunk_r0[k]] = 1;
  m->h_ystart[0] = (m->halo_lo || wrap_local) ? 1 : 0;
  m->h_ystart[m->ney] = (m->halo_hi || wrap_local) ? 1 : 0;
  std::vector<int> ys;
  for (int k = 1; k < m->nchunks; ++k) {
    ys.push_back(m->h_chunk_r0[k] * N - 1);
    ys.push_back(m->h_chunk_r0[k] * N);
  }
  if (wrap_local) {
    ys.push_back(m->nyl - 1);
    ys.push_back(0);
  }
  m->nyseam = (int)ys.size() / 2;
  // all y interfaces (stand-alone gatherScatter), appended after the chunk seams
  for (int r = 1; r < m->ney; ++r) {
    ys.push_back(r * N - 1);
    ys.push_back(r * N);
  }
  if (wrap_local) {
    ys.push_back(m->nyl - 1);
    ys.push_back(0);
  }
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(m->d_chunk_r0);
  cudaFree(m->d_yseam);
  cudaFree(m->d_ystart);
  cudaFree(m->d_grp);
  cudaFree(m->d_tcnt);
  cudaFree(m->d_tpart);
  m->d_chunk_r0 = nullptr, m->d_yseam = nullptr, m->d_ystart = nullptr, m->d_grp = nullptr, m->d_tcnt = nullptr,
  m->d_tpart = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_chunk_r0, (m->nchunks + 1) * sizeof(int)));
  SEMB_CHECK_CUDA(cudaMemcpy(m->d_chunk_r0, m->h_chunk_r0.data(), (m->nchunks + 1) * sizeof(int),
                             cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_ystart, m->ney + 1));
  SEMB_CHECK_CUDA(cudaMemcpy(m->d_ystart, m->h_ystart.data(), m->ney + 1, cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_yseam, std::max<size_t>(ys.size(), 2) * sizeof(int)));
  if (!ys.empty())
    SEMB_CHECK_CUDA(cudaMemcpy(m->d_yseam, ys.data(), ys.size() * sizeof(int), cudaMemcpyHostToDevice));
  // micro-tasks (32 lines each) of an x seam per chunk, as a prefix, appended to the group table
  for (int k = 0, acc = 0; k <= m->nchunks; ++k) {
    grp.push_back(acc);
    if (k < m->nchunks) acc += ((m->h_chunk_r0[k + 1] - m->h_chunk_r0[k]) * N + 31) / 32;
  }
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_grp, grp.size() * sizeof(int)));
  SEMB_CHECK_CUDA(cudaMemcpy(m->d_grp, grp.data(), grp.size() * sizeof(int), cudaMemcpyHostToDevice));
  m->xmic_total = grp.back();
  const SembTailLayout L(m->nstrips, m->nchunks, m->nstrips - 1 + m->perx, m->xmic_total);
  m->ntcnt = L.ntasks() + 1;
  m->ntpart = m->nstrips * m->ngroups + L.nparts();
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_tcnt, (size_t)m->ntcnt * sizeof(unsigned)));
  SEMB_CHECK_CUDA(cudaMemset(m->d_tcnt, 0, (size_t)m->ntcnt * sizeof(unsigned)));
#ifdef SEMB_TAIL_TIMING
  if (!m->d_dbg) {
    SEMB_CHECK_CUDA(cudaMalloc(&m->d_dbg, 8 * (size_t)SEMB_NPARTIALS * sizeof(long long)));
    SEMB_CHECK_CUDA(cudaMemset(m->d_dbg, 0, 8 * (size_t)SEMB_NPARTIALS * sizeof(long long)));
  }
#endif
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_tpart, (size_t)m->ntpart * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemset(m->d_tpart, 0, (size_t)m->ntpart * sizeof(double)));
  return SEMB_OK;
}

// Launch plan of the fused operator: strips x chunks, seam lists, halo neighbours.
static int mesh_build_plan(semb_mesh* m) {
  semb_ctx* c = m->ctx;
  m->fast = (m->nr == m->ns && m->nr >= 2 && m->nr <= SEMB_MAXN);
  m->bx = m->fast ? semb_strip_bx(m->nr) : 32;
  // derivative matrices on symmetric nodes are centro-antisymmetric: the even-odd kernel variant applies
  m->eo = m->fast && semb_strip_defect(m->nr, m->hDr.data(), m->hDs.data()) < 1e-12 && !getenv("SEMB_NO_EVENODD");
  m->nstrips = (m->Ex + m->bx - 1) / m->bx;
  if (m->fast) {  // balanced strips (semb_strip_e0): the fewest strips whose widest member still fits the CTA
    for (;; ++m->nstrips) {
      int widest = 0;
      for (int s = 0; s < m->nstrips; ++s)
        widest = std::max(widest, semb_strip_e0(s + 1, m->nstrips, m->Ex, m->nr) - semb_strip_e0(s, m->nstrips, m->Ex, m->nr));
      if (widest <= m->bx) break;
    }
  }
  // halo neighbours
  const int P = c->nranks, rk = c->rank;
  SEMB_TRY(semb_halo_plan(P, rk, m->pery, &m->halo_lo, &m->halo_hi, &m->rank_lo, &m->rank_hi));
  // x seams (strip boundaries + the periodic wrap): column pairs
  std::vector<int> xs;
  for (int s = 1; s < m->nstrips; ++s) {
    const int e = m->fast ? semb_strip_e0(s, m->nstrips, m->Ex, m->nr) : s * m->bx;
    xs.push_back(e * m->nr - 1);
    xs.push_back(e * m->nr);
  }
  if (m->perx) {
    xs.push_back(m->nxl - 1);
    xs.push_back(0);
  }
  m->nxseam = (int)xs.size() / 2;
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_xseam, std::max<size_t>(xs.size(), 2) * sizeof(int)));
  if (!xs.empty())
    SEMB_CHECK_CUDA(cudaMemcpy(m->d_xseam, xs.data(), xs.size() * sizeof(int), cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_halo_lo, (size_t)m->pitch * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_halo_hi, (size_t)m->pitch * sizeof(double)));
  m->npartials = SEMB_NPARTIALS;
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_partials, 3 * (size_t)m->npartials * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_counters, 8 * sizeof(unsigned)));
  SEMB_CHECK_CUDA(cudaMemset(m->d_counters, 0, 8 * sizeof(unsigned)));
  // Mailbox (layout above).  With more than one rank it is exported through CUDA IPC so that the neighbours'
  // kernels can store into it over NVLink (P2P mode).
  SEMB_CHECK_CUDA(cudaMalloc(&m->d_mailbox, mail_bytes(m)));
  SEMB_CHECK_CUDA(cudaMemset(m->d_mailbox, 0, mail_bytes(m)));
  m->d_scal = (SembScal*)m->d_mailbox;
  m->d_mail_halo = (double*)((char*)m->d_mailbox + mail_scal_bytes());
  m->p2p = false;
  m->peer_mailbox[rk] = m->d_mailbox;
  if (P > 1 && m->fast && !getenv("SEMB_NO_P2P")) {
    // exchange the IPC handles with the communicator itself, then map every peer's mailbox
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, m->d_mailbox);
    int ok = (e == cudaSuccess) ? 1 : 0;
    if (!ok) cudaGetLastError();
    char* d_h = nullptr;
    SEMB_CHECK_CUDA(cudaMalloc(&d_h, (size_t)P * sizeof(cudaIpcMemHandle_t)));
    SEMB_CHECK_CUDA(cudaMemcpy(d_h + (size_t)rk * sizeof(cudaIpcMemHandle_t), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    SEMB_CHECK_NCCL(ncclAllGather(d_h + (size_t)rk * sizeof(cudaIpcMemHandle_t), d_h, sizeof(cudaIpcMemHandle_t), ncclChar,
                                  c->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all(P);
    SEMB_CHECK_CUDA(cudaMemcpyAsync(all.data(), d_h, (size_t)P * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost,
                                    c->stream));
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_h);
    for (int r = 0; r < P && ok; ++r) {
      if (r == rk) continue;
      e = cudaIpcOpenMemHandle(&m->peer_mailbox[r], all[r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
      }
    }
    // every rank must take the same path: agree on success
    double flag = ok ? 0.0 : 1.0;
    SEMB_TRY(semb_comm_allreduce_max(c, &flag, 1));
    m->p2p = (flag == 0.0);
    if (!m->p2p)
      for (int r = 0; r < P; ++r)
        if (r != rk && m->peer_mailbox[r]) {
          cudaIpcCloseMemHandle(m->peer_mailbox[r]);
          m->peer_mailbox[r] = nullptr;
        }
  }
  SEMB_CHECK_CUDA(cudaMallocHost(&m->h_scal, sizeof(SembScal)));
  memset(m->h_scal, 0, sizeof(SembScal));
  m->h_scal->nranks = P;
  m->h_scal->rank = rk;
  m->h_scal->done = 1;
  SEMB_CHECK_CUDA(cudaMemcpy(m->d_scal, m->h_scal, SEMB_SCAL_HOST_BYTES, cudaMemcpyHostToDevice));
  {  // persistent part: the peer table and the bound of every wait on a peer
    SembScal* peers[SEMB_MAX_RANKS] = {nullptr};
    for (int r = 0; r < P; ++r) peers[r] = (SembScal*)m->peer_mailbox[r];
    SEMB_CHECK_CUDA(cudaMemcpy((char*)m->d_scal + offsetof(SembScal, peers), peers, sizeof(peers), cudaMemcpyHostToDevice));
    double ms = 20000.0;
    if (const char* e = getenv("SEMB_PEER_TIMEOUT_MS")) ms = atof(e);
    cudaDeviceProp prop;
    SEMB_CHECK_CUDA(cudaGetDeviceProperties(&prop, c->device));
    const long long ticks = ms > 0 ? (long long)(ms * (double)prop.clockRate) : 0;  // clockRate is in kHz = ticks per ms
    SEMB_CHECK_CUDA(cudaMemcpy((char*)m->d_scal + offsetof(SembScal, spin_limit), &ticks, sizeof(ticks), cudaMemcpyHostToDevice));
  }
  // interface completion inside the strip kernel (one launch per apply): single rank, or peer memory between ranks
  m->tail = m->fast && (P == 1 || m->p2p) && !getenv("SEMB_NO_TAIL");
  int occ = 1;
  if (m->fast) SEMB_TRY(semb_strip_regs(m->nr, false, false, nullptr, nullptr, &occ));
  if (occ < 1) occ = 1;
  // chunks: fill the resident CTA slots; the rule (waves x (rows of the longest chunk + 1/2), measured on chunk sweeps:
  // 256x256 order 8 takes 29 chunks = one full wave, 59.4 us per apply, instead of 59 = two waves of 4-5 rows, 66.5 us;
  // 1112x1112 keeps 22) lives in semb_plan_chunks (semb_host.cpp) so that it is testable without a GPU
  const int slots = c->sm_count * occ;
  int best = 1;
  SEMB_TRY(semb_plan_chunks(m->nstrips, m->ney, slots, &best));
  if (best > m->ney) best = m->ney;
  if ((long long)(best + 1) * m->nstrips > SEMB_NPARTIALS) best = SEMB_NPARTIALS / m->nstrips - 1;
  if (best < 1) best = 1;
  // One rank, several waves of CTAs: the separate seam kernels are 2-3 % faster than the one-launch form (the tasks read
  // the interface values back through L2 while the next wave streams; profiles/r02_ab_tail_r2h.txt: 771 vs 788 us per
  // apply at 1112x1112 order 8, 2012 vs 2023 us per PCG iteration).  Everything that has launches or exchanges to
  // save -- one wave (small meshes), several ranks -- keeps the one-launch form.  SEMB_FORCE_TAIL=1 keeps it always.
  if (m->tail && P == 1 && (long long)m->nstrips * best >= 2LL * slots && !getenv("SEMB_FORCE_TAIL")) m->tail = false;
  // One rank, plain applies (no PCG reduction to save): the seam kernels also win on a single wave of CTAs (256x256
  // order 8 Helmholtz, L2 flushed: 69.8 vs 73.7 us), the PCG iteration does not (156.5 vs 160.0 us) -- so on one rank
  // the one-launch form serves the PCG-mode applies only
  m->tail_plain = m->tail && (P > 1 || getenv("SEMB_FORCE_TAIL"));
  SEMB_TRY(mesh_set_groups(m, best));
  // pipelined host twin (semb_oplhs_host) on several ranks: every rank must take the same path (its halo exchange differs
  // from the one-launch apply's), so the size criterion is agreed on once, here
  m->host_pipe = (size_t)m->nxl * m->nyl >= ((size_t)1 << 22);
  if (P > 1) {
    double small = m->host_pipe ? 0.0 : 1.0;
    SEMB_TRY(semb_comm_allreduce_max(c, &small, 1));
    m->host_pipe = (small == 0.0);
  }
  if (P > 1) SEMB_TRY(semb_comm_barrier(c));  // every mailbox is mapped and initialised before anyone pushes
  return SEMB_OK;
}

static int mesh_new_impl(semb_ctx* c, int nr, int ns, int Ex, int Ey, int perx, int pery, const double* Dr,
                    const double* Ds, const double* wr, const double* ws, semb_mesh** out) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(out, "mesh: null output");
  SEMB_REQUIRE(nr >= 2 && ns >= 2 && nr <= 64 && ns <= 64, "mesh: need 2 <= nr, ns <= 64 (got %d, %d)", nr, ns);
  SEMB_REQUIRE(Ex >= 1 && Ey >= 1, "mesh: need Ex, Ey >= 1");
  SEMB_REQUIRE(c->nranks <= Ey, "mesh: more ranks (%d) than element rows (%d)", c->nranks, Ey);
  SEMB_REQUIRE((long long)nr * Ex < (1ll << 30) && (long long)ns * Ey < (1ll << 30), "mesh: too large");
  semb_mesh* m = new semb_mesh();
  *out = m;  // handed to the caller at once: mesh_new destroys it if anything below fails
  m->ctx = c;
  m->nr = nr;
  m->ns = ns;
  m->Ex = Ex;
  m->Ey = Ey;
  m->perx = perx ? 1 : 0;
  m->pery = pery ? 1 : 0;
  SEMB_TRY(semb_partition(Ey, c->nranks, c->rank, &m->ey0, &m->ney));
  m->nxl = nr * Ex;
  m->nyl = ns * m->ney;
  m->pitch = semb_pitch_for(m->nxl);
  m->nalloc = (size_t)m->pitch * m->nyl;
  if (m->nalloc >= ((size_t)1 << 31)) {
    semb_set_error("mesh: %zu doubles per field on one GPU exceeds the 2^31 index range of the kernels; use more ranks", m->nalloc);
    return SEMB_EINVAL;
  }
  m->hDr.assign((size_t)nr * nr, 0.0);
  m->hDs.assign((size_t)ns * ns, 0.0);
  m->hwr.assign(nr, 0.0);
  m->hws.assign(ns, 0.0);
  if (Dr) {
    std::copy(Dr, Dr + (size_t)nr * nr, m->hDr.begin());
  } else {
    std::vector<double> z(nr), w(nr);
    SEMB_TRY(semb_gausslobatto(nr, z.data(), w.data()));
    SEMB_TRY(semb_deriv_mat(nr, z.data(), m->hDr.data()));
  }
  if (Ds) {
    std::copy(Ds, Ds + (size_t)ns * ns, m->hDs.begin());
  } else {
    std::vector<double> z(ns), w(ns);
    SEMB_TRY(semb_gausslobatto(ns, z.data(), w.data()));
    SEMB_TRY(semb_deriv_mat(ns, z.data(), m->hDs.data()));
  }
  {
    std::vector<double> z(std::max(nr, ns)), w(std::max(nr, ns));
    if (wr) std::copy(wr, wr + nr, m->hwr.begin());
    else {
      SEMB_TRY(semb_gausslobatto(nr, z.data(), w.data()));
      std::copy(w.begin(), w.begin() + nr, m->hwr.begin());
    }
    if (ws) std::copy(ws, ws + ns, m->hws.begin());
    else {
      SEMB_TRY(semb_gausslobatto(ns, z.data(), w.data()));
      std::copy(w.begin(), w.begin() + ns, m->hws.begin());
    }
  }
  // row-major device copies for the generic kernels
  std::vector<double> rm((size_t)nr * nr);
  for (int i = 0; i < nr; ++i)
    for (int k = 0; k < nr; ++k) rm[(size_t)i * nr + k] = m->hDr[i + (size_t)k * nr];
  SEMB_CHECK_CUDA(cudaMalloc(&m->dDr, rm.size() * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(m->dDr, rm.data(), rm.size() * sizeof(double), cudaMemcpyHostToDevice));
  rm.assign((size_t)ns * ns, 0.0);
  for (int i = 0; i < ns; ++i)
    for (int k = 0; k < ns; ++k) rm[(size_t)i * ns + k] = m->hDs[i + (size_t)k * ns];
  SEMB_CHECK_CUDA(cudaMalloc(&m->dDs, rm.size() * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(m->dDs, rm.data(), rm.size() * sizeof(double), cudaMemcpyHostToDevice));
  SEMB_TRY(mesh_build_plan(m));
  SEMB_TRY(mesh_alloc_array(m, SEMB_MULT));
  {  // 1-D factors of mult (mesh.jl:94-96): 1/2 on duplicated lines, 1 elsewhere; pads are 0
    std::vector<double> wx((size_t)m->pitch, 0.0), wy((size_t)m->nyl, 1.0);
    for (int x = 0; x < m->nxl; ++x) {
      const int e = x / nr, i = x % nr;
      const bool d = (i == nr - 1 && e < Ex - 1) || (i == 0 && e > 0) || (m->perx && (x == 0 || x == m->nxl - 1));
      wx[x] = d ? 0.5 : 1.0;
    }
    for (int y = 0; y < m->nyl; ++y) {
      const int rg = m->ey0 + y / ns, j = y % ns;
      const bool d = (j == ns - 1 && rg < Ey - 1) || (j == 0 && rg > 0) ||
                     (m->pery && ((rg == 0 && j == 0) || (rg == Ey - 1 && j == ns - 1)));
      wy[y] = d ? 0.5 : 1.0;
    }
    SEMB_CHECK_CUDA(cudaMalloc(&m->d_wx1d, wx.size() * sizeof(double)));
    SEMB_CHECK_CUDA(cudaMalloc(&m->d_wy1d, wy.size() * sizeof(double)));
    SEMB_CHECK_CUDA(cudaMemcpy(m->d_wx1d, wx.data(), wx.size() * sizeof(double), cudaMemcpyHostToDevice));
    SEMB_CHECK_CUDA(cudaMemcpy(m->d_wy1d, wy.data(), wy.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  SEMB_TRY(semb_launch_mult(c, m->arr[SEMB_MULT], m->pitch, nr, ns, Ex, Ey, m->ey0, m->ney, m->perx, m->pery));
  return SEMB_OK;
}

static int mesh_new(semb_ctx* c, int nr, int ns, int Ex, int Ey, int perx, int pery, const double* Dr,
                    const double* Ds, const double* wr, const double* ws, semb_mesh** out) {
  SEMB_REQUIRE(out, "mesh: null output");
  *out = nullptr;
  const int rc = mesh_new_impl(c, nr, ns, Ex, Ey, perx, pery, Dr, Ds, wr, ws, out);
  if (rc < 0 && *out) {
    semb_mesh_destroy(*out);
    *out = nullptr;
  }
  return rc;
}


static int mesh_geometry(semb_mesh* m) {
  semb_ctx* c = m->ctx;
  double *d_wr = nullptr, *d_ws = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d_wr, m->nr * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&d_ws, m->ns * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpyAsync(d_wr, m->hwr.data(), m->nr * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SEMB_CHECK_CUDA(cudaMemcpyAsync(d_ws, m->hws.data(), m->ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  for (int w = SEMB_JAC; w <= SEMB_G22; ++w) SEMB_TRY(mesh_alloc_array(m, w));
  SEMB_TRY(semb_launch_geom(c, m->arr[SEMB_X], m->arr[SEMB_Y], m->pitch, m->nr, m->ns, m->Ex, m->ney, m->dDr,
                            m->dDs, d_wr, d_ws, m->arr[SEMB_JAC], m->arr[SEMB_JACI], m->arr[SEMB_RX],
                            m->arr[SEMB_RY], m->arr[SEMB_SX], m->arr[SEMB_SY], m->arr[SEMB_B], m->arr[SEMB_BI],
                            m->arr[SEMB_G11], m->arr[SEMB_G12], m->arr[SEMB_G22]));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_wr);
  cudaFree(d_ws);
  return SEMB_OK;
}

extern "C" int semb_mesh_create_xy(semb_ctx* c, int nr, int ns, int Ex, int Ey, int perx, int pery,
                                   const double* Dr, const double* Ds, const double* wr, const double* ws,
                                   const double* x, const double* y, semb_mesh** out) {
  SEMB_REQUIRE(x && y, "semb_mesh_create_xy: x and y are required");
  semb_mesh* m = nullptr;
  SEMB_TRY(mesh_new(c, nr, ns, Ex, Ey, perx, pery, Dr, Ds, wr, ws, &m));
  int rc = mesh_alloc_array(m, SEMB_X);
  if (!rc) rc = mesh_alloc_array(m, SEMB_Y);
  if (!rc) rc = upload_pitched(m, m->arr[SEMB_X], x);
  if (!rc) rc = upload_pitched(m, m->arr[SEMB_Y], y);
  if (!rc) rc = mesh_geometry(m);
  if (rc) {
    semb_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return SEMB_OK;
}

extern "C" int semb_mesh_create_deform(semb_ctx* c, int nr, int ns, int Ex, int Ey, int perx, int pery, int kind,
                                       const double* params, int nparams, semb_mesh** out) {
  SEMB_REQUIRE(kind >= SEMB_DEFORM_IDENTITY && kind <= SEMB_DEFORM_WAVY, "semb_mesh_create_deform: unknown kind %d",
               kind);
  double p[3] = {0.0, 0.0, 0.0};
  if (kind == SEMB_DEFORM_ANNULUS) {
    p[0] = 0.5;
    p[1] = 1.0;
    p[2] = 2.0 * 3.14159265358979323846;
  }
  if (kind == SEMB_DEFORM_WAVY) p[0] = 0.1;
  for (int i = 0; i < nparams && i < 3; ++i) p[i] = params[i];
  semb_mesh* m = nullptr;
  SEMB_TRY(mesh_new(c, nr, ns, Ex, Ey, perx, pery, nullptr, nullptr, nullptr, nullptr, &m));
  std::vector<double> z0r(nr), z0s(ns), w(std::max(nr, ns));
  semb_gausslobatto(nr, z0r.data(), w.data());
  semb_gausslobatto(ns, z0s.data(), w.data());
  for (auto& v : z0r) v = 0.5 * (v + 1.0);  // semmesh.jl:13
  for (auto& v : z0s) v = 0.5 * (v + 1.0);
  double *dzr = nullptr, *dzs = nullptr;
  int rc = SEMB_OK;
  auto body = [&]() -> int {
    SEMB_CHECK_CUDA(cudaMalloc(&dzr, nr * sizeof(double)));
    SEMB_CHECK_CUDA(cudaMalloc(&dzs, ns * sizeof(double)));
    SEMB_CHECK_CUDA(cudaMemcpy(dzr, z0r.data(), nr * sizeof(double), cudaMemcpyHostToDevice));
    SEMB_CHECK_CUDA(cudaMemcpy(dzs, z0s.data(), ns * sizeof(double), cudaMemcpyHostToDevice));
    SEMB_TRY(mesh_alloc_array(m, SEMB_X));
    SEMB_TRY(mesh_alloc_array(m, SEMB_Y));
    SEMB_TRY(semb_launch_grid(c, m->arr[SEMB_X], m->arr[SEMB_Y], m->pitch, nr, ns, Ex, Ey, m->ey0, m->ney, dzr, dzs,
                              kind, p));
    SEMB_TRY(mesh_geometry(m));
    return SEMB_OK;
  };
  rc = body();
  if (dzr) cudaFree(dzr);
  if (dzs) cudaFree(dzs);
  if (rc) {
    semb_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return SEMB_OK;
}

extern "C" int semb_mesh_create_arrays(semb_ctx* c, int nr, int ns, int Ex, int Ey, int perx, int pery,
                                       const double* Dr, const double* Ds, const double* G11, const double* G12,
                                       const double* G22, const double* B, semb_mesh** out) {
  SEMB_REQUIRE(Dr && Ds && G11 && G12 && G22, "semb_mesh_create_arrays: Dr, Ds, G11, G12, G22 are required");
  semb_mesh* m = nullptr;
  SEMB_TRY(mesh_new(c, nr, ns, Ex, Ey, perx, pery, Dr, Ds, nullptr, nullptr, &m));
  const double* src[4] = {G11, G12, G22, B};
  const int which[4] = {SEMB_G11, SEMB_G12, SEMB_G22, SEMB_B};
  for (int i = 0; i < 4; ++i) {
    if (!src[i]) continue;
    int rc = mesh_alloc_array(m, which[i]);
    if (!rc) rc = upload_pitched(m, m->arr[which[i]], src[i]);
    if (rc) {
      semb_mesh_destroy(m);
      return rc;
    }
  }
  *out = m;
  return SEMB_OK;
}

extern "C" int semb_mesh_destroy(semb_mesh* m) {
  if (!m) return SEMB_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  std::vector<semb_field*> fs = m->fields;
  for (semb_field* f : fs) semb_field_destroy(f);
  semb_fdm_free_impl(m->fdm);
  m->fdm = nullptr;
  for (int i = 0; i < SEMB_MESH_ARRAY_COUNT; ++i)
    if (m->arr[i]) cudaFree(m->arr[i]);
  cudaFree(m->dDr);
  cudaFree(m->dDs);
  cudaFree(m->d_wx1d);
  cudaFree(m->d_wy1d);
  cudaFree(m->d_chunk_r0);
  cudaFree(m->d_ystart);
  cudaFree(m->d_grp);
  cudaFree(m->d_tcnt);
  cudaFree(m->d_tpart);
  cudaFree(m->d_dbg);
  cudaFree(m->d_xseam);
  cudaFree(m->d_yseam);
  cudaFree(m->d_halo_lo);
  cudaFree(m->d_halo_hi);
  cudaFree(m->d_partials);
  cudaFree(m->d_counters);
  if (m->p2p) {
    semb_comm_barrier(m->ctx);  // no peer may still be writing into (or mapping) this mailbox
    for (int r = 0; r < m->ctx->nranks; ++r)
      if (r != m->ctx->rank && m->peer_mailbox[r]) cudaIpcCloseMemHandle(m->peer_mailbox[r]);
    semb_comm_barrier(m->ctx);
  }
  cudaFree(m->d_mailbox);
  if (m->h_scal) cudaFreeHost(m->h_scal);
  delete m;
  return SEMB_OK;
}

extern "C" int semb_mesh_dims(semb_mesh* m, int* nr, int* ns, int* Ex, int* Ey, int* nxl, int* nyl, int* ey0,
                              int* ney) {
  SEMB_REQUIRE(m, "null mesh");
  if (nr) *nr = m->nr;
  if (ns) *ns = m->ns;
  if (Ex) *Ex = m->Ex;
  if (Ey) *Ey = m->Ey;
  if (nxl) *nxl = m->nxl;
  if (nyl) *nyl = m->nyl;
  if (ey0) *ey0 = m->ey0;
  if (ney) *ney = m->ney;
  return SEMB_OK;
}

extern "C" int semb_mesh_plan(semb_mesh* m, int* nstrips, int* nchunks, int* nxseam, int* nyseam, int* fast) {
  if (m && fast) *fast = 0;
  SEMB_REQUIRE(m, "null mesh");
  if (nstrips) *nstrips = m->nstrips;
  if (nchunks) *nchunks = m->nchunks;
  if (nxseam) *nxseam = m->nxseam;
  if (nyseam) *nyseam = m->nyseam;
  if (fast) *fast = m->fast ? (m->eo ? 2 : 1) : 0;
  return SEMB_OK;
}

// -DSEMB_TAIL_TIMING builds: 8 values per CTA of the last fused-tail launch (globaltimer ns: start, rows done / announced
// per chunk, tail prologue done, tasks done; [7] = micro-tasks this CTA ran).  SEMB_EINVAL in a normal build.
extern "C" int semb_mesh_debug_read(semb_mesh* m, long long* host, int ncta) {
  SEMB_REQUIRE(m && host && ncta >= 1 && ncta <= SEMB_NPARTIALS, "semb_mesh_debug_read: bad argument");
  SEMB_REQUIRE(m->d_dbg, "semb_mesh_debug_read: library was not built with -DSEMB_TAIL_TIMING (or no fused apply ran yet)");
  SEMB_ENTER(m->ctx);
  SEMB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  SEMB_CHECK_CUDA(cudaMemcpy(host, m->d_dbg, 8 * (size_t)ncta * sizeof(long long), cudaMemcpyDeviceToHost));
  return SEMB_OK;
}

extern "C" int semb_mesh_groups(semb_mesh* m, int* ngroups) {
  SEMB_REQUIRE(m && ngroups, "semb_mesh_groups: null argument");
  *ngroups = m->ngroups;
  return SEMB_OK;
}

extern "C" int semb_mesh_fused_tail(semb_mesh* m, int* on) {
  SEMB_REQUIRE(m && on, "semb_mesh_fused_tail: null argument");
  *on = m->tail ? 1 : 0;
  return SEMB_OK;
}

// test / tuning hook: override the number of y chunks (CTA rows) of the strip kernel
extern "C" int semb_mesh_set_chunks(semb_mesh* m, int nchunks) {
  SEMB_REQUIRE(m && nchunks >= 1 && nchunks <= m->ney, "semb_mesh_set_chunks: 1 <= nchunks <= ney");
  SEMB_REQUIRE((long long)(nchunks + 1) * m->nstrips <= SEMB_NPARTIALS, "semb_mesh_set_chunks: too many CTAs");
  SEMB_ENTER(m->ctx);
  return mesh_set_groups(m, nchunks);
}

extern "C" int semb_mesh_get(semb_mesh* m, int which, double* host) {
  SEMB_REQUIRE(m && host, "semb_mesh_get: null argument");
  SEMB_REQUIRE(which >= 0 && which < SEMB_MESH_ARRAY_COUNT, "semb_mesh_get: bad selector %d", which);
  SEMB_REQUIRE(m->arr[which], "semb_mesh_get: array %d is not held by this mesh", which);
  SEMB_ENTER(m->ctx);
  return download_pitched(m, m->arr[which], host);
}

extern "C" int semb_mesh_set(semb_mesh* m, int which, const double* host) {
  SEMB_REQUIRE(m && host, "semb_mesh_set: null argument");
  SEMB_REQUIRE(which >= 0 && which < SEMB_MESH_ARRAY_COUNT && which != SEMB_MULT, "semb_mesh_set: bad selector %d", which);
  SEMB_ENTER(m->ctx);
  SEMB_TRY(mesh_alloc_array(m, which));
  return upload_pitched(m, m->arr[which], host);
}

extern "C" int semb_mesh_get_D(semb_mesh* m, double* Dr, double* Ds) {
  SEMB_REQUIRE(m, "null mesh");
  if (Dr) std::copy(m->hDr.begin(), m->hDr.end(), Dr);
  if (Ds) std::copy(m->hDs.begin(), m->hDs.end(), Ds);
  return SEMB_OK;
}

struct MaskFlags {
  int mx0, mx1, my0, my1;
};
static int parse_bc(semb_mesh* m, const char* bc, MaskFlags* f) {
  f->mx0 = f->mx1 = f->my0 = f->my1 = 0;
  if (!bc) return SEMB_OK;
  for (int i = 0; i < 4; ++i)
    SEMB_REQUIRE(bc[i] == 'D' || bc[i] == 'N', "bc must be 4 chars of 'D'/'N' (mesh.jl:138), got '%c'", bc[i]);
  // mesh.jl:159-165: periodic direction overrides 'D'; only slabs on the global boundary carry y flags
  f->mx0 = (bc[0] == 'D') && !m->perx;
  f->mx1 = (bc[1] == 'D') && !m->perx;
  f->my0 = (bc[2] == 'D') && !m->pery && m->ey0 == 0;
  f->my1 = (bc[3] == 'D') && !m->pery && (m->ey0 + m->ney == m->Ey);
  return SEMB_OK;
}

extern "C" int semb_generate_mask(semb_mesh* m, const char bc[4], double* host) {
  SEMB_REQUIRE(m && bc && host, "semb_generate_mask: null argument");
  SEMB_ENTER(m->ctx);
  MaskFlags f;
  SEMB_TRY(parse_bc(m, bc, &f));
  double* d = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d, m->nalloc * sizeof(double)));
  int rc = semb_launch_mask_gen(m->ctx, d, m->pitch, m->nxl, m->nyl, f.mx0, f.mx1, f.my0, f.my1);
  if (!rc) rc = download_pitched(m, d, host);
  cudaFree(d);
  return rc;
}

// ---- fields ------------------------------------------------------------------------------------------------
extern "C" int semb_field_create(semb_mesh* m, semb_field** out) {
  SEMB_REQUIRE(m && out, "semb_field_create: null argument");
  SEMB_ENTER(m->ctx);
  semb_field* f = new semb_field();
  f->mesh = m;
  cudaError_t e = cudaMalloc(&f->d, m->nalloc * sizeof(double));
  if (e != cudaSuccess) {
    delete f;
    semb_set_error("semb_field_create: cudaMalloc of %zu bytes failed: %s", m->nalloc * sizeof(double),
                   cudaGetErrorString(e));
    return SEMB_ENOMEM;
  }
  SEMB_CHECK_CUDA(cudaMemsetAsync(f->d, 0, m->nalloc * sizeof(double), m->ctx->stream));
  m->fields.push_back(f);
  *out = f;
  return SEMB_OK;
}

extern "C" int semb_field_destroy(semb_field* f) {
  if (!f) return SEMB_OK;
  semb_mesh* m = f->mesh;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  auto it = std::find(m->fields.begin(), m->fields.end(), f);
  if (it != m->fields.end()) m->fields.erase(it);
  cudaFree(f->d);
  delete f;
  return SEMB_OK;
}

extern "C" int semb_field_upload(semb_field* f, const double* host) {
  SEMB_REQUIRE(f && host, "semb_field_upload: null argument");
  SEMB_ENTER(f->mesh->ctx);
  return upload_pitched(f->mesh, f->d, host);
}

extern "C" int semb_field_download(semb_field* f, double* host) {
  SEMB_REQUIRE(f && host, "semb_field_download: null argument");
  SEMB_ENTER(f->mesh->ctx);
  return download_pitched(f->mesh, f->d, host);
}

extern "C" int semb_field_fill(semb_field* f, double v) {
  SEMB_REQUIRE(f, "null field");
  semb_mesh* m = f->mesh;
  SEMB_ENTER(m->ctx);
  return semb_launch_fill(m->ctx, f->d, v, m->pitch, m->nxl, m->nyl);
}

extern "C" int semb_field_copy(semb_field* dst, const semb_field* src) {
  SEMB_REQUIRE(dst && src && dst->mesh == src->mesh, "semb_field_copy: fields must share a mesh");
  semb_mesh* m = dst->mesh;
  SEMB_ENTER(m->ctx);
  SEMB_CHECK_CUDA(cudaMemcpyAsync(dst->d, src->d, m->nalloc * sizeof(double), cudaMemcpyDeviceToDevice,
                                  m->ctx->stream));
  return SEMB_OK;
}

extern "C" int semb_field_fill_random(semb_field* f, uint64_t seed) {
  SEMB_REQUIRE(f, "null field");
  semb_mesh* m = f->mesh;
  SEMB_ENTER(m->ctx);
  return semb_launch_fill_random(m->ctx, f->d, m->pitch, m->nxl, m->nyl, m->nxl, (long long)m->ey0 * m->ns, seed);
}

extern "C" int semb_field_axpby(double a, const semb_field* x, double b, semb_field* y) {
  SEMB_REQUIRE(x && y && x->mesh == y->mesh, "semb_field_axpby: fields must share a mesh");
  semb_mesh* m = y->mesh;
  SEMB_ENTER(m->ctx);
  return semb_launch_axpby(m->ctx, a, x->d, b, y->d, m->nalloc);
}

extern "C" int semb_field_devptr(semb_field* f, void** dptr, long long* pitch) {
  SEMB_REQUIRE(f, "null field");
  if (dptr) *dptr = f->d;
  if (pitch) *pitch = f->mesh->pitch;
  return SEMB_OK;
}

// ---- operators ------------------------------------------------------------------------------------------------
static int check_field(semb_mesh* m, const semb_field* f, const char* what, bool optional = false) {
  if (!f) {
    SEMB_REQUIRE(optional, "%s: null field", what);
    return SEMB_OK;
  }
  SEMB_REQUIRE(f->mesh == m, "%s: field belongs to another mesh (DimensionMismatch)", what);
  return SEMB_OK;
}

static P2PArgs p2p_args(semb_mesh* m, unsigned long long epoch = 0) {
  P2PArgs x;
  x.on = m->p2p ? 1 : 0;
  x.nranks = m->ctx->nranks;
  x.rank = m->ctx->rank;
  x.epoch = epoch;
  return x;
}

// local / peer halo rows inside a mailbox: [parity][side][pitch], side 0 = row from below, 1 = from above
static double* mail_halo(semb_mesh* m, void* mailbox, int parity, int side) {
  return (double*)((char*)mailbox + mail_scal_bytes()) + (size_t)(2 * parity + side) * m->pitch;
}

// Exchange the slab's boundary rows with the neighbour ranks.  P2P mode: one small kernel stores the rows
// straight into the neighbours' mailboxes over NVLink and releases an epoch flag (the consumer, the y-seam
// kernel, waits on it).  Fallback: grouped ncclSend/ncclRecv.  Returns the epoch used in *epoch.
static int halo_exchange(semb_mesh* m, double* field, int pcg, unsigned long long* epoch) {
  semb_ctx* c = m->ctx;
  if (epoch) *epoch = 0;
  if (!m->halo_lo && !m->halo_hi) return SEMB_OK;
  if (m->p2p) {
    const unsigned long long ep = ++m->ep_halo;
    const int par = (int)(ep & 1ull);
    SembScal* plo = m->halo_lo ? (SembScal*)m->peer_mailbox[m->rank_lo] : nullptr;
    SembScal* phi = m->halo_hi ? (SembScal*)m->peer_mailbox[m->rank_hi] : nullptr;
    if (epoch) *epoch = ep;
    return semb_launch_halo_push(c, field, m->pitch, m->nxl, m->nyl,
                                 plo ? mail_halo(m, plo, par, 1) : nullptr, phi ? mail_halo(m, phi, par, 0) : nullptr,
                                 plo ? &plo->flag_halo[1] : nullptr, phi ? &phi->flag_halo[0] : nullptr, ep,
                                 m->d_counters + 5, m->d_scal, pcg);
  }
  SEMB_REQUIRE(c->comm, "halo exchange without a communicator");
  SEMB_CHECK_NCCL(ncclGroupStart());
  // order matters only when both neighbours are the same rank (2 ranks, periodic y): sends go
  // {last row -> hi, first row -> lo}, receives {lo, hi}
  if (m->halo_hi)
    SEMB_CHECK_NCCL(ncclSend(field + (size_t)(m->nyl - 1) * m->pitch, m->nxl, ncclDouble, m->rank_hi, c->comm,
                             c->stream));
  if (m->halo_lo) SEMB_CHECK_NCCL(ncclSend(field, m->nxl, ncclDouble, m->rank_lo, c->comm, c->stream));
  if (m->halo_lo) SEMB_CHECK_NCCL(ncclRecv(m->d_halo_lo, m->nxl, ncclDouble, m->rank_lo, c->comm, c->stream));
  if (m->halo_hi) SEMB_CHECK_NCCL(ncclRecv(m->d_halo_hi, m->nxl, ncclDouble, m->rank_hi, c->comm, c->stream));
  SEMB_CHECK_NCCL(ncclGroupEnd());
  return SEMB_OK;
}

static int ensure_tmp(semb_mesh* m, semb_field** slot) {
  if (*slot) return SEMB_OK;
  return semb_field_create(m, slot);
}

struct OpSpec {
  const semb_field* nu_arr = nullptr;
  double nu = 1.0;
  const semb_field* k_arr = nullptr;
  double k = 0.0;
  const char* bc = nullptr;
  const semb_field* M_arr = nullptr;
  bool gs = true;
};

static void fill_common(semb_mesh* m, OpArgs& a) {
  a.G11 = m->arr[SEMB_G11];
  a.G12 = m->arr[SEMB_G12];
  a.G22 = m->arr[SEMB_G22];
  a.B = m->arr[SEMB_B];
  a.mult = m->arr[SEMB_MULT];
  a.pitch = m->pitch;
  a.N = m->ns;
  a.Ex = m->Ex;
  a.ney = m->ney;
  a.nxl = m->nxl;
  a.nyl = m->nyl;
  a.perx = m->perx;
  a.chunk_r0 = m->d_chunk_r0;
  a.nchunks = m->nchunks;
  a.ystart = m->d_ystart;
  a.xseam = m->d_xseam;
  a.nxseam = m->nxseam;
  a.yseam = m->d_yseam;
  a.nyseam = m->nyseam;
  a.scal = m->d_scal;
  a.halo_lo = m->d_halo_lo;
  a.halo_hi = m->d_halo_hi;
  a.wx1d = m->d_wx1d;
  a.wy1d = m->d_wy1d;
}

// out = [mask(gs(]  nu .* laplace(u) + k .* B .* u  [))]   -- the one place operators are composed
static int run_operator(semb_mesh* m, const double* u, double* out, const OpSpec& sp, bool pcg, double* p_inout,
                        int precond, double prec_b0) {
  semb_ctx* c = m->ctx;
  SEMB_REQUIRE(m->arr[SEMB_G11] && m->arr[SEMB_G12] && m->arr[SEMB_G22], "operator: mesh has no G11/G12/G22");
  const bool hasmass = (sp.k_arr != nullptr) || (sp.k != 0.0);
  SEMB_REQUIRE(!hasmass || m->arr[SEMB_B], "operator: mass term requested but the mesh has no B");
  // kernel variant with general coefficients: mass term and/or array viscosity
  const bool massterm = hasmass || (sp.nu_arr != nullptr);
  SEMB_REQUIRE(!precond || m->arr[SEMB_B], "operator: preconditioner needs B");
  OpArgs a;
  fill_common(m, a);
  a.u = u;
  a.out = out;
  a.nu_arr = sp.nu_arr ? sp.nu_arr->d : nullptr;
  a.k_arr = sp.k_arr ? sp.k_arr->d : nullptr;
  a.M_arr = sp.M_arr ? sp.M_arr->d : nullptr;
  a.nu = sp.nu;
  a.k = sp.k;
  a.gs = sp.gs ? 1 : 0;
  a.pcg = pcg ? 1 : 0;
  a.precond = precond;
  a.prec_b0 = prec_b0;
  if (pcg) {
    a.pold = p_inout;
    a.pout = p_inout;
  }
  MaskFlags f;
  SEMB_TRY(parse_bc(m, sp.bc, &f));
  a.mx0 = f.mx0;
  a.mx1 = f.mx1;
  a.my0 = f.my0;
  a.my1 = f.my1;
  if (m->fast && m->tail && sp.gs && (pcg || m->tail_plain)) {
    // ONE launch: the strip kernel's CTAs finish the interfaces, exchange the halo rows through peer memory and
    // (PCG) reduce + all-gather sum(p.*Ap.*mult) themselves (semb_tail.cuh)
    a.tail = 1;
    a.grp = m->d_grp;
    a.xmic = m->d_grp + 2 * m->ngroups;
    a.xmic_total = m->xmic_total;
    a.tcnt = m->d_tcnt;
    a.tpart = m->d_tpart;
    a.nxs = m->nstrips - 1 + m->perx;
    a.ywrap = (m->pery && c->nranks == 1) ? 1 : 0;
    a.has_lo = m->halo_lo;
    a.has_hi = m->halo_hi;
    a.ep_host = m->ep_tail_host;
    if (!pcg && (m->halo_lo || m->halo_hi)) ++m->ep_tail_host;  // (PCG-mode applies count on the device: graph replay)
    if (m->halo_lo) {
      a.peer_rows_lo = mail_tail_rows(m, m->peer_mailbox[m->rank_lo]);
    }
    if (m->halo_hi) {
      a.peer_rows_hi = mail_tail_rows(m, m->peer_mailbox[m->rank_hi]);
    }
    a.my_rows = mail_tail_rows(m, m->d_mailbox);
    a.dbg = m->d_dbg;
    return semb_launch_strip(c, a, m->hDr.data(), m->hDs.data(), m->nstrips, m->ngroups, pcg, massterm, m->eo);
  }
  if (m->fast) {
    a.partials = m->d_partials;
    a.counters = m->d_counters + 0;
    SEMB_TRY(semb_launch_strip(c, a, m->hDr.data(), m->hDs.data(), m->nstrips, m->nchunks, pcg, massterm, m->eo));
    if (!sp.gs) return SEMB_OK;
    a.partials = m->d_partials + m->npartials;
    a.counters = m->d_counters + 1;
    SEMB_TRY(semb_launch_seam_x(c, a));
    unsigned long long eph = 0;
    SEMB_TRY(halo_exchange(m, out, pcg ? 1 : 0, &eph));
    if (m->p2p) {  // received rows live in the local mailbox, double-buffered by epoch parity
      a.halo_lo = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 0);
      a.halo_hi = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 1);
    }
    a.partials = m->d_partials + 2 * (size_t)m->npartials;
    a.counters = m->d_counters + 2;
    // P2P + PCG: the y-seam kernel's last block also all-gathers sum(p.*Ap.*mult) over NVLink
    SEMB_TRY(semb_launch_seam_y(c, a, m->halo_lo, m->halo_hi, true, p2p_args(m, eph)));
    return SEMB_OK;
  }
  // generic path (nr != ns, or outside 2..17): separate passes, same arithmetic per node
  SEMB_TRY(ensure_tmp(m, &m->w_tmp));
  SEMB_TRY(ensure_tmp(m, &m->w_t1));
  SEMB_TRY(ensure_tmp(m, &m->w_t2));
  semb_field *t1 = m->w_t1, *t2 = m->w_t2;
  int rc = SEMB_OK;
  auto body = [&]() -> int {
    if (pcg) SEMB_TRY(semb_launch_pcg_dir(c, m, u, p_inout, precond, prec_b0));
    OpArgs g = a;
    g.u = pcg ? p_inout : u;
    g.out = sp.gs ? m->w_tmp->d : out;
    SEMB_TRY(semb_launch_generic_local(c, g, m->nr, m->ns, m->dDr, m->dDs, t1->d, t2->d, hasmass));
    if (!sp.gs) return SEMB_OK;
    SEMB_TRY(semb_launch_gs_x(c, m->w_tmp->d, out, m->pitch, m->nr, m->Ex, m->nxl, m->nyl, m->perx));
    SEMB_TRY(halo_exchange(m, out, 0, nullptr));  // generic meshes never use P2P
    OpArgs y = a;
    y.pcg = 0;
    y.yseam = m->d_yseam + 2 * m->nyseam;  // all y interfaces
    y.nyseam = (m->ney - 1) + ((m->pery && c->nranks == 1) ? 1 : 0);
    SEMB_TRY(semb_launch_seam_y(c, y, m->halo_lo, m->halo_hi, false, P2PArgs()));
    OpArgs md = a;
    md.partials = m->d_partials;
    md.counters = m->d_counters + 0;
    SEMB_TRY(semb_launch_mask_dot(c, m, md));
    return SEMB_OK;
  };
  rc = body();
  return rc;
}

extern "C" int semb_lapl(semb_mesh* m, const semb_field* u, semb_field* out) {
  return semb_hlmz(m, u, nullptr, 1.0, nullptr, 0.0, out);
}

extern "C" int semb_hlmz(semb_mesh* m, const semb_field* u, const semb_field* nu_arr, double nu,
                         const semb_field* k_arr, double k, semb_field* out) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "hlmz(u)"));
  SEMB_TRY(check_field(m, out, "hlmz(out)"));
  SEMB_TRY(check_field(m, nu_arr, "hlmz(nu)", true));
  SEMB_TRY(check_field(m, k_arr, "hlmz(k)", true));
  SEMB_REQUIRE(u != out, "hlmz: out must not alias u");
  OpSpec sp;
  sp.nu_arr = nu_arr;
  sp.nu = nu;
  sp.k_arr = k_arr;
  sp.k = k;
  sp.gs = false;
  return run_operator(m, u->d, out->d, sp, false, nullptr, 0, 1.0);
}

extern "C" int semb_mass(semb_mesh* m, const semb_field* u, semb_field* out) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "mass(u)"));
  SEMB_TRY(check_field(m, out, "mass(out)"));
  SEMB_REQUIRE(m->arr[SEMB_B], "mass: mesh has no B");
  return semb_launch_mask(m->ctx, u->d, m->arr[SEMB_B], out->d, m->nalloc);  // B .* u, mass.jl:17
}

extern "C" int semb_gather_scatter(semb_mesh* m, const semb_field* u, semb_field* out) {
  SEMB_REQUIRE(m, "null mesh");
  semb_ctx* c = m->ctx;
  SEMB_ENTER(c);
  SEMB_TRY(check_field(m, u, "gatherScatter(u)"));
  SEMB_TRY(check_field(m, out, "gatherScatter(out)"));
  SEMB_REQUIRE(u != out, "gatherScatter: out must not alias u");
  SEMB_TRY(semb_launch_gs_x(c, u->d, out->d, m->pitch, m->nr, m->Ex, m->nxl, m->nyl, m->perx));
  unsigned long long eph = 0;
  SEMB_TRY(halo_exchange(m, out->d, 0, &eph));
  OpArgs y;
  fill_common(m, y);
  if (m->p2p) {
    y.halo_lo = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 0);
    y.halo_hi = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 1);
  }
  y.out = out->d;
  y.yseam = m->d_yseam + 2 * m->nyseam;
  y.nyseam = (m->ney - 1) + ((m->pery && c->nranks == 1) ? 1 : 0);
  return semb_launch_seam_y(c, y, m->halo_lo, m->halo_hi, false, p2p_args(m, eph));
}

extern "C" int semb_mask(semb_mesh* m, const semb_field* u, const semb_field* M, semb_field* out) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "mask(u)"));
  SEMB_TRY(check_field(m, out, "mask(out)"));
  SEMB_TRY(check_field(m, M, "mask(M)", true));
  return semb_launch_mask(m->ctx, u->d, M ? M->d : nullptr, out->d, m->nalloc);
}

extern "C" int semb_mask_bc(semb_mesh* m, const semb_field* u, const char bc[4], semb_field* out) {
  SEMB_REQUIRE(m && bc, "semb_mask_bc: null argument");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "mask(u)"));
  SEMB_TRY(check_field(m, out, "mask(out)"));
  if (u != out) SEMB_TRY(semb_field_copy(out, u));
  MaskFlags f;
  SEMB_TRY(parse_bc(m, bc, &f));
  OpArgs a;
  fill_common(m, a);
  a.out = out->d;
  a.mx0 = f.mx0;
  a.mx1 = f.mx1;
  a.my0 = f.my0;
  a.my1 = f.my1;
  a.partials = m->d_partials;
  a.counters = m->d_counters + 0;
  return semb_launch_mask_dot(m->ctx, m, a);
}

extern "C" int semb_oplhs(semb_mesh* m, const semb_field* u, const semb_field* nu_arr, double nu,
                          const semb_field* k_arr, double k, const char* bc, const semb_field* M_arr,
                          semb_field* out) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "opLHS(u)"));
  SEMB_TRY(check_field(m, out, "opLHS(out)"));
  SEMB_TRY(check_field(m, nu_arr, "opLHS(nu)", true));
  SEMB_TRY(check_field(m, k_arr, "opLHS(k)", true));
  SEMB_TRY(check_field(m, M_arr, "opLHS(M)", true));
  SEMB_REQUIRE(u != out, "opLHS: out must not alias u");
  OpSpec sp;
  sp.nu_arr = nu_arr;
  sp.nu = nu;
  sp.k_arr = k_arr;
  sp.k = k;
  sp.bc = bc;
  sp.M_arr = M_arr;
  sp.gs = true;
  return run_operator(m, u->d, out->d, sp, false, nullptr, 0, 1.0);
}

extern "C" int semb_jac(semb_mesh* m, const semb_field* x, const semb_field* y, semb_field* J, semb_field* Ji,
                        semb_field* rx, semb_field* ry, semb_field* sx, semb_field* sy) {
  SEMB_REQUIRE(m, "null mesh");
  semb_ctx* c = m->ctx;
  SEMB_ENTER(c);
  SEMB_TRY(check_field(m, x, "jac(x)"));
  SEMB_TRY(check_field(m, y, "jac(y)"));
  double *d_wr = nullptr, *d_ws = nullptr;
  SEMB_CHECK_CUDA(cudaMalloc(&d_wr, m->nr * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMalloc(&d_ws, m->ns * sizeof(double)));
  SEMB_CHECK_CUDA(cudaMemcpy(d_wr, m->hwr.data(), m->nr * sizeof(double), cudaMemcpyHostToDevice));
  SEMB_CHECK_CUDA(cudaMemcpy(d_ws, m->hws.data(), m->ns * sizeof(double), cudaMemcpyHostToDevice));
  int rc = semb_launch_geom(c, x->d, y->d, m->pitch, m->nr, m->ns, m->Ex, m->ney, m->dDr, m->dDs, d_wr, d_ws,
                            J ? J->d : nullptr, Ji ? Ji->d : nullptr, rx ? rx->d : nullptr, ry ? ry->d : nullptr,
                            sx ? sx->d : nullptr, sy ? sy->d : nullptr, nullptr, nullptr, nullptr, nullptr,
                            nullptr);
  cudaStreamSynchronize(c->stream);
  cudaFree(d_wr);
  cudaFree(d_ws);
  return rc;
}

static int gather_scalars(semb_mesh* m, double* xchg, int per_rank) {
  semb_ctx* c = m->ctx;
  if (c->nranks == 1) return SEMB_OK;
  SEMB_REQUIRE(c->comm, "multi-rank reduction without a communicator");
  SEMB_CHECK_NCCL(ncclAllGather(xchg + (size_t)c->rank * per_rank, xchg, per_rank, ncclDouble, c->comm, c->stream));
  return SEMB_OK;
}

static int read_scal(semb_mesh* m) {
  SEMB_CHECK_CUDA(cudaMemcpyAsync(m->h_scal, m->d_scal, sizeof(SembScal), cudaMemcpyDeviceToHost, m->ctx->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  if (m->h_scal->err) {  // a kernel gave up waiting for a peer (semb_wait_epoch): results since then are undefined
    semb_set_error("peer exchange timed out on rank %d of %d (a neighbour rank is lost or has diverged); "
                   "SEMB_PEER_TIMEOUT_MS sets the bound", m->ctx->rank, m->ctx->nranks);
    return SEMB_ENCCL;
  }
  return SEMB_OK;
}

extern "C" int semb_mesh_peer_status(semb_mesh* m) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  return read_scal(m);
}

static int reduce_common(semb_mesh* m, int which, const semb_field* a, const semb_field* b, double* result,
                         double ref = 0.0) {
  semb_ctx* c = m->ctx;
  SEMB_ENTER(c);
  SEMB_REQUIRE(result, "reduction: null result");
  SEMB_TRY(semb_launch_reduce(c, m, which, a->d, b ? b->d : nullptr, p2p_args(m), ref));
  if (c->nranks > 1 && !m->p2p) {
    SEMB_TRY(gather_scalars(m, SEMB_SCAL_PTR(m, xchg_red), 1));
    SEMB_TRY(semb_launch_reduce_finalize(c, m, which));
  }
  SEMB_TRY(read_scal(m));
  *result = m->h_scal->red[which];
  return SEMB_OK;
}

extern "C" int semb_dot_mult(semb_mesh* m, const semb_field* a, const semb_field* b, double* result) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_TRY(check_field(m, a, "dot(a)"));
  SEMB_TRY(check_field(m, b, "dot(b)"));
  return reduce_common(m, 0, a, b, result);
}

extern "C" int semb_norm_inf(semb_mesh* m, const semb_field* a, double* result) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_TRY(check_field(m, a, "norm(a)"));
  return reduce_common(m, 1, a, nullptr, result);
}

// Is an array coefficient (nu, k) really one constant, on every rank?  The reference's drivers always carry
// nu as an array (diffusion.jl:11) that setVisc! fills with a constant (examples/p2d.jl:21-24, d2d.jl:35-38);
// recognising that lets the solve run the scalar-coefficient kernel variant (same bits: nu .* x is the same
// product either way) instead of streaming 8 more bytes per DOF per iteration.
static int field_constant(semb_mesh* m, const semb_field* f, bool* is_const, double* value) {
  semb_ctx* c = m->ctx;
  double v0 = 0.0;
  SEMB_CHECK_CUDA(cudaMemcpyAsync(&v0, f->d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  double spread = 0.0;
  SEMB_TRY(reduce_common(m, 2, f, nullptr, &spread, v0));  // max |f - f[0,0]| over this rank's nodes (all-reduced)
  double chk[3] = {spread, v0, -v0};
  SEMB_TRY(semb_comm_allreduce_max(c, chk, 3));  // and every rank saw the same f[0,0]
  *is_const = (chk[0] == 0.0) && (chk[1] == -chk[2]) && std::isfinite(v0);
  *value = v0;
  return SEMB_OK;
}

// ---- PCG ---------------------------------------------------------------------------------------------------------
extern "C" int semb_pcg_begin(semb_mesh* m, const semb_pcg_opts* o, const semb_field* b, semb_field* x) {
  SEMB_REQUIRE(m && o, "semb_pcg_begin: null argument");
  semb_ctx* c = m->ctx;
  SEMB_ENTER(c);
  SEMB_TRY(check_field(m, b, "pcg(b)"));
  SEMB_TRY(check_field(m, x, "pcg(x)"));
  SEMB_TRY(check_field(m, o->nu_arr, "pcg(nu)", true));
  SEMB_TRY(check_field(m, o->k_arr, "pcg(k)", true));
  SEMB_TRY(check_field(m, o->M_arr, "pcg(M)", true));
  SEMB_REQUIRE(b != x, "pcg: x must not alias b");
  SEMB_REQUIRE(o->precond >= 0 && o->precond <= 2, "pcg: precond must be 0 (identity), 1 (diagonal) or 2 (FDM)");
  SEMB_REQUIRE(o->precond != 1 || m->arr[SEMB_B], "pcg: diagonal preconditioner needs B");
  SEMB_REQUIRE(o->precond != 2 || (m->fdm && m->fast && !m->pcg_custom),
               "pcg: precond = 2 needs an FDM preconditioner on this mesh (semb_fdm_create) and the fused operator path");
  SEMB_REQUIRE(o->precond != 2 || c->nranks == 1 || m->p2p, "pcg: on several ranks the FDM preconditioner needs the peer-memory transport");
  MaskFlags f;
  SEMB_TRY(parse_bc(m, o->bc, &f));
  SEMB_TRY(ensure_tmp(m, &m->w_r));
  SEMB_TRY(ensure_tmp(m, &m->w_p));
  SEMB_TRY(ensure_tmp(m, &m->w_Ap));
  if (!m->fast) {  // the generic operator's work fields: never allocated while a CUDA graph is being captured
    SEMB_TRY(ensure_tmp(m, &m->w_tmp));
    SEMB_TRY(ensure_tmp(m, &m->w_t1));
    SEMB_TRY(ensure_tmp(m, &m->w_t2));
  }
  m->pcg_opts = *o;
  m->pcg_x = x;
  if (o->nu_arr) {
    bool cst = false;
    double v = 0.0;
    SEMB_TRY(field_constant(m, o->nu_arr, &cst, &v));
    if (cst) {
      m->pcg_opts.nu_arr = nullptr;
      m->pcg_opts.nu = v;
    }
  }
  if (o->k_arr) {
    bool cst = false;
    double v = 0.0;
    SEMB_TRY(field_constant(m, o->k_arr, &cst, &v));
    if (cst) {
      m->pcg_opts.k_arr = nullptr;
      m->pcg_opts.k = v;
    }
  }
  long long maxiter = o->maxiter;
  if (maxiter < 0) maxiter = (long long)m->nxl * ((long long)m->ns * m->Ey);  // length(b), pcg.jl:21
  // reset the device scalars
  memset(m->h_scal, 0, sizeof(SembScal));
  m->h_scal->nranks = c->nranks;
  m->h_scal->rank = c->rank;
  SEMB_CHECK_CUDA(cudaMemcpyAsync(m->d_scal, m->h_scal, SEMB_SCAL_HOST_BYTES, cudaMemcpyHostToDevice, c->stream));
  // diagonal preconditioner on the fused path: init / update keep h = r./B./b0 (they form it anyway for t), and the
  // strip kernel stages h instead of r: same bits, no divisions and no B column at the head of the strip kernel's row
  const bool keep_h = o->precond && m->fast && !m->pcg_custom && (o->precond == 2 || !getenv("SEMB_NO_PCG_H"));
  if (keep_h) SEMB_TRY(ensure_tmp(m, &m->w_h));
  m->pcg_keep_h = keep_h;
  SEMB_TRY(semb_launch_pcg_init(c, m, b->d, x->d, m->w_r->d, m->w_p->d, keep_h ? m->w_h->d : nullptr, o->precond,
                                o->prec_b0, o->tol, maxiter,
                                p2p_args(m)));
  if (c->nranks > 1 && !m->p2p) {
    SEMB_TRY(gather_scalars(m, SEMB_SCAL_PTR(m, xchg_t), 2));
    SEMB_TRY(semb_launch_pcg_finalize(c, m, 1));
  }
  if (o->precond == 2) SEMB_TRY(semb_fdm_apply_impl(m->fdm, m->w_r->d, m->w_h->d, 2));  // h = opM(r), t, state (pcg.jl:37,45)
  m->pcg_active = true;
  return SEMB_OK;
}

static int pcg_one_iteration(semb_mesh* m) {
  semb_ctx* c = m->ctx;
  const semb_pcg_opts& o = m->pcg_opts;
  OpSpec sp;
  sp.nu_arr = o.nu_arr;
  sp.nu = o.nu;
  sp.k_arr = o.k_arr;
  sp.k = o.k;
  sp.bc = o.bc;
  sp.M_arr = o.M_arr;
  sp.gs = true;
  if (m->pcg_custom) {
    SEMB_TRY(semb_launch_pcg_dir(c, m, m->w_r->d, m->w_p->d, o.precond, o.prec_b0));  // p = h + beta*p, pcg.jl:46-50
    SEMB_TRY(m->pcg_custom());                                                         // w_Ap = opA(w_p), pcg.jl:51
    SEMB_TRY(semb_launch_reduce(c, m, 0, m->w_p->d, m->w_Ap->d, p2p_args(m)));  // pcg.jl:52
    if (c->nranks > 1 && !m->p2p) {
      SEMB_TRY(gather_scalars(m, SEMB_SCAL_PTR(m, xchg_red), 1));
      SEMB_TRY(semb_launch_reduce_finalize(c, m, 0));
    }
    SEMB_TRY(semb_launch_pcg_set_pap(c, m));
  } else {
    if (m->pcg_keep_h)
      SEMB_TRY(run_operator(m, m->w_h->d, m->w_Ap->d, sp, true, m->w_p->d, 0, o.prec_b0));
    else
      SEMB_TRY(run_operator(m, m->w_r->d, m->w_Ap->d, sp, true, m->w_p->d, o.precond, o.prec_b0));
    if (c->nranks > 1 && !m->p2p) {
      SEMB_TRY(semb_launch_pcg_pack_pap(c, m));
      SEMB_TRY(gather_scalars(m, SEMB_SCAL_PTR(m, xchg_pap), 1));
      SEMB_TRY(semb_launch_pcg_combine_pap(c, m));
    }
  }
  // P2P: the update kernel's last block all-gathers {t, norm(r,Inf)} over NVLink and advances the state
  SEMB_TRY(semb_launch_pcg_update(c, m, m->pcg_x->d, m->w_r->d, m->w_p->d, m->w_Ap->d,
                                  m->pcg_keep_h ? m->w_h->d : nullptr, o.precond, o.prec_b0,
                                  p2p_args(m)));
  if (c->nranks > 1 && !m->p2p) {
    SEMB_TRY(gather_scalars(m, SEMB_SCAL_PTR(m, xchg_t), 2));
    SEMB_TRY(semb_launch_pcg_finalize(c, m, 0));
  }
  if (o.precond == 2) SEMB_TRY(semb_fdm_apply_impl(m->fdm, m->w_r->d, m->w_h->d, 1));  // h = opM(r), t, state
  return SEMB_OK;
}

extern "C" int semb_pcg_iterate(semb_mesh* m, int n) {
  SEMB_REQUIRE(m && m->pcg_active, "semb_pcg_iterate: call semb_pcg_begin first");
  SEMB_ENTER(m->ctx);
  for (int i = 0; i < n; ++i) SEMB_TRY(pcg_one_iteration(m));
  return SEMB_OK;
}

extern "C" int semb_pcg_status(semb_mesh* m, long long* iters, double* resinf, int* done) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(read_scal(m));
  if (iters) *iters = m->h_scal->iters;
  if (resinf) *resinf = m->h_scal->rmax;
  if (done) *done = m->h_scal->done;
  return m->h_scal->warned ? SEMB_NOT_CONVERGED : SEMB_OK;
}

extern "C" int semb_pcg(semb_mesh* m, const semb_pcg_opts* o, const semb_field* b, semb_field* x, long long* iters,
                        double* resinf) {
  SEMB_TRY(semb_pcg_begin(m, o, b, x));
  semb_ctx* c = m->ctx;
  // polling interval of the done flag: 16 iterations when the kernels themselves stop at `done` (a late poll costs a
  // few empty launches), 4 with a custom operator, whose kernels run in full until the host notices
  const int every = o->check_every > 0 ? o->check_every : (m->pcg_custom ? 4 : 16);
  // The loop is launch-bound on small meshes: capture `every` iterations (4 kernels each, all scalars live in
  // device memory, so the arguments never change) into a CUDA graph and replay it between polls of the
  // device-side done flag.
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long per_graph = 0;
  // Several ranks: only when every exchange of an iteration is a peer-memory one whose epoch lives in device memory
  // (fused tail + in-kernel all-gathers), i.e. nothing host-side changes from one iteration to the next.
  const bool use_graph = (c->nranks == 1 || (m->p2p && m->tail && !m->pcg_custom)) && every >= 2 && !getenv("SEMB_NO_GRAPH");
  int rc = SEMB_OK;
  for (;;) {
    rc = read_scal(m);
    if (rc < 0 || m->h_scal->done) break;
    if (!std::isfinite(m->h_scal->t) || !std::isfinite(m->h_scal->rmax)) {
      semb_set_error("pcg: non-finite residual (t=%g, rmax=%g) at iteration %lld", m->h_scal->t, m->h_scal->rmax,
                     m->h_scal->iters);
      rc = SEMB_EINVAL;
      break;
    }
    // The kernels of a custom opA (Stokes Schur operator) do not test the device done flag, so iterations enqueued
    // past maxiter would run the operator in full: the last batch is cut to the iterations that remain
    // (50 pressure iterations at 256x256, order 10/8: 4 batches of 16 = 48.7 ms -> 16+16+16+2 = 34 ms).
    long long batch = every;
    if (m->pcg_custom) {
      const long long mx = o->maxiter < 0 ? (long long)m->nxl * ((long long)m->ns * m->Ey) : o->maxiter;
      batch = std::min<long long>(every, std::max<long long>(1, mx - m->h_scal->iters));
    }
    if (use_graph && !exec && batch == every) {
      const long long l0 = c->launches;
      const bool prof = c->profile;
      c->profile = false;
      cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        rc = semb_pcg_iterate(m, every);
        e = cudaStreamEndCapture(c->stream, &graph);
        if (rc >= 0 && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
      }
      c->profile = prof;
      per_graph = c->launches - l0;
      c->launches = l0;
      if (rc < 0) break;
      if (e != cudaSuccess || !exec) {
        semb_set_error("pcg: CUDA graph capture failed: %s", cudaGetErrorString(e));
        rc = SEMB_ECUDA;
        break;
      }
    }
    if (exec && batch == every) {
      cudaError_t e = cudaGraphLaunch(exec, c->stream);
      if (e != cudaSuccess) {
        semb_set_error("pcg: cudaGraphLaunch failed: %s", cudaGetErrorString(e));
        rc = SEMB_ECUDA;
        break;
      }
      c->launches += per_graph;
    } else {
      rc = semb_pcg_iterate(m, (int)batch);
      if (rc < 0) break;
    }
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  m->pcg_active = false;
  if (rc < 0) return rc;
  if (iters) *iters = m->h_scal->iters;
  if (resinf) *resinf = m->h_scal->rmax;
  return m->h_scal->warned ? SEMB_NOT_CONVERGED : SEMB_OK;
}

// ---- grad / dealiased advection (grad.jl, advect.jl) ---------------------------------------------------------------
extern "C" int semb_grad(semb_mesh* m, const semb_field* u, semb_field* ux, semb_field* uy) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, u, "grad(u)"));
  SEMB_TRY(check_field(m, ux, "grad(ux)"));
  SEMB_TRY(check_field(m, uy, "grad(uy)"));
  SEMB_REQUIRE(u != ux && u != uy && ux != uy, "grad: outputs must not alias");
  SEMB_REQUIRE(m->arr[SEMB_RX] && m->arr[SEMB_SY], "grad: mesh has no metric terms (create it from x,y)");
  return semb_launch_grad(m->ctx, m, u->d, ux->d, uy->d);
}

// work space of advect for one (mshV, mshD) pair
struct AdvectWork {
  semb_mesh *V = nullptr, *D = nullptr;
  double *dJr = nullptr, *dJs = nullptr, *dJrT = nullptr, *dJsT = nullptr;  // column-major interpolation matrices
  semb_field *Tx = nullptr, *Ty = nullptr;                                 // on V
  semb_field *JTx = nullptr, *JTy = nullptr, *Jux = nullptr, *Juy = nullptr, *JCu = nullptr;  // on D
  double* mid = nullptr;  // mixed-resolution intermediate of the two-pass ABu
};

static int advect_work_free(AdvectWork* w) {
  if (!w) return SEMB_OK;
  cudaFree(w->dJr);
  cudaFree(w->dJs);
  cudaFree(w->dJrT);
  cudaFree(w->dJsT);
  cudaFree(w->mid);
  semb_field* fs[] = {w->Tx, w->Ty, w->JTx, w->JTy, w->Jux, w->Juy, w->JCu};
  for (semb_field* f : fs) semb_field_destroy(f);
  delete w;
  return SEMB_OK;
}

static int advect_work_fill(AdvectWork* w, semb_mesh* V, semb_mesh* D);
static int advect_work_create(semb_mesh* V, semb_mesh* D, AdvectWork** out) {
  SEMB_REQUIRE(V->arr[SEMB_RX] && V->arr[SEMB_B], "advect: mshV has no metric terms / B");
  AdvectWork* w = new AdvectWork();
  w->V = V;
  w->D = D;
  const int rc = advect_work_fill(w, V, D);  // any failure below releases the struct and what it already holds
  if (rc < 0) {
    advect_work_free(w);
    return rc;
  }
  *out = w;
  return SEMB_OK;
}

static int advect_work_fill(AdvectWork* w, semb_mesh* V, semb_mesh* D) {
  if (D) {
    SEMB_REQUIRE(D->ctx == V->ctx && D->Ex == V->Ex && D->Ey == V->Ey && D->ney == V->ney && D->perx == V->perx &&
                     D->pery == V->pery,
                 "advect: mshD must match mshV in Ex, Ey, periodicity and partition");
    SEMB_REQUIRE(D->arr[SEMB_B], "advect: mshD has no B");
    // Jr = interpMat(mshD.zr, mshV.zr), Js = interpMat(mshD.zs, mshV.zs)  (advect.jl:72-73)
    auto nodes = [](int n, std::vector<double>& z) {
      std::vector<double> wts(n);
      z.resize(n);
      return semb_gausslobatto(n, z.data(), wts.data());
    };
    std::vector<double> zrV, zsV, zrD, zsD;
    SEMB_TRY(nodes(V->nr, zrV));
    SEMB_TRY(nodes(V->ns, zsV));
    SEMB_TRY(nodes(D->nr, zrD));
    SEMB_TRY(nodes(D->ns, zsD));
    std::vector<double> Jr((size_t)D->nr * V->nr), Js((size_t)D->ns * V->ns), JrT(Jr.size()), JsT(Js.size());
    SEMB_TRY(semb_interp_mat(D->nr, zrD.data(), V->nr, zrV.data(), Jr.data()));
    SEMB_TRY(semb_interp_mat(D->ns, zsD.data(), V->ns, zsV.data(), Js.data()));
    for (int i = 0; i < D->nr; ++i)
      for (int k = 0; k < V->nr; ++k) JrT[k + (size_t)i * V->nr] = Jr[i + (size_t)k * D->nr];
    for (int i = 0; i < D->ns; ++i)
      for (int k = 0; k < V->ns; ++k) JsT[k + (size_t)i * V->ns] = Js[i + (size_t)k * D->ns];
    auto up = [&](const std::vector<double>& h, double** d) -> int {
      SEMB_CHECK_CUDA(cudaMalloc(d, h.size() * sizeof(double)));
      SEMB_CHECK_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
      return SEMB_OK;
    };
    SEMB_TRY(up(Jr, &w->dJr));
    SEMB_TRY(up(Js, &w->dJs));
    SEMB_TRY(up(JrT, &w->dJrT));
    SEMB_TRY(up(JsT, &w->dJsT));
  }
  return SEMB_OK;
}

// intermediate fields of the multi-pass form, allocated on first use (the fused kernel needs none)
static int advect_work_fields(AdvectWork* w) {
  semb_mesh *V = w->V, *D = w->D;
  if (!w->Tx) {
    SEMB_TRY(semb_field_create(V, &w->Tx));
    SEMB_TRY(semb_field_create(V, &w->Ty));
  }
  if (D && !w->JCu) {
    semb_field** fd[] = {&w->JTx, &w->JTy, &w->Jux, &w->Juy, &w->JCu};
    for (semb_field** f : fd) SEMB_TRY(semb_field_create(D, f));
    const size_t nmid = std::max((size_t)D->pitch * V->nyl, (size_t)V->pitch * D->nyl);
    SEMB_CHECK_CUDA(cudaMalloc(&w->mid, nmid * sizeof(double)));
  }
  return SEMB_OK;
}

// out (on V) = advect(T, ux, uy, ...) ; advect.jl:45-64 (dealiased) or :27-43
static int advect_run(AdvectWork* w, const double* T, const double* ux, const double* uy, double* out) {
  semb_mesh *V = w->V, *D = w->D;
  semb_ctx* c = V->ctx;
  if (D && !getenv("SEMB_NO_FUSED_ADVECT")) {  // one fused kernel when nr == ns and nrd == nsd
    int done = 0;
    if (!getenv("SEMB_NO_TILED_ADVECT")) SEMB_TRY(semb_launch_advect_tile(c, V, D, 1, &T, ux, uy, w->dJr, w->dJs, &out, &done));
    if (done) return SEMB_OK;
    SEMB_TRY(semb_launch_advect_fused(c, V, D, T, ux, uy, w->dJr, w->dJs, out, &done));
    if (done) return SEMB_OK;
  }
  SEMB_TRY(advect_work_fields(w));
  SEMB_TRY(semb_launch_grad(c, V, T, w->Tx->d, w->Ty->d));  // Tx,Ty = grad(T,mshV)
  if (!D) return semb_launch_advect_pointwise(c, V, ux, w->Tx->d, uy, w->Ty->d, out);
  auto interp = [&](const double* in, double* outD) -> int {  // ABu(Js,Jr,in): Br = Jr first, then As = Js (ABu.jl:14-33)
    SEMB_TRY(semb_launch_abu_r(c, w->dJr, D->nr, V->nr, in, V->nxl, V->nyl, V->pitch, w->mid, D->pitch));
    return semb_launch_abu_s(c, w->dJs, D->ns, V->ns, w->mid, D->nxl, V->nyl, D->pitch, outD, D->pitch);
  };
  SEMB_TRY(interp(w->Tx->d, w->JTx->d));
  SEMB_TRY(interp(w->Ty->d, w->JTy->d));
  SEMB_TRY(interp(ux, w->Jux->d));
  SEMB_TRY(interp(uy, w->Juy->d));
  SEMB_TRY(semb_launch_advect_pointwise(c, D, w->Jux->d, w->JTx->d, w->Juy->d, w->JTy->d, w->JCu->d));  // :59-60
  // Cu = ABu(Js',Jr',JCu), advect.jl:61
  SEMB_TRY(semb_launch_abu_r(c, w->dJrT, V->nr, D->nr, w->JCu->d, D->nxl, D->nyl, D->pitch, w->mid, V->pitch));
  return semb_launch_abu_s(c, w->dJsT, V->ns, D->ns, w->mid, V->nxl, D->nyl, V->pitch, out, V->pitch);
}

extern "C" int semb_advect(semb_mesh* mV, semb_mesh* mD, const semb_field* T, const semb_field* ux, const semb_field* uy,
                           semb_field* out) {
  SEMB_REQUIRE(mV, "null mesh");
  SEMB_ENTER(mV->ctx);
  SEMB_TRY(check_field(mV, T, "advect(T)"));
  SEMB_TRY(check_field(mV, ux, "advect(ux)"));
  SEMB_TRY(check_field(mV, uy, "advect(uy)"));
  SEMB_TRY(check_field(mV, out, "advect(out)"));
  SEMB_REQUIRE(out != T && out != ux && out != uy, "advect: out must not alias an input");
  AdvectWork* w = nullptr;
  int rc = advect_work_create(mV, mD, &w);
  if (rc >= 0) rc = advect_run(w, T->d, ux->d, uy->d, out->d);
  cudaStreamSynchronize(mV->ctx->stream);
  advect_work_free(w);
  return rc;
}

// ---- device-resident Diffusion driver (diffusion.jl) --------------------------------------------------------------
struct semb_diffusion {
  semb_mesh* m = nullptr;
  char bc[5] = {0, 0, 0, 0, 0};
  int k = 3;
  double Ti = 0, Tf = 0, dt = 0;
  long long istep = 0;
  std::vector<double> time, bdfA, bdfB;
  semb_field *u = nullptr, *ub = nullptr, *nu = nullptr, *f = nullptr, *rhs = nullptr, *tmp = nullptr, *x = nullptr;
  std::vector<semb_field*> uh;
  // ConvectionDiffusion (convectionDiffusion.jl): advecting velocity, advect(uh[i]) and the dealiasing work space
  bool conv = false;
  semb_field *vx = nullptr, *vy = nullptr;
  std::vector<semb_field*> adv;
  AdvectWork* work = nullptr;
  int precond_kind = 0;  // 0: the reference's opM (identity for Diffusion, u./B./b0 for ConvectionDiffusion), 2: FDM
};

extern "C" int semb_diffusion_create(semb_mesh* m, const char bc[4], double Ti, double Tf, double dt, int k,
                                     semb_diffusion** out) {
  SEMB_REQUIRE(m && bc && out, "semb_diffusion_create: null argument");
  SEMB_REQUIRE(k >= 1 && k <= 4, "semb_diffusion_create: history length k must be 1..4 (got %d)", k);
  SEMB_REQUIRE(m->arr[SEMB_B] && m->arr[SEMB_G11], "semb_diffusion_create: mesh needs B and G factors");
  SEMB_ENTER(m->ctx);
  MaskFlags fl;
  SEMB_TRY(parse_bc(m, bc, &fl));
  semb_diffusion* d = new semb_diffusion();
  d->m = m;
  memcpy(d->bc, bc, 4);
  d->k = k;
  d->Ti = Ti;
  d->Tf = Tf;
  d->dt = dt;
  d->time.assign(k + 1, Ti);  // time.jl:87
  d->bdfA.assign(k, 0.0);
  d->bdfB.assign(k + 1, 0.0);
  SEMB_TRY(semb_bdf_ext_k(k + 1, d->time.data(), k, d->bdfA.data(), d->bdfB.data()));
  semb_field** all[] = {&d->u, &d->ub, &d->nu, &d->f, &d->rhs, &d->tmp, &d->x};
  for (semb_field** p : all) SEMB_TRY(semb_field_create(m, p));
  d->uh.resize(k);
  for (int i = 0; i < k; ++i) SEMB_TRY(semb_field_create(m, &d->uh[i]));
  *out = d;
  return SEMB_OK;
}

extern "C" int semb_convdiff_create(semb_mesh* mV, semb_mesh* mD, const char bc[4], double Ti, double Tf, double dt,
                                    int k, semb_diffusion** out) {
  SEMB_REQUIRE(mV && mD, "semb_convdiff_create: null mesh");
  semb_diffusion* d = nullptr;
  SEMB_TRY(semb_diffusion_create(mV, bc, Ti, Tf, dt, k, &d));
  d->conv = true;
  int rc = semb_field_create(mV, &d->vx);
  if (rc >= 0) rc = semb_field_create(mV, &d->vy);
  d->adv.assign(k, nullptr);
  for (int i = 0; i < k && rc >= 0; ++i) rc = semb_field_create(mV, &d->adv[i]);
  if (rc >= 0) rc = advect_work_create(mV, mD, &d->work);
  if (rc < 0) {
    semb_diffusion_destroy(d);
    return rc;
  }
  *out = d;
  return SEMB_OK;
}

extern "C" int semb_diffusion_destroy(semb_diffusion* d) {
  if (!d) return SEMB_OK;
  cudaStreamSynchronize(d->m->ctx->stream);
  semb_field_destroy(d->vx);
  semb_field_destroy(d->vy);
  for (semb_field* p : d->adv) semb_field_destroy(p);
  advect_work_free(d->work);
  semb_field* all[] = {d->u, d->ub, d->nu, d->f, d->rhs, d->tmp, d->x};
  for (semb_field* p : all) semb_field_destroy(p);
  for (semb_field* p : d->uh) semb_field_destroy(p);
  delete d;
  return SEMB_OK;
}

extern "C" int semb_diffusion_field(semb_diffusion* d, int which, semb_field** f) {
  SEMB_REQUIRE(d && f, "semb_diffusion_field: null argument");
  semb_field* base[] = {d->u, d->ub, d->nu, d->f, d->rhs, d->vx, d->vy};
  if (which >= 0 && which < 7) {
    SEMB_REQUIRE(base[which], "semb_diffusion_field: field %d exists only for ConvectionDiffusion", which);
    *f = base[which];
    return SEMB_OK;
  }
  SEMB_REQUIRE(which >= SEMB_DFN_UH0 && which < SEMB_DFN_UH0 + d->k, "semb_diffusion_field: bad selector %d", which);
  *f = d->uh[which - SEMB_DFN_UH0];
  return SEMB_OK;
}

extern "C" int semb_diffusion_begin_step(semb_diffusion* d, double* time, long long* istep) {
  SEMB_REQUIRE(d, "null diffusion");
  SEMB_ENTER(d->m->ctx);
  // updateHist!(fld), mesh.jl:207-215: uh[i] .= uh[i-1]; uh[1] .= u  (pointer rotation + one copy)
  semb_field* last = d->uh[d->k - 1];
  for (int i = d->k - 1; i >= 1; --i) d->uh[i] = d->uh[i - 1];
  d->uh[0] = last;
  SEMB_TRY(semb_field_copy(d->uh[0], d->u));
  // updateHist!(time), mesh.jl:217-224; istep += 1; time[1] += dt; bdfExtK!, diffusion.jl:92-95
  for (int i = d->k; i >= 1; --i) d->time[i] = d->time[i - 1];
  d->time[0] = d->time[1];
  d->istep += 1;
  d->time[0] += d->dt;
  SEMB_TRY(semb_bdf_ext_k(d->k + 1, d->time.data(), d->k, d->bdfA.data(), d->bdfB.data()));
  if (time) *time = d->time[0];
  if (istep) *istep = d->istep;
  return SEMB_OK;
}

extern "C" int semb_diffusion_finish_step(semb_diffusion* d, double tol, long long* iters, double* resinf) {
  SEMB_REQUIRE(d, "null diffusion");
  semb_mesh* m = d->m;
  semb_ctx* c = m->ctx;
  SEMB_ENTER(c);
  MaskFlags fl;
  SEMB_TRY(parse_bc(m, d->bc, &fl));
  // makeRHS!, diffusion.jl:51-65
  SEMB_TRY(semb_lapl(m, d->ub, d->tmp));
  const double* uh[4] = {nullptr, nullptr, nullptr, nullptr};
  double b[4] = {0, 0, 0, 0};
  for (int i = 0; i < d->k; ++i) {
    uh[i] = d->uh[i]->d;
    b[i] = d->bdfB[1 + i];
  }
  const double* adv[4] = {nullptr, nullptr, nullptr, nullptr};
  if (d->conv) {  // exH[i] = -advect(uh[i],vx,vy,mshV,mshD,JrVD,JsVD), convectionDiffusion.jl:102
    // all k history levels share vx, vy: one launch of the tiled kernel when it serves (nr, nrd)
    const double* Ts[4] = {nullptr, nullptr, nullptr, nullptr};
    double* outs[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < d->k; ++i) {
      Ts[i] = d->uh[i]->d;
      outs[i] = d->adv[i]->d;
      adv[i] = d->adv[i]->d;
    }
    int done = 0;
    if (d->work->D && d->k <= 4 && !getenv("SEMB_NO_FUSED_ADVECT") && !getenv("SEMB_NO_TILED_ADVECT"))
      SEMB_TRY(semb_launch_advect_tile(c, d->work->V, d->work->D, d->k, Ts, d->vx->d, d->vy->d, d->work->dJr,
                                       d->work->dJs, outs, &done));
    if (!done)
      for (int i = 0; i < d->k; ++i) SEMB_TRY(advect_run(d->work, Ts[i], d->vx->d, d->vy->d, outs[i]));
  }
  SEMB_TRY(semb_launch_rhs(c, m, d->f->d, d->nu->d, d->tmp->d, d->k, uh, b, d->conv ? adv : nullptr,
                           d->conv ? d->bdfA.data() : nullptr, fl.mx0, fl.mx1, fl.my0, fl.my1, d->x->d));
  SEMB_TRY(semb_gather_scatter(m, d->x, d->rhs));
  // solve!, diffusion.jl:67-77
  semb_pcg_opts o;
  memset(&o, 0, sizeof(o));
  o.nu = 1.0;
  o.nu_arr = d->nu;
  o.k = d->bdfB[0];
  o.bc = d->bc;
  o.tol = tol;
  o.maxiter = -1;
  if (d->conv) {  // opPrecond(u) = u ./ B ./ bdfB[1], convectionDiffusion.jl:87-91
    o.precond = 1;
    o.prec_b0 = d->bdfB[0];
  }
  if (d->precond_kind == 2) {
    // opt-in: the FDM preconditioner (lapl.jl:105-119) of nu*lapl + bdfB[1]*mass instead of the reference's opM.  It needs
    // a constant viscosity; the tables do not depend on (nu, k), so following the BDF start-up only resets two scalars
    bool cst = false;
    double v = 0.0;
    SEMB_TRY(field_constant(m, d->nu, &cst, &v));
    if (cst && v > 0.0 && d->bdfB[0] >= 0.0 && m->fast && m->nr >= 3 && (c->nranks == 1 || m->p2p)) {
      if (!m->fdm || strncmp(m->fdm_bc, d->bc, 4) != 0) {
        semb_fdm* h = nullptr;
        SEMB_TRY(semb_fdm_create(m, d->bc, v, d->bdfB[0], &h));
      } else {
        SEMB_TRY(semb_fdm_set_coeffs_impl(m->fdm, v, d->bdfB[0]));
      }
      o.precond = 2;
      o.prec_b0 = 1.0;
    }
  }
  int rc = semb_pcg(m, &o, d->rhs, d->x, iters, resinf);
  if (rc < 0) return rc;
  SEMB_TRY(semb_field_copy(d->u, d->x));
  SEMB_TRY(semb_field_axpby(1.0, d->ub, 1.0, d->u));  // u .+= ub
  return rc;
}

// kind 0: the reference's preconditioner (pcg.jl:37 opM = identity in diffusion.jl:71, opPrecond in convectionDiffusion.jl:118);
// kind 2: the FDM preconditioner (lapl.jl:105-119, semb_fdm_create) -- fewer iterations, other iterates within the tolerance
extern "C" int semb_diffusion_set_precond(semb_diffusion* d, int kind) {
  SEMB_REQUIRE(d, "null diffusion");
  SEMB_REQUIRE(kind == 0 || kind == 2, "semb_diffusion_set_precond: kind must be 0 (the reference's) or 2 (FDM)");
  d->precond_kind = kind;
  return SEMB_OK;
}

extern "C" int semb_diffusion_state(semb_diffusion* d, double* time, double* bdfA, double* bdfB, long long* istep) {
  SEMB_REQUIRE(d, "null diffusion");
  if (time) std::copy(d->time.begin(), d->time.end(), time);
  if (bdfA) std::copy(d->bdfA.begin(), d->bdfA.end(), bdfA);
  if (bdfB) std::copy(d->bdfB.begin(), d->bdfB.end(), bdfB);
  if (istep) *istep = d->istep;
  return SEMB_OK;
}

// ---- host-pointer twins -----------------------------------------------------------------------------------------
// Device fields backing the *_host twins are cached on the mesh (slot order = request order), so a
// twin call costs copies + kernels, not cudaMalloc/cudaFree.
struct TmpFields {
  semb_mesh* m;
  size_t next = 0;
  explicit TmpFields(semb_mesh* mm) : m(mm) {}
  int make(const double* host, semb_field** out) {
    *out = nullptr;
    if (next >= m->host_tmp.size()) {
      semb_field* p = nullptr;
      SEMB_TRY(semb_field_create(m, &p));
      m->host_tmp.push_back(p);
    }
    semb_field* p = m->host_tmp[next++];
    if (host) SEMB_TRY(semb_field_upload(p, host));
    *out = p;
    return SEMB_OK;
  }
  int maybe(const double* host, semb_field** out) {
    *out = nullptr;
    return host ? make(host, out) : SEMB_OK;
  }
};

extern "C" int semb_lapl_host(semb_mesh* m, const double* u, double* out) {
  return semb_hlmz_host(m, u, nullptr, 1.0, nullptr, 0.0, out);
}

extern "C" int semb_hlmz_host(semb_mesh* m, const double* u, const double* nu_arr, double nu, const double* k_arr,
                              double k, double* out) {
  SEMB_REQUIRE(m && u && out, "hlmz_host: null argument");
  TmpFields t(m);
  semb_field *fu, *fo, *fn, *fk;
  SEMB_TRY(t.make(u, &fu));
  SEMB_TRY(t.make(nullptr, &fo));
  SEMB_TRY(t.maybe(nu_arr, &fn));
  SEMB_TRY(t.maybe(k_arr, &fk));
  SEMB_TRY(semb_hlmz(m, fu, fn, nu, fk, k, fo));
  return semb_field_download(fo, out);
}

extern "C" int semb_mass_host(semb_mesh* m, const double* u, double* out) {
  SEMB_REQUIRE(m && u && out, "mass_host: null argument");
  TmpFields t(m);
  semb_field *fu, *fo;
  SEMB_TRY(t.make(u, &fu));
  SEMB_TRY(t.make(nullptr, &fo));
  SEMB_TRY(semb_mass(m, fu, fo));
  return semb_field_download(fo, out);
}

extern "C" int semb_gather_scatter_host(semb_mesh* m, const double* u, double* out) {
  SEMB_REQUIRE(m && u && out, "gatherScatter_host: null argument");
  TmpFields t(m);
  semb_field *fu, *fo;
  SEMB_TRY(t.make(u, &fu));
  SEMB_TRY(t.make(nullptr, &fo));
  SEMB_TRY(semb_gather_scatter(m, fu, fo));
  return semb_field_download(fo, out);
}

extern "C" int semb_mask_host(semb_mesh* m, const double* u, const double* M, double* out) {
  SEMB_REQUIRE(m && u && out, "mask_host: null argument");
  TmpFields t(m);
  semb_field *fu, *fo, *fm;
  SEMB_TRY(t.make(u, &fu));
  SEMB_TRY(t.make(nullptr, &fo));
  SEMB_TRY(t.maybe(M, &fm));
  SEMB_TRY(semb_mask(m, fu, fm, fo));
  return semb_field_download(fo, out);
}

// Pipelined form of the host twin for large single-rank meshes: the element rows are cut into slabs of chunks;
// slab s is uploaded (H2D stream) while slab s-1 is computed (strip + seam kernels restricted to the slab,
// compute stream) and slab s-2 is downloaded (D2H stream), so the PCIe link is busy in both directions.
// The operator is element-local and a slab's last line becomes final once the next slab's first line exists
// (the y seam between them), hence the one-slab lag of the download.  Same kernels, same bits.
static int oplhs_host_pipelined(semb_mesh* m, const double* u, double nu, double k, const char* bc, double* out,
                                semb_field* fu, semb_field* fo) {
  semb_ctx* c = m->ctx;
  const int N = m->ns;
  // fill + drain of the three-stage pipeline cost 2 slab times: 4 slabs 23.6 ms, 8: 20.8, 22: 19.3 ms per 1e8-DOF apply
  int want = 24;
  if (const char* e = getenv("SEMB_HOST_SLABS")) want = std::max(1, atoi(e));
  int nslab = std::min(want, m->nchunks);
  OpArgs a;
  fill_common(m, a);
  a.u = fu->d;
  a.out = fo->d;
  a.nu = nu;
  a.k = k;
  a.gs = 1;
  MaskFlags f;
  SEMB_TRY(parse_bc(m, bc, &f));
  a.mx0 = f.mx0;
  a.mx1 = f.mx1;
  a.my0 = f.my0;
  a.my1 = f.my1;
  a.partials = m->d_partials;
  a.counters = m->d_counters;
  const bool massterm = (k != 0.0);
  std::vector<cudaEvent_t> ev_in(nslab), ev_cmp(nslab);
  for (int s = 0; s < nslab; ++s) {
    SEMB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_in[s], cudaEventDisableTiming));
    SEMB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_cmp[s], cudaEventDisableTiming));
  }
  auto rows_of = [&](int s, int* y0, int* y1) {
    const int c0 = (int)((long long)s * m->nchunks / nslab), c1 = (int)((long long)(s + 1) * m->nchunks / nslab);
    *y0 = m->h_chunk_r0[c0] * N;
    *y1 = m->h_chunk_r0[c1] * N;
  };
  auto copy_rows = [&](bool h2d, int y0, int y1, cudaStream_t st) -> int {
    const size_t hw = (size_t)m->nxl * sizeof(double), dw = (size_t)m->pitch * sizeof(double);
    if (h2d)
      SEMB_CHECK_CUDA(cudaMemcpy2DAsync(fu->d + (size_t)y0 * m->pitch, dw, u + (size_t)y0 * m->nxl, hw, hw, y1 - y0,
                                        cudaMemcpyHostToDevice, st));
    else
      SEMB_CHECK_CUDA(cudaMemcpy2DAsync(out + (size_t)y0 * m->nxl, hw, fo->d + (size_t)y0 * m->pitch, dw, hw, y1 - y0,
                                        cudaMemcpyDeviceToHost, st));
    return SEMB_OK;
  };
  int rc = SEMB_OK;
  auto body = [&]() -> int {
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < nslab; ++s) {
      int y0, y1;
      rows_of(s, &y0, &y1);
      const int c0 = (int)((long long)s * m->nchunks / nslab), c1 = (int)((long long)(s + 1) * m->nchunks / nslab);
      SEMB_TRY(copy_rows(true, y0, y1, c->in_stream));
      SEMB_CHECK_CUDA(cudaEventRecord(ev_in[s], c->in_stream));
      SEMB_CHECK_CUDA(cudaStreamWaitEvent(c->stream, ev_in[s], 0));
      OpArgs b = a;
      b.chunk0 = c0;
      SEMB_TRY(semb_launch_strip(c, b, m->hDr.data(), m->hDs.data(), m->nstrips, c1 - c0, false, massterm, m->eo));
      b.y_begin = y0;
      b.y_end = y1;
      SEMB_TRY(semb_launch_seam_x(c, b));
      // y seams: q is the seam between chunk q and q+1; this slab closes seams c0-1 .. c1-2
      const int q0 = std::max(c0 - 1, 0), q1 = c1 - 1;
      if (q1 > q0) {
        OpArgs y = a;
        y.yseam = m->d_yseam + 2 * q0;
        y.nyseam = q1 - q0;
        SEMB_TRY(semb_launch_seam_y(c, y, 0, 0, true, P2PArgs()));
      }
      SEMB_CHECK_CUDA(cudaEventRecord(ev_cmp[s], c->stream));
      if (s > 0) {  // slab s-1 is final now (but for a line shared with a neighbour rank: that one goes last)
        int p0, p1;
        rows_of(s - 1, &p0, &p1);
        SEMB_CHECK_CUDA(cudaStreamWaitEvent(c->out_stream, ev_cmp[s], 0));
        SEMB_TRY(copy_rows(false, p0 + ((s == 1 && m->halo_lo) ? 1 : 0), p1, c->out_stream));
      }
    }
    int p0, p1;
    rows_of(nslab - 1, &p0, &p1);
    SEMB_CHECK_CUDA(cudaStreamWaitEvent(c->out_stream, ev_cmp[nslab - 1], 0));
    SEMB_TRY(copy_rows(false, p0 + ((nslab == 1 && m->halo_lo) ? 1 : 0), p1 - (m->halo_hi ? 1 : 0), c->out_stream));
    if (m->halo_lo || m->halo_hi) {
      // several ranks: the slab's first / last line is a y seam with the neighbour rank.  Both are x-complete now:
      // exchange them (peer memory or NCCL), close the two seams, download the two lines
      unsigned long long eph = 0;
      SEMB_TRY(halo_exchange(m, fo->d, 0, &eph));
      OpArgs y = a;
      y.halo_lo = m->d_halo_lo;
      y.halo_hi = m->d_halo_hi;
      if (m->p2p) {
        y.halo_lo = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 0);
        y.halo_hi = mail_halo(m, m->d_mailbox, (int)(eph & 1ull), 1);
      }
      y.yseam = m->d_yseam;
      y.nyseam = 0;
      SEMB_TRY(semb_launch_seam_y(c, y, m->halo_lo, m->halo_hi, true, p2p_args(m, eph)));
      SEMB_CHECK_CUDA(cudaEventRecord(ev_cmp[0], c->stream));
      SEMB_CHECK_CUDA(cudaStreamWaitEvent(c->out_stream, ev_cmp[0], 0));
      if (m->halo_lo) SEMB_TRY(copy_rows(false, 0, 1, c->out_stream));
      if (m->halo_hi) SEMB_TRY(copy_rows(false, m->nyl - 1, m->nyl, c->out_stream));
    }
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->out_stream));
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return SEMB_OK;
  };
  rc = body();
  if (rc < 0) {  // never leave copies in flight on the caller's buffers
    cudaStreamSynchronize(c->in_stream);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->out_stream);
  }
  for (int s = 0; s < nslab; ++s) {
    cudaEventDestroy(ev_in[s]);
    cudaEventDestroy(ev_cmp[s]);
  }
  return rc;
}

extern "C" int semb_oplhs_host(semb_mesh* m, const double* u, const double* nu_arr, double nu, const double* k_arr,
                               double k, const char* bc, const double* M_arr, double* out) {
  SEMB_REQUIRE(m && u && out, "oplhs_host: null argument");
  SEMB_ENTER(m->ctx);
  TmpFields t(m);
  semb_field *fu, *fo, *fn, *fk, *fm;
  // (several ranks: only rank-independent criteria -- the two paths exchange the halo rows differently)
  const bool pipe = m->fast && !m->pery && !nu_arr && !k_arr && !M_arr && (m->nchunks >= 4 || m->ctx->nranks > 1) &&
                    m->host_pipe && (k == 0.0 || m->arr[SEMB_B]) && !getenv("SEMB_NO_PIPELINE");
  if (pipe) {
    SEMB_TRY(t.make(nullptr, &fu));
    SEMB_TRY(t.make(nullptr, &fo));
    return oplhs_host_pipelined(m, u, nu, k, bc, out, fu, fo);
  }
  SEMB_TRY(t.make(u, &fu));
  SEMB_TRY(t.make(nullptr, &fo));
  SEMB_TRY(t.maybe(nu_arr, &fn));
  SEMB_TRY(t.maybe(k_arr, &fk));
  SEMB_TRY(t.maybe(M_arr, &fm));
  SEMB_TRY(semb_oplhs(m, fu, fn, nu, fk, k, bc, fm, fo));
  return semb_field_download(fo, out);
}

extern "C" int semb_pcg_host(semb_mesh* m, const semb_pcg_opts* o, const double* nu_arr, const double* k_arr,
                             const double* M_arr, const double* b, double* x, long long* iters, double* resinf) {
  SEMB_REQUIRE(m && o && b && x, "pcg_host: null argument");
  TmpFields t(m);
  semb_field *fb, *fx, *fn, *fk, *fm;
  SEMB_TRY(t.make(b, &fb));
  SEMB_TRY(t.make(nullptr, &fx));
  SEMB_TRY(t.maybe(nu_arr, &fn));
  SEMB_TRY(t.maybe(k_arr, &fk));
  SEMB_TRY(t.maybe(M_arr, &fm));
  semb_pcg_opts oo = *o;
  oo.nu_arr = fn;
  oo.k_arr = fk;
  oo.M_arr = fm;
  int rc = semb_pcg(m, &oo, fb, fx, iters, resinf);
  if (rc < 0) return rc;
  SEMB_TRY(semb_field_download(fx, x));
  return rc;
}

extern "C" int semb_abu_host(semb_ctx* c, const double* As, int ma, int na, const double* Br, int mb, int nb,
                             const double* u, int mrows, int ncols, double* out) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(u && out && mrows >= 1 && ncols >= 1, "ABu: bad u");
  const bool hasB = Br && mb > 0 && nb > 0, hasA = As && ma > 0 && na > 0;
  // Julia: Int(m*mb/nb) / Int(Ey*ma) throw InexactError when not integral (ABu.jl:16,26)
  SEMB_REQUIRE(!hasB || mrows % nb == 0, "ABu: InexactError: rows %d not a multiple of size(Br,2)=%d", mrows, nb);
  SEMB_REQUIRE(!hasA || ncols % na == 0, "ABu: InexactError: cols %d not a multiple of size(As,2)=%d", ncols, na);
  const int m1 = hasB ? mrows / nb * mb : mrows;
  const int n1 = hasA ? ncols / na * ma : ncols;
  double *du = nullptr, *dt = nullptr, *dout = nullptr, *dA = nullptr, *dB = nullptr;
  int rc = SEMB_OK;
  auto body = [&]() -> int {
    SEMB_CHECK_CUDA(cudaMalloc(&du, (size_t)mrows * ncols * sizeof(double)));
    SEMB_CHECK_CUDA(cudaMemcpyAsync(du, u, (size_t)mrows * ncols * sizeof(double), cudaMemcpyHostToDevice,
                                    c->stream));
    const double* cur = du;
    if (hasB) {
      SEMB_CHECK_CUDA(cudaMalloc(&dB, (size_t)mb * nb * sizeof(double)));
      SEMB_CHECK_CUDA(cudaMemcpyAsync(dB, Br, (size_t)mb * nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      SEMB_CHECK_CUDA(cudaMalloc(&dt, (size_t)m1 * ncols * sizeof(double)));
      SEMB_TRY(semb_launch_abu_r(c, dB, mb, nb, cur, mrows, ncols, mrows, dt, m1));
      cur = dt;
    }
    if (hasA) {
      SEMB_CHECK_CUDA(cudaMalloc(&dA, (size_t)ma * na * sizeof(double)));
      SEMB_CHECK_CUDA(cudaMemcpyAsync(dA, As, (size_t)ma * na * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      SEMB_CHECK_CUDA(cudaMalloc(&dout, (size_t)m1 * n1 * sizeof(double)));
      SEMB_TRY(semb_launch_abu_s(c, dA, ma, na, cur, m1, ncols, m1, dout, m1));
      cur = dout;
    }
    SEMB_CHECK_CUDA(cudaMemcpyAsync(out, cur, (size_t)m1 * n1 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return SEMB_OK;
  };
  rc = body();
  cudaFree(du);
  cudaFree(dt);
  cudaFree(dout);
  cudaFree(dA);
  cudaFree(dB);
  return rc;
}

// ---- FDM preconditioner (lapl.jl:105-119, examples/p2d_explicit.jl:109-141; kernels in semb_fdm.cu) ------
extern "C" int semb_fdm_create(semb_mesh* m, const char bc[4], double nu, double k, semb_fdm** out) {
  SEMB_REQUIRE(m && out, "semb_fdm_create: null argument");
  *out = nullptr;
  SEMB_ENTER(m->ctx);
  SEMB_REQUIRE(m->nr == m->ns && m->nr >= 3 && m->nr <= SEMB_MAXN, "fdm: needs nr == ns in 3..%d", SEMB_MAXN);
  SEMB_REQUIRE(m->arr[SEMB_B] && m->arr[SEMB_G11] && m->arr[SEMB_G22], "fdm: mesh needs B, G11, G22");
  SEMB_REQUIRE(!(m->perx && m->Ex < 2) && !(m->pery && m->Ey < 2), "fdm: a periodic direction needs at least 2 elements");
  SEMB_REQUIRE(nu > 0.0 && k >= 0.0, "fdm: needs nu > 0, k >= 0");
  MaskFlags f;
  SEMB_TRY(parse_bc(m, bc, &f));
  const int gy0 = (bc && bc[2] == 'D' && !m->pery) ? 1 : 0, gy1 = (bc && bc[3] == 'D' && !m->pery) ? 1 : 0;
  semb_fdm_free_impl(m->fdm);  // one per mesh: the new one replaces it
  m->fdm = nullptr;
  semb_fdm* h = nullptr;
  const int rc = semb_fdm_create_impl(m, nu, k, f.mx0, f.mx1, f.my0, f.my1, gy0, gy1, &h);
  if (rc < 0) {
    semb_fdm_free_impl(h);
    return rc;
  }
  m->fdm = h;
  memcpy(m->fdm_bc, bc ? bc : "NNNN", 4);
  *out = h;
  return SEMB_OK;
}

extern "C" int semb_fdm_destroy(semb_mesh* m) {
  SEMB_REQUIRE(m, "null mesh");
  SEMB_ENTER(m->ctx);
  SEMB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  semb_fdm_free_impl(m->fdm);
  m->fdm = nullptr;
  return SEMB_OK;
}

extern "C" int semb_fdm_apply(semb_mesh* m, const semb_field* r, semb_field* out) {
  SEMB_REQUIRE(m && m->fdm, "semb_fdm_apply: no FDM preconditioner on this mesh (semb_fdm_create)");
  SEMB_ENTER(m->ctx);
  SEMB_TRY(check_field(m, r, "fdm(r)"));
  SEMB_TRY(check_field(m, out, "fdm(out)"));
  SEMB_REQUIRE(r != out, "fdm: out must not alias r");
  return semb_fdm_apply_impl(m->fdm, r->d, out->d, 0);
}

extern "C" int semb_fdm_apply_host(semb_mesh* m, const double* r, double* out) {
  SEMB_REQUIRE(m && r && out, "fdm_host: null argument");
  TmpFields t(m);
  semb_field *fr, *fo;
  SEMB_TRY(t.make(r, &fr));
  SEMB_TRY(t.make(nullptr, &fo));
  SEMB_TRY(semb_fdm_apply(m, fr, fo));
  return semb_field_download(fo, out);
}

#include "semb_stokes_api.cuh"
#include "semb_explicit_api.cuh"
