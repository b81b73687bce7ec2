// One translation unit per polynomial size: compiled with -DSEMB_INST_N=<nr> (see build.py) so the
// fully unrolled strip kernels build in parallel.
#include "semb_strip.cuh"

#ifndef SEMB_INST_N
#error "compile with -DSEMB_INST_N=<n>"
#endif

#define SEMB_CAT2(a, b) a##b
#define SEMB_CAT(a, b) SEMB_CAT2(a, b)

namespace {
constexpr int N = SEMB_INST_N;

// row-major copies of Dr, Ds and their transposes from the column-major host matrices
void row_major_set(const double* hDr, const double* hDs, double (&A)[4][N * N]) {
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < N; ++k) {
      A[0][i * N + k] = hDs[i + k * N];  // A1 = Ds
      A[1][i * N + k] = hDr[i + k * N];  // A2 = Dr
      A[2][i * N + k] = hDs[k + i * N];  // A3 = Ds^T
      A[3][i * N + k] = hDr[k + i * N];  // A4 = Dr^T
    }
}

template <bool PCGM, bool MASS, bool EO>
int launch_variant(semb_ctx* ctx, const OpArgs& a, const double* hDr, const double* hDs, int nstrips,
                   int nchunks) {
  using C = StripCfg<N>;
  StripParams<N> P;
  double A[4][N * N];
  row_major_set(hDr, hDs, A);
  for (int q = 0; q < 4; ++q) StripTab<N>::fill(A[q], EO, P.tab[q]);
  P.a = a;
  auto kern = semb_strip_kernel<N, PCGM, MASS, EO>;
  static bool attr_done[64] = {false};  // per device: function attributes belong to the device's context
  const int dev = ctx->device & 63;
  if (!attr_done[dev]) {
    SEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done[dev] = true;
  }
  const bool prof = ctx->profile && ctx->prof_used + 2 <= ctx->prof_ev.size();
  if (prof) SEMB_CHECK_CUDA(cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
  kern<<<dim3(nstrips, nchunks), C::T, C::SMEM, ctx->stream>>>(P);
  SEMB_CHECK_CUDA(cudaGetLastError());
  if (prof) {
    SEMB_CHECK_CUDA(cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
    ctx->prof_used += 2;
  }
  ctx->launches++;
  return SEMB_OK;
}

template <bool PCGM, bool MASS>
int attr_variant(int* regs, int* smem, int* occ) {
  using C = StripCfg<N>;
  auto kern = semb_strip_kernel<N, PCGM, MASS, true>;
  cudaFuncAttributes fa;
  SEMB_CHECK_CUDA(cudaFuncGetAttributes(&fa, kern));
  SEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  int nb = 0;
  SEMB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::T, C::SMEM));
  if (regs) *regs = fa.numRegs;
  if (smem) *smem = C::SMEM + (int)fa.sharedSizeBytes;
  if (occ) *occ = nb;
  return SEMB_OK;
}
}  // namespace

int SEMB_CAT(semb_launch_strip_n, SEMB_INST_N)(semb_ctx* ctx, const OpArgs& a, const double* hDr,
                                               const double* hDs, int nstrips, int nchunks, bool pcg,
                                               bool massterm, bool eo) {
#define SEMB_GO(P_, M_, E_) return launch_variant<P_, M_, E_>(ctx, a, hDr, hDs, nstrips, nchunks)
  if (eo) {
    if (pcg) { if (massterm) SEMB_GO(true, true, true); else SEMB_GO(true, false, true); }
    if (massterm) SEMB_GO(false, true, true); else SEMB_GO(false, false, true);
  }
  if (pcg) { if (massterm) SEMB_GO(true, true, false); else SEMB_GO(true, false, false); }
  if (massterm) SEMB_GO(false, true, false); else SEMB_GO(false, false, false);
#undef SEMB_GO
}

// largest relative centro-antisymmetry defect of Dr and Ds (decides the even-odd kernel variant)
double SEMB_CAT(semb_strip_defect_n, SEMB_INST_N)(const double* hDr, const double* hDs) {
  double A[4][N * N];
  row_major_set(hDr, hDs, A);
  const double d0 = StripTab<N>::antisymmetry_defect(A[0]), d1 = StripTab<N>::antisymmetry_defect(A[1]);
  return d0 > d1 ? d0 : d1;
}

int SEMB_CAT(semb_strip_attr_n, SEMB_INST_N)(bool pcg, bool massterm, int* regs, int* smem, int* occ) {
  if (pcg) return massterm ? attr_variant<true, true>(regs, smem, occ) : attr_variant<true, false>(regs, smem, occ);
  return massterm ? attr_variant<false, true>(regs, smem, occ) : attr_variant<false, false>(regs, smem, occ);
}
