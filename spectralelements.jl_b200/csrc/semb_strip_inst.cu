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

template <bool PCGM, bool MASS>
int launch_variant(semb_ctx* ctx, const OpArgs& a, const double* hDr, const double* hDs, int nstrips,
                   int nchunks) {
  using C = StripCfg<N>;
  StripParams<N> P;
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < N; ++k) {
      P.Dr[i * N + k] = hDr[i + k * N];  // host copies are column-major
      P.Ds[i * N + k] = hDs[i + k * N];
    }
  P.a = a;
  auto kern = semb_strip_kernel<N, PCGM, MASS>;
  static bool attr_done = false;
  if (!attr_done) {
    SEMB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_done = true;
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
  auto kern = semb_strip_kernel<N, PCGM, MASS>;
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
                                               bool massterm) {
  if (pcg) {
    return massterm ? launch_variant<true, true>(ctx, a, hDr, hDs, nstrips, nchunks)
                    : launch_variant<true, false>(ctx, a, hDr, hDs, nstrips, nchunks);
  }
  return massterm ? launch_variant<false, true>(ctx, a, hDr, hDs, nstrips, nchunks)
                  : launch_variant<false, false>(ctx, a, hDr, hDs, nstrips, nchunks);
}

int SEMB_CAT(semb_strip_attr_n, SEMB_INST_N)(bool pcg, bool massterm, int* regs, int* smem, int* occ) {
  if (pcg) return massterm ? attr_variant<true, true>(regs, smem, occ) : attr_variant<true, false>(regs, smem, occ);
  return massterm ? attr_variant<false, true>(regs, smem, occ) : attr_variant<false, false>(regs, smem, occ);
}
