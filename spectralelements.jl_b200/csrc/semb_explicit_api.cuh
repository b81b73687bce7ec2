// Mesh-free host twins of the explicit-argument operator forms (textually included at the end of semb_api.cu):
//   laplace(u,Dr,Ds,G11,G12,G22)              lapl.jl:70-81
//   laplace(u,Jr,Js,Dr,Ds,G11,G12,G22)        lapl.jl:83-103  (dealiased: G on the finer grid)
//   mass(u,M,B,Jr,Js,QQtx,QQty,mult) core     mass.jl:32-50   (Bu = ABu(Js',Jr', B .* ABu(Js,Jr,u)))
//   a .* b                                    mask.jl:14 / the `mult` hook (lapl.jl:62, mass.jl:44) for plain arrays
// as examples/p2d_explicit.jl:183-188 and examples/semPS.jl:168-172 call them with hand-built operator arrays that belong
// to no Mesh.  Everything runs on the device (generic ABu kernels + two pointwise kernels); arrays are dense column-major.
#pragma once

namespace {

// vr = G11.*a + G12.*b ; vs = G12.*a + G22.*b   (lapl.jl:75-76 / :94-95), written over a and b
__global__ void semb_gmix_kernel(const double* __restrict__ G11, const double* __restrict__ G12,
                                 const double* __restrict__ G22, double* a, double* b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double ur = a[i], us = b[i], g12 = G12[i];
    a[i] = __dadd_rn(__dmul_rn(G11[i], ur), __dmul_rn(g12, us));
    b[i] = __dadd_rn(__dmul_rn(g12, ur), __dmul_rn(G22[i], us));
  }
}
// out = a .* b, or out = a + b
__global__ void semb_pointwise2_kernel(const double* a, const double* b, double* out, size_t n, int add) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = add ? __dadd_rn(a[i], b[i]) : __dmul_rn(a[i], b[i]);
}

struct DevBufs {  // device scratch of one call, freed on scope exit
  std::vector<void*> p;
  ~DevBufs() {
    for (void* q : p) cudaFree(q);
  }
  int alloc(double** d, size_t n) {
    SEMB_CHECK_CUDA(cudaMalloc(d, (n ? n : 1) * sizeof(double)));
    p.push_back(*d);
    return SEMB_OK;
  }
  int upload(semb_ctx* c, double** d, const double* h, size_t n) {
    SEMB_TRY(alloc(d, n));
    SEMB_CHECK_CUDA(cudaMemcpyAsync(*d, h, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return SEMB_OK;
  }
};

int explicit_grid(semb_ctx* c, size_t n) {
  size_t b = (n + 255) / 256, cap = (size_t)c->sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// column-major transpose on the host (small operator matrices)
std::vector<double> transposed(const double* A, int rows, int cols) {
  std::vector<double> T((size_t)rows * cols);
  for (int i = 0; i < rows; ++i)
    for (int k = 0; k < cols; ++k) T[k + (size_t)i * cols] = A[i + (size_t)k * rows];
  return T;
}

}  // namespace

extern "C" int semb_laplace_host(semb_ctx* c, int m, int n, const double* Dr, int nr, const double* Ds, int ns,
                                 const double* Jr, int nrd, const double* Js, int nsd, const double* G11,
                                 const double* G12, const double* G22, const double* u, double* out) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(u && out && Dr && Ds && G11 && G12 && G22 && m >= 1 && n >= 1 && nr >= 1 && ns >= 1, "laplace: bad argument");
  SEMB_REQUIRE(m % nr == 0 && n % ns == 0, "laplace: InexactError: %d x %d is not a whole number of %d x %d elements", m, n,
               nr, ns);
  const bool dealias = Jr && Js && nrd > 0 && nsd > 0;
  SEMB_REQUIRE(dealias || (!Jr && !Js), "laplace: pass both Jr and Js, or neither");
  const int md = dealias ? m / nr * nrd : m, nd = dealias ? n / ns * nsd : n;  // grid the G factors live on
  const size_t nv = (size_t)m * n, ndd = (size_t)md * nd;
  DevBufs B;
  double *du, *dDr, *dDs, *dDrT, *dDsT, *dG11, *dG12, *dG22, *ur, *us, *t1, *a, *b, *dJr = nullptr, *dJs = nullptr,
                                                                                    *dJrT = nullptr, *dJsT = nullptr;
  SEMB_TRY(B.upload(c, &du, u, nv));
  SEMB_TRY(B.upload(c, &dDr, Dr, (size_t)nr * nr));
  SEMB_TRY(B.upload(c, &dDs, Ds, (size_t)ns * ns));
  const std::vector<double> hDrT = transposed(Dr, nr, nr), hDsT = transposed(Ds, ns, ns);
  SEMB_TRY(B.upload(c, &dDrT, hDrT.data(), hDrT.size()));
  SEMB_TRY(B.upload(c, &dDsT, hDsT.data(), hDsT.size()));
  SEMB_TRY(B.upload(c, &dG11, G11, ndd));
  SEMB_TRY(B.upload(c, &dG12, G12, ndd));
  SEMB_TRY(B.upload(c, &dG22, G22, ndd));
  SEMB_TRY(B.alloc(&ur, nv));
  SEMB_TRY(B.alloc(&us, nv));
  SEMB_TRY(semb_launch_abu_r(c, dDr, nr, nr, du, m, n, m, ur, m));  // ur = ABu([],Dr,u)
  SEMB_TRY(semb_launch_abu_s(c, dDs, ns, ns, du, m, n, m, us, m));  // us = ABu(Ds,[],u)
  std::vector<double> hJrT, hJsT;
  if (dealias) {
    SEMB_TRY(B.upload(c, &dJr, Jr, (size_t)nrd * nr));
    SEMB_TRY(B.upload(c, &dJs, Js, (size_t)nsd * ns));
    hJrT = transposed(Jr, nrd, nr);
    hJsT = transposed(Js, nsd, ns);
    SEMB_TRY(B.upload(c, &dJrT, hJrT.data(), hJrT.size()));
    SEMB_TRY(B.upload(c, &dJsT, hJsT.data(), hJsT.size()));
    SEMB_TRY(B.alloc(&t1, (size_t)md * std::max(n, nd)));
    SEMB_TRY(B.alloc(&a, ndd));
    SEMB_TRY(B.alloc(&b, ndd));
    SEMB_TRY(semb_launch_abu_r(c, dJr, nrd, nr, ur, m, n, m, t1, md));  // Jur = ABu(Js,Jr,ur), lapl.jl:91
    SEMB_TRY(semb_launch_abu_s(c, dJs, nsd, ns, t1, md, n, md, a, md));
    SEMB_TRY(semb_launch_abu_r(c, dJr, nrd, nr, us, m, n, m, t1, md));  // Jus, :92
    SEMB_TRY(semb_launch_abu_s(c, dJs, nsd, ns, t1, md, n, md, b, md));
  } else {
    a = ur;
    b = us;
  }
  semb_gmix_kernel<<<explicit_grid(c, ndd), 256, 0, c->stream>>>(dG11, dG12, dG22, a, b, ndd);  // :75-76 / :94-95
  SEMB_CHECK_CUDA(cudaGetLastError());
  c->launches++;
  if (dealias) {  // wr = ABu(Js',Jr',vr), ws = ABu(Js',Jr',vs), :97-98
    SEMB_TRY(semb_launch_abu_r(c, dJrT, nr, nrd, a, md, nd, md, t1, m));
    SEMB_TRY(semb_launch_abu_s(c, dJsT, ns, nsd, t1, m, nd, m, ur, m));
    SEMB_TRY(semb_launch_abu_r(c, dJrT, nr, nrd, b, md, nd, md, t1, m));
    SEMB_TRY(semb_launch_abu_s(c, dJsT, ns, nsd, t1, m, nd, m, us, m));
  }
  // Au = ABu([],Dr',wr) + ABu(Ds',[],ws), :78 / :100  (du is free again)
  double* w2;
  SEMB_TRY(B.alloc(&w2, nv));
  SEMB_TRY(semb_launch_abu_r(c, dDrT, nr, nr, ur, m, n, m, du, m));
  SEMB_TRY(semb_launch_abu_s(c, dDsT, ns, ns, us, m, n, m, w2, m));
  semb_pointwise2_kernel<<<explicit_grid(c, nv), 256, 0, c->stream>>>(du, w2, ur, nv, 1);
  SEMB_CHECK_CUDA(cudaGetLastError());
  c->launches++;
  SEMB_CHECK_CUDA(cudaMemcpyAsync(out, ur, nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return SEMB_OK;
}

extern "C" int semb_mass_explicit_host(semb_ctx* c, int m, int n, const double* Jr, int nrd, int nr, const double* Js,
                                       int nsd, int ns, const double* Bm, const double* u, double* out) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(u && out && m >= 1 && n >= 1, "mass: bad argument");
  const bool hasJr = Jr && nrd > 0 && nr > 0, hasJs = Js && nsd > 0 && ns > 0;
  SEMB_REQUIRE(!hasJr || m % nr == 0, "mass: InexactError: rows %d not a multiple of size(Jr,2)=%d", m, nr);
  SEMB_REQUIRE(!hasJs || n % ns == 0, "mass: InexactError: cols %d not a multiple of size(Js,2)=%d", n, ns);
  const int md = hasJr ? m / nr * nrd : m, nd = hasJs ? n / ns * nsd : n;
  const size_t nv = (size_t)m * n, ndd = (size_t)md * nd;
  DevBufs B;
  double *cur, *t, *dJ, *dB;
  SEMB_TRY(B.upload(c, &cur, u, nv));
  if (hasJr) {  // Ju = ABu(Js,Jr,u), mass.jl:36
    SEMB_TRY(B.upload(c, &dJ, Jr, (size_t)nrd * nr));
    SEMB_TRY(B.alloc(&t, (size_t)md * n));
    SEMB_TRY(semb_launch_abu_r(c, dJ, nrd, nr, cur, m, n, m, t, md));
    cur = t;
  }
  if (hasJs) {
    SEMB_TRY(B.upload(c, &dJ, Js, (size_t)nsd * ns));
    SEMB_TRY(B.alloc(&t, ndd));
    SEMB_TRY(semb_launch_abu_s(c, dJ, nsd, ns, cur, md, n, md, t, md));
    cur = t;
  }
  if (Bm) {  // BJu = B .* Ju, :38-40 (length(B)==0 keeps Ju)
    SEMB_TRY(B.upload(c, &dB, Bm, ndd));
    semb_pointwise2_kernel<<<explicit_grid(c, ndd), 256, 0, c->stream>>>(dB, cur, cur, ndd, 0);
    SEMB_CHECK_CUDA(cudaGetLastError());
    c->launches++;
  }
  if (hasJr) {  // Bu = ABu(Js',Jr',BJu), :42
    const std::vector<double> hT = transposed(Jr, nrd, nr);
    SEMB_TRY(B.upload(c, &dJ, hT.data(), hT.size()));
    SEMB_TRY(B.alloc(&t, (size_t)m * nd));
    SEMB_TRY(semb_launch_abu_r(c, dJ, nr, nrd, cur, md, nd, md, t, m));
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));  // hT leaves scope
    cur = t;
  }
  if (hasJs) {
    const std::vector<double> hT = transposed(Js, nsd, ns);
    SEMB_TRY(B.upload(c, &dJ, hT.data(), hT.size()));
    SEMB_TRY(B.alloc(&t, nv));
    SEMB_TRY(semb_launch_abu_s(c, dJ, ns, nsd, cur, m, nd, m, t, m));
    SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    cur = t;
  }
  SEMB_CHECK_CUDA(cudaMemcpyAsync(out, cur, nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return SEMB_OK;
}

extern "C" int semb_mul_host(semb_ctx* c, size_t n, const double* a, const double* b, double* out) {
  SEMB_ENTER(c);
  SEMB_REQUIRE(a && b && out, "mul: null argument");
  if (n == 0) return SEMB_OK;
  DevBufs B;
  double *da, *db;
  SEMB_TRY(B.upload(c, &da, a, n));
  SEMB_TRY(B.upload(c, &db, b, n));
  semb_pointwise2_kernel<<<explicit_grid(c, n), 256, 0, c->stream>>>(da, db, da, n, 0);
  SEMB_CHECK_CUDA(cudaGetLastError());
  c->launches++;
  SEMB_CHECK_CUDA(cudaMemcpyAsync(out, da, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return SEMB_OK;
}
