// amx_logpdf.cu -- K4: batched mixture / MVN-Cholesky log-density.
// Replaces lnormprob (automix.c:1727-1750) and det (:1752-1761) for n points at once; the
// same device routine (solve_lower, amx_targets.cuh) is what K2 and K3 call per sample.
#include <vector>

#include "amx_internal.cuh"
#include "amx_targets.cuh"

namespace amx {

constexpr int kPdfThreads = 256;

// One thread per point.  The component records (one "model" of a proposal family blob) are
// staged in shared memory; x is read row-major (a warp touches 32*d consecutive doubles, so
// every fetched sector is consumed).  comp_out is n x L row-major, mix_out is
// log(sum_l exp(log wt_l + lpd_l)) formed the way the reference forms it (:847-859).
template <int DMAX>
__global__ void __launch_bounds__(kPdfThreads) mix_logpdf_kernel(const void *blob, int blob_bytes, long n,
                                                                 const double *__restrict__ x,
                                                                 double *__restrict__ comp_out,
                                                                 double *__restrict__ mix_out) {
  extern __shared__ double smem[];
  const double *src = reinterpret_cast<const double *>(blob);
  for (int i = threadIdx.x; i < blob_bytes / 8; i += blockDim.x) smem[i] = src[i];
  __syncthreads();
  const amx_fam_hdr *h = reinterpret_cast<const amx_fam_hdr *>(smem);
  const double *D = reinterpret_cast<const double *>(h + 1);
  const int d = h->dims[0], L = h->ncomp[0], st = h->stride[0];
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double v[DMAX], r[DMAX];
#pragma unroll
    for (int j = 0; j < DMAX; j++) v[j] = (j < d) ? x[i * d + j] : 0.0;
    double s = 0.0;
    for (int l = 0; l < L; l++) {
      const double *rec = D + h->off[0] + l * st;
      const double lpd = fma(-0.5, solve_lower<DMAX>(rec, d, v, r), rec[3]);
      if (comp_out) comp_out[i * L + l] = lpd;
      s += exp(rec[1] + lpd);
    }
    if (mix_out) mix_out[i] = log(s);
  }
}

static int launch_logpdf(int d, int L, const double *wt, const double *mean, const double *tri, long n,
                         const double *x_dev, double *comp_dev, double *mix_dev) {
  if (d < 1 || d > AMX_MAX_DIM || L < 1 || L > AMX_MAX_COMPS || n < 1)
    return fail(AMX_EINVAL, "amx_mix_logpdf: d=%d L=%d n=%ld out of range", d, L, n);
  amx_fam_hdr h;
  int zero = 0;
  if (amx_fam_plan(&h, 1, &d, &L, &zero) < 0) return fail(AMX_EINVAL, "amx_mix_logpdf: bad shape");
  std::vector<double> data(h.total);
  amx_fam_pack(&h, AMX_FAM_PROPOSAL, wt, mean, tri, nullptr, data.data());
  void *blob = nullptr;
  int bytes = 0;
  if (int rc = upload_blob(h, data.data(), &blob, &bytes)) return rc;
  int dev = 0, sms = 0;
  AMX_CUDA(cudaGetDevice(&dev));
  AMX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long want = (n + kPdfThreads - 1) / kPdfThreads;
  const unsigned grid = (unsigned)(want < (long)sms * 8 ? want : (long)sms * 8);
  if (d <= 8) {
    auto kern = mix_logpdf_kernel<8>;
    if (bytes > 48 * 1024) AMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    kern<<<grid, kPdfThreads, bytes, stream()>>>(blob, bytes, n, x_dev, comp_dev, mix_dev);
  } else {
    auto kern = mix_logpdf_kernel<AMX_MAX_DIM>;
    if (bytes > 48 * 1024) AMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    kern<<<grid, kPdfThreads, bytes, stream()>>>(blob, bytes, n, x_dev, comp_dev, mix_dev);
  }
  count_launch();
  AMX_CUDA(cudaGetLastError());
  AMX_CUDA(cudaStreamSynchronize(stream()));
  cudaFree(blob);
  return AMX_OK;
}

}  // namespace amx

using namespace amx;

extern "C" {

int amx_mix_logpdf_dev(int d, int L, const double *wt, const double *mean, const double *tri, long n,
                       const double *x_dev, double *comp_out_dev, double *mix_out_dev) {
  if (int rc = require_device()) return rc;
  return launch_logpdf(d, L, wt, mean, tri, n, x_dev, comp_out_dev, mix_out_dev);
}

int amx_mix_logpdf(int d, int L, const double *wt, const double *mean, const double *tri, long n,
                   const double *x, double *comp_out, double *mix_out) {
  if (int rc = require_device()) return rc;
  if (n < 1 || d < 1 || L < 1) return fail(AMX_EINVAL, "amx_mix_logpdf: empty input");
  double *x_dev = nullptr, *c_dev = nullptr, *m_dev = nullptr;
  AMX_CUDA(cudaMalloc(&x_dev, sizeof(double) * (size_t)n * d));
  if (comp_out) AMX_CUDA(cudaMalloc(&c_dev, sizeof(double) * (size_t)n * L));
  if (mix_out) AMX_CUDA(cudaMalloc(&m_dev, sizeof(double) * (size_t)n));
  AMX_CUDA(cudaMemcpyAsync(x_dev, x, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, stream()));
  int rc = launch_logpdf(d, L, wt, mean, tri, n, x_dev, c_dev, m_dev);
  if (rc == AMX_OK) {
    if (comp_out) AMX_CUDA(cudaMemcpyAsync(comp_out, c_dev, sizeof(double) * (size_t)n * L, cudaMemcpyDeviceToHost, stream()));
    if (mix_out) AMX_CUDA(cudaMemcpyAsync(mix_out, m_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, stream()));
    AMX_CUDA(cudaStreamSynchronize(stream()));
  }
  cudaFree(x_dev);
  cudaFree(c_dev);
  cudaFree(m_dev);
  return rc;
}

}  // extern "C"
