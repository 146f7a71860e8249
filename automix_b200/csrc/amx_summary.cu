// amx_summary.cu -- posterior summaries on the device.
//
// What a LibAutoMix user reads out of runStats after rjmcmc_samples is (a) the integrated
// autocorrelation time of the model-index series (the reference's report writer runs Sokal's
// adaptive truncated periodogram over runStats.xr: user_examples/logwrite.c:354-403, series
// recorded at automix.c:122-124, length chosen at :367-371) and (b) per-model sample means and
// covariances formed from theta_summary (automix.c:105-120).  With a population of 1e6 chains the
// per-sweep rows those are computed from cannot be materialised, so both are computed where the
// state lives:
//
//  * sokal_kernel: one CTA per series.  The reference takes two length-n FFTs to get the circular
//    autocovariance at ALL n lags and then reads the first m ~ 3 tau of them; here the CTA forms the
//    lags it needs directly, 16 at a time with the window test in between, and stops where the
//    reference's loop stops: O(n m) multiply-adds from L1 instead of O(n log n) with two passes over
//    scratch memory, no scratch, no bit reversal.  Same definition (mean removed, circular,
//    normalised by lag 0), so var, tau agree to rounding and m exactly.
//  * moments_kernel: per-model count, sum of log-posterior, and pivot-shifted first and second
//    moments of the population's current states, reduced in a fixed order (bitwise reproducible),
//    added into running totals so that a caller can accumulate over thinned sweeps.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "amx_internal.cuh"
#include "amx_layout.h"

namespace amx {

// ---- Sokal integrated autocorrelation time -------------------------------------------------------
constexpr int kSkThreads = 256;
constexpr int kSkWarps = kSkThreads / 32;
constexpr int kSkLags = 16;

__device__ __forceinline__ double block_sum_fixed(double v, double *s_w) {
  // fixed tree inside a warp, warps added in index order: reproducible run to run
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kSkWarps; w++) t += s_w[w];
  __syncthreads();
  return t;
}

// x: [nseries][n] doubles; y: scratch of the same shape (centred series).  n is a power of two.
__global__ void __launch_bounds__(kSkThreads) sokal_kernel(long n, const double *__restrict__ x, double *__restrict__ y,
                                                           double *var, double *tau, int *m) {
  __shared__ double s_w[kSkWarps];
  __shared__ double s_part[kSkWarps][kSkLags];
  __shared__ double s_lag[kSkLags];
  __shared__ int s_stop;
  __shared__ double s_sum;
  const long series = blockIdx.x;
  const double *xs = x + series * n;
  double *ys = y + series * n;
  const long mask = n - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  double acc = 0.0;
  for (long j = threadIdx.x; j < n; j += kSkThreads) acc += xs[j];
  const double mean = block_sum_fixed(acc, s_w) / (double)n;  // the reference zeroes the DC term (:377)
  acc = 0.0;
  for (long j = threadIdx.x; j < n; j += kSkThreads) {
    const double v = xs[j] - mean;
    ys[j] = v;
    acc += v * v;
  }
  const double a0 = block_sum_fixed(acc, s_w);  // (their xreal[0] after the second transform) / n
  if (threadIdx.x == 0) {
    var[series] = a0 / (double)(n - 1);  // :380
    s_stop = 0;
    s_sum = -(1.0 / 3.0);  // :390
  }
  __syncthreads();
  if (!(a0 > 0.0)) {
    // constant series: the reference's 1/0 makes every rho NaN, the window never closes (:391-396)
    if (threadIdx.x == 0) {
      tau[series] = nan("");
      m[series] = (int)n + 1;
    }
    return;
  }
  const double c = 1.0 / a0;  // :381
  for (long t0 = 0; t0 < n; t0 += kSkLags) {
    double r[kSkLags];
#pragma unroll
    for (int b = 0; b < kSkLags; b++) r[b] = 0.0;
    for (long j = threadIdx.x; j < n; j += kSkThreads) {
      const double yj = ys[j];
#pragma unroll
      for (int b = 0; b < kSkLags; b++) r[b] = fma(yj, ys[(j + t0 + b) & mask], r[b]);
    }
#pragma unroll
    for (int b = 0; b < kSkLags; b++) {
      const double v = warp_sum(r[b]);
      if (lane == 0) s_part[warp][b] = v;
    }
    __syncthreads();
    if (threadIdx.x < kSkLags) {
      double t = 0.0;
      for (int w = 0; w < kSkWarps; w++) t += s_part[w][threadIdx.x];
      s_lag[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double sum = s_sum;
      for (int b = 0; b < kSkLags; b++) {
        const long i = t0 + b;
        if (i >= n) break;
        sum += s_lag[b] * c - (1.0 / 6.0);  // :392
        if (sum < 0.0) {                    // :393
          tau[series] = 2.0 * (sum + (double)i / 6.0);  // :399
          m[series] = (int)i + 1;                        // :400
          s_stop = 1;
          break;
        }
      }
      s_sum = sum;
    }
    __syncthreads();
    if (s_stop) return;
  }
  if (threadIdx.x == 0) {  // the window never closed: i1 == n on exit of the reference's loop
    tau[series] = 2.0 * (s_sum + (double)n / 6.0);
    m[series] = (int)n + 1;
  }
}

// model-index series of the trace chains: the last nkeep sweeps of the last amx_rj_sweeps call
__global__ void ktrace_to_series_kernel(const int *tr_k, long nsweeps, long nkeep, int ntrace, double *x) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkeep * ntrace) return;
  const long c = i / nkeep, s = i % nkeep;
  x[i] = (double)tr_k[c * nsweeps + (nsweeps - nkeep) + s];
}

static int sokal_check(int nseries, long n) {
  if (nseries < 1) return fail(AMX_EINVAL, "sokal: nseries %d", nseries);
  // the reference's transform wants a power of two in [4, 2^20] (logwrite.c:358-361, :424-439)
  if (n < 4 || (n & (n - 1)) || n > (1L << 20)) return fail(AMX_EINVAL, "sokal: length %ld is not a power of two in [4, 2^20]", n);
  return AMX_OK;
}

// x_dev: [nseries][n] on the device; outputs are host arrays
int sokal_device(int nseries, long n, const double *x_dev, double *var, double *tau, int *m) {
  if (int rc = sokal_check(nseries, n)) return rc;
  double *y = nullptr, *o = nullptr;
  int *om = nullptr;
  AMX_CUDA(cudaMalloc(&y, sizeof(double) * (size_t)nseries * n));
  AMX_CUDA(cudaMalloc(&o, sizeof(double) * 2 * nseries));
  AMX_CUDA(cudaMalloc(&om, sizeof(int) * nseries));
  sokal_kernel<<<nseries, kSkThreads, 0, stream()>>>(n, x_dev, y, o, o + nseries, om);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && var) e = cudaMemcpyAsync(var, o, sizeof(double) * nseries, cudaMemcpyDeviceToHost, stream());
  if (e == cudaSuccess && tau) e = cudaMemcpyAsync(tau, o + nseries, sizeof(double) * nseries, cudaMemcpyDeviceToHost, stream());
  if (e == cudaSuccess && m) e = cudaMemcpyAsync(m, om, sizeof(int) * nseries, cudaMemcpyDeviceToHost, stream());
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream());
  cudaFree(y);
  cudaFree(o);
  cudaFree(om);
  if (e != cudaSuccess) return fail(AMX_ECUDA, "sokal: %s", cudaGetErrorString(e));
  return AMX_OK;
}

int sokal_ktrace(const int *tr_k, long nsweeps, long nkeep, int ntrace, double *var, double *tau, int *m) {
  if (int rc = sokal_check(ntrace, nkeep)) return rc;
  if (nkeep > nsweeps) return fail(AMX_EINVAL, "sokal: nkeep %ld exceeds the %ld recorded sweeps", nkeep, nsweeps);
  double *x = nullptr;
  AMX_CUDA(cudaMalloc(&x, sizeof(double) * (size_t)ntrace * nkeep));
  const long tot = nkeep * ntrace;
  ktrace_to_series_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream()>>>(tr_k, nsweeps, nkeep, ntrace, x);
  count_launch();
  const int rc = sokal_device(ntrace, nkeep, x, var, tau, m);
  cudaFree(x);
  return rc;
}

// ---- per-model posterior moments of the population ------------------------------------------------
constexpr int kMomThreads = 128;  // = samples per tile

// pair code: model << 16 | type << 8 | j; type 0 count, 1 sum lp, 2 S1[j], 3+i S2[i][j] (j <= i)
__global__ void __launch_bounds__(kMomThreads) moments_kernel(const int *__restrict__ k, const double *__restrict__ theta,
                                                              const double *__restrict__ lp, long C, int dmax,
                                                              const void *prop_blob, const int *__restrict__ pairs, int P,
                                                              double *__restrict__ partial) {
  __shared__ int sk[kMomThreads];
  __shared__ double slp[kMomThreads];
  extern __shared__ double sdx[];  // [dmax][128]
  const amx_fam_hdr *h = (const amx_fam_hdr *)prop_blob;
  const double *data = (const double *)(h + 1);
  double *row = partial + (size_t)blockIdx.x * P;
  const int tid = threadIdx.x;
  for (int p = tid; p < P; p += kMomThreads) row[p] = 0.0;  // each entry is only ever touched by this thread
  const long ntiles = (C + kMomThreads - 1) / kMomThreads;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long i = tile * kMomThreads + tid;
    int kk = -1;
    if (i < C) {
      kk = k[i];
      slp[tid] = lp[i];
      const int d = h->dims[kk];
      const double *piv = data + h->off[kk] + AMX_REC_HEAD;  // mean of the model's first proposal component
      for (int j = 0; j < dmax; j++) sdx[j * kMomThreads + tid] = (j < d) ? theta[(long)j * C + i] - piv[j] : 0.0;
    }
    sk[tid] = kk;
    __syncthreads();
    for (int p = tid; p < P; p += kMomThreads) {
      const int code = pairs[p], mdl = code >> 16, ty = (code >> 8) & 0xff, j = code & 0xff;
      double acc = 0.0;
      if (ty == 0) {
        for (int s = 0; s < kMomThreads; s++) acc += (sk[s] == mdl) ? 1.0 : 0.0;
      } else if (ty == 1) {
        for (int s = 0; s < kMomThreads; s++) acc += (sk[s] == mdl) ? slp[s] : 0.0;
      } else if (ty == 2) {
        const double *a = sdx + j * kMomThreads;
        for (int s = 0; s < kMomThreads; s++) acc += (sk[s] == mdl) ? a[s] : 0.0;
      } else {
        const double *a = sdx + (ty - 3) * kMomThreads, *b = sdx + j * kMomThreads;
        for (int s = 0; s < kMomThreads; s++) acc += (sk[s] == mdl) ? a[s] * b[s] : 0.0;
      }
      row[p] += acc;
    }
    __syncthreads();
  }
}

__global__ void moments_fold_kernel(const double *__restrict__ partial, int nblocks, int P, double *__restrict__ total) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double t = 0.0;
  for (int b = 0; b < nblocks; b++) t += partial[(size_t)b * P + p];  // block order: reproducible
  total[p] += t;
}

void moments_free(MomentsBuf *mb) {
  if (!mb) return;
  cudaFree(mb->pairs_dev);
  cudaFree(mb->partial_dev);
  cudaFree(mb->total_dev);
  delete mb;
}

int moments_reset(MomentsBuf **pmb, const amx_fam_hdr &h) {
  MomentsBuf *mb = *pmb;
  if (!mb) {
    mb = new MomentsBuf();
    memset(mb, 0, sizeof(*mb));
    std::vector<int> pairs;
    for (int k = 0; k < h.nmodels; k++) {
      mb->off[k] = (int)pairs.size();
      const int d = h.dims[k];
      pairs.push_back(k << 16 | 0 << 8);
      pairs.push_back(k << 16 | 1 << 8);
      for (int j = 0; j < d; j++) pairs.push_back(k << 16 | 2 << 8 | j);
      for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++) pairs.push_back(k << 16 | (3 + i) << 8 | j);
    }
    mb->P = (int)pairs.size();
    int dev = 0, sms = 0;
    AMX_CUDA(cudaGetDevice(&dev));
    AMX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    mb->nblocks = sms * 4;
    AMX_CUDA(cudaMalloc(&mb->pairs_dev, sizeof(int) * mb->P));
    AMX_CUDA(cudaMalloc(&mb->partial_dev, sizeof(double) * (size_t)mb->P * mb->nblocks));
    AMX_CUDA(cudaMalloc(&mb->total_dev, sizeof(double) * mb->P));
    AMX_CUDA(cudaMemcpyAsync(mb->pairs_dev, pairs.data(), sizeof(int) * mb->P, cudaMemcpyHostToDevice, stream()));
    AMX_CUDA(cudaStreamSynchronize(stream()));
    *pmb = mb;
  }
  AMX_CUDA(cudaMemsetAsync(mb->total_dev, 0, sizeof(double) * mb->P, stream()));
  mb->snapshots = 0;
  return AMX_OK;
}

int moments_accumulate(MomentsBuf *mb, const int *k, const double *theta, const double *lp, long C, int dmax,
                       const void *prop_blob) {
  const long ntiles = (C + kMomThreads - 1) / kMomThreads;
  const int grid = (int)(ntiles < mb->nblocks ? ntiles : mb->nblocks);
  const size_t smem = sizeof(double) * (size_t)dmax * kMomThreads;
  moments_kernel<<<grid, kMomThreads, smem, stream()>>>(k, theta, lp, C, dmax, prop_blob, mb->pairs_dev, mb->P, mb->partial_dev);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  moments_fold_kernel<<<(mb->P + 127) / 128, 128, 0, stream()>>>(mb->partial_dev, grid, mb->P, mb->total_dev);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  mb->snapshots++;
  return AMX_OK;
}

int moments_get(const MomentsBuf *mb, const amx_proposal *prop, int model, unsigned long long *count, double *mean,
                double *cov, double *mean_lp) {
  const amx_fam_hdr &h = prop->hdr;
  if (model < 0 || model >= h.nmodels) return fail(AMX_EINVAL, "moments: model %d of %d", model, h.nmodels);
  const int d = h.dims[model], np = 2 + d + d * (d + 1) / 2;
  std::vector<double> t(np), piv(d);
  AMX_CUDA(cudaMemcpyAsync(t.data(), mb->total_dev + mb->off[model], sizeof(double) * np, cudaMemcpyDeviceToHost, stream()));
  const double *data = (const double *)((const char *)prop->blob_dev + sizeof(amx_fam_hdr));
  AMX_CUDA(cudaMemcpyAsync(piv.data(), data + h.off[model] + AMX_REC_HEAD, sizeof(double) * d, cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  const double n = t[0];
  if (count) *count = (unsigned long long)n;
  if (mean_lp) *mean_lp = n > 0 ? t[1] / n : 0.0;
  const double *s1 = t.data() + 2, *s2 = t.data() + 2 + d;
  for (int i = 0; i < d; i++) {
    if (mean) mean[i] = n > 0 ? piv[i] + s1[i] / n : 0.0;
    if (cov)
      for (int j = 0; j <= i; j++) {
        // unbiased sample covariance from the shifted sums
        const double v = n > 1 ? (s2[AMX_TRI(i, j)] - s1[i] * s1[j] / n) / (n - 1.0) : 0.0;
        cov[i * d + j] = v;
        cov[j * d + i] = v;
      }
  }
  return AMX_OK;
}

}  // namespace amx

using namespace amx;

extern "C" int amx_sokal(int nseries, long n, const double *x, double *var, double *tau, int *m) {
  if (int rc = require_device()) return rc;
  if (!x) return fail(AMX_EINVAL, "sokal: null series");
  if (int rc = sokal_check(nseries, n)) return rc;
  double *xd = nullptr;
  const size_t bytes = sizeof(double) * (size_t)nseries * n;
  AMX_CUDA(cudaMalloc(&xd, bytes));
  cudaError_t e = cudaMemcpyAsync(xd, x, bytes, cudaMemcpyHostToDevice, stream());
  int rc = e == cudaSuccess ? sokal_device(nseries, n, xd, var, tau, m) : fail(AMX_ECUDA, "sokal: %s", cudaGetErrorString(e));
  cudaFree(xd);
  return rc;
}

extern "C" int amx_sokal_dev(int nseries, long n, const double *x_dev, double *var, double *tau, int *m) {
  if (int rc = require_device()) return rc;
  if (!x_dev) return fail(AMX_EINVAL, "sokal: null series");
  return sokal_device(nseries, n, x_dev, var, tau, m);
}
