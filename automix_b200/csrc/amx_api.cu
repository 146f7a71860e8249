// amx_api.cu -- runtime state, error channel, plug-in and proposal objects of the C-ABI
// (include/amx.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "amx_common.cuh"
#include <dlfcn.h>

#include "amx_internal.cuh"
#include "amx_targets.cuh"

namespace amx {

static thread_local char g_err[512] = "";
static cudaStream_t g_stream = nullptr;
static int g_defer_sync = 0;
static unsigned long long g_launches = 0;

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
cudaStream_t stream() { return g_stream; }
bool defer_sync() { return g_defer_sync != 0; }
void count_launch(unsigned n) { g_launches += n; }

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    cudaGetLastError();
    return fail(AMX_ENODEV, "no CUDA device available (%s); automix-b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  return AMX_OK;
}

// upload a [header | doubles] blob
int upload_blob(const amx_fam_hdr &h, const double *data, void **dev, int *bytes) {
  const size_t nb = sizeof(amx_fam_hdr) + sizeof(double) * (size_t)h.total;
  std::vector<char> host(nb);
  memcpy(host.data(), &h, sizeof(h));
  if (h.total) memcpy(host.data() + sizeof(h), data, sizeof(double) * (size_t)h.total);
  AMX_CUDA(cudaMalloc(dev, nb));
  AMX_CUDA(cudaMemcpyAsync(*dev, host.data(), nb, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  *bytes = (int)nb;
  return AMX_OK;
}

}  // namespace amx

using namespace amx;

extern "C" {

const char *amx_last_error(void) { return g_err; }
const char *amx_version(void) { return "automix-b200 0.1 (sm_100a)"; }

int amx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int amx_set_device(int ordinal) {
  AMX_CUDA(cudaSetDevice(ordinal));
  return AMX_OK;
}
int amx_set_stream(void *s) {
  g_stream = reinterpret_cast<cudaStream_t>(s);
  return AMX_OK;
}
int amx_set_deferred_sync(int on) {
  g_defer_sync = on ? 1 : 0;
  return AMX_OK;
}
int amx_synchronize(void) {
  AMX_CUDA(cudaStreamSynchronize(g_stream));
  return AMX_OK;
}
int amx_copy_dev(void *dst, const void *src, size_t bytes) {
  AMX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
  return AMX_OK;
}
unsigned long long amx_launch_count(int reset) {
  unsigned long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// ---- plug-ins ------------------------------------------------------------------------------
static amx_target *new_target(int kind, int nmodels, const int *dims) {
  if (nmodels < 1 || nmodels > AMX_MAX_MODELS) {
    fail(AMX_EINVAL, "nmodels=%d outside [1,%d]", nmodels, AMX_MAX_MODELS);
    return nullptr;
  }
  amx_target *t = (amx_target *)calloc(1, sizeof(amx_target));
  t->d.kind = kind;
  t->d.nmodels = nmodels;
  for (int k = 0; k < nmodels; k++) {
    t->d.dims[k] = dims[k];
    if (dims[k] > t->d.dmax) t->d.dmax = dims[k];
    if (dims[k] < 1 || dims[k] > AMX_MAX_DIM) {
      fail(AMX_EINVAL, "model %d has dimension %d outside [1,%d]", k, dims[k], AMX_MAX_DIM);
      free(t);
      return nullptr;
    }
  }
  return t;
}

amx_target *amx_target_gaussmix(int nmodels, const int *dims, const int *ncomp, const double *modw,
                                const double *wt, const double *mean, const double *tri, int flags) {
  if (require_device()) return nullptr;
  amx_target *t = new_target(kTargetGaussMix, nmodels, dims);
  if (!t) return nullptr;
  amx_fam_hdr h;
  int extlen[AMX_MAX_MODELS];
  for (int k = 0; k < nmodels; k++) extlen[k] = 1;
  if (amx_fam_plan(&h, nmodels, dims, ncomp, extlen) < 0) {
    fail(AMX_EINVAL, "bad Gaussian-mixture target shape");
    free(t);
    return nullptr;
  }
  std::vector<double> data(h.total);
  amx_fam_pack(&h, AMX_FAM_TARGET, wt, mean, tri, modw, data.data());
  t->d.flags = flags;
  if (upload_blob(h, data.data(), &t->d.blob_dev, &t->d.blob_bytes)) {
    free(t);
    return nullptr;
  }
  return t;
}

amx_target *amx_target_quad(int nmodels, const int *dims, const double *center, const double *scale,
                            const double *lo, const double *hi) {
  if (require_device()) return nullptr;
  amx_target *t = new_target(kTargetQuad, nmodels, dims);
  if (!t) return nullptr;
  amx_fam_hdr h;
  memset(&h, 0, sizeof(h));
  h.nmodels = nmodels;
  int pos = 0, src = 0;
  std::vector<double> data;
  for (int k = 0; k < nmodels; k++) {
    const int d = dims[k];
    h.dims[k] = d;
    h.ncomp[k] = 1;
    h.off[k] = pos;
    h.stride[k] = 4 * d;
    if (d > h.dmax) h.dmax = d;
    for (int i = 0; i < d; i++) data.push_back(center[src + i]);
    for (int i = 0; i < d; i++) data.push_back(scale[src + i]);
    for (int i = 0; i < d; i++) data.push_back(lo ? lo[src + i] : -INFINITY);
    for (int i = 0; i < d; i++) data.push_back(hi ? hi[src + i] : INFINITY);
    pos += 4 * d;
    src += d;
  }
  h.Lmax = 1;
  h.total = pos;
  if (upload_blob(h, data.data(), &t->d.blob_dev, &t->d.blob_bytes)) {
    free(t);
    return nullptr;
  }
  return t;
}

amx_target *amx_target_coalmine(void) {
  if (require_device()) return nullptr;
  int dims[6];
  for (int k = 0; k < 6; k++) dims[k] = 2 * k + 3;
  amx_target *t = new_target(kTargetCoal, 6, dims);
  if (!t) return nullptr;
  amx_fam_hdr h;
  memset(&h, 0, sizeof(h));
  h.nmodels = 6;
  h.dmax = 13;
  h.Lmax = 1;
  std::vector<double> data;
  const double alpha = 1.0, beta = 200.0, lam = 3.0;
  for (int k = 0; k < 6; k++) {
    const int ns = k + 1;
    h.dims[k] = dims[k];
    h.ncomp[k] = 1;
    h.off[k] = 3 * k;
    h.stride[k] = 3;
    // the two groups of terms of usercpt.c:99 and :106 that depend on k only
    data.push_back(-lam + ns * log(lam) - lgamma((double)(ns + 1)));
    data.push_back(lgamma(2.0 * (ns + 1)) - (2.0 * ns + 1.0) * log(AMX_COAL_T));
    data.push_back(alpha * log(beta) - lgamma(alpha));
  }
  h.total = (int)data.size();
  if (upload_blob(h, data.data(), &t->d.blob_dev, &t->d.blob_bytes)) {
    free(t);
    return nullptr;
  }
  return t;
}

amx_target *amx_target_mixnorm(int nmodels, const int *ncomp, int ndata, const double *y, const double *prior5) {
  if (require_device()) return nullptr;
  if (nmodels < 1 || nmodels > AMX_MAX_MODELS || !ncomp || ndata < 1 || !y || !prior5) {
    fail(AMX_EINVAL, "amx_target_mixnorm: bad arguments");
    return nullptr;
  }
  int dims[AMX_MAX_MODELS];
  for (int k = 0; k < nmodels; k++) {
    if (ncomp[k] < 1 || ncomp[k] > kMixNormKmax) {
      fail(AMX_EINVAL, "amx_target_mixnorm: 1..%d components per model (model %d has %d)", kMixNormKmax, k, ncomp[k]);
      return nullptr;
    }
    dims[k] = 3 * ncomp[k] - 1;
  }
  if (!(prior5[0] > 0.0 && prior5[2] > 0.0 && prior5[4] > 0.0)) {
    fail(AMX_EINVAL, "amx_target_mixnorm: prior standard deviations must be positive");
    return nullptr;
  }
  amx_target *t = new_target(kTargetMixNorm, nmodels, dims);
  if (!t) return nullptr;
  amx_fam_hdr h;
  memset(&h, 0, sizeof(h));
  h.nmodels = nmodels;
  for (int k = 0; k < nmodels; k++) {
    h.dims[k] = dims[k];
    h.ncomp[k] = ncomp[k];
    if (dims[k] > h.dmax) h.dmax = dims[k];
    if (ncomp[k] > h.Lmax) h.Lmax = ncomp[k];
  }
  std::vector<double> data;
  data.push_back((double)ndata);
  for (int i = 0; i < 5; i++) data.push_back(prior5[i]);
  for (int i = 0; i < ndata; i++) data.push_back(y[i]);
  h.total = (int)data.size();
  if (upload_blob(h, data.data(), &t->d.blob_dev, &t->d.blob_bytes)) {
    free(t);
    return nullptr;
  }
  return t;
}

amx_target *amx_target_plugin(const char *so_path, int nmodels, const int *dims, const void *blob, size_t blob_bytes,
                              int flags) {
  if (require_device()) return nullptr;
  if (!so_path || nmodels < 1 || nmodels > AMX_MAX_MODELS || !dims || (blob_bytes && !blob) || (blob_bytes & 7)) {
    fail(AMX_EINVAL, "amx_target_plugin: bad arguments (the parameter blob must be a multiple of 8 bytes)");
    return nullptr;
  }
  void *dl = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) {
    fail(AMX_EINVAL, "amx_target_plugin: %s", dlerror());
    return nullptr;
  }
  typedef const PluginVtbl *(*entry_fn)(void);
  entry_fn entry = (entry_fn)dlsym(dl, "amx_plugin_entry");
  const PluginVtbl *v = entry ? entry() : nullptr;
  if (!v || v->abi != kPluginAbi) {
    fail(AMX_EINVAL, "amx_target_plugin: %s is not a plug-in of this library version (rebuild it with python -m automix_b200.plugin)", so_path);
    dlclose(dl);
    return nullptr;
  }
  amx_target *t = new_target(kTargetPlugin, nmodels, dims);
  if (!t) {
    dlclose(dl);
    return nullptr;
  }
  t->d.plugin = v;
  t->d.plugin_dl = dl;
  t->d.flags = flags;
  const size_t nb = blob_bytes ? blob_bytes : 8;
  if (cudaMalloc(&t->d.blob_dev, nb) != cudaSuccess ||
      (blob_bytes && cudaMemcpy(t->d.blob_dev, blob, blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess)) {
    fail(AMX_ECUDA, "amx_target_plugin: parameter upload failed");
    if (t->d.blob_dev) cudaFree(t->d.blob_dev);
    dlclose(dl);
    free(t);
    return nullptr;
  }
  t->d.blob_bytes = (int)blob_bytes;
  return t;
}

amx_target *amx_target_host_scalar(int nmodels, const int *dims, amx_scalar_fn f) {
  amx_target *t = new_target(kTargetHostScalar, nmodels, dims);
  if (t) t->d.scalar = f;
  return t;
}
amx_target *amx_target_host_batched(int nmodels, const int *dims, amx_batched_fn f, void *user) {
  amx_target *t = new_target(kTargetHostBatched, nmodels, dims);
  if (t) {
    t->d.batched = f;
    t->d.user = user;
  }
  return t;
}
void amx_target_destroy(amx_target *t) {
  if (!t) return;
  if (t->d.blob_dev) cudaFree(t->d.blob_dev);
  if (t->d.plugin_dl) dlclose(t->d.plugin_dl);
  free(t);
}

// ---- proposal --------------------------------------------------------------------------------
amx_proposal *amx_proposal_create(int nmodels, const int *dims, const int *ncomp, const double *wt,
                                  const double *mean, const double *tri, const double *sig) {
  if (require_device()) return nullptr;
  amx_proposal *p = (amx_proposal *)calloc(1, sizeof(amx_proposal));
  if (amx_fam_plan(&p->hdr, nmodels, dims, ncomp, dims) < 0) {
    fail(AMX_EINVAL, "bad proposal shape (nmodels<=%d, d<=%d, L<=%d)", AMX_MAX_MODELS, AMX_MAX_DIM,
         AMX_MAX_COMPS);
    free(p);
    return nullptr;
  }
  std::vector<double> data(p->hdr.total);
  amx_fam_pack(&p->hdr, AMX_FAM_PROPOSAL, wt, mean, tri, sig, data.data());
  if (upload_blob(p->hdr, data.data(), &p->blob_dev, &p->blob_bytes)) {
    free(p);
    return nullptr;
  }
  return p;
}
void amx_proposal_destroy(amx_proposal *p) {
  if (!p) return;
  if (p->blob_dev) cudaFree(p->blob_dev);
  free(p);
}

// ---- fp64 peak micro-benchmark --------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
  // 16 independent DFMA chains per thread: enough ILP to cover the pipe latency at 8 warps/SMSP
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7, x8 = x0 + 8, x9 = x0 + 9, xa = x0 + 10, xb = x0 + 11, xc = x0 + 12, xd = x0 + 13,
         xe = x0 + 14, xf = x0 + 15;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    x8 = fma(x8, a, b); x9 = fma(x9, a, b); xa = fma(xa, a, b); xb = fma(xb, a, b);
    xc = fma(xc, a, b); xd = fma(xd, a, b); xe = fma(xe, a, b); xf = fma(xf, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + ((x8 + x9) + (xa + xb)) +
                   ((xc + xd) + (xe + xf));
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

int amx_measure_fp64_peak(double *flops_per_s) {
  if (int rc = require_device()) return rc;
  int dev = 0, sms = 0;
  AMX_CUDA(cudaGetDevice(&dev));
  AMX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *out = nullptr;
  AMX_CUDA(cudaMalloc(&out, sizeof(double)));
  cudaEvent_t e0, e1;
  AMX_CUDA(cudaEventCreate(&e0));
  AMX_CUDA(cudaEventCreate(&e1));
  const int iters = 20000, blocks = sms * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    AMX_CUDA(cudaEventRecord(e0, stream()));
    dfma_peak_kernel<<<blocks, threads, 0, stream()>>>(out, iters, 0.999999, 1e-9);
    count_launch();
    AMX_CUDA(cudaEventRecord(e1, stream()));
    AMX_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    AMX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 16.0 * (double)iters * blocks * threads / (ms * 1e-3);
    if (rep > 0 && fl > best) best = fl;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *flops_per_s = best;
  return AMX_OK;
}

}  // extern "C"
