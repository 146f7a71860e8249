// amx_plugin_tu.cu -- translation unit of a USER-SUPPLIED __device__ log-posterior (the device variant of the
// reference's plug-in contract `double f(int model_k, double *x)`, automix.h:46).
//
// Built by `python -m automix_b200.plugin build my_target.cuh` into libamx_plugin_<name>.so:
//     nvcc -shared -gencode arch=compute_100a,code=sm_100a -DAMX_PLUGIN_SOURCE='"/abs/my_target.cuh"' amx_plugin_tu.cu ...
// The user's file defines, after the SDK headers below are in scope,
//
//     struct AmxUserTarget {
//       __device__ void bind(const void *blob, int flags);   // blob: the bytes given to amx_target_plugin (in shared
//                                                            // memory when they fit, else in global memory)
//       __device__ int flops(int model_k) const;             // cost of one evaluation (roofline accounting only)
//       template <int DMAX>
//       __device__ double eval(int model_k, const double (&x)[DMAX]) const;   // the log-posterior
//     };
//
// and may use everything the built-in families use (amx_fam_hdr, solve_lower, aget, ...).  This file instantiates for
// it the very kernel templates of the built-in families -- the fused reversible-jump sweep, the chain start, both
// stage-1 RWM kernels, the batched evaluation -- and exports them through one table.
#include "amx_rj_kernels.cuh"
#include "amx_rwm_kernels.cuh"

#ifndef AMX_PLUGIN_SOURCE
#error "define AMX_PLUGIN_SOURCE to the path of the source that defines struct AmxUserTarget"
#endif
#include AMX_PLUGIN_SOURCE

namespace {
using T = AmxUserTarget;
using namespace amx;

int p_rj_sweeps(const RjLaunch *a, int dmax, int Lmax, int nm, int tape) {
  return tape ? launch_cfg<T, TapeStream>(*a, dmax, Lmax, nm) : launch_cfg<T, PhiloxStream>(*a, dmax, Lmax, nm);
}
int p_rj_init(const RjLaunch *a, const double *init_dev, int tape) {
  const unsigned grid = (unsigned)((a->st.C + kRjThreads - 1) / kRjThreads);
  if (tape) rj_init_kernel<T, TapeStream><<<grid, kRjThreads, 0, stream()>>>(*a, init_dev);
  else rj_init_kernel<T, PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(*a, init_dev);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  return AMX_OK;
}
int p_rwm(const RwmArgs *a, int tape) { return tape ? rwm_launch_d<T, TapeStream>(*a) : rwm_launch_d<T, PhiloxStream>(*a); }
int p_eval(const void *blob, int flags, const int *dims, long n, const int *k, const double *x, long ldx, double *out) {
  const unsigned grid = (unsigned)((n + kRjThreads - 1) / kRjThreads);
  EvalDims ed;
  for (int q = 0; q < AMX_MAX_MODELS; q++) ed.dims[q] = dims[q];
  target_eval_kernel<T><<<grid, kRjThreads, 0, stream()>>>(blob, flags, ed, n, k, x, ldx, out);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  return AMX_OK;
}
const PluginVtbl g_vtbl = {kPluginAbi, p_rj_sweeps, p_rj_init, p_rwm, p_eval};
}  // namespace

extern "C" const amx::PluginVtbl *amx_plugin_entry(void) { return &g_vtbl; }
