// amx_targets.cuh -- __device__ log-posterior plug-ins (the "device" tier beside the
// reference's scalar callback `double f(int model_k, double *x)`, automix.h:46).
//
// A plug-in is a small view struct with
//     template <int DMAX> double eval(int k, const double (&x)[DMAX]) const
//     int flops(int k) const                      (C_lp of SURVEY.md 8d)
// whose parameters sit in one contiguous blob that a CTA stages in shared memory.
#pragma once

#include <float.h>

#include "amx_coal_data.h"
#include "amx_common.cuh"

namespace amx {

enum TargetKind { kTargetGaussMix = 1, kTargetQuad = 2, kTargetCoal = 3, kTargetMixNorm = 4, kTargetPlugin = 50, kTargetHostScalar = 100,
                  kTargetHostBatched = 101 };

// Solve T r = x - mu for the lower-triangular factor in a family record and return |r|^2.
// (forward substitution of lnormprob, automix.c:1735-1747, with the diagonal divisions
// replaced by multiplications with the precomputed reciprocals.)
template <int DMAX, bool UNROLL = (DMAX <= kRegArrayMax)>
__device__ __forceinline__ double solve_lower(const double *rec, int d, const double (&x)[DMAX],
                                              double (&r)[DMAX]) {
  const double *mu = rec + AMX_REC_HEAD, *rd = mu + d, *T = rd + d;
  double q = 0.0;
  if constexpr (UNROLL) {
#pragma unroll
    for (int i = 0; i < DMAX; i++) {
      if (i < d) {
        double v = x[i] - mu[i];
#pragma unroll
        for (int j = 0; j < i; j++) v = fma(-T[AMX_TRI(i, j)], r[j], v);
        r[i] = v * rd[i];
        q = fma(r[i], r[i], q);
      }
    }
  } else {
    for (int i = 0; i < d; i++) {
      double v = x[i] - mu[i];
      const double *Ti = T + AMX_TRI(i, 0);
      for (int j = 0; j < i; j++) v = fma(-Ti[j], r[j], v);
      r[i] = v * rd[i];
      q = fma(r[i], r[i], q);
    }
  }
  return q;
}

// ---- |T^-1 (x - mu)|^2 with the dimension as a compile-time constant ----------------------------------------------
// In the wide configurations (d > 8) x and r are run-time indexed local-memory vectors and the row loop above costs
// ~13 instructions per multiply-add (ncu, C5-RJ: generic 64-bit address arithmetic for T, LDL/STL of r, the remainder
// ladders of a partially unrolled loop with trip counts of a few).  With d fixed, r lives in registers, every load has
// an immediate offset and the body is one load + one DFMA per term -- operation for operation the same arithmetic in
// the same order, so the value is bit-identical.  One out-of-line copy per translation unit, selected by a switch on d:
// after the sort (rj_sort_*_kernel) the lanes of a warp share d, so the switch does not diverge.
constexpr int kQuadFixedMax = 20;
template <int D>
__device__ __forceinline__ double quad_fixed(const double *__restrict__ rec, const double *__restrict__ x) {
  const double *mu = rec + AMX_REC_HEAD, *rd = mu + D, *T = rd + D;
  double r[D];
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    double v = x[i] - mu[i];
#pragma unroll
    for (int j = 0; j < i; j++) v = fma(-T[AMX_TRI(i, j)], r[j], v);
    r[i] = v * rd[i];
    q = fma(r[i], r[i], q);
  }
  return q;
}
static __device__ __noinline__ double quad_form_wide(const double *rec, int d, const double *x) {
  switch (d) {
#define AMX_QF(D) case D: return quad_fixed<D>(rec, x);
    AMX_QF(1) AMX_QF(2) AMX_QF(3) AMX_QF(4) AMX_QF(5) AMX_QF(6) AMX_QF(7) AMX_QF(8) AMX_QF(9) AMX_QF(10)
    AMX_QF(11) AMX_QF(12) AMX_QF(13) AMX_QF(14) AMX_QF(15) AMX_QF(16) AMX_QF(17) AMX_QF(18) AMX_QF(19) AMX_QF(20)
#undef AMX_QF
  }
  // wider than the unrolled forms: the row loop
  const double *mu = rec + AMX_REC_HEAD, *rd = mu + d, *T = rd + d;
  double r[AMX_MAX_DIM];
  double q = 0.0;
  for (int i = 0; i < d; i++) {
    double v = x[i] - mu[i];
    const double *Ti = T + AMX_TRI(i, 0);
    for (int j = 0; j < i; j++) v = fma(-Ti[j], r[j], v);
    r[i] = v * rd[i];
    q = fma(r[i], r[i], q);
  }
  return q;
}
// the quadratic form alone (callers that do not need r).  `fixed`: the lanes of the warp share d (sorted population, or
// one model per kernel) -- take the fixed-dimension forms; in a warp of mixed models the switch would serialise every
// distinct d (measured: 3.2e7 against 7.9e7 chain-sweeps/s on C5-RJ), so the row loop stays the form for those.
template <int DMAX>
__device__ __forceinline__ double quad_form(const double *rec, int d, const double (&x)[DMAX], bool fixed) {
  if constexpr (DMAX <= kRegArrayMax) {
    double r[DMAX];
    return solve_lower<DMAX>(rec, d, x, r);
  } else {
    if (fixed) return quad_form_wide(rec, d, &x[0]);
    double r[DMAX];
    return solve_lower<DMAX>(rec, d, x, r);
  }
}
constexpr int kTargetFlagUniformDims = 0x4000;  // bind() flag: the lanes of a warp evaluate the same model

// ---- Gaussian-mixture family (toy1, toy2, the synthetic scaling targets) -------------
struct GaussMixTarget {
  const amx_fam_hdr *h;
  const double *D;
  int flags;
  bool fixed;
  __device__ __forceinline__ void bind(const void *blob, int fl) {
    h = reinterpret_cast<const amx_fam_hdr *>(blob);
    D = reinterpret_cast<const double *>(h + 1);
    flags = fl & ~kTargetFlagUniformDims;
    fixed = (fl & kTargetFlagUniformDims) != 0;
  }
  __device__ __forceinline__ int flops(int k) const {
    const int d = h->dims[k];
    return h->ncomp[k] * (d * d + 3 * d + 6) + 2;
  }
  template <int DMAX>
  __device__ __forceinline__ double eval(int k, const double (&x)[DMAX]) const {
    const int d = h->dims[k], G = h->ncomp[k], st = h->stride[k];
    const double *rec = D + h->off[k];
    const double modw = D[h->ext[k]];
    if (flags == AMX_GM_PLAIN) {  // log(modw * sum_g c_g exp(-q_g/2)), as usertoy1.c:72-100
      double s = 0.0;
      for (int g = 0; g < G; g++) s = fma(rec[g * st + 2], exp(-0.5 * quad_form<DMAX>(rec + g * st, d, x, fixed)), s);
      return log(modw * s);
    }
    // log-sum-exp form: running maximum, rescale on the fly
    double m = -DBL_MAX, s = 0.0;
    for (int g = 0; g < G; g++) {
      const double a = rec[g * st + 3] - 0.5 * quad_form<DMAX>(rec + g * st, d, x, fixed);
      if (a > m) {
        s = s * exp(m - a) + 1.0;
        m = a;
      } else {
        s += exp(a - m);
      }
    }
    return log(modw) + m + log(s);
  }
};

// ---- separable quadratic with box support (README 1-D Normal, truncated Normal) --------
struct QuadTarget {
  const amx_fam_hdr *h;  // dims[k], off[k] -> center[d], scale[d], lo[d], hi[d]
  const double *D;
  __device__ __forceinline__ void bind(const void *blob, int) {
    h = reinterpret_cast<const amx_fam_hdr *>(blob);
    D = reinterpret_cast<const double *>(h + 1);
  }
  __device__ __forceinline__ int flops(int k) const { return 4 * h->dims[k]; }
  template <int DMAX>
  __device__ __forceinline__ double eval(int k, const double (&x)[DMAX]) const {
    const int d = h->dims[k];
    const double *c = D + h->off[k], *sc = c + d, *lo = sc + d, *hi = lo + d;
    double s = 0.0;
    bool out = false;
    if constexpr (DMAX <= kRegArrayMax) {
#pragma unroll
      for (int i = 0; i < DMAX; i++)
        if (i < d) {
          out |= (x[i] <= lo[i]) | (x[i] >= hi[i]);
          s += -(x[i] - c[i]) * (x[i] - c[i]) / (2.0 * sc[i] * sc[i]);
        }
    } else {
      for (int i = 0; i < d; i++) {
        out |= (x[i] <= lo[i]) | (x[i] >= hi[i]);
        s += -(x[i] - c[i]) * (x[i] - c[i]) / (2.0 * sc[i] * sc[i]);
      }
    }
    return out ? -DBL_MAX : s;
  }
};

// ---- coal-mining change-point posterior (usercpt.c:46-134) ------------------------------
__constant__ double c_coal_y[AMX_COAL_N] = {AMX_COAL_VALUES};   // uniform index across lanes: the sequential walk
__device__ const double g_coal_y[AMX_COAL_N] = {AMX_COAL_VALUES};  // divergent index: the binary searches

struct CoalTarget {
  // blob: amx_fam_hdr (dims only) + per model [c_prior, c_tail]: the two log-gamma
  // groups of usercpt.c:99 and :106, which depend on k alone and are formed on the host.
  const amx_fam_hdr *h;
  const double *D;
  __device__ __forceinline__ void bind(const void *blob, int) {
    h = reinterpret_cast<const amx_fam_hdr *>(blob);
    D = reinterpret_cast<const double *>(h + 1);
  }
  __device__ __forceinline__ int flops(int k) const { return 12 * (k + 2) + 2 * AMX_COAL_N; }
  template <int DMAX>
  __device__ __forceinline__ double eval(int k, const double (&x)[DMAX]) const {
    if constexpr (DMAX < 13) {
      return 0.0 / 0.0;  // needs d up to 13: only the general configuration is valid
    } else {
      const double alpha = 1.0, beta = 200.0, T = AMX_COAL_T;
      const int ns = k + 1;
      double hh[8], s[9], ds[8];
      hh[0] = x[0];
      s[0] = 0.0;
      for (int i = 1; i <= ns; i++) {
        hh[i] = x[i];
        s[i] = x[ns + i];
        ds[i - 1] = s[i] - s[i - 1];
      }
      ds[ns] = T - s[ns];
      s[ns + 1] = T;
      bool bad = false;
      for (int i = 0; i <= ns; i++) bad |= (hh[i] <= 0.0) | (ds[i] <= 0.0);
      if (bad) return -10000.0;
      const double abcon = D[h->off[k] + 2];
      double lp = D[h->off[k] + 0];
      double lh[8];  // log h_i: needed by the prior and again by the likelihood (one evaluation, same value)
      for (int i = 0; i <= ns; i++) {
        lh[i] = log(hh[i]);
        lp += (abcon + (alpha - 1.0) * lh[i] - beta * hh[i]);
        lp += log(ds[i]);
      }
      lp += D[h->off[k] + 1];
      // Likelihood.  The reference walks the 191 sorted data once and lets the segment index advance by at
      // most one per datum (usercpt.c:114-127).  If every change point is followed by a datum of its own
      // segment -- first index beyond s_m strictly increasing in m and below 191 -- that walk visits the
      // segments exactly at those indices and its sum is  sum_j (idx_{j+1}-idx_j) log h_j - h_j ds_j  with the
      // same terms in the same order, so ns binary searches replace the 191-step walk.  Otherwise (two change
      // points in one data gap, or none of the data beyond one) the walk's late advances matter: fall back.
      int idx[9];
      idx[0] = 0;
      bool regular = true;
      for (int m = 1; m <= ns; m++) {
        int lo = 0, hi = AMX_COAL_N;  // first i with y[i] > s[m]
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (g_coal_y[mid] > s[m]) hi = mid;
          else lo = mid + 1;
        }
        idx[m] = lo;
        regular &= (lo > idx[m - 1] || m == 1) && (lo < AMX_COAL_N);
      }
      if (regular) {
        idx[ns + 1] = AMX_COAL_N;
        double llh = 0.0;
        for (int j = 0; j <= ns; j++) llh += ((idx[j + 1] - idx[j]) * lh[j] - hh[j] * ds[j]);
        return lp + llh;
      }
      int seen = 0, j = 0;
      double top = s[1], llh = 0.0;
      for (int i = 0; i < AMX_COAL_N; i++) {
        if (c_coal_y[i] > top) {  // at most one segment advance per datum (usercpt.c:114-127)
          const int nj = i - seen;
          seen = i;
          llh += (nj * lh[j] - hh[j] * ds[j]);
          j++;
          if (j > ns) return lp;
          top = s[j + 1];
        }
      }
      llh += (AMX_COAL_N - seen) * lh[j] - hh[j] * ds[j];
      return lp + llh;
    }
  }
};

// ---- finite mixture of normals with an unknown number of components (BASELINE config 4) ------------------------
// "Enzyme-style" data y_1..y_n; model k has K = ncomp[k] components and d = 3K - 1 parameters
//   theta = (a_1 .. a_{K-1} | m_1 .. m_K | s_1 .. s_K):  stick-breaking logits, means, log standard deviations,
//   w_j = v_j prod_{i<j} (1 - v_i), v_j = 1 / (1 + exp(-a_j)), w_K = prod_{i<K} (1 - v_i),
//   log-likelihood sum_i log sum_j w_j N(y_i; m_j, exp(2 s_j)), independent normal priors (normalised, so that models of
//   different dimension are comparable): a_j ~ N(0, pa^2), m_j ~ N(pm, pms^2), s_j ~ N(ps, pss^2).
// The reference has no such example; the definition is ours, stated once in
// automix_b200/workloads.py (c4_mixnorm) and restated for the host in oracle/host_targets.c.  Blob: amx_fam_hdr
// (dims, ncomp) + [n, pa, pm, pms, ps, pss, y_1 .. y_n].
constexpr int kMixNormKmax = 10;
struct MixNormTarget {
  const amx_fam_hdr *h;
  const double *D;
  __device__ __forceinline__ void bind(const void *blob, int) {
    h = reinterpret_cast<const amx_fam_hdr *>(blob);
    D = reinterpret_cast<const double *>(h + 1);
  }
  __device__ __forceinline__ int flops(int k) const { return (int)D[0] * (8 * h->ncomp[k] + 4) + 30 * h->ncomp[k]; }
  static __device__ __forceinline__ double softplus(double x) {  // log(1 + exp(x)) without overflow
    return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x));
  }
  template <int DMAX>
  __device__ __forceinline__ double eval(int k, const double (&x)[DMAX]) const {
    const int K = h->ncomp[k], n = (int)D[0];
    const double pa = D[1], pm = D[2], pms = D[3], ps = D[4], pss = D[5];
    const double *y = D + 6;
    const double hl2pi = 0.9189385332046727;  // log(2 pi) / 2
    double cj[kMixNormKmax], mj[kMixNormKmax], isj[kMixNormKmax];
    double lrem = 0.0, lprior = 0.0;
    for (int j = 0; j < K; j++) {
      double lw = lrem;
      if (j < K - 1) {
        const double a = aget(x, j);
        lw = lrem - softplus(-a);   // + log v_j
        lrem -= softplus(a);        // + log (1 - v_j)
        const double za = a / pa;
        lprior += -0.5 * (za * za) - log(pa) - hl2pi;
      }
      const double m = aget(x, K - 1 + j), sl = aget(x, 2 * K - 1 + j);
      const double zm = (m - pm) / pms, zs = (sl - ps) / pss;
      lprior += (-0.5 * (zm * zm) - log(pms) - hl2pi) + (-0.5 * (zs * zs) - log(pss) - hl2pi);
      cj[j] = lw - sl - hl2pi;
      mj[j] = m;
      isj[j] = exp(-sl);
    }
    double ll = 0.0;
    for (int i = 0; i < n; i++) {
      const double yi = y[i];
      double mx = -DBL_MAX, t[kMixNormKmax];
      for (int j = 0; j < K; j++) {
        const double z = (yi - mj[j]) * isj[j];
        t[j] = cj[j] - 0.5 * (z * z);
        mx = t[j] > mx ? t[j] : mx;
      }
      double ssum = 0.0;
      for (int j = 0; j < K; j++) ssum += exp(t[j] - mx);
      ll += mx + log(ssum);
    }
    return ll + lprior;
  }
};

// Plug-ins whose models are wide (or whose evaluation is long): only the large kernel configurations are built for them.
template <class TGT>
struct TargetIsWide {
  static constexpr bool value = false;
};
template <>
struct TargetIsWide<CoalTarget> {
  static constexpr bool value = true;
};
template <>
struct TargetIsWide<MixNormTarget> {
  static constexpr bool value = true;
};

// A user-supplied __device__ log-posterior: a shared object built from the user's source and amx_plugin_tu.cu
// (python -m automix_b200.plugin), holding the K1/K3 kernels instantiated for the user's struct, reached through this
// table (amx_target_plugin, include/amx.h).
struct RjLaunch;
struct RwmArgs;
constexpr int kPluginAbi = 3;
struct PluginVtbl {
  int abi;
  int (*rj_sweeps)(const RjLaunch *a, int dmax, int Lmax, int nm, int tape);
  int (*rj_init)(const RjLaunch *a, const double *init_dev, int tape);
  int (*rwm)(const RwmArgs *a, int tape);
  int (*eval)(const void *blob_dev, int flags, const int *dims, long n, const int *k_dev, const double *x_dev, long ldx,
              double *out_dev);
};

// Host-side description of a plug-in (amx_api.cu owns these).
struct TargetDesc {
  const PluginVtbl *plugin;  // kTargetPlugin
  void *plugin_dl;
  int kind;
  int flags;
  int nmodels;
  int dmax;
  int dims[AMX_MAX_MODELS];
  void *blob_dev;  // device copy of the parameter blob
  int blob_bytes;
  amx_scalar_fn scalar;
  amx_batched_fn batched;
  void *user;
};

}  // namespace amx

struct amx_target {
  amx::TargetDesc d;
};
