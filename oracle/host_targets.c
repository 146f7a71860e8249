/*
 * host_targets.c -- host-side log-posterior callbacks with the reference's
 * contract `double f(int model_k, double *x)` (automix.h:46).
 *
 * TEST INFRASTRUCTURE: these are the workload definitions handed to the oracle
 * (amx_oracle.c), to the compiled reference (oracle/_ref) and to the CPU
 * baseline leg of bench.py.  The device plug-ins in
 * automix_b200/csrc/amx_targets.cuh are written separately and are checked
 * against these (tests/test_targets.py); these in turn are checked against the
 * reference's own example files where such a file exists (usertoy1.c,
 * usertoy2.c, usercpt.c -- tests/test_oracle_vs_ref.py).
 *
 * The callback has no user pointer, so the selected target is process-global,
 * exactly like user state in the reference's examples.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "amx_layout.h"

enum { T_NONE = 0, T_GAUSSMIX, T_QUAD, T_COALMINE, T_MIXNORM };

static int g_kind = T_NONE;
static long g_calls = 0;

/* --- Gaussian mixture family --------------------------------------------- */
static amx_fam_hdr g_hdr;
static double *g_data = NULL;
static int g_flags = 0;

int amxh_select_gaussmix(int nmodels, const int *dims, const int *ncomp,
                         const double *modw, const double *wt,
                         const double *mean, const double *tri, int flags) {
  int extlen[AMX_MAX_MODELS];
  for (int k = 0; k < nmodels && k < AMX_MAX_MODELS; k++) extlen[k] = 1;
  int total = amx_fam_plan(&g_hdr, nmodels, dims, ncomp, extlen);
  if (total < 0) return -1;
  free(g_data);
  g_data = malloc(sizeof(double) * total);
  amx_fam_pack(&g_hdr, AMX_FAM_TARGET, wt, mean, tri, modw, g_data);
  g_flags = flags;
  g_kind = T_GAUSSMIX;
  return 0;
}

static double quad_form(const double *rec, int d, const double *x) {
  const double *mu = rec + AMX_REC_HEAD, *rd = mu + d, *T = rd + d;
  double r[AMX_MAX_DIM], q = 0.0;
  for (int i = 0; i < d; i++) {
    double v = x[i] - mu[i];
    for (int j = 0; j < i; j++) v -= T[AMX_TRI(i, j)] * r[j];
    r[i] = v * rd[i];
    q += r[i] * r[i];
  }
  return q;
}

static double gaussmix_eval(int k, const double *x) {
  int d = g_hdr.dims[k], G = g_hdr.ncomp[k], st = g_hdr.stride[k];
  const double *rec = g_data + g_hdr.off[k];
  double modw = g_data[g_hdr.ext[k]];
  if (g_flags == 0) { /* plain, the form of usertoy1.c:72-100 */
    double s = 0.0;
    for (int g = 0; g < G; g++)
      s += rec[g * st + 2] * exp(-0.5 * quad_form(rec + g * st, d, x));
    return log(modw * s);
  }
  double a[AMX_MAX_COMPS], m = -DBL_MAX, s = 0.0;
  for (int g = 0; g < G; g++) {
    a[g] = rec[g * st + 3] - 0.5 * quad_form(rec + g * st, d, x);
    if (a[g] > m) m = a[g];
  }
  for (int g = 0; g < G; g++) s += exp(a[g] - m);
  return log(modw) + m + log(s);
}

/* --- separable quadratic --------------------------------------------------- */
static int q_nmodels = 0, q_dims[AMX_MAX_MODELS], q_off[AMX_MAX_MODELS];
static double *q_center = NULL, *q_scale = NULL, *q_lo = NULL, *q_hi = NULL;

int amxh_select_quad(int nmodels, const int *dims, const double *center,
                     const double *scale, const double *lo, const double *hi) {
  int tot = 0;
  if (nmodels > AMX_MAX_MODELS) return -1;
  for (int k = 0; k < nmodels; k++) {
    q_dims[k] = dims[k];
    q_off[k] = tot;
    tot += dims[k];
  }
  q_nmodels = nmodels;
  free(q_center); free(q_scale); free(q_lo); free(q_hi);
  q_center = malloc(sizeof(double) * tot);
  q_scale = malloc(sizeof(double) * tot);
  q_lo = malloc(sizeof(double) * tot);
  q_hi = malloc(sizeof(double) * tot);
  for (int i = 0; i < tot; i++) {
    q_center[i] = center[i];
    q_scale[i] = scale[i];
    q_lo[i] = lo ? lo[i] : -INFINITY;
    q_hi[i] = hi ? hi[i] : INFINITY;
  }
  g_kind = T_QUAD;
  return 0;
}

static double quad_eval(int k, const double *x) {
  int o = q_off[k];
  double s = 0.0;
  for (int i = 0; i < q_dims[k]; i++) {
    if (x[i] <= q_lo[o + i] || x[i] >= q_hi[o + i]) return -DBL_MAX;
    /* the form of README.md:70-71 */
    s += -(x[i] - q_center[o + i]) * (x[i] - q_center[o + i]) /
         (2.0 * q_scale[o + i] * q_scale[o + i]);
  }
  return s;
}

/* --- coal-mining change points (follows usercpt.c:46-134) ------------------ */
#include "../include/amx_coal_data.h"

int amxh_select_coalmine(void) {
  g_kind = T_COALMINE;
  return 0;
}

static double coal_eval(int k, const double *th) {
  const double alpha = 1.0, beta = 200.0, lam = 3.0, T = AMX_COAL_T;
  int ns = k + 1; /* change points; ns+1 rates */
  double h[8], s[9], ds[8];
  h[0] = th[0];
  s[0] = 0.0;
  for (int i = 1; i <= ns; i++) {
    h[i] = th[i];
    s[i] = th[ns + i];
    ds[i - 1] = s[i] - s[i - 1];
  }
  ds[ns] = T - s[ns];
  s[ns + 1] = T;
  for (int i = 0; i <= ns; i++)
    if (h[i] <= 0.0 || ds[i] <= 0.0) return -10000.0;
  double abcon = alpha * log(beta) - lgamma(alpha);
  double lp = -lam + ns * log(lam) - lgamma((double)(ns + 1));
  for (int i = 0; i <= ns; i++) {
    lp += (abcon + (alpha - 1.0) * log(h[i]) - beta * h[i]);
    lp += log(ds[i]);
  }
  lp += (lgamma(2.0 * (ns + 1)) - (2.0 * ns + 1.0) * log(T));
  /* likelihood: one segment advance per datum at most (usercpt.c:114-127) */
  int seen = 0, j = 0;
  double top = s[1], llh = 0.0;
  for (int i = 0; i < AMX_COAL_N; i++) {
    if (amx_coal_y[i] > top) {
      int nj = i - seen;
      seen = i;
      llh += (nj * log(h[j]) - h[j] * ds[j]);
      j++;
      if (j > ns) return lp; /* prior only, as the reference does */
      top = s[j + 1];
    }
  }
  llh += (AMX_COAL_N - seen) * log(h[j]) - h[j] * ds[j];
  return lp + llh;
}

/* --- finite mixture of normals, unknown number of components (BASELINE config 4) ------------
 * Not in the reference; the definition is automix_b200/workloads.py:c4_mixnorm.  Model k has K = ncomp[k] components,
 * theta = (a_1..a_{K-1} stick-breaking logits | m_1..m_K means | s_1..s_K log standard deviations). */
#define MN_KMAX 10
static int mn_nmodels = 0, mn_ncomp[AMX_MAX_MODELS], mn_n = 0;
static double mn_prior[5], *mn_y = NULL;

int amxh_select_mixnorm(int nmodels, const int *ncomp, int ndata, const double *y, const double *prior5) {
  if (nmodels > AMX_MAX_MODELS) return -1;
  mn_nmodels = nmodels;
  for (int k = 0; k < nmodels; k++) mn_ncomp[k] = ncomp[k];
  mn_n = ndata;
  free(mn_y);
  mn_y = malloc(sizeof(double) * ndata);
  memcpy(mn_y, y, sizeof(double) * ndata);
  memcpy(mn_prior, prior5, sizeof(mn_prior));
  g_kind = T_MIXNORM;
  return 0;
}

static double softplus(double x) { return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x)); }

static double mixnorm_eval(int k, const double *x) {
  const int K = mn_ncomp[k];
  const double pa = mn_prior[0], pm = mn_prior[1], pms = mn_prior[2], ps = mn_prior[3], pss = mn_prior[4];
  const double hl2pi = 0.9189385332046727;
  double cj[MN_KMAX], mj[MN_KMAX], isj[MN_KMAX], lrem = 0.0, lprior = 0.0;
  for (int j = 0; j < K; j++) {
    double lw = lrem;
    if (j < K - 1) {
      double a = x[j];
      lw = lrem - softplus(-a);
      lrem -= softplus(a);
      double za = a / pa;
      lprior += -0.5 * (za * za) - log(pa) - hl2pi;
    }
    double m = x[K - 1 + j], sl = x[2 * K - 1 + j];
    double zm = (m - pm) / pms, zs = (sl - ps) / pss;
    lprior += (-0.5 * (zm * zm) - log(pms) - hl2pi) + (-0.5 * (zs * zs) - log(pss) - hl2pi);
    cj[j] = lw - sl - hl2pi;
    mj[j] = m;
    isj[j] = exp(-sl);
  }
  double ll = 0.0;
  for (int i = 0; i < mn_n; i++) {
    double mx = -DBL_MAX, t[MN_KMAX], ssum = 0.0;
    for (int j = 0; j < K; j++) {
      double z = (mn_y[i] - mj[j]) * isj[j];
      t[j] = cj[j] - 0.5 * (z * z);
      if (t[j] > mx) mx = t[j];
    }
    for (int j = 0; j < K; j++) ssum += exp(t[j] - mx);
    ll += mx + log(ssum);
  }
  return ll + lprior;
}

/* --- the callbacks ---------------------------------------------------------- */
double amxh_logpost(int k, double *x) {
  g_calls++;
  switch (g_kind) {
    case T_GAUSSMIX: return gaussmix_eval(k, x);
    case T_QUAD: return quad_eval(k, x);
    case T_COALMINE: return coal_eval(k, x);
    case T_MIXNORM: return mixnorm_eval(k, x);
  }
  return NAN;
}

void amxh_batched(long n, const int *k, const double *x, long ldx, double *lp,
                  void *user) {
  (void)user;
  for (long i = 0; i < n; i++) lp[i] = amxh_logpost(k[i], (double *)x + i * ldx);
}

long amxh_calls(int reset) {
  long c = g_calls;
  if (reset) g_calls = 0;
  return c;
}

/* function-pointer getters for ctypes */
void *amxh_logpost_ptr(void) { return (void *)amxh_logpost; }
void *amxh_batched_ptr(void) { return (void *)amxh_batched; }
