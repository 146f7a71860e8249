/*
 * test_dropin.c -- end-to-end checks of the drop-in LibAutoMix API (include/automix.h) in the
 * style of the reference's own test program (tests/test_automix.c there): every case is the
 * pipeline initAMSampler -> estimate_conditional_probs -> burn_samples -> rjmcmc_samples with a
 * user callback `double f(int, double*)`, followed by a statistical check on am.st.*.
 * The scenarios are ours (same families: samplers, parameter estimation, two-model selection);
 * tolerances are tighter than the reference's +-0.5 where Monte-Carlo error allows.
 *
 * Build: gcc tests/c/test_dropin.c -Iinclude -Lautomix_b200/lib -lautomix -lm
 */
#include "automix.h"
#include "amx.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int failures = 0;
#define CHECK(cond, ...)                              \
  do {                                                \
    if (!(cond)) {                                    \
      failures++;                                     \
      printf("  FAIL %s:%d: ", __FILE__, __LINE__);   \
      printf(__VA_ARGS__);                            \
      printf("\n");                                   \
    }                                                 \
  } while (0)

static const double ydata[10] = {0.50613293, 0.70961096, 0.28166951, 0.12532996, 0.46374168,
                                 0.58337466, 0.52458217, 0.56052633, 0.57215576, 0.68698825};

/* N(0.5, 1) up to a constant */
static double lp_normal(int k, double *x) { (void)k; return -(x[0] - 0.5) * (x[0] - 0.5) / 2.0; }
/* N(1, 1) truncated to (0, 10) */
static double lp_truncnormal(int k, double *x) {
  (void)k;
  if (x[0] <= 0.0 || x[0] >= 10.0) return -DBL_MAX;
  return -(x[0] - 1.0) * (x[0] - 1.0) / 2.0;
}
/* Beta(2, 2) through the exported loggamma */
static double lp_beta22(int k, double *x) {
  (void)k;
  if (x[0] <= 0.0 || x[0] >= 1.0) return -DBL_MAX;
  return log(x[0]) + log(1.0 - x[0]) + loggamma(4.0) - 2.0 * loggamma(2.0);
}
/* likelihood of ydata under N(mu, sigma), theta = (sigma, mu) */
static double lp_normal_params(int k, double *t) {
  (void)k;
  if (t[0] <= 0.0) return -DBL_MAX;
  double s = 0.0;
  for (int i = 0; i < 10; i++) s += -(ydata[i] - t[1]) * (ydata[i] - t[1]);
  return -10.0 * log(t[0]) + s / (2.0 * t[0] * t[0]);
}
/* likelihood of ydata under Beta(a, b) */
static double lp_beta_params(int k, double *t) {
  (void)k;
  if (t[0] <= 0.0 || t[1] <= 0.0) return -DBL_MAX;
  double s = 0.0;
  for (int i = 0; i < 10; i++) s += (t[0] - 1.0) * log(ydata[i]) + (t[1] - 1.0) * log(1.0 - ydata[i]);
  return s + 10.0 * (loggamma(t[0] + t[1]) - loggamma(t[0]) - loggamma(t[1]));
}
static double lp_normal_vs_beta(int k, double *t) { return k == 0 ? lp_normal_params(0, t) : lp_beta_params(1, t); }

static void run_pipeline(amSampler *am, int nmodels, int *dims, targetDist f, double *init, int nrj) {
  int rc = initAMSampler(am, nmodels, dims, f, init);
  CHECK(rc == EXIT_SUCCESS, "initAMSampler returned %d", rc);
  amx_sampler_set_seed(am, 20261018u);
  amx_sampler_set_chains(am, 64, 1);
  estimate_conditional_probs(am, 20000);
  burn_samples(am, 2000);
  rjmcmc_samples(am, nrj);
  const amx_sampler_stats *s = amx_sampler_stats_get(am);
  CHECK(s && s->last_error == 0, "GPU stage failed: %s", amx_last_error());
}

static void moments_1d(amSampler *am, int k, int comp, double *mean, double *sd, double *lo, double *hi) {
  int n = am->st.theta_summary_len[k];
  double s = 0, ss = 0;
  *lo = DBL_MAX;
  *hi = -DBL_MAX;
  for (int i = 0; i < n; i++) {
    double v = am->st.theta_summary[k][i][comp];
    s += v;
    ss += v * v;
    if (v < *lo) *lo = v;
    if (v > *hi) *hi = v;
  }
  *mean = s / n;
  *sd = sqrt(ss / n - (*mean) * (*mean));
}

static void test_sampler(const char *name, targetDist f, double init0, double mean, double sd, double lo, double hi) {
  printf("%s ...\n", name);
  amSampler am;
  int d = 1;
  double init[1] = {init0};
  const int n = 20000;
  run_pipeline(&am, 1, &d, f, init, n);
  double m, s, a, b;
  CHECK(am.st.ksummary && am.st.ksummary[0] == n, "ksummary[0]=%d", am.st.ksummary ? am.st.ksummary[0] : -1);
  CHECK(am.st.theta_summary_len[0] == n, "theta_summary_len");
  moments_1d(&am, 0, 0, &m, &s, &a, &b);
  printf("  mean %.4f (want %.4f)  sd %.4f (want %.4f)  range [%.3f, %.3f]  L=%d sig=%.3f\n", m, mean, s, sd, a, b,
         am.jd.nMixComps[0], am.jd.sig[0][0]);
  CHECK(fabs(m - mean) < 0.08, "mean");
  CHECK(fabs(s - sd) < 0.08, "sd");
  CHECK(a > lo && b < hi, "support");
  for (int i = 0; i < n; i += 997) CHECK(am.st.k_which_summary[i] == 1, "k_which_summary is 1-based");
  CHECK(am.cpstats.nfitmix[0] > 0 && am.cpstats.fitmix_Lkk[0][am.cpstats.nfitmix[0] - 1] >= 1, "EM trace");
  CHECK(am.st.timesecs_rjmcmc > 0 && am.cpstats.timesecs_condprobs > 0, "timers");
  CHECK(am.ch.sweep_i == 1u + 2000u + (unsigned long)n, "sweep_i persists across burn -> sample (%lu)", am.ch.sweep_i);
  freeAMSampler(&am);
}

static void test_two_models(void) {
  printf("Normal-vs-Beta model selection ...\n");
  amSampler am;
  int dims[2] = {2, 2};
  double init[4] = {0.2, 0.5, 3.0, 3.0};
  const int n = 30000;
  run_pipeline(&am, 2, dims, lp_normal_vs_beta, init, n);
  const amx_sampler_stats *s = amx_sampler_stats_get(&am);
  double tot = (double)(s->visits[0] + s->visits[1]);
  double frac_pop = s->visits[0] / tot;
  double frac0 = am.st.ksummary[0] / (double)n;
  double m_sig, m_mu, sd, lo, hi;
  moments_1d(&am, 0, 0, &m_sig, &sd, &lo, &hi);
  moments_1d(&am, 0, 1, &m_mu, &sd, &lo, &hi);
  printf("  P(normal) chain0 %.3f, population %.3f (reference test expects 0.95 +- 0.5); sigma %.3f mu %.3f\n", frac0,
         frac_pop, m_sig, m_mu);
  CHECK(tot == 64.0 * n, "population visits add up (%g)", tot);
  CHECK(frac_pop > 0.85 && frac_pop < 0.995, "posterior model probability");
  CHECK(fabs(frac0 - frac_pop) < 0.05, "chain 0 agrees with the population");
  CHECK(fabs(m_mu - 0.5) < 0.05 && fabs(m_sig - 0.21) < 0.06, "posterior means of (sigma, mu)");
  CHECK(am.st.ntrytd == 64ul * n, "jump tries");
  freeAMSampler(&am);
}

static void test_device_plugin(void) {
  printf("toy1 with the __device__ Gaussian-mixture plug-in ...\n");
  int dims[2] = {1, 2}, ncomp[2] = {2, 3};
  double modw[2] = {0.3, 0.7};
  double wt[5] = {0.2, 0.8, 1.0 / 3, 1.0 / 3, 1.0 / 3};
  double mean[8] = {-3, 2, 0, 3, -4, 1, 4, 1};
  double tri[11] = {2, 1, 2, 0, 0.7071068, 1.414214, 1.060660, 0.9354143, 1.414214, -1.060660, 0.9354143};
  amx_target *t = amx_target_gaussmix(2, dims, ncomp, modw, wt, mean, tri, AMX_GM_PLAIN);
  CHECK(t != NULL, "amx_target_gaussmix: %s", amx_last_error());
  amSampler am;
  double init[3] = {0.3, -0.2, 0.4};
  initAMSampler(&am, 2, dims, NULL, init);
  amx_sampler_set_target(&am, t);
  amx_sampler_set_seed(&am, 99);
  amx_sampler_set_chains(&am, 32768, 1);
  estimate_conditional_probs(&am, 100000);
  burn_samples(&am, 1000);
  rjmcmc_samples(&am, 2000);
  const amx_sampler_stats *s = amx_sampler_stats_get(&am);
  double p0 = s->visits[0] / (double)(s->visits[0] + s->visits[1]);
  printf("  fitted L = (%d, %d); P(model 1) = %.4f (true 0.3; thesis 0.2997)\n", am.jd.nMixComps[0], am.jd.nMixComps[1], p0);
  CHECK(s->last_error == 0, "GPU stage failed: %s", amx_last_error());
  CHECK(fabs(p0 - 0.3) < 0.01, "posterior model probability");
  CHECK(am.jd.nMixComps[0] >= 1 && am.jd.nMixComps[1] >= 2, "mixture fit");
  /* posterior summaries: Sokal's tau of chain 0's model-index series (what the reference's report writer prints)
   * and per-model moments over the population.  toy1 model 1: 0.2 N(-3, 2^2) + 0.8 N(2, 1): mean 1, var 5.6 */
  unsigned long long c0 = 0, c1 = 0;
  double pm[2], pc[4], plp;
  CHECK(am.st.nkeep == 512 && am.st.m >= 2 && am.st.tau > 0.5 && am.st.tau < 200.0 && am.st.var > 0.0 && am.st.var < 0.3,
        "Sokal summary: nkeep %d var %g tau %g m %d", am.st.nkeep, am.st.var, am.st.tau, am.st.m);
  CHECK(amx_sampler_posterior(&am, 0, &c0, pm, pc, &plp) == 0, "posterior moments: %s", amx_last_error());
  printf("  tau = %.2f (m = %d); model 1 over %llu chains: mean %.3f (true 1.0) var %.3f (true 5.6)\n", am.st.tau, am.st.m, c0,
         pm[0], pc[0]);
  CHECK(fabs(pm[0] - 1.0) < 0.12 && fabs(pc[0] - 5.6) < 0.4, "model-1 posterior mean and variance");
  CHECK(amx_sampler_posterior(&am, 1, &c1, pm, pc, NULL) == 0 && c0 + c1 == 32768ull, "one draw per chain (%llu + %llu)", c0, c1);
  CHECK(pc[1] == pc[2], "covariance is symmetric");
  /* the fitted proposal on disk, and a second sampler that starts from it: stages 1-2 are skipped */
  CHECK(amx_sampler_save_proposal(&am, "/tmp/amx_toy1_mix.data") == 0, "save proposal");
  amSampler am2;
  initAMSampler(&am2, 2, dims, NULL, init);
  amx_sampler_set_target(&am2, t);
  amx_sampler_set_seed(&am2, 100);
  amx_sampler_set_chains(&am2, 32768, 1);
  CHECK(amx_sampler_load_proposal(&am2, "/tmp/amx_toy1_mix.data") == 0, "load proposal");
  am2.student_T_dof = 5; /* optional modes through the reference's own fields */
  am2.doPerm = 1;
  burn_samples(&am2, 500);
  rjmcmc_samples(&am2, 1500);
  const amx_sampler_stats *s2 = amx_sampler_stats_get(&am2);
  double q0 = s2->visits[0] / (double)(s2->visits[0] + s2->visits[1]);
  printf("  from the saved proposal, Student-t(5) + permutation: P(model 1) = %.4f, stage-1/2 kernel time %.1f ms\n", q0,
         s2->kernel_ms_rwm + s2->kernel_ms_em);
  CHECK(s2->last_error == 0 && fabs(q0 - 0.3) < 0.012, "posterior model probability from a loaded proposal");
  CHECK(s2->kernel_ms_rwm == 0.0 && s2->kernel_ms_em == 0.0, "stages 1-2 were skipped");
  /* factors and means come back bit for bit; the weights are renormalised by the loader when they do not sum to
   * exactly one (as the reference's reader does), which can move the last bit */
  CHECK(fabs(am2.jd.lambda[1][0] - am.jd.lambda[1][0]) <= 4e-16 && am2.jd.B[1][0][1][0] == am.jd.B[1][0][1][0] &&
            am2.jd.mu[1][0][1] == am.jd.mu[1][0][1],
        "proposal survived the file");
  freeAMSampler(&am2);
  freeAMSampler(&am);
  amx_target_destroy(t);
}

/* BASELINE config 3: coal-mining change points (usercpt.c), models with 1..6 change points.
 * Posterior model probabilities of the reference's sampler: 9.6e7 sweeps of the unmodified reference
 * (oracle/gen_golden_posterior.py -> tests/golden/coalmine_posterior.npz: truth_p, truth_se); the thesis (p.177)
 * prints 0.058 0.250 0.296 0.234 0.118 0.044.  Bound: 4 standard errors -- the population's own Monte-Carlo error
 * (amx_sampler_stats.visit_se: spread between 64 disjoint groups of independent chains) combined with the fixture's. */
static void test_coalmine(void) {
  printf("coal-mining change points with the __device__ plug-in ...\n");
  const double want[6] = {0.05812314, 0.25126416, 0.29729483, 0.23361300, 0.11625657, 0.04344830};
  const double want_se[6] = {5.1e-05, 2.3e-04, 3.6e-04, 3.7e-04, 2.5e-04, 1.6e-04};
  const long C = 16384;
  int dims[6];
  double init[48];
  int pos = 0;
  for (int k = 0; k < 6; k++) {
    dims[k] = 2 * k + 3;
    for (int j = 0; j < k + 2; j++) init[pos + j] = 1.0 / 200.0;                      /* usercpt.c:35-37 */
    for (int j = 1; j < k + 2; j++) init[pos + k + 1 + j] = (40907.0 * j) / (k + 2);  /* usercpt.c:38-40 */
    pos += dims[k];
  }
  amx_target *t = amx_target_coalmine();
  CHECK(t != NULL, "amx_target_coalmine: %s", amx_last_error());
  amSampler am;
  initAMSampler(&am, 6, dims, NULL, init);
  amx_sampler_set_target(&am, t);
  /* The seed picks the stage-1 chains.  The reference's scale rule sig <- max(0, sig - gamma * 0.25) (automix.c:633-637)
   * clamps at zero when a coordinate's natural scale (the rates h ~ 0.005 here) is of the order of the step gamma, so
   * some seeds end stage 1 with a scale of 0 or ~1e-6 -- that coordinate then hardly moves within its model, with the
   * reference as with this library, and the sampler needs ~4e5 burn-in sweeps instead of 1e4 (measured:
   * profiles/posterior_from_mixfile.py; seeds 7, 11, 13, 19 pass the check below, 17 and 1851 need the longer burn-in).
   * The statistical check uses a seed whose scales are all above 1e-5, and says so if that ever changes. */
  const char *sd = getenv("AMX_TEST_SEED");
  amx_sampler_set_seed(&am, sd ? strtoull(sd, NULL, 10) : 11);
  amx_sampler_set_chains(&am, C, 1); /* more than one chain: the shared (population) pk rule is the default */
  estimate_conditional_probs(&am, 100000);
  burn_samples(&am, 10000);
  rjmcmc_samples(&am, 4000);
  const amx_sampler_stats *s = amx_sampler_stats_get(&am);
  CHECK(s->last_error == 0, "GPU stage failed: %s", amx_last_error());
  double tot = 0, sigmin = 1e300;
  for (int k = 0; k < 6; k++) tot += (double)s->visits[k];
  for (int k = 0; k < 6; k++)
    for (int i = 0; i < dims[k]; i++) sigmin = am.jd.sig[k][i] < sigmin ? am.jd.sig[k][i] : sigmin;
  CHECK(sigmin > 1e-5, "stage 1 ended with a (near-)zero RWM scale for this seed (the reference's clamp): pick another seed");
  printf("  smallest RWM scale %.3g\n", sigmin);
  printf("  fitted L = (%d %d %d %d %d %d); P(k) =", am.jd.nMixComps[0], am.jd.nMixComps[1], am.jd.nMixComps[2],
         am.jd.nMixComps[3], am.jd.nMixComps[4], am.jd.nMixComps[5]);
  for (int k = 0; k < 6; k++) printf(" %.4f", s->visits[k] / tot);
  printf("\n  stage times: RWM %.0f ms, EM %.0f ms, RJ %.0f ms (kernels)\n", s->kernel_ms_rwm, s->kernel_ms_em,
         s->kernel_ms_rj);
  for (int k = 0; k < 6; k++) {
    const double se = sqrt(s->visit_se[k] * s->visit_se[k] + want_se[k] * want_se[k]);
    CHECK(s->visit_se[k] > 0.0 && s->visit_se[k] < 0.0015, "Monte-Carlo error of P(model %d): %g", k, s->visit_se[k]);
    CHECK(fabs(s->visits[k] / tot - want[k]) < 4.0 * se, "P(model %d) = %.4f, reference %.4f, 4 se = %.4f", k,
          s->visits[k] / tot, want[k], 4.0 * se);
  }
  printf("  Monte-Carlo s.e.:");
  for (int k = 0; k < 6; k++) printf(" %.5f", s->visit_se[k]);
  printf("\n");
  CHECK(tot == (double)C * 4000.0, "visits add up");
  /* the reference's own rule, chain by chain, under this schedule: its finite-time adaptation bias is reproduced
   * (P(k=5) = 0.101 +- 0.001 from 4000 chains of the reference, sched_p of the fixture), which the shared rule removed */
  amSampler am2;
  initAMSampler(&am2, 6, dims, NULL, init);
  amx_sampler_set_target(&am2, t);
  amx_sampler_set_seed(&am2, 1852);
  amx_sampler_set_chains(&am2, C, 1);
  amx_sampler_set_pk_mode(&am2, AMX_PK_PER_CHAIN);
  amx_sampler_save_proposal(&am, "/tmp/amx_cpt_mix.data");
  CHECK(amx_sampler_load_proposal(&am2, "/tmp/amx_cpt_mix.data") == 0, "load proposal");
  burn_samples(&am2, 2000);
  rjmcmc_samples(&am2, 2000);
  const amx_sampler_stats *s2 = amx_sampler_stats_get(&am2);
  double tot2 = 0;
  for (int k = 0; k < 6; k++) tot2 += (double)s2->visits[k];
  printf("  per-chain pk rule, same schedule: P(k) =");
  for (int k = 0; k < 6; k++) printf(" %.4f", s2->visits[k] / tot2);
  printf("\n");
  /* (how far below the posterior's 0.1163 depends on the proposal: 0.095 .. 0.106 over seeds; tests/test_gpu_posterior.py
   * compares with the reference itself on one and the same proposal) */
  CHECK(s2->visits[4] / tot2 < 0.111 && s2->visits[4] / tot2 > 0.085,
        "per-chain rule P(model 4) = %.4f: expected the reference's under-visit of the large models", s2->visits[4] / tot2);
  freeAMSampler(&am2);
  freeAMSampler(&am);
  amx_target_destroy(t);
}

int main(void) {
  amSampler bad;
  int d1 = 1;
  printf("negative nmodels ...\n");
  CHECK(initAMSampler(&bad, -1, &d1, lp_normal, NULL) == EXIT_FAILURE, "negative nmodels must fail");
  unsigned long seed = 12345;
  sdrni(&seed);
  double u0 = sdrand();
  CHECK(fabs(u0 - 0.25515066366218653) < 1e-16, "sdrand after sdrni(12345) = %.17g", u0);
  CHECK(fabs(loggamma(7.25) - 7.0521854507385395) < 1e-13, "loggamma");
  CHECK(loggamma(-1.5) == 1.79E308 && loggamma(0.0) == 1.79E308, "loggamma outside (0, XBIG] returns the reference's XINF");

  if (getenv("AMX_TEST_ONLY_COALMINE")) {
    test_coalmine();
    printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
  }
  test_sampler("Normal(0.5,1) sampler", lp_normal, 0.5, 0.5, 1.0, -DBL_MAX, DBL_MAX);
  test_sampler("truncated Normal sampler", lp_truncnormal, 1.0, 1.2876, 0.7939, 0.0, 10.0);
  test_sampler("Beta(2,2) sampler", lp_beta22, 0.5, 0.5, 0.2236, 0.0, 1.0);
  test_two_models();
  test_device_plugin();
  test_coalmine();
  printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
  return failures ? 1 : 0;
}
