/*
 * ref_bench.c -- TEST / BENCH INFRASTRUCTURE.  Times the UNMODIFIED reference
 * (oracle/_ref/libautomix.so, its own generator, its own public API) on the
 * host cores: the cpu_baseline / `--impl reference` legs of bench.py.
 *
 * stage 3: initAMSampler -> the proposal distribution is filled in from a flat
 * mixture (the same one the GPU arm uses) and marked as estimated, exactly what
 * the reference's "mode 1" intends (main.c:81-87) -> sdrni(seed) ->
 * burn_samples -> rjmcmc_samples; the time is the reference's own
 * st.timesecs_rjmcmc plus a wall clock around the call.
 * stage 2: fit_mixture_from_samples called directly on a sample array.
 */
#include "automix.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

void sdrni(unsigned long *seed);
void fit_mixture_from_samples(int model_k, proposalDist jd, double **samples,
                              int nsamples, condProbStats *cpstats,
                              int NUM_MIX_COMPS_MAX, int NUM_FITMIX_MAX);
int initProposalDist(proposalDist *jd, int nmodels, int *model_dims,
                     int NUM_MIX_COMPS_MAX);
void freeProposalDist(proposalDist jd);

#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int refbench_rj(int nmodels, const int *dims, const int *ncomp,
                const double *lam, const double *mu, const double *B,
                const double *sig, const double *init_flat, targetDist f,
                int nburn, int nsweeps, unsigned long seed, long *visits,
                double *secs_reported, double *secs_wall,
                unsigned long *counters6) {
  amSampler am;
  if (initAMSampler(&am, nmodels, (int *)dims, f, (double *)init_flat)) return -1;
  long a = 0, b = 0, c = 0, e = 0;
  for (int k = 0; k < nmodels; k++) {
    int d = dims[k], tri = d * (d + 1) / 2;
    am.jd.nMixComps[k] = ncomp[k];
    for (int l = 0; l < ncomp[k]; l++) {
      am.jd.lambda[k][l] = lam[a + l];
      memcpy(am.jd.mu[k][l], mu + b + (long)l * d, sizeof(double) * d);
      for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++)
          am.jd.B[k][l][i][j] = B[c + (long)l * tri + TRI(i, j)];
    }
    memcpy(am.jd.sig[k], sig + e, sizeof(double) * d);
    a += ncomp[k];
    b += (long)ncomp[k] * d;
    c += (long)ncomp[k] * tri;
    e += d;
  }
  am.cpstats.isInitialized = 1; /* the proposal is given: skip stages 1-2 */
  /* freeCondProbStats walks these arrays when isInitialized is set */
  am.cpstats.sig_k_rwm_summary = NULL;
  am.cpstats.nacc_ntry_rwm = NULL;
  am.cpstats.nfitmix = NULL;
  am.cpstats.fitmix_annulations = NULL;
  am.cpstats.fitmix_costfnnew = NULL;
  am.cpstats.fitmix_lpn = NULL;
  am.cpstats.fitmix_Lkk = NULL;
  sdrni(&seed);
  burn_samples(&am, nburn);
  double t0 = now();
  rjmcmc_samples(&am, nsweeps);
  *secs_wall = now() - t0;
  *secs_reported = am.st.timesecs_rjmcmc;
  for (int k = 0; k < nmodels; k++) visits[k] = am.st.ksummary[k];
  counters6[0] = am.st.naccrwmb;
  counters6[1] = am.st.ntryrwmb;
  counters6[2] = am.st.naccrwms;
  counters6[3] = am.st.ntryrwms;
  counters6[4] = am.st.nacctd;
  counters6[5] = am.st.ntrytd;
  am.cpstats.isInitialized = 0; /* nothing of cpstats was allocated */
  freeAMSampler(&am);
  return 0;
}

int refbench_em(int d, int n, const double *x, int Lmax, int maxit,
                unsigned long seed, int *iters, int *L_out, double *secs_wall) {
  proposalDist jd;
  int dims1 = d;
  initProposalDist(&jd, 1, &dims1, Lmax);
  condProbStats cp;
  memset(&cp, 0, sizeof(cp));
  int cap = maxit + 4, nfit = 0;
  int *ann = calloc(cap, sizeof(int)), *Ltr = calloc(cap, sizeof(int));
  double *cost = calloc(cap, sizeof(double)), *ll = calloc(cap, sizeof(double));
  cp.nfitmix = &nfit;
  cp.fitmix_annulations = &ann;
  cp.fitmix_costfnnew = &cost;
  cp.fitmix_lpn = &ll;
  cp.fitmix_Lkk = &Ltr;
  double **rows = malloc(sizeof(double *) * n);
  for (int i = 0; i < n; i++) rows[i] = (double *)x + (long)i * d;
  sdrni(&seed);
  double t0 = now();
  fit_mixture_from_samples(0, jd, rows, n, &cp, Lmax, maxit);
  *secs_wall = now() - t0;
  *iters = nfit;
  *L_out = jd.nMixComps[0];
  free(ann); free(Ltr); free(cost); free(ll); free(rows);
  freeProposalDist(jd);
  return 0;
}
