/*
 * ref_population.c -- TEST INFRASTRUCTURE.  The UNMODIFIED reference (oracle/_ref/libautomix.so, its own generator and
 * public API) run with the POPULATION schedule the drop-in uses: R independent chains, each started by the reference's
 * own initChain (automix.c:423-449) and advanced by burn_samples(nburn) + rjmcmc_samples(nsweep), on the coal-mining
 * posterior of the reference's usercpt.c.  It separates the estimator (many short chains from a common start) from the
 * kernel: if the reference itself shows the same model probabilities under this schedule, a deviation from its single
 * long chain belongs to the schedule.  Prints per-replicate visit counts (between-chain variance gives the Monte-Carlo
 * error) -- and writes the proposal it fitted so that the GPU side can run on the identical proposal.
 *
 * usage: ref_population R nburn nsweep seed_fit seed_chains [mix_out.data]   (NOADAPT=1 in the environment: doAdapt = 0)
 */
#include "automix.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void sdrni(unsigned long *seed);
double ref_cpt(int k, double *x);
void ref_cpt_init(int k, int mdim, double *rwm);

int main(int argc, char **argv) {
  int R = argc > 1 ? atoi(argv[1]) : 64, nburn = argc > 2 ? atoi(argv[2]) : 2000, nsweep = argc > 3 ? atoi(argv[3]) : 2000;
  unsigned long seed = argc > 4 ? strtoul(argv[4], 0, 10) : 1851;
  unsigned long seed2 = argc > 5 ? strtoul(argv[5], 0, 10) : 0;
  int dims[6];
  double init[48];
  int pos = 0;
  for (int k = 0; k < 6; k++) {
    dims[k] = 2 * k + 3;
    ref_cpt_init(k, dims[k], init + pos);
    pos += dims[k];
  }
  amSampler am;
  initAMSampler(&am, 6, dims, ref_cpt, init);
  if (getenv("NOADAPT")) am.doAdapt = 0;
  sdrni(&seed);
  estimate_conditional_probs(&am, 100000);
  fprintf(stderr, "fitted L = %d %d %d %d %d %d (%.1f s)\n", am.jd.nMixComps[0], am.jd.nMixComps[1], am.jd.nMixComps[2],
          am.jd.nMixComps[3], am.jd.nMixComps[4], am.jd.nMixComps[5], am.cpstats.timesecs_condprobs);
  if (argc > 6) {
    FILE *f = fopen(argv[6], "w");
    fprintf(f, "%d\n", 6);
    for (int k = 0; k < 6; k++) fprintf(f, "%d\n", dims[k]);
    for (int k = 0; k < 6; k++) {
      int d = dims[k], L = am.jd.nMixComps[k];
      for (int i = 0; i < d; i++) fprintf(f, "%.17g\n", am.jd.sig[k][i]);
      fprintf(f, "%d\n", L);
      for (int l = 0; l < L; l++) {
        fprintf(f, "%.17g\n", am.jd.lambda[k][l]);
        for (int i = 0; i < d; i++) fprintf(f, "%.17g\n", am.jd.mu[k][l][i]);
        for (int i = 0; i < d; i++)
          for (int j = 0; j <= i; j++) fprintf(f, "%.17g\n", am.jd.B[k][l][i][j]);
      }
    }
    fclose(f);
  }
  if (seed2) sdrni(&seed2); /* the chains' stream, independent of the fit's */
  double tot[6] = {0}, tot2[6] = {0};
  for (int r = 0; r < R; r++) {
    am.ch.isInitialized = 0; /* a fresh chain: initChain runs again */
    burn_samples(&am, nburn);
    rjmcmc_samples(&am, nsweep);
    printf("%d", r);
    for (int k = 0; k < 6; k++) {
      double p = am.st.ksummary[k] / (double)nsweep;
      tot[k] += p;
      tot2[k] += p * p;
      printf(" %d", am.st.ksummary[k]);
    }
    printf("\n");
  }
  fprintf(stderr, "P(k) over %d chains x (%d burn + %d):", R, nburn, nsweep);
  for (int k = 0; k < 6; k++) fprintf(stderr, " %.4f", tot[k] / R);
  fprintf(stderr, "\n  MC s.e.:");
  for (int k = 0; k < 6; k++) {
    double m = tot[k] / R, v = (tot2[k] / R - m * m) / (R > 1 ? R - 1 : 1);
    fprintf(stderr, " %.4f", sqrt(v > 0 ? v : 0));
  }
  fprintf(stderr, "\n");
  return 0;
}
