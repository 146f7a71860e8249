/*
 * ref_harness.c -- flat-array doorway into the UNMODIFIED reference.
 *
 * TEST INFRASTRUCTURE (oracle side).  Linked by oracle/Makefile together with
 * an object compiled from /root/reference/src/libautomix/automix.c where it
 * lies (no reference source is copied into this repository) into
 * oracle/_ref/libautomix_tape.so.  The only thing done to the reference object
 * is `objcopy --weaken-symbol=sdrand --weaken-symbol=sdrni`, so that the strong
 * definitions below take over the uniform source: every internal draw of the
 * reference (26 call sites, all through the PLT) then reads the injected tape.
 *
 * Every ref_* function has an orc_* twin with the same signature in
 * amx_oracle.c; the tests call both with the same arguments and the same tape.
 */
#include "automix.h" /* the reference's own header, found via -I at build time */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

/* reference internals (exported, automix.c:19-69) */
void gauss(double *z, int n);
void rt(double *z, int n, int dof);
void chol(int n, double **B);
void perm(double *work, int n);
double ltprob(int dof, double z);
double lnormprob(int n, double *mu_k_l, double **B_k_l, double *datai);
double det(int n, double **B_k_l);
double rgamma(double s);
double loggamma(double x);
void rwm_within_model(int k1, int mdim, int nsweep2, condProbStats *cpstats,
                      double *sig_k, int dof, double **samples,
                      targetDist logpost, double *initRWM);
void fit_mixture_from_samples(int model_k, proposalDist jd, double **samples,
                              int nsamples, condProbStats *cpstats,
                              int NUM_MIX_COMPS_MAX, int NUM_FITMIX_MAX);
void fit_autorj(int model_k, proposalDist jd, double **samples, int nsamples);
void reversible_jump_move(bool doPerm, bool doAdapt, chainState *ch,
                          proposalDist jd, int dof, runStats *st,
                          targetDist logpost);
int initProposalDist(proposalDist *jd, int nmodels, int *model_dims,
                     int NUM_MIX_COMPS_MAX);
void freeProposalDist(proposalDist jd);

/* ---- the interposed uniform source --------------------------------------- */
static const double *g_tape = NULL;
static long g_len = 0, g_pos = 0;
static int g_over = 0;

void ref_tape_set(const double *tape, long n) {
  g_tape = tape;
  g_len = n;
  g_pos = 0;
  g_over = 0;
}
long ref_tape_used(void) { return g_pos; }
int ref_tape_overrun(void) { return g_over; }

double sdrand(void) { /* strong: overrides the weakened reference symbol */
  if (g_tape != NULL && g_pos < g_len) return g_tape[g_pos++];
  g_over = 1;
  g_pos++;
  return 0.5;
}
void sdrni(unsigned long *seed) { (void)seed; }

/* ---- helpers: packed <-> row-pointer ------------------------------------- */
static double **rows_alloc(int d) {
  double **r = malloc(sizeof(double *) * (d > 0 ? d : 1));
  for (int i = 0; i < d; i++) r[i] = calloc(d, sizeof(double));
  return r;
}
static void rows_free(double **r, int d) {
  for (int i = 0; i < d; i++) free(r[i]);
  free(r);
}
static void rows_from_packed(double **r, const double *p, int d) {
  for (int i = 0; i < d; i++)
    for (int j = 0; j <= i; j++) r[i][j] = p[TRI(i, j)];
}
static void packed_from_rows(double *p, double **r, int d) {
  for (int i = 0; i < d; i++)
    for (int j = 0; j <= i; j++) p[TRI(i, j)] = r[i][j];
}

void ref_gauss(double *z, int n) { gauss(z, n); }
void ref_rt(double *z, int n, int dof) { rt(z, n, dof); }
void ref_perm(double *v, int n) { perm(v, n); }
double ref_rgamma(double s) { return rgamma(s); }
double ref_loggamma(double x) { return loggamma(x); }
double ref_ltprob(int dof, double z) { return ltprob(dof, z); }

void ref_chol(int d, double *A) {
  double **r = rows_alloc(d);
  rows_from_packed(r, A, d);
  chol(d, r);
  packed_from_rows(A, r, d);
  rows_free(r, d);
}
double ref_det(int d, const double *B) {
  double **r = rows_alloc(d);
  rows_from_packed(r, B, d);
  double v = det(d, r);
  rows_free(r, d);
  return v;
}
double ref_lnormprob(int d, const double *mu, const double *B,
                     const double *x) {
  double **r = rows_alloc(d);
  rows_from_packed(r, B, d);
  double v = lnormprob(d, (double *)mu, r, (double *)x);
  rows_free(r, d);
  return v;
}

void ref_mix_logpdf(int d, int L, const double *lam, const double *mu,
                    const double *B, long n, const double *x, double *comp,
                    double *mix) {
  int tri = d * (d + 1) / 2;
  double ***r = malloc(sizeof(double **) * L);
  for (int l = 0; l < L; l++) {
    r[l] = rows_alloc(d);
    rows_from_packed(r[l], B + (long)l * tri, d);
  }
  for (long i = 0; i < n; i++) {
    double s = 0.0;
    for (int l = 0; l < L; l++) {
      double v = lnormprob(d, (double *)mu + (long)l * d, r[l],
                           (double *)x + i * d);
      if (comp) comp[i * L + l] = v;
      s += exp(log(lam[l]) + v);
    }
    if (mix) mix[i] = log(s);
  }
  for (int l = 0; l < L; l++) rows_free(r[l], d);
  free(r);
}

/* ---- stage 1 -------------------------------------------------------------- */
int ref_rwm_within_model(int model_k, int d, int nsweep2, int dof,
                         targetDist f, const double *init, double *sig,
                         double *samples_out, double *sig_trace,
                         double *acc_trace, double *final_state,
                         double *final_lp) {
  int nsweepr = nsweep2 > 10000 * d ? nsweep2 : 10000 * d;
  int total = nsweepr + nsweepr / 10;
  int ntrace = total / 100 + 1;
  int nsamp = 1000 * d;
  /* cpstats sized for this single model only; rwm_within_model indexes
   * [model_k][row][i] */
  condProbStats cp;
  memset(&cp, 0, sizeof(cp));
  cp.sig_k_rwm_summary = calloc(model_k + 1, sizeof(double **));
  cp.nacc_ntry_rwm = calloc(model_k + 1, sizeof(double **));
  double **a = malloc(sizeof(double *) * ntrace);
  double **b = malloc(sizeof(double *) * ntrace);
  a[0] = calloc((size_t)ntrace * d, sizeof(double));
  b[0] = calloc((size_t)ntrace * d, sizeof(double));
  for (int i = 1; i < ntrace; i++) {
    a[i] = a[i - 1] + d;
    b[i] = b[i - 1] + d;
  }
  cp.sig_k_rwm_summary[model_k] = a;
  cp.nacc_ntry_rwm[model_k] = b;
  double **samples = malloc(sizeof(double *) * nsamp);
  for (int i = 0; i < nsamp; i++) samples[i] = samples_out + (long)i * d;
  double *start = malloc(sizeof(double) * d);
  memcpy(start, init, sizeof(double) * d);

  rwm_within_model(model_k, d, nsweep2, &cp, sig, dof, samples, f, start);

  if (sig_trace) memcpy(sig_trace, a[0], sizeof(double) * (size_t)(total / 100) * d);
  if (acc_trace) memcpy(acc_trace, b[0], sizeof(double) * (size_t)(total / 100) * d);
  /* the reference does not hand back the final state; the last stored sample
   * is the state after the last sweep (remain==0 satisfies the storage rule) */
  if (final_state)
    memcpy(final_state, samples_out + (long)(nsamp - 1) * d, sizeof(double) * d);
  if (final_lp) *final_lp = NAN;
  free(a[0]);
  free(b[0]);
  free(a);
  free(b);
  free(cp.sig_k_rwm_summary);
  free(cp.nacc_ntry_rwm);
  free(samples);
  free(start);
  return total;
}

/* ---- stage 2 -------------------------------------------------------------- */
int ref_fit_mixture(int d, int n, const double *x, int Lmax, int maxit,
                    double *lam, double *mu, double *B, int *L_out,
                    int *trace_L, double *trace_loglik, double *trace_cost,
                    int *trace_ann, int *init_idx, double *cur_lam,
                    double *cur_mu, double *cur_B, int *cur_L, double *cur_w,
                    double *cur_lpd) {
  (void)init_idx; (void)cur_lam; (void)cur_mu; (void)cur_B; (void)cur_L;
  (void)cur_w; (void)cur_lpd; /* internals the reference does not expose */
  int tri = d * (d + 1) / 2;
  proposalDist jd;
  int dims1 = d;
  initProposalDist(&jd, 1, &dims1, Lmax);
  condProbStats cp;
  memset(&cp, 0, sizeof(cp));
  int cap = maxit + 4;
  int nfit = 0;
  int *ann = calloc(cap, sizeof(int)), *Ltr = calloc(cap, sizeof(int));
  double *cost = calloc(cap, sizeof(double)), *ll = calloc(cap, sizeof(double));
  cp.nfitmix = &nfit;
  cp.fitmix_annulations = &ann;
  cp.fitmix_costfnnew = &cost;
  cp.fitmix_lpn = &ll;
  cp.fitmix_Lkk = &Ltr;
  double **samples = malloc(sizeof(double *) * n);
  for (int i = 0; i < n; i++) samples[i] = (double *)x + (long)i * d;

  fit_mixture_from_samples(0, jd, samples, n, &cp, Lmax, maxit);

  int L = jd.nMixComps[0];
  *L_out = L;
  for (int l = 0; l < L; l++) {
    lam[l] = jd.lambda[0][l];
    memcpy(mu + (long)l * d, jd.mu[0][l], sizeof(double) * d);
    packed_from_rows(B + (long)l * tri, jd.B[0][l], d);
  }
  for (int i = 0; i < nfit; i++) {
    trace_L[i] = Ltr[i];
    trace_loglik[i] = ll[i];
    trace_cost[i] = cost[i];
    trace_ann[i] = ann[i];
  }
  free(ann);
  free(Ltr);
  free(cost);
  free(ll);
  free(samples);
  freeProposalDist(jd);
  return nfit;
}

void ref_fit_autorj(int d, int n, const double *x, double *lam, double *mu,
                    double *B) {
  proposalDist jd;
  int dims1 = d;
  initProposalDist(&jd, 1, &dims1, 1);
  double **samples = malloc(sizeof(double *) * n);
  for (int i = 0; i < n; i++) samples[i] = (double *)x + (long)i * d;
  fit_autorj(0, jd, samples, n);
  lam[0] = jd.lambda[0][0];
  memcpy(mu, jd.mu[0][0], sizeof(double) * d);
  packed_from_rows(B, jd.B[0][0], d);
  free(samples);
  freeProposalDist(jd);
}

/* ---- stage 3 -------------------------------------------------------------- */
static void jd_fill(proposalDist *jd, int nmodels, const int *dims,
                    const int *ncomp, const double *lam, const double *mu,
                    const double *B, const double *sig) {
  int Lcap = 1;
  for (int k = 0; k < nmodels; k++)
    if (ncomp[k] > Lcap) Lcap = ncomp[k];
  initProposalDist(jd, nmodels, (int *)dims, Lcap);
  long a = 0, b = 0, c = 0, e = 0;
  for (int k = 0; k < nmodels; k++) {
    int d = dims[k], tri = d * (d + 1) / 2;
    jd->nMixComps[k] = ncomp[k];
    for (int l = 0; l < ncomp[k]; l++) {
      jd->lambda[k][l] = lam[a + l];
      memcpy(jd->mu[k][l], mu + b + (long)l * d, sizeof(double) * d);
      rows_from_packed(jd->B[k][l], B + c + (long)l * tri, d);
    }
    memcpy(jd->sig[k], sig + e, sizeof(double) * d);
    a += ncomp[k];
    b += (long)ncomp[k] * d;
    c += (long)ncomp[k] * tri;
    e += d;
  }
}

int ref_chain_init(int nmodels, const int *dims, const double *init_flat,
                   targetDist f, double *theta, double *pk, double *lp, int *k,
                   int *nreinit, double *pkllim, unsigned long *sweep_i) {
  /* initChain (:423-449) needs a proposalDist only for nmodels/model_dims/
   * nMixComps; restated here through the reference's own entry point would
   * require a fitted jd, so call it with a minimal one. */
  void initChain(chainState * ch, proposalDist jd, double **initRWM,
                 targetDist logposterior);
  void freeChain(chainState * ch);
  proposalDist jd;
  initProposalDist(&jd, nmodels, (int *)dims, 1);
  for (int j = 0; j < nmodels; j++) jd.nMixComps[j] = 1;
  double **init = malloc(sizeof(double *) * nmodels);
  long off = 0;
  for (int j = 0; j < nmodels; j++) {
    init[j] = (double *)init_flat + off;
    off += dims[j];
  }
  chainState ch;
  memset(&ch, 0, sizeof(ch));
  initChain(&ch, jd, init, f);
  for (int i = 0; i < ch.mdim; i++) theta[i] = ch.theta[i];
  for (int j = 0; j < nmodels; j++) pk[j] = ch.pk[j];
  *lp = ch.log_posterior;
  *k = ch.current_model_k;
  *nreinit = ch.nreinit;
  *pkllim = ch.pkllim;
  *sweep_i = ch.sweep_i;
  freeChain(&ch);
  free(init);
  freeProposalDist(jd);
  return 0;
}

int ref_rj_sweeps(int nmodels, const int *dims, const int *ncomp,
                  const double *lam, const double *mu, const double *B,
                  const double *sig, targetDist f, long nsweeps, int burning,
                  int do_adapt, int do_perm, int dof, double *theta, double *pk,
                  double *lp, int *k, int *nreinit, double *pkllim,
                  unsigned long *sweep_i, int *tr_k, double *tr_lp,
                  double *tr_theta, double *tr_pk, unsigned long *cnt6,
                  long *visits) {
  proposalDist jd;
  jd_fill(&jd, nmodels, dims, ncomp, lam, mu, B, sig);
  int dmax = 0;
  for (int j = 0; j < nmodels; j++)
    if (dims[j] > dmax) dmax = dims[j];
  chainState ch;
  memset(&ch, 0, sizeof(ch));
  ch.theta = malloc(sizeof(double) * dmax);
  ch.pk = malloc(sizeof(double) * nmodels);
  memcpy(ch.theta, theta, sizeof(double) * dmax);
  memcpy(ch.pk, pk, sizeof(double) * nmodels);
  ch.log_posterior = *lp;
  ch.current_model_k = *k;
  ch.mdim = dims[*k];
  ch.current_Lkk = ncomp[*k];
  ch.nreinit = *nreinit;
  ch.reinit = 0;
  ch.pkllim = *pkllim;
  ch.sweep_i = *sweep_i;
  ch.isBurning = burning;
  ch.isInitialized = 1;
  runStats st;
  memset(&st, 0, sizeof(st));

  for (long s = 0; s < nsweeps; s++, ch.sweep_i++) {
    ch.doBlockRWM = (ch.sweep_i % 10 == 0); /* :95 / :148 */
    reversible_jump_move(do_perm, do_adapt, &ch, jd, dof, &st, f);
    if (visits) visits[ch.current_model_k]++;
    if (tr_k) tr_k[s] = ch.current_model_k;
    if (tr_lp) tr_lp[s] = ch.log_posterior;
    if (tr_theta)
      for (int i = 0; i < dmax; i++)
        tr_theta[s * dmax + i] = i < ch.mdim ? ch.theta[i] : 0.0;
    if (tr_pk)
      for (int j = 0; j < nmodels; j++) tr_pk[s * nmodels + j] = ch.pk[j];
  }
  memcpy(theta, ch.theta, sizeof(double) * dmax);
  memcpy(pk, ch.pk, sizeof(double) * nmodels);
  *lp = ch.log_posterior;
  *k = ch.current_model_k;
  *nreinit = ch.nreinit;
  *pkllim = ch.pkllim;
  *sweep_i = ch.sweep_i;
  if (cnt6) {
    cnt6[0] += st.naccrwmb;
    cnt6[1] += st.ntryrwmb;
    cnt6[2] += st.naccrwms;
    cnt6[3] += st.ntryrwms;
    cnt6[4] += st.nacctd;
    cnt6[5] += st.ntrytd;
  }
  free(ch.theta);
  free(ch.pk);
  freeProposalDist(jd);
  return 0;
}
