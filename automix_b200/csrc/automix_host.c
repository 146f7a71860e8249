/*
 * automix_host.c -- the LibAutoMix 2.1 public API (include/automix.h) implemented in plain C on
 * top of the CUDA C-ABI (include/amx.h).  This is the host side of the drop-in: same function
 * names, argument meaning, ownership rules and error conventions as the reference
 * (reference automix.c:77-252), with the three stages executed by the sm_100a kernels.
 *
 *   initAMSampler ............ reference :197-240   (allocation, defaults, clock seed)
 *   estimate_conditional_probs  reference :157-195   -> amx_rwm_adapt + amx_em_fit / amx_autorj_fit
 *   burn_samples .............. reference :135-155   -> amx_rj_sweeps(burning)
 *   rjmcmc_samples ............ reference :77-133    -> amx_rj_sweeps + trace of chain 0
 *   freeAMSampler ............. reference :242-252
 *   sdrand / sdrni / loggamma . reference :1297-1316, :1323-1579 (exported because the reference's
 *                               example programs call them)
 *
 * The sweeps run for a population of independent chains; chain 0 is traced and fills the legacy
 * per-sweep arrays of runStats, the population totals go to amx_sampler_stats (64-bit).
 * There is no CPU path: if the GPU work fails the call prints the reason to stderr, records it in
 * amx_sampler_stats.last_error and returns (the reference's entry points are void).
 */
#include "automix.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "amx.h"

/* ---- per-sampler extension record, kept outside the ABI-frozen structs ---------------------------- */
typedef struct sampler_ext {
  const amSampler *am;
  const amx_target *user_target; /* set by amx_sampler_set_target, not owned */
  amx_target *own_target;        /* wraps am->logposterior */
  amx_proposal *prop;
  amx_rj *rj;
  long rj_chains, rwm_chains;
  uint64_t seed;
  int seed_set;
  int pk_mode; /* -1: default (population when more than one chain) */
  int st_alloc_nsweep; /* runStats arrays we own */
  amx_sampler_stats stats;
  struct sampler_ext *next;
} sampler_ext;

static sampler_ext *g_ext = NULL;

static sampler_ext *ext_of(const amSampler *am, int create) {
  for (sampler_ext *e = g_ext; e; e = e->next)
    if (e->am == am) return e;
  if (!create) return NULL;
  sampler_ext *e = (sampler_ext *)calloc(1, sizeof(*e));
  e->am = am;
  e->pk_mode = -1;
  e->next = g_ext;
  g_ext = e;
  return e;
}

static void ext_drop(const amSampler *am) {
  sampler_ext **pp = &g_ext;
  while (*pp) {
    if ((*pp)->am == am) {
      sampler_ext *e = *pp;
      *pp = e->next;
      if (e->rj) amx_rj_destroy(e->rj);
      if (e->prop) amx_proposal_destroy(e->prop);
      if (e->own_target) amx_target_destroy(e->own_target);
      free(e);
      return;
    }
    pp = &(*pp)->next;
  }
}

static int report(sampler_ext *e, const char *where, int rc) {
  if (rc != AMX_OK) {
    fprintf(stderr, "automix-b200: %s failed (%d): %s\n", where, rc, amx_last_error());
    if (e) e->stats.last_error = rc;
  }
  return rc;
}

static double wall_seconds(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* ---- the SuperDuper generator the reference exports (LCG x69069 mod 2^32 xor Tausworthe 15/17) --- */
static unsigned long sd_lcg = 1, sd_taus = 1;

double sdrand(void) {
  sd_lcg = (sd_lcg * 69069UL) & 0xFFFFFFFFUL;
  sd_taus ^= sd_taus >> 15;
  sd_taus ^= (sd_taus << 17) & 0xFFFFFFFFUL;
  return (double)((sd_taus ^ sd_lcg) >> 1) * 4.656612873E-10;
}

void sdrni(unsigned long *seed) {
  unsigned long s = *seed;
  if (s == 0) s = (unsigned long)time(0);
  sd_taus = s / 65536;
  sd_lcg = s - 65536 * sd_taus;
  sd_taus = 65536 * sd_taus + 1;
  sd_lcg = 32768 * sd_lcg + 1;
  *seed = s;
}

/* The reference's loggamma (Cody & Hillstrom's ALGAMA, automix.c:1323-1579) is defined for 0 < x <= XBIG and returns
 * XINF = 1.79E308 for every other argument.  User log-posteriors lean on that: tests/test_automix.c:311-321 evaluates
 * alpha * log(beta) - loggamma(alpha) with no support check, and it is the 1.79E308 that keeps the chain out of
 * alpha <= 0 (lgamma is finite there). */
double loggamma(double x) { return (x > 0.0 && x <= 2.55E305) ? lgamma(x) : 1.79E308; }

/* ---- construction / destruction ---------------------------------------------------------------------- */
static int alloc_proposal(proposalDist *jd, int nmodels, const int *dims, int Lcap) {
  jd->nmodels = nmodels;
  jd->NUM_MIX_COMPS_MAX = Lcap;
  jd->model_dims = (int *)malloc(sizeof(int) * (nmodels > 0 ? nmodels : 1));
  jd->nMixComps = (int *)calloc(nmodels > 0 ? nmodels : 1, sizeof(int));
  jd->lambda = (double **)calloc(nmodels > 0 ? nmodels : 1, sizeof(double *));
  jd->mu = (double ***)calloc(nmodels > 0 ? nmodels : 1, sizeof(double **));
  jd->B = (double ****)calloc(nmodels > 0 ? nmodels : 1, sizeof(double ***));
  jd->sig = (double **)calloc(nmodels > 0 ? nmodels : 1, sizeof(double *));
  if (!jd->model_dims || !jd->nMixComps || !jd->lambda || !jd->mu || !jd->B || !jd->sig) return EXIT_FAILURE;
  for (int k = 0; k < nmodels; k++) {
    const int d = dims[k];
    jd->model_dims[k] = d;
    jd->lambda[k] = (double *)calloc(Lcap, sizeof(double));
    jd->mu[k] = (double **)calloc(Lcap, sizeof(double *));
    jd->B[k] = (double ***)calloc(Lcap, sizeof(double **));
    jd->sig[k] = (double *)calloc(d, sizeof(double));
    if (!jd->lambda[k] || !jd->mu[k] || !jd->B[k] || !jd->sig[k]) return EXIT_FAILURE;
    for (int l = 0; l < Lcap; l++) {
      jd->mu[k][l] = (double *)calloc(d, sizeof(double));
      jd->B[k][l] = (double **)calloc(d, sizeof(double *));
      if (!jd->mu[k][l] || !jd->B[k][l]) return EXIT_FAILURE;
      for (int i = 0; i < d; i++) {
        jd->B[k][l][i] = (double *)calloc(d, sizeof(double));
        if (!jd->B[k][l][i]) return EXIT_FAILURE;
      }
    }
  }
  jd->isInitialized = true;
  return EXIT_SUCCESS;
}

static void free_proposal(proposalDist *jd) {
  if (!jd->model_dims) return;
  for (int k = 0; k < jd->nmodels; k++) {
    const int d = jd->model_dims[k];
    for (int l = 0; l < jd->NUM_MIX_COMPS_MAX; l++) {
      for (int i = 0; i < d; i++) free(jd->B[k][l][i]);
      free(jd->B[k][l]);
      free(jd->mu[k][l]);
    }
    free(jd->B[k]);
    free(jd->mu[k]);
    free(jd->lambda[k]);
    free(jd->sig[k]);
  }
  free(jd->B);
  free(jd->mu);
  free(jd->lambda);
  free(jd->sig);
  free(jd->model_dims);
  free(jd->nMixComps);
  memset(jd, 0, sizeof(*jd));
}

int initAMSampler(amSampler *am, int nmodels, int *model_dims, targetDist logpost, double *initRWM) {
  if (nmodels < 0) {
    printf("Error: negative number of models.\n");
    return EXIT_FAILURE;
  }
  if (nmodels < 1 || nmodels > AMX_MAX_MODELS) {
    printf("Error: automix-b200 supports 1..%d models (got %d).\n", AMX_MAX_MODELS, nmodels);
    return EXIT_FAILURE;
  }
  for (int k = 0; k < nmodels; k++)
    if (model_dims[k] < 1 || model_dims[k] > AMX_MAX_DIM) {
      printf("Error: automix-b200 supports model dimensions 1..%d (model %d has %d).\n", AMX_MAX_DIM, k, model_dims[k]);
      return EXIT_FAILURE;
    }
  memset(am, 0, sizeof(*am));
  am->NMODELS_MAX = 15;
  am->NUM_MIX_COMPS_MAX = 30;
  am->NUM_FITMIX_MAX = 5000;
  am->seed = 0;
  sdrni(&am->seed); /* clock seed, as the reference does at :207-208 */
  if (alloc_proposal(&am->jd, nmodels, model_dims, am->NUM_MIX_COMPS_MAX) != EXIT_SUCCESS) return EXIT_FAILURE;
  am->logposterior = logpost;
  am->initRWM = (double **)malloc(sizeof(double *) * nmodels);
  int pos = 0;
  for (int k = 0; k < nmodels; k++) {
    am->initRWM[k] = (double *)malloc(sizeof(double) * model_dims[k]);
    for (int i = 0; i < model_dims[k]; i++) am->initRWM[k][i] = initRWM ? initRWM[pos++] : sdrand();
  }
  am->cpstats.isInitialized = false;
  am->ch.isInitialized = false;
  am->st.isInitialized = false;
  am->doAdapt = true;
  am->doPerm = false;
  am->student_T_dof = 0;
  am->am_mixfit = FIGUEREIDO_MIX_FIT;
  ext_drop(am);
  ext_of(am, 1);
  /* AMX_SEED: a reproducible run of an UNCHANGED program (the reference seeds from the clock at :207-208 and offers
   * no way around it: the legacy driver's -s is overwritten) */
  const char *sv = getenv("AMX_SEED");
  if (sv && *sv) amx_sampler_set_seed(am, strtoull(sv, NULL, 10));
  return EXIT_SUCCESS;
}

static void free_cpstats(condProbStats *cp, int nmodels) {
  if (!cp->isInitialized) return;
  for (int k = 0; k < nmodels; k++) {
    if (cp->sig_k_rwm_summary && cp->sig_k_rwm_summary[k]) {
      free(cp->sig_k_rwm_summary[k][0]);
      free(cp->sig_k_rwm_summary[k]);
    }
    if (cp->nacc_ntry_rwm && cp->nacc_ntry_rwm[k]) {
      free(cp->nacc_ntry_rwm[k][0]);
      free(cp->nacc_ntry_rwm[k]);
    }
    if (cp->fitmix_annulations) free(cp->fitmix_annulations[k]);
    if (cp->fitmix_costfnnew) free(cp->fitmix_costfnnew[k]);
    if (cp->fitmix_lpn) free(cp->fitmix_lpn[k]);
    if (cp->fitmix_Lkk) free(cp->fitmix_Lkk[k]);
  }
  free(cp->sig_k_rwm_summary);
  free(cp->nacc_ntry_rwm);
  free(cp->nfitmix);
  free(cp->fitmix_annulations);
  free(cp->fitmix_costfnnew);
  free(cp->fitmix_lpn);
  free(cp->fitmix_Lkk);
  memset(cp, 0, sizeof(*cp));
}

static void free_runstats(runStats *st, int nmodels) {
  free(st->xr);
  free(st->ksummary);
  if (st->pk_summary) free(st->pk_summary[0]);
  free(st->pk_summary);
  free(st->k_which_summary);
  if (st->logp_summary) free(st->logp_summary[0]);
  free(st->logp_summary);
  if (st->theta_summary) {
    for (int k = 0; k < nmodels; k++) {
      if (st->theta_summary[k]) free(st->theta_summary[k][0]);
      free(st->theta_summary[k]);
    }
    free(st->theta_summary);
  }
  free(st->theta_summary_len);
  free(st->theta_summary_size);
  memset(st, 0, sizeof(*st));
}

void freeAMSampler(amSampler *am) {
  const int nmodels = am->jd.nmodels;
  sampler_ext *e = ext_of(am, 0);
  if (am->initRWM) {
    for (int k = 0; k < nmodels; k++) free(am->initRWM[k]);
    free(am->initRWM);
    am->initRWM = NULL;
  }
  free_cpstats(&am->cpstats, nmodels);
  if (am->ch.isInitialized) {
    free(am->ch.theta);
    free(am->ch.pk);
    am->ch.theta = am->ch.pk = NULL;
    am->ch.isInitialized = false;
  }
  /* the reference leaks runStats (its freeRunStats is never called, :242-252); users may read
   * am.st only before freeAMSampler, so releasing our arrays here keeps the contract */
  if (e && e->st_alloc_nsweep) free_runstats(&am->st, nmodels);
  free_proposal(&am->jd);
  ext_drop(am);
}

/* ---- helpers: flat views of the nested structs -------------------------------------------------------- */
static const amx_target *target_of(amSampler *am, sampler_ext *e) {
  if (e->user_target) return e->user_target;
  if (!e->own_target) e->own_target = amx_target_host_scalar(am->jd.nmodels, am->jd.model_dims, am->logposterior);
  return e->own_target;
}

static long env_long(const char *name, long dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  long x = atol(v);
  return x > 0 ? x : dflt;
}

static int unsupported_modes(amSampler *am, sampler_ext *e, const char *where) {
  if (am->student_T_dof < 0) {
    fprintf(stderr, "automix-b200: %s: negative student_T_dof\n", where);
    e->stats.last_error = AMX_EINVAL;
    return 1;
  }
  return 0;
}

/* ---- stages 1 + 2 --------------------------------------------------------------------------------------- */
void estimate_conditional_probs(amSampler *am, int nsweep2) {
  const double t0 = wall_seconds();
  sampler_ext *e = ext_of(am, 1);
  proposalDist *jd = &am->jd;
  condProbStats *cp = &am->cpstats;
  const int nm = jd->nmodels;
  if (unsupported_modes(am, e, "estimate_conditional_probs")) return;
  if (nsweep2 < 1) nsweep2 = 1; /* the reference runs max(nsweep2, 10000 d) sweeps whatever the sign (:584) */
  const amx_target *tgt = target_of(am, e);
  if (!tgt) {
    report(e, "plug-in creation", AMX_EINVAL);
    return;
  }
  /* Rows of the stage-1 traces (sigma and acceptance every 100 sweeps, :648-655).  The reference sizes each model's
   * arrays by that model's own sweep count but keeps ONE length, the last model's (initCondProbStats :254-299,
   * rwm_summary_len :266): its report writer then walks every model with that length (logwrite.c:151-154) and reads
   * past the shorter arrays -- `amcpt` segfaults on the reference.  Here every model gets the longest length, zero
   * beyond its own rows, and the arrays grow if a later call asks for more sweeps. */
  int rows_max = 1;
  for (int k = 0; k < nm; k++) {
    const int d = jd->model_dims[k];
    const int nsw = nsweep2 > 10000 * d ? nsweep2 : 10000 * d;
    const int rows = (nsw + nsw / 10) / 100 > 1 ? (nsw + nsw / 10) / 100 : 1;
    if (rows > rows_max) rows_max = rows;
  }
  const int fresh = !cp->isInitialized || cp->nfitmix == NULL;
  if (fresh) {
    cp->sig_k_rwm_summary = (double ***)calloc(nm, sizeof(double **));
    cp->nacc_ntry_rwm = (double ***)calloc(nm, sizeof(double **));
    cp->nfitmix = (int *)calloc(nm, sizeof(int));
    cp->fitmix_annulations = (int **)calloc(nm, sizeof(int *));
    cp->fitmix_costfnnew = (double **)calloc(nm, sizeof(double *));
    cp->fitmix_lpn = (double **)calloc(nm, sizeof(double *));
    cp->fitmix_Lkk = (int **)calloc(nm, sizeof(int *));
    for (int k = 0; k < nm; k++) {
      const int cap = am->NUM_FITMIX_MAX + 1;
      cp->fitmix_annulations[k] = (int *)calloc(cap, sizeof(int));
      cp->fitmix_costfnnew[k] = (double *)calloc(cap, sizeof(double));
      cp->fitmix_lpn[k] = (double *)calloc(cap, sizeof(double));
      cp->fitmix_Lkk[k] = (int *)calloc(cap, sizeof(int));
    }
  }
  if (fresh || rows_max > cp->rwm_summary_len) {
    for (int k = 0; k < nm; k++) {
      const int d = jd->model_dims[k];
      if (cp->sig_k_rwm_summary[k]) {
        free(cp->sig_k_rwm_summary[k][0]);
        free(cp->sig_k_rwm_summary[k]);
        free(cp->nacc_ntry_rwm[k][0]);
        free(cp->nacc_ntry_rwm[k]);
      }
      cp->sig_k_rwm_summary[k] = (double **)malloc(sizeof(double *) * rows_max);
      cp->nacc_ntry_rwm[k] = (double **)malloc(sizeof(double *) * rows_max);
      cp->sig_k_rwm_summary[k][0] = (double *)calloc((size_t)rows_max * d, sizeof(double));
      cp->nacc_ntry_rwm[k][0] = (double *)calloc((size_t)rows_max * d, sizeof(double));
      for (int r = 1; r < rows_max; r++) {
        cp->sig_k_rwm_summary[k][r] = cp->sig_k_rwm_summary[k][r - 1] + d;
        cp->nacc_ntry_rwm[k][r] = cp->nacc_ntry_rwm[k][r - 1] + d;
      }
    }
    cp->rwm_summary_len = rows_max;
    cp->isInitialized = true;
  }
  const long P = e->rwm_chains > 0 ? e->rwm_chains : env_long("AMX_RWM_CHAINS", 1);
  const uint64_t seed = e->seed_set ? e->seed : (uint64_t)am->seed;
  /* stage 1 for all models at once (reference :176, one model after another there): the models'
   * chains are independent, their kernels overlap on the GPU */
  size_t tot_s = 0, tot_x = 0, tot_i = 0;
  size_t off_s[AMX_MAX_MODELS], off_x[AMX_MAX_MODELS];
  for (int k = 0; k < nm; k++) {
    const int d = jd->model_dims[k];
    off_s[k] = tot_s;
    off_x[k] = tot_x;
    tot_s += (size_t)P * d;
    tot_x += (size_t)P * 1000 * d * d;
    tot_i += d;
  }
  double *sig_all = (double *)malloc(sizeof(double) * tot_s);
  double *samples_all = (double *)malloc(sizeof(double) * tot_x);
  double *init_all = (double *)malloc(sizeof(double) * tot_i);
  double *tr_sig[AMX_MAX_MODELS], *tr_acc[AMX_MAX_MODELS];
  {
    size_t q = 0;
    for (int k = 0; k < nm; k++) {
      for (int i = 0; i < jd->model_dims[k]; i++) init_all[q++] = am->initRWM[k][i];
      tr_sig[k] = cp->sig_k_rwm_summary[k][0];
      tr_acc[k] = cp->nacc_ntry_rwm[k][0];
    }
  }
  {
    double ms = 0.0;
    amx_rwm_set_dof(am->student_T_dof);
    int rc = amx_rwm_adapt_all(tgt, nsweep2, P, init_all, seed, sig_all, samples_all, tr_sig, tr_acc, &ms);
    e->stats.kernel_ms_rwm += ms;
    free(init_all);
    if (report(e, "amx_rwm_adapt_all", rc)) {
      free(sig_all);
      free(samples_all);
      return;
    }
  }
  for (int k = 0; k < nm; k++) {
    const int d = jd->model_dims[k];
    const int tri = d * (d + 1) / 2;
    const long ns = 1000L * d;
    double *sig = sig_all + off_s[k];
    double *samples = samples_all + off_x[k];
    int rc = AMX_OK;
    double *fit = samples; /* chain 0 = the reference's single chain */
    if (P > 1) {           /* pool the tails of all chains into one n x d sample set */
      fit = (double *)malloc(sizeof(double) * (size_t)ns * d);
      for (long i = 0; i < ns; i++) {
        const long c = i % P, row = ns - 1 - i / P;
        memcpy(fit + i * d, samples + ((size_t)c * ns + row) * d, sizeof(double) * d);
      }
      for (int i = 0; i < d; i++) { /* average the adapted scales */
        double s = 0.0;
        for (long c = 0; c < P; c++) s += sig[c * d + i];
        sig[i] = s / (double)P;
      }
    }
    memcpy(jd->sig[k], sig, sizeof(double) * d);
    /* stage 2 (reference :179-189) */
    const int Lmax = am->NUM_MIX_COMPS_MAX < AMX_MAX_COMPS ? am->NUM_MIX_COMPS_MAX : AMX_MAX_COMPS;
    double *wt = (double *)calloc(Lmax, sizeof(double));
    double *mean = (double *)calloc((size_t)Lmax * d, sizeof(double));
    double *tr = (double *)calloc((size_t)Lmax * tri, sizeof(double));
    int L = 0;
    if (am->am_mixfit == FIGUEREIDO_MIX_FIT) {
      int idx[AMX_MAX_COMPS];
      for (int l = 0; l < Lmax;) { /* distinct start rows from the library's uniform stream (:682-697) */
        idx[l] = (int)floor((double)ns * sdrand());
        int dup = 0;
        for (int m = 0; m < l; m++) dup |= (idx[m] == idx[l]);
        if (!dup) l++;
      }
      amx_em_result res;
      rc = amx_em_fit(d, ns, fit, Lmax, am->NUM_FITMIX_MAX, idx, wt, mean, tr, cp->fitmix_Lkk[k],
                      cp->fitmix_lpn[k], cp->fitmix_costfnnew[k], cp->fitmix_annulations[k], NULL, NULL, NULL, NULL,
                      NULL, &res);
      if (!report(e, "amx_em_fit", rc)) {
        L = res.L;
        cp->nfitmix[k] = res.iters;
        e->stats.kernel_ms_em += res.kernel_ms;
      }
    } else {
      rc = amx_autorj_fit(d, ns, fit, wt, mean, tr);
      if (!report(e, "amx_autorj_fit", rc)) L = 1;
    }
    if (rc == AMX_OK) {
      jd->nMixComps[k] = L;
      for (int l = 0; l < L; l++) {
        jd->lambda[k][l] = wt[l];
        memcpy(jd->mu[k][l], mean + (size_t)l * d, sizeof(double) * d);
        for (int i = 0; i < d; i++)
          for (int j = 0; j <= i; j++) jd->B[k][l][i][j] = tr[(size_t)l * tri + AMX_TRI(i, j)];
      }
    }
    if (fit != samples) free(fit);
    free(wt);
    free(mean);
    free(tr);
    if (rc != AMX_OK) {
      free(sig_all);
      free(samples_all);
      return;
    }
  }
  free(sig_all);
  free(samples_all);
  /* a new proposal invalidates a population built on the old one */
  if (e->rj) {
    amx_rj_destroy(e->rj);
    e->rj = NULL;
  }
  if (e->prop) {
    amx_proposal_destroy(e->prop);
    e->prop = NULL;
  }
  cp->timesecs_condprobs = wall_seconds() - t0;
}

/* ---- stage 3 ----------------------------------------------------------------------------------------------- */
static int ensure_population(amSampler *am, sampler_ext *e, int n_trace) {
  proposalDist *jd = &am->jd;
  const int nm = jd->nmodels;
  if (e->rj) return AMX_OK;
  int tw = 0, tm = 0, tt = 0, ts = 0, dmax = 0;
  for (int k = 0; k < nm; k++) {
    const int d = jd->model_dims[k], L = jd->nMixComps[k];
    tw += L;
    tm += L * d;
    tt += L * (d * (d + 1) / 2);
    ts += d;
    if (d > dmax) dmax = d;
  }
  double *wt = (double *)malloc(sizeof(double) * (tw + 1)), *mean = (double *)malloc(sizeof(double) * (tm + 1));
  double *tri = (double *)malloc(sizeof(double) * (tt + 1)), *sig = (double *)malloc(sizeof(double) * (ts + 1));
  double *init = (double *)malloc(sizeof(double) * (ts + 1));
  int a = 0, b = 0, c = 0, s = 0;
  for (int k = 0; k < nm; k++) {
    const int d = jd->model_dims[k], L = jd->nMixComps[k];
    for (int l = 0; l < L; l++) {
      wt[a++] = jd->lambda[k][l];
      for (int i = 0; i < d; i++) mean[b++] = jd->mu[k][l][i];
      for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++) tri[c++] = jd->B[k][l][i][j];
    }
    for (int i = 0; i < d; i++) {
      sig[s] = jd->sig[k][i];
      init[s++] = am->initRWM[k][i];
    }
  }
  const amx_target *tgt = target_of(am, e);
  int rc = AMX_EINVAL;
  if (tgt) {
    e->prop = amx_proposal_create(nm, jd->model_dims, jd->nMixComps, wt, mean, tri, sig);
    if (e->prop) {
      long C = e->rj_chains > 0 ? e->rj_chains : env_long("AMX_CHAINS", e->user_target ? 65536 : 64);
      const uint64_t seed = e->seed_set ? e->seed : (uint64_t)am->seed;
      e->rj = amx_rj_create(e->prop, tgt, C, init, seed ^ 0x9E3779B97F4A7C15ull, n_trace);
      if (e->rj) {
        e->stats.nchains = C;
        const int mode = e->pk_mode >= 0 ? e->pk_mode : (C > 1 ? AMX_PK_POPULATION : AMX_PK_PER_CHAIN);
        amx_rj_set_pk_mode(e->rj, mode, 0);
        rc = amx_rj_init_chains(e->rj); /* initChain (reference :423-449) for every chain */
      }
    }
  }
  free(wt);
  free(mean);
  free(tri);
  free(sig);
  free(init);
  if (rc == AMX_OK && !am->ch.isInitialized) {
    am->ch.theta = (double *)calloc(dmax, sizeof(double));
    am->ch.pk = (double *)calloc(nm, sizeof(double));
    am->ch.isInitialized = true;
  }
  return rc;
}

/* copy chain 0 into the legacy chainState */
static void mirror_chain0(amSampler *am, sampler_ext *e, int burning) {
  int k = 0, nre = 0;
  double lp = 0, lim = 0;
  unsigned long long sw = 0;
  int dmax = 0;
  for (int q = 0; q < am->jd.nmodels; q++)
    if (am->jd.model_dims[q] > dmax) dmax = am->jd.model_dims[q];
  if (amx_rj_get_state(e->rj, 0, 1, am->ch.theta, am->ch.pk, &lp, &k, &nre, &lim, &sw) != AMX_OK) return;
  am->ch.log_posterior = lp;
  am->ch.current_model_k = k;
  am->ch.mdim = am->jd.model_dims[k];
  am->ch.current_Lkk = am->jd.nMixComps[k];
  am->ch.nreinit = nre;
  am->ch.reinit = 0;
  am->ch.pkllim = lim;
  am->ch.sweep_i = (unsigned long)sw;
  am->ch.isBurning = burning;
  am->ch.doBlockRWM = (sw % 10 == 0);
}

static void collect_population(sampler_ext *e, int nm) {
  unsigned long long vis[AMX_MAX_MODELS];
  amx_rj_stats rs;
  memset(vis, 0, sizeof(vis));
  int rc = amx_rj_collect(e->rj, vis, &rs, 1);
  report(e, "amx_rj_collect", rc);
  for (int k = 0; k < nm && k < 32; k++) e->stats.visits[k] = vis[k];
  amx_rj_visit_se(e->rj, NULL, e->stats.visit_se, NULL);
  e->stats.acc_block = rs.acc_block;
  e->stats.try_block = rs.try_block;
  e->stats.acc_single = rs.acc_single;
  e->stats.try_single = rs.try_single;
  e->stats.acc_jump = rs.acc_jump;
  e->stats.try_jump = rs.try_jump;
  e->stats.kernel_ms_rj += rs.kernel_ms;
}

void burn_samples(amSampler *am, int nburn) {
  const double t0 = wall_seconds();
  sampler_ext *e = ext_of(am, 1);
  if (!am->cpstats.isInitialized) estimate_conditional_probs(am, 100000); /* reference :137-139 */
  if (!am->cpstats.isInitialized || unsupported_modes(am, e, "burn_samples") || e->stats.last_error != 0) return;
  if (report(e, "population setup", ensure_population(am, e, 1))) return;
  amx_rj_set_modes(e->rj, am->student_T_dof, am->doPerm);
  if (nburn > 0) {
    if (report(e, "amx_rj_sweeps", amx_rj_sweeps(e->rj, nburn, 1, am->doAdapt))) return;
    collect_population(e, am->jd.nmodels);
  }
  mirror_chain0(am, e, 1);
  am->st.timesecs_burn = wall_seconds() - t0;
}

void rjmcmc_samples(amSampler *am, int nsweep) {
  const double t0 = wall_seconds();
  sampler_ext *e = ext_of(am, 1);
  runStats *st = &am->st;
  const int nm = am->jd.nmodels;
  if (!am->cpstats.isInitialized) estimate_conditional_probs(am, 100000); /* reference :79-81 */
  if (nsweep < 1) return;
  int dmax = 0;
  for (int k = 0; k < nm; k++)
    if (am->jd.model_dims[k] > dmax) dmax = am->jd.model_dims[k];
  int setup_ok = am->cpstats.isInitialized && !unsupported_modes(am, e, "rjmcmc_samples") && e->stats.last_error == 0;
  if (setup_ok && report(e, "population setup", ensure_population(am, e, 1))) setup_ok = 0;

  /* the reference re-creates its statistics on every call (initRunStats never sets
   * isInitialized, :357-394); keep that observable behaviour without its leak */
  if (e->st_alloc_nsweep) free_runstats(st, nm);
  memset(st, 0, sizeof(*st));
  st->nsokal = 1;
  {
    int p = (int)(log((double)(nsweep / (2 * st->nsokal))) / log(2.0) + 0.001);
    if (p > 15) p = 15;
    if (p < 0) p = 0;
    st->nkeep = 1 << p;
  }
  st->keep = nsweep - st->nkeep * st->nsokal;
  st->xr = (double *)calloc(st->nkeep, sizeof(double));
  st->ksummary = (int *)calloc(nm, sizeof(int));
  st->pk_summary = (double **)malloc(sizeof(double *) * nsweep);
  st->pk_summary[0] = (double *)calloc((size_t)nsweep * nm, sizeof(double));
  st->logp_summary = (double **)malloc(sizeof(double *) * nsweep);
  st->logp_summary[0] = (double *)calloc((size_t)nsweep * 2, sizeof(double));
  for (int i = 1; i < nsweep; i++) {
    st->pk_summary[i] = st->pk_summary[i - 1] + nm;
    st->logp_summary[i] = st->logp_summary[i - 1] + 2;
  }
  st->k_which_summary = (int *)calloc(nsweep, sizeof(int));
  st->theta_summary_len = (int *)calloc(nm, sizeof(int));
  st->theta_summary_size = (int *)calloc(nm, sizeof(int));
  st->theta_summary = (double ***)calloc(nm, sizeof(double **));
  e->st_alloc_nsweep = nsweep;

  if (setup_ok) amx_rj_set_modes(e->rj, am->student_T_dof, am->doPerm);
  if (!setup_ok || report(e, "amx_rj_sweeps", amx_rj_sweeps(e->rj, nsweep, 0, am->doAdapt))) {
    if (!setup_ok) fprintf(stderr, "automix-b200: rjmcmc_samples: an earlier stage failed (%s); no sweeps were run\n", amx_last_error());
    /* The entry points are void (reference automix.h:86-100) and callers index am.st.theta_summary[k][i] right
     * away: after a failed GPU stage leave every row pointing at zeros instead of NULL (the error has been printed
     * and is in amx_sampler_stats.last_error). */
    for (int k = 0; k < nm; k++) {
      st->theta_summary[k] = (double **)malloc(sizeof(double *) * nsweep);
      st->theta_summary[k][0] = (double *)calloc(dmax > 0 ? dmax : 1, sizeof(double));
      for (int r = 1; r < nsweep; r++) st->theta_summary[k][r] = st->theta_summary[k][0];
      st->theta_summary_size[k] = nsweep;
    }
    return;
  }
  collect_population(e, nm);
  e->stats.sweeps_per_chain = (unsigned long long)nsweep;

  /* chain 0's per-sweep record -> the legacy arrays (reference :100-124) */
  int *tk = (int *)malloc(sizeof(int) * nsweep);
  double *tlp = (double *)malloc(sizeof(double) * nsweep);
  double *tth = (double *)malloc(sizeof(double) * (size_t)nsweep * dmax);
  if (!report(e, "amx_rj_get_trace", amx_rj_get_trace(e->rj, tk, tlp, tth, st->pk_summary[0]))) {
    for (int s = 0; s < nsweep; s++) st->ksummary[tk[s]]++;
    for (int k = 0; k < nm; k++) {
      const int len = st->ksummary[k], d = am->jd.model_dims[k];
      st->theta_summary_size[k] = len;
      st->theta_summary[k] = (double **)malloc(sizeof(double *) * (len > 0 ? len : 1));
      st->theta_summary[k][0] = (double *)malloc(sizeof(double) * (size_t)(len > 0 ? len : 1) * d);
      for (int r = 1; r < len; r++) st->theta_summary[k][r] = st->theta_summary[k][r - 1] + d;
    }
    int xr_i = 0;
    for (int s = 0; s < nsweep; s++) {
      const int k = tk[s], d = am->jd.model_dims[k];
      st->k_which_summary[s] = k + 1; /* 1-based, reference :101 */
      st->logp_summary[s][0] = tlp[s];
      memcpy(st->theta_summary[k][st->theta_summary_len[k]++], tth + (size_t)s * dmax, sizeof(double) * d);
      if (s > st->keep && ((s - st->keep) % st->nsokal == 0) && xr_i < st->nkeep) st->xr[xr_i++] = k;
    }
  }
  free(tk);
  free(tlp);
  free(tth);
  /* Posterior summaries, computed where the data is.  The reference leaves var/tau/m to its report writer
   * (sokal(st.nkeep, st.xr, ...), logwrite.c:228); the same numbers are filled in here, from the same xr
   * (its last entry is never written by the reference's loop, :122-124; it is zero here), without
   * overwriting xr as sokal() does.  The population's final states give one posterior draw per chain:
   * their per-model moments are kept for amx_sampler_posterior. */
  if (st->nkeep >= 4) report(e, "amx_sokal", amx_sokal(1, st->nkeep, st->xr, &st->var, &st->tau, &st->m));
  if (!report(e, "amx_rj_moments_reset", amx_rj_moments_reset(e->rj)))
    report(e, "amx_rj_moments_accumulate", amx_rj_moments_accumulate(e->rj));
  /* acceptance counters: population totals (64-bit fields) */
  st->naccrwmb = e->stats.acc_block;
  st->ntryrwmb = e->stats.try_block;
  st->naccrwms = e->stats.acc_single;
  st->ntryrwms = e->stats.try_single;
  st->nacctd = e->stats.acc_jump;
  st->ntrytd = e->stats.try_jump;
  mirror_chain0(am, e, 0);
  st->timesecs_rjmcmc = wall_seconds() - t0;
}

/* ---- extension ------------------------------------------------------------------------------------------------ */
int amx_sampler_set_target(amSampler *am, const struct amx_target *t) {
  sampler_ext *e = ext_of(am, 1);
  e->user_target = t;
  return AMX_OK;
}
int amx_sampler_set_chains(amSampler *am, long rj_chains, long rwm_chains) {
  sampler_ext *e = ext_of(am, 1);
  e->rj_chains = rj_chains;
  e->rwm_chains = rwm_chains;
  return AMX_OK;
}
int amx_sampler_set_pk_mode(amSampler *am, int mode) {
  sampler_ext *e = ext_of(am, 1);
  if (mode != AMX_PK_PER_CHAIN && mode != AMX_PK_POPULATION) return AMX_EINVAL;
  e->pk_mode = mode;
  if (e->rj) amx_rj_set_pk_mode(e->rj, mode, 0);
  return AMX_OK;
}
int amx_sampler_set_seed(amSampler *am, uint64_t seed) {
  sampler_ext *e = ext_of(am, 1);
  e->seed = seed;
  e->seed_set = 1;
  /* the start rows of the mixture fit come from the library's sdrand stream (:682-697), which initAMSampler seeds
   * from the clock: reseed it too, so that a seeded run is reproducible end to end */
  unsigned long s = (unsigned long)(seed ? seed : 1u);
  sdrni(&s);
  return AMX_OK;
}
int amx_sampler_posterior(const amSampler *am, int model, unsigned long long *count, double *mean, double *cov,
                          double *mean_lp) {
  sampler_ext *e = ext_of(am, 0);
  if (!e || !e->rj) return AMX_EINVAL;
  return amx_rj_moments_get(e->rj, model, count, mean, cov, mean_lp);
}
const amx_sampler_stats *amx_sampler_stats_get(const amSampler *am) {
  sampler_ext *e = ext_of(am, 0);
  return e ? &e->stats : NULL;
}

/* ---- proposal distribution on disk --------------------------------------------------
 * Same token order as the reference's <stem>_mix.data (logwrite.c:247-277: nmodels; the dimensions; per model
 * the RWM scales, the component count, then per component weight, mean and the lower triangle of B row by
 * row), so either side reads the other's files -- but written with %.17g, which round-trips a double exactly
 * (the reference prints %lf, six decimals). */
int amx_sampler_save_proposal(const amSampler *am, const char *path) {
  const proposalDist *jd = &am->jd;
  FILE *f = fopen(path, "w");
  if (!f) return AMX_EINVAL;
  fprintf(f, "%d\n", jd->nmodels);
  for (int k = 0; k < jd->nmodels; k++) fprintf(f, "%d\n", jd->model_dims[k]);
  for (int k = 0; k < jd->nmodels; k++) {
    const int d = jd->model_dims[k], L = jd->nMixComps[k];
    for (int i = 0; i < d; i++) fprintf(f, "%.17g\n", jd->sig[k][i]);
    fprintf(f, "%d\n", L);
    for (int l = 0; l < L; l++) {
      fprintf(f, "%.17g\n", jd->lambda[k][l]);
      for (int i = 0; i < d; i++) fprintf(f, "%.17g\n", jd->mu[k][l][i]);
      for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++) fprintf(f, "%.17g\n", jd->B[k][l][i][j]);
    }
  }
  return fclose(f) == 0 ? AMX_OK : AMX_EINVAL;
}

/* Reads a file in that layout into am->jd with the reference reader's checks (logwrite.c:27-109: model count
 * and dimensions must match, weights must sum to one within 1e-5 and are renormalised), and marks the
 * conditional probabilities as estimated so that burn_samples / rjmcmc_samples go straight to stage 3 -- what
 * the reference's "mode 1" intends but does not do (it forgets the flag and re-estimates). */
int amx_sampler_load_proposal(amSampler *am, const char *path) {
  proposalDist *jd = &am->jd;
  sampler_ext *e = ext_of(am, 1);
  FILE *f = fopen(path, "r");
  if (!f) return AMX_EINVAL;
  int rc = AMX_OK, v = 0;
  if (fscanf(f, "%d", &v) != 1 || v != jd->nmodels) rc = AMX_EINVAL;
  for (int k = 0; rc == AMX_OK && k < jd->nmodels; k++)
    if (fscanf(f, "%d", &v) != 1 || v != jd->model_dims[k]) rc = AMX_EINVAL;
  for (int k = 0; rc == AMX_OK && k < jd->nmodels; k++) {
    const int d = jd->model_dims[k];
    for (int i = 0; i < d && rc == AMX_OK; i++)
      if (fscanf(f, "%lf", &jd->sig[k][i]) != 1) rc = AMX_EINVAL;
    int L = 0;
    if (rc == AMX_OK && (fscanf(f, "%d", &L) != 1 || L < 1 || L > jd->NUM_MIX_COMPS_MAX || L > AMX_MAX_COMPS)) rc = AMX_EINVAL;
    if (rc != AMX_OK) break;
    jd->nMixComps[k] = L;
    double sum = 0.0;
    for (int l = 0; l < L && rc == AMX_OK; l++) {
      if (fscanf(f, "%lf", &jd->lambda[k][l]) != 1) rc = AMX_EINVAL;
      for (int i = 0; i < d && rc == AMX_OK; i++)
        if (fscanf(f, "%lf", &jd->mu[k][l][i]) != 1) rc = AMX_EINVAL;
      for (int i = 0; i < d && rc == AMX_OK; i++)
        for (int j = 0; j <= i && rc == AMX_OK; j++)
          if (fscanf(f, "%lf", &jd->B[k][l][i][j]) != 1) rc = AMX_EINVAL;
      /* weights and Cholesky diagonals feed logarithms (amx_fam_pack): refuse what would turn into NaN later */
      if (rc == AMX_OK && !(jd->lambda[k][l] > 0.0)) rc = AMX_EINVAL;
      for (int i = 0; rc == AMX_OK && i < d; i++)
        if (!(jd->B[k][l][i][i] > 0.0)) rc = AMX_EINVAL;
      sum += jd->lambda[k][l];
    }
    if (rc == AMX_OK && fabs(sum - 1.0) > 1E-5) rc = AMX_EINVAL;
    if (rc == AMX_OK && sum != 1.0)
      for (int l = 0; l < L; l++) jd->lambda[k][l] /= sum;
  }
  fclose(f);
  if (rc != AMX_OK) {
    fprintf(stderr, "automix-b200: %s is not a proposal file for this sampler\n", path);
    e->stats.last_error = rc;
    return rc;
  }
  am->cpstats.isInitialized = true;
  if (e->rj) { /* a population built on the old proposal is stale */
    amx_rj_destroy(e->rj);
    e->rj = NULL;
  }
  if (e->prop) {
    amx_proposal_destroy(e->prop);
    e->prop = NULL;
  }
  return AMX_OK;
}
