/*
 * automix.h -- drop-in header of automix-b200 for programs written against
 * LibAutoMix 2.1 (quatrope/AutoMix).
 *
 * A program that includes the reference's src/libautomix/automix.h and links
 * -lautomix recompiles against this header and links against
 * automix_b200/lib/libautomix.so unchanged: the five public functions have the
 * reference's names, argument meaning and return conventions
 * (reference automix.h:86-100), and the five structs below have the reference's
 * field order, field types and therefore its x86-64 layout (reference
 * automix.h:113-229; sizes 80/64/80/168/448 bytes, checked by the
 * _Static_asserts at the end of this file), because user code reads results
 * straight out of them (am.st.ksummary, am.st.theta_summary, am.jd.*).
 *
 * What is different underneath: the sweeps run on the GPU for a POPULATION of
 * independent chains (include/amx.h).  The legacy arrays in runStats are filled
 * from chain 0 of the population -- a chain with exactly the reference's
 * semantics -- and the population-wide totals (64-bit) are reached through
 * amx_sampler_stats().  New knobs never change these structs; they live behind
 * the amx_sampler_* calls declared at the bottom.
 */
#ifndef AUTOMIX_B200_DROPIN_H
#define AUTOMIX_B200_DROPIN_H

#define AUTOMIX_MAJOR_VERSION 2
#define AUTOMIX_MINOR_VERSION 1
#define AUTOMIX_REVISION 0
#define AUTOMIX_VERSION "2.1"

#include <stdint.h>
#include <time.h>

/* log-posterior callback: model index (0-based) and a pointer to model_dims[k]
 * doubles owned by the library; returns the log target up to a constant. */
typedef double (*targetDist)(int model_k, double *x);

/* The reference spells its flags `bool` and defines it as int (its header
 * line 49); the flags below are therefore 4-byte ints. */
#if !defined(__cplusplus) && !defined(bool)
typedef int bool;
#define true 1
#define false 0
#endif
#ifdef __cplusplus
typedef int am_bool;
#else
typedef bool am_bool;
#endif

typedef enum { FIGUEREIDO_MIX_FIT = 0, AUTORJ_MIX_FIT } automix_mix_fit;

typedef struct amSampler amSampler;
typedef struct chainState chainState;
typedef struct proposalDist proposalDist;
typedef struct condProbStats condProbStats;
typedef struct runStats runStats;

#ifdef __cplusplus
extern "C" {
#endif

int initAMSampler(amSampler *am, int nmodels, int *model_dims, targetDist logpost, double *initRWM);
void freeAMSampler(amSampler *am);
void estimate_conditional_probs(amSampler *am, int nsweeps);
void burn_samples(amSampler *am, int nsweeps);
void rjmcmc_samples(amSampler *am, int nsweeps);

/* exported by the reference library and used by its example programs
 * (usertoy1.c:8, main.c:21, usercpt.c:13, tests/test_automix.c:8) */
double sdrand(void);
void sdrni(unsigned long *seed);
double loggamma(double x);

#ifdef __cplusplus
}
#endif

struct chainState {
  double *theta;
  double *pk;
  double log_posterior;
  int current_model_k;
  int mdim;
  int current_Lkk;
  int nreinit;
  int reinit;
  double pkllim;
  am_bool doBlockRWM;
  am_bool isBurning;
  unsigned long sweep_i;
  am_bool isInitialized;
};

struct proposalDist {
  int nmodels;
  int *nMixComps;        /* L_k: fitted mixture components per model */
  int *model_dims;       /* d_k */
  double **lambda;       /* lambda[k][l] */
  double ***mu;          /* mu[k][l][i] */
  double ****B;          /* B[k][l][i][j], j<=i: Cholesky factor of the component covariance */
  double **sig;          /* sig[k][i]: adapted RWM scales */
  int NUM_MIX_COMPS_MAX;
  am_bool isInitialized;
};

struct condProbStats {
  int rwm_summary_len;
  double ***sig_k_rwm_summary;
  double ***nacc_ntry_rwm;
  int *nfitmix;
  int **fitmix_annulations;
  double **fitmix_costfnnew;
  double **fitmix_lpn;
  int **fitmix_Lkk;
  double timesecs_condprobs;
  am_bool isInitialized;
};

struct runStats {
  unsigned long naccrwmb, ntryrwmb; /* block RWM */
  unsigned long naccrwms, ntryrwms; /* single-coordinate RWM */
  unsigned long nacctd, ntrytd;     /* reversible-jump moves */
  double ***theta_summary;          /* theta_summary[k][visit][i] */
  int *theta_summary_len;
  int *theta_summary_size;
  int nsokal;
  int nkeep;
  int keep;
  int m;
  double *xr;
  double var;
  double tau;
  int *ksummary;                    /* visits per model (chain 0) */
  double **pk_summary;
  int *k_which_summary;             /* 1-based model index per sweep */
  double **logp_summary;
  double timesecs_rjmcmc;
  double timesecs_burn;
  am_bool isInitialized;
};

struct amSampler {
  int NMODELS_MAX;
  int NUM_MIX_COMPS_MAX;
  int NUM_FITMIX_MAX;
  chainState ch;
  proposalDist jd;
  condProbStats cpstats;
  runStats st;
  am_bool doAdapt;
  am_bool doPerm;
  targetDist logposterior;
  double **initRWM;
  int student_T_dof;
  automix_mix_fit am_mixfit;
  unsigned long seed;
};

#if defined(__x86_64__) && !defined(__cplusplus)
_Static_assert(sizeof(chainState) == 80, "chainState layout");
_Static_assert(sizeof(proposalDist) == 64, "proposalDist layout");
_Static_assert(sizeof(condProbStats) == 80, "condProbStats layout");
_Static_assert(sizeof(runStats) == 168, "runStats layout");
_Static_assert(sizeof(amSampler) == 448, "amSampler layout");
_Static_assert(__builtin_offsetof(amSampler, ch) == 16 && __builtin_offsetof(amSampler, jd) == 96 &&
                   __builtin_offsetof(amSampler, cpstats) == 160 && __builtin_offsetof(amSampler, st) == 240 &&
                   __builtin_offsetof(amSampler, doAdapt) == 408 && __builtin_offsetof(amSampler, logposterior) == 416 &&
                   __builtin_offsetof(amSampler, initRWM) == 424 && __builtin_offsetof(amSampler, student_T_dof) == 432 &&
                   __builtin_offsetof(amSampler, am_mixfit) == 436 && __builtin_offsetof(amSampler, seed) == 440,
               "amSampler field offsets (SURVEY.md appendix D)");
#endif

/* ---- extension: knobs and results that must not live in the structs above ----------- */
#ifdef __cplusplus
extern "C" {
#endif

struct amx_target; /* include/amx.h */

typedef struct amx_sampler_stats {
  long nchains;
  unsigned long long sweeps_per_chain;      /* of the last rjmcmc_samples call */
  unsigned long long visits[32];            /* model visits over ALL chains (64-bit) */
  unsigned long long acc_block, try_block, acc_single, try_single, acc_jump, try_jump;
  double kernel_ms_rwm, kernel_ms_em, kernel_ms_rj;
  int last_error;                           /* AMX_OK or the code of the last failure */
  double visit_se[32];                      /* Monte-Carlo standard error of visits[k] / sum(visits), from the spread
                                               between disjoint groups of chains (NaN below 256 chains) */
} amx_sampler_stats;

/* Use a __device__ plug-in (amx_target_gaussmix / _quad / _coalmine) instead of the scalar
 * callback: the whole sweep loop then runs on the GPU.  Call after initAMSampler.  The sampler
 * does not take ownership of the plug-in. */
int amx_sampler_set_target(amSampler *am, const struct amx_target *t);
/* Population size for stage 3 (default: $AMX_CHAINS, else 65536 with a device plug-in and
 * 64 with a host callback: up to there the callbacks hide behind the PCIe latency of an exchange,
 * so the run costs what one chain would) and number of independent stage-1 chains per model whose
 * stored samples are pooled for the fit (default 1 = the reference's single chain). */
int amx_sampler_set_chains(amSampler *am, long rj_chains, long rwm_chains);
/* Seed of the counter-based per-chain streams (default: am->seed, which initAMSampler takes
 * from the clock as the reference does).  Also reseeds the library's sdrand() stream, which draws
 * the start rows of the mixture fit, so a seeded run is reproducible end to end. */
int amx_sampler_set_seed(amSampler *am, uint64_t seed);
/* How the adaptive jump probabilities pk are kept by a population (include/amx.h, amx_rj_set_pk_mode):
 * 0 = per chain, the reference's rule for its one chain; 1 = one pk shared by the population, adapted from the
 * population's model-visit histogram.  Default: 1 when the population has more than one chain (many short chains
 * that each adapt their own pk inherit the reference's finite-time adaptation bias; the shared rule has none),
 * 0 for a single chain (then the run is the reference's chain). */
int amx_sampler_set_pk_mode(amSampler *am, int mode);
const amx_sampler_stats *amx_sampler_stats_get(const amSampler *am);
/* Per-model posterior moments over the population after rjmcmc_samples: one draw per chain (its final
 * state); count = chains in that model, mean[d], unbiased cov[d*d], mean log-posterior (any may be NULL).
 * rjmcmc_samples also fills am->st.var / tau / m with Sokal's integrated autocorrelation time of chain 0's
 * model-index series am->st.xr -- the numbers the reference's report writer prints (logwrite.c:228, :327). */
int amx_sampler_posterior(const amSampler *am, int model, unsigned long long *count, double *mean, double *cov,
                          double *mean_lp);
/* The fitted proposal distribution (am->jd) on disk, in the token order of the reference's <stem>_mix.data
 * (logwrite.c:247-277) but lossless (%.17g).  Loading marks the conditional probabilities as estimated, so
 * burn_samples / rjmcmc_samples skip stages 1-2 (the intent of the reference's mode 1). */
int amx_sampler_save_proposal(const amSampler *am, const char *path);
int amx_sampler_load_proposal(amSampler *am, const char *path);

#ifdef __cplusplus
}
#endif

#endif /* AUTOMIX_B200_DROPIN_H */
