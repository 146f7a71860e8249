/*
 * amx.h -- C-ABI of automix-b200: the flat, torch-free doorway to the sm_100a
 * kernels that replace LibAutoMix's sampling hot path.
 *
 * Plain C: pointers, sizes, opaque handles.  No CUDA or torch types appear in
 * any signature (streams travel as void*).  The drop-in LibAutoMix API
 * (include/automix.h: initAMSampler, estimate_conditional_probs, burn_samples,
 * rjmcmc_samples, freeAMSampler) is implemented in C on top of exactly these
 * calls (automix_b200/csrc/automix_host.c); a binding from another language
 * binds these.
 *
 * Each block names the reference interface it replaces; citations are to
 * /root/reference/src/libautomix/automix.c unless another file is given.
 *
 * Conventions
 *   - all functions return 0 on success, a negative AMX_E* code on failure;
 *     amx_last_error() gives the message (the reference has void returns and no
 *     error channel: automix.h:86-100).
 *   - "flat" mixture arguments follow include/amx_layout.h.
 *   - *_dev entry points take device pointers and enqueue on the library stream
 *     (amx_set_stream); the others take host pointers and include the copies.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with AMX_ENODEV.
 */
#ifndef AMX_H
#define AMX_H

#include <stddef.h>
#include <stdint.h>

#include "amx_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

#define AMX_OK 0
#define AMX_EINVAL (-1)
#define AMX_ENODEV (-2)
#define AMX_ECUDA (-3)
#define AMX_ENOMEM (-4)
#define AMX_ETAPE (-5)   /* injected uniform tape exhausted */
#define AMX_ENUMERIC (-6) /* non-positive-definite scatter, NaN log-posterior */

/* ---- runtime ------------------------------------------------------------ */
/* Threading: the library is driven by ONE host thread per process (one process per GPU, as the reference is a
 * single-threaded library with global generator state).  The current stream, the deferred-sync flag, the launch
 * counter and the drop-in layer's sampler table are process-wide and unsynchronised; only the error text is
 * per thread.  Host callbacks are called on the calling thread. */
const char *amx_last_error(void);
const char *amx_version(void);
int amx_device_count(void);
int amx_set_device(int ordinal);
/* Library work is enqueued on this stream (a cudaStream_t passed as void*;
 * NULL = the legacy default stream). */
int amx_set_stream(void *cuda_stream);
/* The mixture fit keeps its device workspace (about n (2d + Lmax + 2) doubles) for the next fit of the same shape
 * instead of returning it to the driver after every call; this hands the idle blocks back. */
int amx_release_workspace(void);
/* Deferred synchronisation for pipelines (default off).  When on, amx_rj_set_state and amx_rj_get_state only
 * ENQUEUE their transfers on the current stream: the host buffers must be pinned and must not be touched until
 * amx_synchronize() (or a sync of that stream) returns.  With two populations on two streams the transfers of
 * one overlap the sweeps of the other. */
int amx_set_deferred_sync(int on);
int amx_synchronize(void);
/* Kernels launched by this library since the counter was last reset. */
unsigned long long amx_launch_count(int reset);
/* Device-to-device copy on the library stream (lets a torch tensor receive library results
 * for an NCCL collective without a trip through the host). */
int amx_copy_dev(void *dst_dev, const void *src_dev, size_t bytes);
/* Dependent-free DFMA micro-benchmark: measured fp64 FLOP/s of this GPU
 * (the roofline denominator for the fp64-bound kernels). */
int amx_measure_fp64_peak(double *flops_per_s);

/* ---- K4: mixture / MVN-Cholesky log-density ----------------------------- */
/* Replaces lnormprob (:1727-1750) + det (:1752-1761), batched.
 *   x        n x d row-major
 *   comp_out n x L row-major component log-densities (may be NULL)
 *   mix_out  n   log sum_l wt_l N(x; mean_l, tri_l tri_l^T)      (may be NULL)
 */
int amx_mix_logpdf(int d, int L, const double *wt, const double *mean,
                   const double *tri, long n, const double *x, double *comp_out,
                   double *mix_out);
int amx_mix_logpdf_dev(int d, int L, const double *wt, const double *mean,
                       const double *tri, long n, const double *x_dev,
                       double *comp_out_dev, double *mix_out_dev);

/* ---- log-posterior plug-ins --------------------------------------------- */
/* The reference contract is `double f(int model_k, double *x)` (automix.h:46).
 * Beside it: built-in __device__ families (full speed) and a batched host
 * callback. */
typedef struct amx_target amx_target;
typedef double (*amx_scalar_fn)(int model_k, double *x);
typedef void (*amx_batched_fn)(long n, const int *model_k, const double *x,
                               long ldx, double *lp_out, void *user);

#define AMX_GM_PLAIN 0 /* log(modw * sum_g c_g exp(-q_g/2)), as usertoy1.c:72-100 */
#define AMX_GM_LSE 1   /* same value through log-sum-exp (no underflow) */

/* Gaussian-mixture target: model k is modw[k] * sum_g wt N(mean, tri tri^T). */
amx_target *amx_target_gaussmix(int nmodels, const int *dims, const int *ncomp,
                                const double *modw, const double *wt,
                                const double *mean, const double *tri,
                                int flags);
/* Separable quadratic: -sum_i (x_i-c_i)^2/(2 s_i^2) inside (lo_i,hi_i), else
 * -DBL_MAX (tests/test_automix.c:242-265, README.md:66-73).  Arrays are the
 * concatenation over models; lo/hi may be NULL (unbounded). */
amx_target *amx_target_quad(int nmodels, const int *dims, const double *center,
                            const double *scale, const double *lo,
                            const double *hi);
/* Coal-mining change-point posterior (src/user_examples/usercpt.c:46-134),
 * models k=0..5, d=2k+3. */
amx_target *amx_target_coalmine(void);
/* Finite mixture of normals with an unknown number of components (BASELINE config 4, "enzyme-style" data): model k
 * has ncomp[k] <= 10 components and d = 3 ncomp[k] - 1 parameters (stick-breaking logits | means | log standard
 * deviations); y[ndata] are the observations; prior5 = (sd of the logits, mean and sd of the component means, mean and
 * sd of the log standard deviations).  The reference ships no such example; the definition is in
 * automix_b200/csrc/amx_targets.cuh (MixNormTarget) and automix_b200/workloads.py (c4_mixnorm). */
amx_target *amx_target_mixnorm(int nmodels, const int *ncomp, int ndata, const double *y, const double *prior5);
/* A USER-SUPPLIED __device__ log-posterior: the device variant of the reference's `double f(int model_k, double *x)`.
 * The user writes a CUDA header defining `struct AmxUserTarget { bind(blob, flags); flops(k); template <int DMAX>
 * double eval(k, const double (&x)[DMAX]); }` (see automix_b200/csrc/amx_plugin_tu.cu and tests/plugins/toy1_user.cuh),
 * builds it once with `python -m automix_b200.plugin build my_target.cuh` (nvcc; the same kernel templates the built-in
 * families use are instantiated for it: full speed, chain state on-chip for all sweeps), and hands the resulting shared
 * object here with the models' dimensions and an opaque parameter blob (a multiple of 8 bytes; bind() receives it).
 * Everything that takes an amx_target -- amx_rwm_adapt, amx_rj_*, amx_target_eval, amx_sampler_set_target -- accepts it. */
amx_target *amx_target_plugin(const char *so_path, int nmodels, const int *dims, const void *blob, size_t blob_bytes,
                              int flags);
amx_target *amx_target_host_scalar(int nmodels, const int *dims,
                                   amx_scalar_fn f);
amx_target *amx_target_host_batched(int nmodels, const int *dims,
                                    amx_batched_fn f, void *user);
void amx_target_destroy(amx_target *t);
/* lp_out[i] = f(model_k[i], x[i*ldx ...]) evaluated by the plug-in on the GPU
 * (device families) or through the callback (host families). */
int amx_target_eval(const amx_target *t, long n, const int *model_k,
                    const double *x, long ldx, double *lp_out);

/* ---- proposal distribution (fitted mixtures + RWM scales) --------------- */
/* The flat mirror of proposalDist (automix.h:134-153). */
typedef struct amx_proposal amx_proposal;
amx_proposal *amx_proposal_create(int nmodels, const int *dims,
                                  const int *ncomp, const double *wt,
                                  const double *mean, const double *tri,
                                  const double *sig);
void amx_proposal_destroy(amx_proposal *p);

/* ---- K3: reversible-jump sweeps over a population of chains ------------- */
/* Replaces initChain (:423-449), reversible_jump_move (:1035-1288) and the
 * sweep loops of burn_samples / rjmcmc_samples (:77-155).  Every chain is an
 * independent copy of the reference's single chain. */
typedef struct amx_rj amx_rj;

typedef struct amx_rj_stats {
  unsigned long long acc_block, try_block;   /* runStats.naccrwmb/ntryrwmb */
  unsigned long long acc_single, try_single; /* naccrwms/ntryrwms */
  unsigned long long acc_jump, try_jump;     /* nacctd/ntrytd */
  unsigned long long flops;                  /* F_RJ of SURVEY.md 8d, summed */
  unsigned long long draws;                  /* uniforms consumed */
  double kernel_ms;                          /* device time of the sweep kernels */
} amx_rj_stats;

/* init_flat: concatenated per-model start vectors (amSampler.initRWM).
 * n_trace: the first n_trace chains record per-sweep (k, lp, theta, pk). */
amx_rj *amx_rj_create(const amx_proposal *p, const amx_target *t, long nchains,
                      const double *init_flat, uint64_t seed, int n_trace);
void amx_rj_destroy(amx_rj *rj);
/* Global id of this population's first chain.  The Philox stream of a chain is keyed by
 * (seed, global chain id), so a population sharded over GPUs gives results that do not depend
 * on the number of shards. */
int amx_rj_set_chain_base(amx_rj *rj, uint64_t first_chain_id);
/* Optional modes of the sweep (amSampler.student_T_dof, amSampler.doPerm; automix.c:1174-1203): Student-t
 * innovations with their variable-length gamma rejection draws, and random permutation of the standardised
 * vector.  Defaults 0 / 0. */
int amx_rj_set_modes(amx_rj *rj, int student_t_dof, int do_perm);
/* How the adaptive model-jump probabilities pk (automix.c:1258-1282) are kept.
 *   AMX_PK_PER_CHAIN   every chain adapts its own pk after every sweep: the reference's rule, bit for bit per chain
 *                      (the parity mode).  As an estimator over many SHORT chains it carries the reference's own
 *                      finite-time adaptation bias (a chain's pk is correlated with its recent path).
 *   AMX_PK_POPULATION  one pk shared by all chains, moved between segments of `segment_sweeps` sweeps by the
 *                      reference's update with the population's model-visit fractions (the warp-aggregated
 *                      histogram of the sweep kernel) in place of the chain's indicator, then the reference's
 *                      re-initialisation rule.  Within a segment every sweep leaves the posterior invariant, so
 *                      population estimates are unbiased; judged by posterior model probabilities, not per step.
 * segment_sweeps = 0 keeps the current value (default 25). */
#define AMX_PK_PER_CHAIN 0
#define AMX_PK_POPULATION 1
int amx_rj_set_pk_mode(amx_rj *rj, int mode, int segment_sweeps);
/* Sorted mode of the sweep kernel.  One thread per chain means a warp pays for its widest chain; with models of very
 * different dimension the population is counting-sorted on the device before every `sweeps_per_sort` sweeps by
 * (current model, model the coming jump proposes), widest first, so that the lanes of a warp run the same trip counts.
 * Per-chain arithmetic and random streams are untouched (results are bit-identical).  -1 = automatic (default: on with
 * one sweep per sort for populations of >= 65536 chains whose widest model has more than 8 coordinates), 0 = off. */
int amx_rj_set_sort(amx_rj *rj, int sweeps_per_sort);
/* The shared pk of the population mode (pk[nmodels]) and its re-initialisation state; any may be NULL. */
int amx_rj_get_pk_shared(const amx_rj *rj, double *pk, int *nreinit, double *pkllim);
/* Parity mode: chain c draws tape[c*stride + i] instead of its Philox stream. */
int amx_rj_set_tape(amx_rj *rj, const double *tape, long stride);
/* Start every chain as initChain does (one uniform picks the model). */
int amx_rj_init_chains(amx_rj *rj);
/* Overwrite / read the state of chains [first, first+count) (host arrays):
 * theta count x dmax, pk count x nmodels, lp, k, nreinit, pkllim; sweep_i is
 * shared by the population. */
int amx_rj_set_state(amx_rj *rj, long first, long count, const double *theta,
                     const double *pk, const double *lp, const int *k,
                     const int *nreinit, const double *pkllim,
                     unsigned long long sweep_i);
int amx_rj_get_state(const amx_rj *rj, long first, long count, double *theta,
                     double *pk, double *lp, int *k, int *nreinit,
                     double *pkllim, unsigned long long *sweep_i);
/* Advance every chain by nsweeps sweeps.  burning: no pk adaptation (:1258).
 * Enqueues only; counters are valid after amx_rj_collect. */
int amx_rj_sweeps(amx_rj *rj, long nsweeps, int burning, int do_adapt);
/* Wait, reduce, and read back: visits[nmodels] (64-bit model-visit counts over
 * all chains and all sweeps since the last reset) and the counters. */
int amx_rj_collect(amx_rj *rj, unsigned long long *visits, amx_rj_stats *st,
                   int reset);
/* Monte-Carlo error of the model-visit fractions.  The sweep kernels keep the visit counts of AMX_RJ_GROUPS disjoint
 * groups of chains apart (chains are independent, so group means are); after amx_rj_collect this returns for each
 * model the visit fraction p[k] over the counts collected and its standard error se[k] from the spread of the
 * group fractions (NaN with fewer than two non-empty groups: populations below 256 chains), and the number of
 * non-empty groups.  Any output may be NULL. */
#define AMX_RJ_GROUPS 64
int amx_rj_visit_se(const amx_rj *rj, double *p, double *se, int *ngroups);
/* Per-sweep records of the trace chains for the last amx_rj_sweeps call:
 * k[n_trace*nsweeps], lp[...], theta[n_trace*nsweeps*dmax], pk[...*nmodels]. */
int amx_rj_get_trace(const amx_rj *rj, int *k, double *lp, double *theta,
                     double *pk);
/* Device pointer to the 64-bit visit histogram (for a torch/NCCL all-reduce
 * that never leaves HBM). */
void *amx_rj_visits_dev(amx_rj *rj);

/* ---- posterior summaries on the device (what users form from runStats) ---- */
/* Integrated autocorrelation time by Sokal's adaptive truncated periodogram, the computation of
 * the reference's report writer (user_examples/logwrite.c:354-403: sokal(n, xr, &var, &tau, &m)
 * over the model-index series recorded at automix.c:122-124), for nseries independent series of
 * length n each (x: [nseries][n], a power of two in [4, 2^20] as the reference requires).
 * Outputs per series: var (sample variance), tau (the reference's convention: twice Sokal's),
 * m (window length + 1; n + 1 when the window never closes; tau is NaN for a constant series,
 * as with the reference).  The input is not modified (the reference overwrites it). */
int amx_sokal(int nseries, long n, const double *x, double *var, double *tau, int *m);
int amx_sokal_dev(int nseries, long n, const double *x_dev, double *var, double *tau, int *m);
/* The same over the model-index series of the population's trace chains: the last nkeep sweeps
 * of the last amx_rj_sweeps call, one (var, tau, m) per trace chain; nothing leaves the device
 * but the results. */
int amx_rj_sokal(const amx_rj *rj, long nkeep, double *var, double *tau, int *m);
/* Per-model posterior moments over the population (the means/covariances a user computes from
 * runStats.theta_summary, automix.c:105-120, without materialising per-sweep rows):
 * _accumulate adds the population's CURRENT states (one draw per chain) to running totals, so a
 * caller alternates amx_rj_sweeps(thin) and _accumulate; _get returns for one model the number
 * of draws, their mean[d], unbiased covariance cov[d*d] and mean log-posterior (any may be NULL).
 * Reduction order is fixed: results are bitwise reproducible. */
int amx_rj_moments_reset(amx_rj *rj);
int amx_rj_moments_accumulate(amx_rj *rj);
int amx_rj_moments_get(const amx_rj *rj, int model, unsigned long long *count, double *mean,
                       double *cov, double *mean_lp);

/* ---- K2: Figueiredo-Jain component-wise EM mixture fit ------------------- */
/* Replaces fit_mixture_from_samples (:664-1006) and fit_autorj (:1008-1033). */
typedef struct amx_em_result {
  int L;            /* components in the minimum-cost mixture */
  int iters;        /* outer iterations performed */
  int status;       /* 0 or AMX_ENUMERIC */
  long comp_steps;  /* component steps performed (sum over iterations of L) */
  double kernel_ms; /* device time of the fit kernel */
  double flops;     /* F_EM of SURVEY.md 8d, summed over component steps */
  double bytes;     /* 8*d bytes per sample-component-step, summed */
  double bytes_requested; /* what the kernel itself asked HBM for: rows copied into the ring + rows written, summed over
                             the passes by the kernel (0 from the first-generation kernel) */
} amx_em_result;

/*
 * x: n x d row-major samples.  Lmax <= AMX_MAX_COMPS start components whose
 * means are the rows init_idx[0..Lmax) (distinct; the reference draws them by
 * rejection from its uniform stream, :682-697 -- amx_em_draw_init reproduces
 * that from a uniform tape).  maxit is NUM_FITMIX_MAX: maxit+1 outer iterations
 * run when the cap binds (:961).  Outputs: wt[Lmax], mean[Lmax*d],
 * tri[Lmax*d(d+1)/2] (first res->L valid) and per-iteration traces of capacity
 * maxit+1 (any may be NULL): trace_L, trace_loglik, trace_cost, trace_ann.
 * cur_* (optional) receive the working state at exit (not the best one), and
 * cur_w the n x Lmax responsibilities, for step-parity tests.
 */
int amx_em_fit(int d, long n, const double *x, int Lmax, int maxit,
               const int *init_idx, double *wt, double *mean, double *tri,
               int *trace_L, double *trace_loglik, double *trace_cost,
               int *trace_ann, double *cur_wt, double *cur_mean,
               double *cur_tri, int *cur_L, double *cur_w,
               amx_em_result *res);
int amx_em_fit_dev(int d, long n, const double *x_dev, int Lmax, int maxit,
                   const int *init_idx, double *wt, double *mean, double *tri,
                   int *trace_L, double *trace_loglik, double *trace_cost,
                   int *trace_ann, amx_em_result *res);
/* The same fit with the samples sharded over ndev GPUs of one box (contiguous blocks of rows).  Every GPU
 * runs the same persistent kernel on its shard; per pass they exchange only the partial sufficient statistics
 * (column sums, log-likelihood, first moment or centred scatter: <= 2 KB per GPU) through NVLink peer memory
 * inside the kernel, reduced in fixed GPU order, so every GPU takes the same annihilation / convergence
 * branches.  devices: CUDA ordinals.  One host thread drives all GPUs. */
int amx_em_fit_multi(int ndev, const int *devices, int d, long n, const double *x, int Lmax, int maxit,
                     const int *init_idx, double *wt, double *mean, double *tri, int *trace_L,
                     double *trace_loglik, double *trace_cost, int *trace_ann, double *cur_wt,
                     double *cur_mean, double *cur_tri, int *cur_L, double *cur_w, amx_em_result *res);
/* Distinct start rows from a uniform stream exactly as :682-697; returns the
 * number of uniforms consumed. */
long amx_em_draw_init(long n, int Lmax, const double *uniforms, long nuniforms,
                      int *init_idx);
/* Single-Gaussian fit (AUTORJ_MIX_FIT). */
int amx_autorj_fit(int d, long n, const double *x, double *wt, double *mean,
                   double *tri);

/* ---- K1: adaptive random-walk Metropolis within each model -------------- */
/* Replaces rwm_within_model (:575-662).  One launch advances nchains
 * independent adaptive chains of model_k for the reference's full schedule
 * (1.1*max(nsweep2, 10000 d) sweeps); chain 0 with an injected tape is the
 * reference chain.  samples_out: nchains x (1000 d) x d; sig_out: nchains x d. */
int amx_rwm_adapt(const amx_target *t, int model_k, int nsweep2, long nchains,
                  const double *init, uint64_t seed, const double *tape,
                  long tape_stride, double *sig_out, double *samples_out,
                  double *sig_trace0, double *acc_trace0, double *kernel_ms);

/* Student-t proposals for the stage-1 chains started after this call (process-wide, 0 = Gaussian). */
int amx_rwm_set_dof(int student_t_dof);
/* Stage 1 for every model at once, the models' kernels overlapping on separate streams.  init_flat,
 * sig_out (nchains x d_k) and samples_out (nchains x 1000 d_k x d_k) are concatenated in model order;
 * sig_trace0 / acc_trace0 are arrays of nmodels pointers (or NULL).  kernel_ms: device time of the stage. */
int amx_rwm_adapt_all(const amx_target *t, int nsweep2, long nchains, const double *init_flat, uint64_t seed,
                      double *sig_out, double *samples_out, double **sig_trace0, double **acc_trace0,
                      double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* AMX_H */
