// amx_em.cu -- K2: the Figueiredo-Jain component-wise EM mixture fit with annihilation,
// one persistent cooperative kernel per fit.  Replaces fit_mixture_from_samples
// (automix.c:664-1006); amx_autorj_fit replaces fit_autorj (:1008-1033).
//
// Data layout in HBM (all fp64, structure-of-arrays so that every pass is coalesced):
//   xT   [d][n]      samples, transposed once from the caller's row-major n x d
//   E    [Lmax][n]   E[slot][i] = exp(log N(x_i; mu_l, B_l B_l^T)) for the component living in
//                    `slot`; annihilation only edits the slot map, no column is moved
//                    (the reference shifts its lpdatagivenl columns, :832-834)
//   wnxt [n]         responsibility of the component that is updated next
//   part [grid][NV]  per-CTA partial sums of the pass in flight
//   EmCtrl           the sequential state of the algorithm (weights, means, factors, costs,
//                    traces), touched only by the "leader"
//
// Passes over the samples (thread per sample, grid-stride, identical sample->thread mapping in
// every pass):
//   SCATTER       S2 = sum_i wnxt_i (x_i-mu)(x_i-mu)^T                     reads  8(d+1) B/sample
//   DENS_REFRESH  E[c] = exp(lnormprob), then the responsibility refresh   reads  8(d+L), writes 16
//   REFRESH       w_il = lam_l E_il / sum, column sums, log-likelihood, and the first moment
//                 S1 = sum_i w_i,next x_i and wnxt for the NEXT component -- so the reference's
//                 four passes per component step (:774-778, :796-811, :815-818, :848-867)
//                 become two, with its exact centred covariance formula kept.
// Between passes the CTAs meet at a grid barrier whose last arriver (the leader) reduces the
// partials in fixed CTA order (bitwise reproducible), and runs the scalar part of the
// algorithm: weight update, annihilation, Cholesky, MML cost, convergence, traces.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "amx_internal.cuh"
#include "amx_targets.cuh"

namespace cg = cooperative_groups;

namespace amx {

constexpr int kEmThreads = 256;
constexpr int kEmWarps = kEmThreads / 32;
constexpr int kEmLmax = AMX_MAX_COMPS;
constexpr int kEmDmax = AMX_MAX_DIM;
constexpr int kEmTriMax = kEmDmax * (kEmDmax + 1) / 2;
constexpr int kEmRecMax = AMX_REC_HEAD + 2 * kEmDmax + kEmTriMax;
constexpr int kEmNV = kEmTriMax;  // values per CTA partial row (>= Lmax + Dmax + 2)

enum EmPass { kPassInitStats = 0, kPassInitDens, kPassScatter, kPassDensRefresh, kPassRefresh, kPassStop };

struct EmCtrl {
  unsigned int arrive, gen;
  int pass;        // next pass every CTA must run
  int L, c, next;  // live components, component in progress, component the refresh prepares
  int forced_pending, natural, forced;
  int iters, stop, status;
  long comp_steps;
  double flops;
  int slot[kEmLmax];
  int free_slot[kEmLmax];
  double lam[kEmLmax];
  double mu[kEmLmax][kEmDmax];
  double B[kEmLmax][kEmTriMax];
  double colsum[kEmLmax];
  double S1[kEmDmax];
  double loglik;
  double cost, cost_prev, cost_best;
  double s2;
  double rec[kEmRecMax];  // family record of the component in progress (for solve_lower)
  int best_L;
  double best_lam[kEmLmax];
  double best_mu[kEmLmax][kEmDmax];
  double best_B[kEmLmax][kEmTriMax];
};

struct EmArgs {
  int d, Lmax, maxit;
  long n, npad;
  const double *x;  // n x d row-major (device)
  double *xT, *E, *wnxt, *part;
  EmCtrl *ctrl;
  const int *init_idx;  // device
  int *trace_L, *trace_ann;
  double *trace_loglik, *trace_cost;
  double *w_out;  // optional n x Lmax responsibilities at exit
};

template <typename T>
__device__ __forceinline__ T ld_cg(const T *p) {
  return __ldcg(p);
}

// ---- grid barrier with a leader ---------------------------------------------------------------------
// Every CTA arrives; the last one returns true and must call barrier_release() when it has
// finished its serial section; the others wait inside.  Counters only grow, so no reset races.
__device__ __forceinline__ bool barrier_arrive(EmCtrl *ctrl, unsigned &epoch) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(&ctrl->arrive, 1u);
    s_last = (t == gridDim.x * (epoch + 1u) - 1u) ? 1 : 0;
    if (s_last) __threadfence();
  }
  __syncthreads();
  return s_last != 0;
}
__device__ __forceinline__ void barrier_release(EmCtrl *ctrl, unsigned &epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicExch(&ctrl->gen, epoch + 1u);
  }
  epoch++;
}
__device__ __forceinline__ void barrier_wait(EmCtrl *ctrl, unsigned &epoch) {
  if (threadIdx.x == 0) {
    while (ld_cg(&ctrl->gen) <= epoch) __nanosleep(40);
    __threadfence();
  }
  epoch++;
  __syncthreads();
}

// ---- CTA-level reduction of NV per-thread values into part[blockIdx][*] -----------------------------
template <int NVAL>
__device__ __forceinline__ void block_reduce_store(const double (&v)[NVAL], int nv, double *s_red, double *part_row) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NVAL; q++) {
    if (q < nv) {
      const double r = warp_sum(v[q]);
      if (lane == 0) s_red[warp * NVAL + q] = r;
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nv; q += blockDim.x) {
    double t = 0.0;
    for (int w = 0; w < kEmWarps; w++) t += s_red[w * NVAL + q];
    part_row[q] = t;
  }
}

// leader: sum the per-CTA rows in CTA order into s_tot[0..nv)
__device__ __forceinline__ void leader_reduce(const double *part, int nv, double *s_tot) {
  for (int q = threadIdx.x; q < nv; q += blockDim.x) {
    double t = 0.0;
    for (unsigned b = 0; b < gridDim.x; b++) t += ld_cg(part + (size_t)b * kEmNV + q);
    s_tot[q] = t;
  }
  __syncthreads();
}

// ---- leader's scalar algorithm (thread 0 only) ------------------------------------------------------
__device__ void em_renorm(EmCtrl *c) {
  double s = 0.0;
  for (int l = 0; l < c->L; l++) s += c->lam[l];
  for (int l = 0; l < c->L; l++) c->lam[l] /= s;
}
// remove component `gone`, shifting the later ones down (:823-836, :908-921)
__device__ void em_drop(EmCtrl *c, int d, int gone) {
  const int tri = d * (d + 1) / 2;
  for (int l = gone; l < c->L - 1; l++) {
    c->lam[l] = c->lam[l + 1];
    c->slot[l] = c->slot[l + 1];
    for (int j = 0; j < d; j++) c->mu[l][j] = c->mu[l + 1][j];
    for (int j = 0; j < tri; j++) c->B[l][j] = c->B[l + 1][j];
  }
  c->L--;
}
__device__ double em_cost(const EmCtrl *c, long n, int nparams) {  // :870-876
  double s = 0.0;
  for (int l = 0; l < c->L; l++) s += log((double)n * c->lam[l] / 12.0);
  return (nparams / 2.0) * s + (c->L / 2.0) * log((double)n / 12.0) + c->L * (nparams + 1) / 2.0 - c->loglik;
}
// column-oriented in-place Cholesky on packed storage (:1682-1701)
__device__ bool em_chol(int d, double *A) {
  bool ok = true;
  for (int cidx = 0; cidx < d; cidx++) {
    double s = A[AMX_TRI(cidx, cidx)];
    for (int j = 0; j < cidx; j++) s -= A[AMX_TRI(cidx, j)] * A[AMX_TRI(cidx, j)];
    if (!(s > 0.0)) ok = false;
    const double p = sqrt(s);
    A[AMX_TRI(cidx, cidx)] = p;
    for (int r = cidx + 1; r < d; r++) {
      double t = A[AMX_TRI(r, cidx)];
      for (int j = 0; j < cidx; j++) t -= A[AMX_TRI(r, j)] * A[AMX_TRI(cidx, j)];
      A[AMX_TRI(r, cidx)] = t / p;
    }
  }
  return ok;
}
// family record of component l for solve_lower (proposal flavour, include/amx_layout.h)
__device__ void em_make_rec(EmCtrl *c, int d, int l) {
  const int tri = d * (d + 1) / 2;
  double prod = 1.0;
  for (int i = 0; i < d; i++) prod *= c->B[l][AMX_TRI(i, i)];
  double *r = c->rec;
  r[0] = c->lam[l];
  r[1] = 0.0;
  r[2] = log(prod);
  r[3] = -(d / 2.0) * log(2.0 * 3.14159265358979323846) - r[2];
  for (int i = 0; i < d; i++) {
    r[AMX_REC_HEAD + i] = c->mu[l][i];
    r[AMX_REC_HEAD + d + i] = 1.0 / c->B[l][AMX_TRI(i, i)];
  }
  for (int i = 0; i < tri; i++) r[AMX_REC_HEAD + 2 * d + i] = c->B[l][i];
}

// start the update of component c->c from the current column sums and first moment (:773-801)
__device__ void em_plan_step(EmCtrl *c, const EmArgs &a) {
  const int d = a.d, nparams = d + d * (d + 1) / 2;
  for (;;) {
    const int cc = c->c;
    double tot = 0.0, wkeep = 0.0;
    for (int l = 0; l < c->L; l++) {
      const double wl = max_m(0.0, (c->colsum[l] - nparams / 2.0));
      if (l == cc) wkeep = wl;
      tot += wl;
    }
    c->lam[cc] = wkeep / tot;
    em_renorm(c);
    c->comp_steps++;
    c->flops += (double)a.n * (2.0 * d * d + 8.0 * d + 4.0 * c->L + 7.0);
    if (c->lam[cc] > 0.005) {
      for (int j = 0; j < d; j++) {
        c->mu[cc][j] = c->S1[j] / c->colsum[cc];
        c->rec[AMX_REC_HEAD + j] = c->mu[cc][j];  // the scatter pass centres on the NEW mean (:803-809)
      }
      c->pass = kPassScatter;
      return;
    }
    // natural annihilation (:821-845); the responsibilities must be refreshed before the
    // next component can be looked at
    c->natural = 1;
    em_drop(c, d, cc);
    em_renorm(c);
    c->next = (cc < c->L) ? cc : 0;
    c->pass = kPassRefresh;
    return;
  }
}

__device__ void em_finish_iteration(EmCtrl *c, const EmArgs &a) {
  if (c->iters > a.maxit) c->stop = 1;  // :961-963
  c->cost_prev = c->cost;
  const int t = c->iters - 1;
  if (a.trace_ann) a.trace_ann[t] = c->natural + c->forced;
  if (a.trace_cost) a.trace_cost[t] = c->cost;
  if (a.trace_loglik) a.trace_loglik[t] = c->loglik;
  if (a.trace_L) a.trace_L[t] = c->L;
  if (c->stop) {
    c->pass = kPassStop;
    return;
  }
  c->iters++;
  c->natural = c->forced = 0;
  c->c = 0;
  em_plan_step(c, a);
}

__device__ void em_end_of_sweep(EmCtrl *c, const EmArgs &a) {
  const int d = a.d, nparams = d + d * (d + 1) / 2, tri = d * (d + 1) / 2;
  c->cost = em_cost(c, a.n, nparams);
  if (c->iters == 1) c->cost_prev = c->cost;
  if (c->iters == 1 || c->cost < c->cost_best) {  // :881-893
    c->best_L = c->L;
    c->cost_best = c->cost;
    for (int l = 0; l < c->L; l++) {
      c->best_lam[l] = c->lam[l];
      for (int j = 0; j < d; j++) c->best_mu[l][j] = c->mu[l][j];
      for (int j = 0; j < tri; j++) c->best_B[l][j] = c->B[l][j];
    }
  }
  if (fabs(c->cost_prev - c->cost) < min_m(1E-5 * fabs(c->cost_prev), 0.01) && c->iters > 1) {  // :894
    if (c->L == 1) {
      c->stop = 1;
    } else {
      c->forced = 2;
      double lo = c->lam[0];
      int gone = 0;
      for (int l = 1; l < c->L; l++)
        if (lo > c->lam[l]) {
          lo = c->lam[l];
          gone = l;
        }
      em_drop(c, d, gone);
      em_renorm(c);
      c->forced_pending = 1;
      c->next = 0;
      c->pass = kPassRefresh;
      return;
    }
  }
  em_finish_iteration(c, a);
}

// leader after a pass: s_tot holds the reduced partial row
__device__ void em_leader(EmCtrl *c, const EmArgs &a, int pass, const double *s_tot) {
  const int d = a.d, tri = d * (d + 1) / 2;
  if (pass == kPassInitStats) {
    // :700-723 common isotropic start; s_tot = [sum x_j (d) | sum x_j^2 (d)]
    double s2 = 0.0;
    const double len = (double)a.n;
    for (int j = 0; j < d; j++) s2 += (s_tot[d + j] - s_tot[j] * s_tot[j] / len) / len;
    s2 /= (10.0 * d);
    c->s2 = s2;
    c->L = a.Lmax;
    for (int l = 0; l < c->L; l++) {
      c->slot[l] = l;
      c->lam[l] = 1.0 / c->L;
      for (int j = 0; j < d; j++) c->mu[l][j] = a.x[(size_t)a.init_idx[l] * d + j];
      for (int j = 0; j < tri; j++) c->B[l][j] = 0.0;
      for (int j = 0; j < d; j++) c->B[l][AMX_TRI(j, j)] = s2;
      if (!em_chol(d, c->B[l])) c->status = AMX_ENUMERIC;
    }
    c->c = 0;
    em_make_rec(c, d, 0);
    c->pass = kPassInitDens;
    return;
  }
  if (pass == kPassInitDens) {  // one component's start densities are in place
    c->c++;
    if (c->c < c->L) {
      em_make_rec(c, d, c->c);
      c->pass = kPassInitDens;
    } else {
      c->next = 0;
      c->iters = 0;  // the refresh that follows is the initial E-step (:733-744)
      c->pass = kPassRefresh;
    }
    return;
  }
  if (pass == kPassScatter) {
    // :803-813 centred scatter / sum of weights, then Cholesky
    const int cc = c->c;
    for (int j = 0; j < tri; j++) c->B[cc][j] = s_tot[j] / c->colsum[cc];
    if (!em_chol(d, c->B[cc])) {
      c->status = AMX_ENUMERIC;
      c->stop = 1;
      c->pass = kPassStop;
      return;
    }
    em_make_rec(c, d, cc);
    c->next = (cc + 1 < c->L) ? cc + 1 : 0;
    c->pass = kPassDensRefresh;
    return;
  }
  // a refresh finished: s_tot = [colsum (L) | loglik | fallbacks | S1 (d)]
  for (int l = 0; l < c->L; l++) c->colsum[l] = s_tot[l];
  c->loglik = s_tot[kEmLmax] - 500.0 * s_tot[kEmLmax + 1];
  for (int j = 0; j < d; j++) c->S1[j] = s_tot[kEmLmax + 2 + j];
  if (c->iters == 0) {  // initial E-step done: start outer iteration 1
    c->iters = 1;
    c->natural = c->forced = 0;
    c->c = 0;
    em_plan_step(c, a);
    return;
  }
  if (c->forced_pending) {  // refresh after a forced annihilation (:931-958)
    c->forced_pending = 0;
    const int nparams = d + d * (d + 1) / 2;
    c->cost = em_cost(c, a.n, nparams);
    em_finish_iteration(c, a);
    return;
  }
  if (pass == kPassDensRefresh) c->c++;  // component kept: move on (:819)
  if (c->c < c->L) em_plan_step(c, a);
  else em_end_of_sweep(c, a);
}

// ---- the fit kernel -------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(kEmThreads) em_fit_kernel(EmArgs a) {
  constexpr int TRI = DMAX * (DMAX + 1) / 2;
  constexpr int NVC = kEmLmax + 2 + DMAX;       // refresh partial row
  constexpr int NVMAX = TRI > NVC ? TRI : NVC;  // scatter partial row is TRI
  __shared__ double s_red[kEmWarps * NVMAX];
  __shared__ double s_tot[kEmNV];
  __shared__ double s_rec[AMX_REC_HEAD + 2 * DMAX + TRI];
  __shared__ double s_lam[kEmLmax];
  __shared__ int s_slot[kEmLmax];
  __shared__ int s_pass, s_L, s_c, s_next;

  EmCtrl *ctrl = a.ctrl;
  const int d = a.d;
  const long n = a.n, np = a.npad;
  const long stride = (long)gridDim.x * blockDim.x;
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double *part_row = a.part + (size_t)blockIdx.x * kEmNV;
  unsigned epoch = 0;
  int pass = kPassInitStats;

  for (;;) {
    // ------------------------------------------------------------------ the data pass
    int nv = 0;
    if (pass == kPassInitStats) {
      // transpose x -> xT and accumulate sum x_j, sum x_j^2 (:700-711)
      double v[2 * DMAX];
#pragma unroll
      for (int j = 0; j < 2 * DMAX; j++) v[j] = 0.0;
      for (long i = i0; i < n; i += stride) {
#pragma unroll
        for (int j = 0; j < DMAX; j++)
          if (j < d) {
            const double xv = a.x[i * d + j];
            __stcg(a.xT + (size_t)j * np + i, xv);
            v[j] += xv;
            v[DMAX + j] = fma(xv, xv, v[DMAX + j]);
          }
      }
      // pack as [sum (d) | sumsq (d)]
      double w[2 * DMAX];
#pragma unroll
      for (int j = 0; j < 2 * DMAX; j++) w[j] = 0.0;
#pragma unroll
      for (int j = 0; j < DMAX; j++)
        if (j < d) {
          aset(w, j, v[j]);
          aset(w, d + j, v[DMAX + j]);
        }
      nv = 2 * d;
      block_reduce_store<2 * DMAX>(w, nv, s_red, part_row);
    } else if (pass == kPassInitDens) {
      const int slot = s_slot[s_c];
      for (long i = i0; i < n; i += stride) {
        double xv[DMAX], r[DMAX];
#pragma unroll
        for (int j = 0; j < DMAX; j++) xv[j] = (j < d) ? __ldcg(a.xT + (size_t)j * np + i) : 0.0;
        const double lpd = fma(-0.5, solve_lower<DMAX>(s_rec, d, xv, r), s_rec[3]);
        __stcg(a.E + (size_t)slot * np + i, exp(lpd));
      }
      nv = 0;
    } else if (pass == kPassScatter) {
      double acc[TRI];
#pragma unroll
      for (int q = 0; q < TRI; q++) acc[q] = 0.0;
      const double *mu = s_rec + AMX_REC_HEAD;
      for (long i = i0; i < n; i += stride) {
        const double w = __ldcg(a.wnxt + i);
        double dx[DMAX];
#pragma unroll
        for (int j = 0; j < DMAX; j++) dx[j] = (j < d) ? __ldcg(a.xT + (size_t)j * np + i) - mu[j] : 0.0;
#pragma unroll
        for (int j = 0; j < DMAX; j++) {
          const double wd = w * dx[j];
#pragma unroll
          for (int k = 0; k <= j; k++) acc[AMX_TRI(j, k)] = fma(wd, dx[k], acc[AMX_TRI(j, k)]);
        }
      }
      nv = d * (d + 1) / 2;  // rows j<d of the packed triangle are exactly the first tri(d) entries
      block_reduce_store<TRI>(acc, nv, s_red, part_row);
    } else if (pass == kPassDensRefresh || pass == kPassRefresh) {
      const int L = s_L, nx = s_next, cc = s_c;
      const bool dens = (pass == kPassDensRefresh);
      const int cslot = dens ? s_slot[cc] : 0;
      double v[NVC];
#pragma unroll
      for (int q = 0; q < NVC; q++) v[q] = 0.0;
      for (long i = i0; i < n; i += stride) {
        double xv[DMAX];
#pragma unroll
        for (int j = 0; j < DMAX; j++) xv[j] = (j < d) ? __ldcg(a.xT + (size_t)j * np + i) : 0.0;
        double enew = 0.0;
        if (dens) {
          double r[DMAX];
          enew = exp(fma(-0.5, solve_lower<DMAX>(s_rec, d, xv, r), s_rec[3]));
          __stcg(a.E + (size_t)cslot * np + i, enew);
        }
        double t[kEmLmax];
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < kEmLmax; l++) {
          if (l < L) {
            const double e = (dens && l == cc) ? enew : __ldcg(a.E + (size_t)s_slot[l] * np + i);
            t[l] = s_lam[l] * e;
            s += t[l];
          }
        }
        double wn;
        if (s > 0) {  // the reference's guard (:855-866)
          const double inv = 1.0 / s;
          v[kEmLmax] += log(s);
          wn = 0.0;
#pragma unroll
          for (int l = 0; l < kEmLmax; l++)
            if (l < L) {
              const double w = t[l] * inv;
              v[l] += w;
              wn = (l == nx) ? w : wn;
            }
        } else {
          const double w = 1.0 / L;
          v[kEmLmax + 1] += 1.0;
#pragma unroll
          for (int l = 0; l < kEmLmax; l++)
            if (l < L) v[l] += w;
          wn = w;
        }
        __stcg(a.wnxt + i, wn);
#pragma unroll
        for (int j = 0; j < DMAX; j++) v[kEmLmax + 2 + j] = fma(wn, xv[j], v[kEmLmax + 2 + j]);
      }
      nv = kEmLmax + 2 + d;
      block_reduce_store<NVC>(v, nv, s_red, part_row);
    }

    // ------------------------------------------------------------------ barrier + leader
    if (barrier_arrive(ctrl, epoch)) {
      if (nv > 0) leader_reduce(a.part, nv, s_tot);
      if (threadIdx.x == 0) em_leader(ctrl, a, pass, s_tot);
      barrier_release(ctrl, epoch);
      __syncthreads();
    } else {
      barrier_wait(ctrl, epoch);
    }
    // ------------------------------------------------------------------ reload the control state
    if (threadIdx.x == 0) {
      s_pass = ld_cg(&ctrl->pass);
      s_L = ld_cg(&ctrl->L);
      s_c = ld_cg(&ctrl->c);
      s_next = ld_cg(&ctrl->next);
    }
    if (threadIdx.x < kEmLmax) {
      s_lam[threadIdx.x] = ld_cg(&ctrl->lam[threadIdx.x]);
      s_slot[threadIdx.x] = ld_cg(&ctrl->slot[threadIdx.x]);
    }
    const int reclen = AMX_REC_HEAD + 2 * d + d * (d + 1) / 2;
    for (int q = threadIdx.x; q < reclen; q += blockDim.x) s_rec[q] = ld_cg(&ctrl->rec[q]);
    __syncthreads();
    pass = s_pass;
    if (pass == kPassStop) break;
  }

  // optional dump of the responsibilities of the working state (step-parity tests)
  if (a.w_out != nullptr) {
    const int L = s_L;
    for (long i = i0; i < n; i += stride) {
      double s = 0.0;
      for (int l = 0; l < L; l++) s += s_lam[l] * __ldcg(a.E + (size_t)s_slot[l] * np + i);
      for (int l = 0; l < L; l++) {
        const double e = __ldcg(a.E + (size_t)s_slot[l] * np + i);
        a.w_out[(size_t)i * a.Lmax + l] = (s > 0) ? s_lam[l] * e / s : 1.0 / L;
      }
    }
  }
}

// single-Gaussian fit: mean, unbiased covariance, Cholesky (:1008-1033).  Two-pass, one CTA per
// block of samples, partials reduced on the host in fixed order (n is 1000 d in the product).
template <int DMAX>
__global__ void __launch_bounds__(kEmThreads) autorj_moment_kernel(int d, long n, const double *x, const double *mu,
                                                                   double *part) {
  constexpr int TRI = DMAX * (DMAX + 1) / 2;
  constexpr int NVAL = TRI > DMAX ? TRI : DMAX;
  __shared__ double s_red[kEmWarps * NVAL];
  double acc[NVAL];
#pragma unroll
  for (int q = 0; q < NVAL; q++) acc[q] = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double dx[DMAX];
#pragma unroll
    for (int j = 0; j < DMAX; j++) dx[j] = (j < d) ? x[i * d + j] - (mu ? mu[j] : 0.0) : 0.0;
    if (mu == nullptr) {
#pragma unroll
      for (int j = 0; j < DMAX; j++) acc[j] += dx[j];
    } else {
#pragma unroll
      for (int j = 0; j < DMAX; j++)
#pragma unroll
        for (int k = 0; k <= j; k++) acc[AMX_TRI(j, k)] = fma(dx[j], dx[k], acc[AMX_TRI(j, k)]);
    }
  }
  const int nv = mu ? d * (d + 1) / 2 : d;
  block_reduce_store<NVAL>(acc, nv, s_red, part + (size_t)blockIdx.x * kEmNV);
}

}  // namespace amx

using namespace amx;

template <int DMAX>
static int em_launch(EmArgs &a, int sms, float *ms) {
  int per_sm = 0;
  AMX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em_fit_kernel<DMAX>, kEmThreads, 0));
  if (per_sm < 1) return fail(AMX_ECUDA, "EM kernel does not fit on an SM");
  long want = (a.n + kEmThreads - 1) / kEmThreads;
  long cap = (long)sms * per_sm;
  unsigned grid = (unsigned)(want < cap ? want : cap);
  if (grid < 1) grid = 1;
  AMX_CUDA(cudaMalloc(&a.part, sizeof(double) * (size_t)grid * kEmNV));
  void *args[] = {&a};
  cudaEvent_t e0, e1;
  AMX_CUDA(cudaEventCreate(&e0));
  AMX_CUDA(cudaEventCreate(&e1));
  AMX_CUDA(cudaEventRecord(e0, stream()));
  AMX_CUDA(cudaLaunchCooperativeKernel((void *)em_fit_kernel<DMAX>, dim3(grid), dim3(kEmThreads), args, 0, stream()));
  count_launch();
  AMX_CUDA(cudaEventRecord(e1, stream()));
  AMX_CUDA(cudaEventSynchronize(e1));
  AMX_CUDA(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return AMX_OK;
}

static int em_fit_impl(int d, long n, const double *x_dev, int Lmax, int maxit, const int *init_idx, double *wt,
                       double *mean, double *tri, int *trace_L, double *trace_loglik, double *trace_cost,
                       int *trace_ann, double *cur_wt, double *cur_mean, double *cur_tri, int *cur_L, double *cur_w,
                       amx_em_result *res) {
  if (d < 1 || d > kEmDmax || Lmax < 1 || Lmax > kEmLmax || n < 1 || maxit < 0 || !init_idx || !wt || !mean || !tri)
    return fail(AMX_EINVAL, "amx_em_fit: need 1<=d<=%d, 1<=Lmax<=%d, n>=1, maxit>=0 (got d=%d Lmax=%d n=%ld maxit=%d)",
                kEmDmax, kEmLmax, d, Lmax, n, maxit);
  if (n < Lmax) return fail(AMX_EINVAL, "amx_em_fit: fewer samples (%ld) than start components (%d)", n, Lmax);
  for (int l = 0; l < Lmax; l++) {
    if (init_idx[l] < 0 || init_idx[l] >= n) return fail(AMX_EINVAL, "amx_em_fit: init_idx[%d]=%d out of range", l, init_idx[l]);
    for (int m = 0; m < l; m++)
      if (init_idx[m] == init_idx[l]) return fail(AMX_EINVAL, "amx_em_fit: start rows must be distinct");
  }
  int dev = 0, sms = 0, coop = 0;
  AMX_CUDA(cudaGetDevice(&dev));
  AMX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  AMX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return fail(AMX_ECUDA, "device lacks cooperative launch");
  const int tlen = d * (d + 1) / 2, cap = maxit + 1;
  EmArgs a;
  memset(&a, 0, sizeof(a));
  a.d = d;
  a.Lmax = Lmax;
  a.maxit = maxit;
  a.n = n;
  a.npad = (n + 31) / 32 * 32;
  a.x = x_dev;
  int *idx_dev = nullptr;
  AMX_CUDA(cudaMalloc(&a.xT, sizeof(double) * (size_t)d * a.npad));
  AMX_CUDA(cudaMalloc(&a.E, sizeof(double) * (size_t)Lmax * a.npad));
  AMX_CUDA(cudaMalloc(&a.wnxt, sizeof(double) * (size_t)a.npad));
  AMX_CUDA(cudaMalloc(&a.ctrl, sizeof(EmCtrl)));
  AMX_CUDA(cudaMemsetAsync(a.ctrl, 0, sizeof(EmCtrl), stream()));
  AMX_CUDA(cudaMalloc(&idx_dev, sizeof(int) * Lmax));
  AMX_CUDA(cudaMemcpyAsync(idx_dev, init_idx, sizeof(int) * Lmax, cudaMemcpyHostToDevice, stream()));
  a.init_idx = idx_dev;
  AMX_CUDA(cudaMalloc(&a.trace_L, sizeof(int) * cap));
  AMX_CUDA(cudaMalloc(&a.trace_ann, sizeof(int) * cap));
  AMX_CUDA(cudaMalloc(&a.trace_loglik, sizeof(double) * cap));
  AMX_CUDA(cudaMalloc(&a.trace_cost, sizeof(double) * cap));
  if (cur_w) AMX_CUDA(cudaMalloc(&a.w_out, sizeof(double) * (size_t)n * Lmax));
  float ms = 0;
  int rc;
  if (d <= 4) rc = em_launch<4>(a, sms, &ms);
  else if (d <= 8) rc = em_launch<8>(a, sms, &ms);
  else if (d <= 12) rc = em_launch<12>(a, sms, &ms);
  else return fail(AMX_EINVAL, "amx_em_fit: d=%d > 12 is not built yet", d);
  if (rc) return rc;
  std::vector<char> hc(sizeof(EmCtrl));
  AMX_CUDA(cudaMemcpy(hc.data(), a.ctrl, sizeof(EmCtrl), cudaMemcpyDeviceToHost));
  const EmCtrl *c = reinterpret_cast<const EmCtrl *>(hc.data());
  for (int l = 0; l < c->best_L; l++) {
    wt[l] = c->best_lam[l];
    for (int j = 0; j < d; j++) mean[(size_t)l * d + j] = c->best_mu[l][j];
    for (int j = 0; j < tlen; j++) tri[(size_t)l * tlen + j] = c->best_B[l][j];
  }
  const int it = c->iters;
  if (trace_L) AMX_CUDA(cudaMemcpy(trace_L, a.trace_L, sizeof(int) * it, cudaMemcpyDeviceToHost));
  if (trace_ann) AMX_CUDA(cudaMemcpy(trace_ann, a.trace_ann, sizeof(int) * it, cudaMemcpyDeviceToHost));
  if (trace_loglik) AMX_CUDA(cudaMemcpy(trace_loglik, a.trace_loglik, sizeof(double) * it, cudaMemcpyDeviceToHost));
  if (trace_cost) AMX_CUDA(cudaMemcpy(trace_cost, a.trace_cost, sizeof(double) * it, cudaMemcpyDeviceToHost));
  if (cur_L) *cur_L = c->L;
  for (int l = 0; l < c->L; l++) {
    if (cur_wt) cur_wt[l] = c->lam[l];
    if (cur_mean)
      for (int j = 0; j < d; j++) cur_mean[(size_t)l * d + j] = c->mu[l][j];
    if (cur_tri)
      for (int j = 0; j < tlen; j++) cur_tri[(size_t)l * tlen + j] = c->B[l][j];
  }
  if (cur_w) AMX_CUDA(cudaMemcpy(cur_w, a.w_out, sizeof(double) * (size_t)n * Lmax, cudaMemcpyDeviceToHost));
  if (res) {
    res->L = c->best_L;
    res->iters = it;
    res->status = c->status;
    res->comp_steps = c->comp_steps;
    res->kernel_ms = ms;
    res->flops = c->flops;
    res->bytes = 8.0 * d * (double)n * (double)c->comp_steps;
  }
  const int status = c->status;
  cudaFree(a.xT); cudaFree(a.E); cudaFree(a.wnxt); cudaFree(a.ctrl); cudaFree(idx_dev); cudaFree(a.part);
  cudaFree(a.trace_L); cudaFree(a.trace_ann); cudaFree(a.trace_loglik); cudaFree(a.trace_cost); cudaFree(a.w_out);
  if (status) return fail(status, "EM fit: scatter matrix not positive definite");
  return AMX_OK;
}

extern "C" {

long amx_em_draw_init(long n, int Lmax, const double *uniforms, long nuniforms, int *init_idx) {
  long used = 0;
  int l = 0;
  while (l < Lmax) {  // :682-697
    if (used >= nuniforms) return -1;
    init_idx[l] = (int)floor((double)n * uniforms[used++]);
    bool dup = false;
    for (int m = 0; m < l; m++)
      if (init_idx[m] == init_idx[l]) dup = true;
    if (!dup) l++;
  }
  return used;
}

int amx_em_fit_dev(int d, long n, const double *x_dev, int Lmax, int maxit, const int *init_idx, double *wt,
                   double *mean, double *tri, int *trace_L, double *trace_loglik, double *trace_cost, int *trace_ann,
                   amx_em_result *res) {
  if (int rc = require_device()) return rc;
  return em_fit_impl(d, n, x_dev, Lmax, maxit, init_idx, wt, mean, tri, trace_L, trace_loglik, trace_cost, trace_ann,
                     nullptr, nullptr, nullptr, nullptr, nullptr, res);
}

int amx_em_fit(int d, long n, const double *x, int Lmax, int maxit, const int *init_idx, double *wt, double *mean,
               double *tri, int *trace_L, double *trace_loglik, double *trace_cost, int *trace_ann, double *cur_wt,
               double *cur_mean, double *cur_tri, int *cur_L, double *cur_w, amx_em_result *res) {
  if (int rc = require_device()) return rc;
  if (d < 1 || n < 1 || !x) return fail(AMX_EINVAL, "amx_em_fit: empty input");
  double *x_dev = nullptr;
  AMX_CUDA(cudaMalloc(&x_dev, sizeof(double) * (size_t)n * d));
  AMX_CUDA(cudaMemcpyAsync(x_dev, x, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, stream()));
  int rc = em_fit_impl(d, n, x_dev, Lmax, maxit, init_idx, wt, mean, tri, trace_L, trace_loglik, trace_cost, trace_ann,
                       cur_wt, cur_mean, cur_tri, cur_L, cur_w, res);
  cudaFree(x_dev);
  return rc;
}

int amx_autorj_fit(int d, long n, const double *x, double *wt, double *mean, double *tri) {
  if (int rc = require_device()) return rc;
  if (d < 1 || d > 12 || n < 2 || !x) return fail(AMX_EINVAL, "amx_autorj_fit: need 1<=d<=12, n>=2");
  double *x_dev = nullptr, *part = nullptr, *mu_dev = nullptr;
  const unsigned grid = (unsigned)((n + kEmThreads - 1) / kEmThreads < 592 ? (n + kEmThreads - 1) / kEmThreads : 592);
  AMX_CUDA(cudaMalloc(&x_dev, sizeof(double) * (size_t)n * d));
  AMX_CUDA(cudaMalloc(&part, sizeof(double) * (size_t)grid * kEmNV));
  AMX_CUDA(cudaMalloc(&mu_dev, sizeof(double) * d));
  AMX_CUDA(cudaMemcpyAsync(x_dev, x, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, stream()));
  std::vector<double> hp((size_t)grid * kEmNV);
  const int tlen = d * (d + 1) / 2;
  for (int phase = 0; phase < 2; phase++) {
    const double *mu_arg = phase ? mu_dev : nullptr;
    if (d <= 4) autorj_moment_kernel<4><<<grid, kEmThreads, 0, stream()>>>(d, n, x_dev, mu_arg, part);
    else autorj_moment_kernel<12><<<grid, kEmThreads, 0, stream()>>>(d, n, x_dev, mu_arg, part);
    count_launch();
    AMX_CUDA(cudaGetLastError());
    AMX_CUDA(cudaMemcpyAsync(hp.data(), part, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, stream()));
    AMX_CUDA(cudaStreamSynchronize(stream()));
    const int nv = phase ? tlen : d;
    for (int q = 0; q < nv; q++) {
      double t = 0.0;
      for (unsigned b = 0; b < grid; b++) t += hp[(size_t)b * kEmNV + q];
      if (phase == 0) mean[q] = t / (double)n;
      else tri[q] = t / (double)(n - 1);
    }
    if (phase == 0) AMX_CUDA(cudaMemcpyAsync(mu_dev, mean, sizeof(double) * d, cudaMemcpyHostToDevice, stream()));
  }
  wt[0] = 1.0;
  // Cholesky of a d x d matrix (:1682-1701): scalar work, done where the result is needed
  for (int c = 0; c < d; c++) {
    double s = tri[AMX_TRI(c, c)];
    for (int j = 0; j < c; j++) s -= tri[AMX_TRI(c, j)] * tri[AMX_TRI(c, j)];
    if (!(s > 0.0)) {
      cudaFree(x_dev); cudaFree(part); cudaFree(mu_dev);
      return fail(AMX_ENUMERIC, "amx_autorj_fit: covariance not positive definite");
    }
    tri[AMX_TRI(c, c)] = sqrt(s);
    for (int r = c + 1; r < d; r++) {
      double t = tri[AMX_TRI(r, c)];
      for (int j = 0; j < c; j++) t -= tri[AMX_TRI(r, j)] * tri[AMX_TRI(c, j)];
      tri[AMX_TRI(r, c)] = t / tri[AMX_TRI(c, c)];
    }
  }
  cudaFree(x_dev);
  cudaFree(part);
  cudaFree(mu_dev);
  return AMX_OK;
}

}  // extern "C"
