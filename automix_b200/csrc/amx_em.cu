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

#include <mutex>
#include <vector>

#include "amx_internal.cuh"
#include "amx_targets.cuh"

namespace cg = cooperative_groups;

namespace amx {

constexpr int kEmThreads = 128;
constexpr int kEmWarps = kEmThreads / 32;
constexpr int kEmLmax = AMX_MAX_COMPS;
constexpr int kEmDmax = AMX_MAX_DIM;
constexpr int kEmTriMax = kEmDmax * (kEmDmax + 1) / 2;
constexpr int kEmRecMax = AMX_REC_HEAD + 2 * kEmDmax + kEmTriMax;
constexpr int kEmNV = kEmTriMax;  // values per CTA partial row (>= Lmax + Dmax + 2)

// kPassRefresh0 is the first E-step: the reference forms those responsibilities WITHOUT its `sum > 0` guard (:737-745)
enum EmPass { kPassInitStats = 0, kPassInitDens, kPassScatter, kPassDensRefresh, kPassRefresh, kPassRefresh0, kPassStop };

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
  long long dbg[8];  // cycle counters (AMX_EM_DEBUG): data pass, arrive->leader, reduce, leader logic, wait, reload
  int best_L;
  double best_lam[kEmLmax];
  double best_mu[kEmLmax][kEmDmax];
  double best_B[kEmLmax][kEmTriMax];
};

constexpr int kEmMaxDev = 8;

// What every CTA needs to know after a barrier; the leader pushes a copy to every GPU (peer stores), so the
// reload that follows the barrier reads local memory only.
struct EmPublic {
  int pass, L, c, next;
  int slot[kEmLmax];
  double lam[kEmLmax];
  double rec[kEmRecMax];
  double pivot[kEmDmax];  // current mean of component `next`: the shift of the fused second moments
};

// The per-GPU pieces of a sample-sharded fit.  All pointers are valid on every participating GPU (peer
// access over NVLink); only the leader touches another GPU's memory.
struct EmDev {
  double *part;      // [NV][grid] value-major per-CTA partials of the pass in flight
  double *devrow;    // [NV] this GPU's partials reduced by its last-arriving CTA
  unsigned *flags;   // [grid * 8] per-CTA release flags
  unsigned *arrive;  // arrival counter of this GPU's CTAs
  EmPublic *pub;     // local mirror of the public control state
  int grid;
};

struct EmArgs {
  int ndev, rank;        // GPUs sharing the fit, and which one this launch runs on
  long n_total;          // samples over all GPUs (n below is this GPU's shard)
  const double *init_rows;  // [Lmax][d] the data rows that start the components (on GPU 0)
  EmDev dev[kEmMaxDev];
  int d, Lmax, maxit;
  int fused;    // 1: the refresh pass also accumulates the (pivot-shifted) second moments of the next component,
                // so a component step is ONE pass and ONE barrier; 0: separate centred scatter pass
  int use_tma;  // 1: tile rows by cp.async.bulk + mbarrier; 0: coalesced per-thread loads into the tile
  int nbuf;  // tile buffers per CTA: 2 = fetch of tile k+1 overlaps tile k, 1 = more resident CTAs
  int smem_doubles;  // dynamic shared memory of the launch, in doubles
  long n, npad;
  const double *x;  // n x d row-major (device)
  double *xT, *E, *wnxt;
  EmCtrl *ctrl;  // the sequential state of the algorithm (on GPU 0)
  const int *init_idx;  // device
  int *trace_L, *trace_ann;
  double *trace_loglik, *trace_cost;
  double *w_out;  // optional n x Lmax responsibilities at exit
};

template <typename T>
__device__ __forceinline__ T ld_cg(const T *p) {
  return __ldcg(p);
}

// ---- barrier with a leader, across the CTAs of one GPU and across GPUs ---------------------------------
// Every CTA arrives at its GPU's counter.  The last CTA of a GPU reduces that GPU's partials into its
// `devrow` and arrives at the global counter (a system-scope atomic on GPU 0's memory, over NVLink when
// remote).  The last GPU's CTA is the leader: it sums the per-GPU rows through peer loads in GPU order,
// runs the sequential section, pushes the public state and the release flags to every GPU with peer stores.
// Counters only grow, so there are no reset races; waiters poll a flag in their own GPU's memory.
__device__ __forceinline__ uint32_t smem_u32(const void *p);
__device__ __forceinline__ void leader_reduce(const double *part, int nv, double *s_tot, double *s_chunk, int chunk_cap);

__device__ __forceinline__ int barrier_arrive(const EmArgs &a, unsigned &epoch, int nv, double *s_tot, double *s_chunk, int chunk_cap) {
  __shared__ int s_last;
  const EmDev &me = a.dev[a.rank];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(me.arrive, 1u);
    s_last = (t == (unsigned)me.grid * (epoch + 1u) - 1u) ? 1 : 0;
    if (s_last) __threadfence();
  }
  __syncthreads();
  if (!s_last) return 0;
  if (nv > 0) leader_reduce(me.part, nv, s_tot, s_chunk, chunk_cap);  // this GPU's CTAs, coalesced, fixed order
  if (a.ndev == 1) return 2;
  // The leader is always GPU 0's last CTA: the control block lives in GPU 0's memory, so the sequential
  // section runs on local memory; the other GPUs post their reduced row and a remote arrival and go to wait.
  if (a.rank != 0) {
    for (int q = threadIdx.x; q < nv; q += blockDim.x) __stcg(me.devrow + q, s_tot[q]);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd_system(&a.ctrl->arrive, 1u);
    return 1;
  }
  if (threadIdx.x == 0) {  // wait (on local memory) until every other GPU has arrived
    const unsigned want = (unsigned)(a.ndev - 1) * (epoch + 1u);
    const long long t0 = clock64();
    unsigned ns = 32;
    s_last = 1;
    while (ld_cg(&a.ctrl->arrive) < want) {
      __nanosleep(ns);
      if (ns < 256) ns *= 2;
      if (clock64() - t0 > 40000000000LL) {
        s_last = 0;  // a GPU never arrived: let the watchdog of the waiters end the kernel
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (!s_last) return 1;
  for (int q = threadIdx.x; q < nv; q += blockDim.x) {  // GPU 0's own sum first, then the others' rows in GPU order
    double t = s_tot[q];
    for (int g = 1; g < a.ndev; g++) t += ld_cg(a.dev[g].devrow + q);
    s_tot[q] = t;
  }
  __syncthreads();
  return 2;
}

// Leader: publish the state to every GPU, then release every CTA.
__device__ __forceinline__ void barrier_release(const EmArgs &a, int d, unsigned &epoch) {
  const EmCtrl *c = a.ctrl;
  if (a.ndev > 1) __threadfence_system();
  else __threadfence();
  __syncthreads();
  const int reclen = AMX_REC_HEAD + 2 * d + d * (d + 1) / 2;
  for (int g = 0; g < a.ndev; g++) {
    EmPublic *p = a.dev[g].pub;
    if (threadIdx.x == 0) {
      __stcg(&p->pass, ld_cg(&c->pass));
      __stcg(&p->L, ld_cg(&c->L));
      __stcg(&p->c, ld_cg(&c->c));
      __stcg(&p->next, ld_cg(&c->next));
    }
    if (threadIdx.x < kEmLmax) {
      __stcg(&p->slot[threadIdx.x], ld_cg(&c->slot[threadIdx.x]));
      __stcg(&p->lam[threadIdx.x], ld_cg(&c->lam[threadIdx.x]));
    }
    for (int q = threadIdx.x; q < reclen; q += blockDim.x) __stcg(&p->rec[q], ld_cg(&c->rec[q]));
    if (threadIdx.x < d) __stcg(&p->pivot[threadIdx.x], ld_cg(&c->mu[ld_cg(&c->next)][threadIdx.x]));
  }
  if (a.ndev > 1) __threadfence_system();
  else __threadfence();
  __syncthreads();
  for (int g = 0; g < a.ndev; g++)
    for (unsigned b = threadIdx.x; b < (unsigned)a.dev[g].grid; b += blockDim.x) __stcg(a.dev[g].flags + 8u * b, epoch + 1u);
  epoch++;
}

// Waiters poll their own flag (32-byte stride: no hot L2 line).  A watchdog turns a lost partner into an
// error instead of a hang.
__device__ __forceinline__ bool barrier_wait(const EmArgs &a, unsigned &epoch) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const unsigned *f = a.dev[a.rank].flags + 8u * blockIdx.x;
    unsigned ns = 64;
    const long long t0 = clock64();
    int ok = 1;
    while (ld_cg(f) <= epoch) {
      __nanosleep(ns);
      if (ns < 512) ns *= 2;
      if (clock64() - t0 > 40000000000LL) {  // ~20 s at 2 GHz
        ok = 0;
        break;
      }
    }
    if (a.ndev > 1) __threadfence_system();
    else __threadfence();
    s_ok = ok;
  }
  epoch++;
  __syncthreads();
  return s_ok != 0;
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

// leader: sum the per-CTA partials into s_tot[0..nv).  Partials are stored value-major, part[q*grid + b], so
// that the 32 lanes of a warp read 32 CONSECUTIVE CTAs of one value: one coalesced 256-byte request instead
// of 32 scattered sectors (the scattered form kept one SM's L1TEX busy for ~60k cycles per pass at 592 CTAs).
// Lane sums run over b = lane, lane+32, ... and are combined by a fixed shuffle tree, so the order of
// additions depends only on the launch geometry (bitwise reproducible).
// One batch: NVB values x KB loads per lane, at most 20 doubles -- what the register budget of this kernel lets
// ptxas keep in flight as ONE group.  (With 40 it issued 17 loads and then trickled the rest between dependent
// adds: eight L2 round trips per batch.  The section is a chain of such round trips, ~150-200 ns each.)
template <int NVB, int KB>
__device__ __forceinline__ void leader_reduce_batch(const double *part, int G, int q0, int nv, double *s_tot) {
  const int lane = threadIdx.x & 31;
  double v[NVB][KB];
#pragma unroll
  for (int a = 0; a < NVB; a++) {
    const int q = q0 + a;
    const double *src = part + (size_t)(q < nv ? q : q0) * G;
#pragma unroll
    for (int k = 0; k < KB; k++) {
      const int b = lane + 32 * k;
      v[a][k] = (q < nv && b < G) ? ld_cg(src + b) : 0.0;
    }
  }
#pragma unroll
  for (int a = 0; a < NVB; a++) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < KB; k++) t += v[a][k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && q0 + a < nv) s_tot[q0 + a] = t;
  }
}
template <int NVB, int KB>
__device__ __forceinline__ void leader_reduce_all(const double *part, int G, int nv, double *s_tot) {
  const int warp = threadIdx.x >> 5;
  for (int q = NVB * warp; q < nv; q += NVB * kEmWarps) leader_reduce_batch<NVB, KB>(part, G, q, nv, s_tot);
}
__device__ __forceinline__ void leader_reduce(const double *part, int nv, double *s_tot, double *s_chunk, int chunk_cap) {
  (void)s_chunk;
  (void)chunk_cap;
  const int G = (int)gridDim.x, kb = (G + 31) / 32;  // loads per value and lane
  if (kb <= 2) leader_reduce_all<8, 2>(part, G, nv, s_tot);
  else if (kb <= 4) leader_reduce_all<4, 4>(part, G, nv, s_tot);
  else if (kb <= 5) leader_reduce_all<3, 5>(part, G, nv, s_tot);
  else if (kb <= 8) leader_reduce_all<2, 8>(part, G, nv, s_tot);
  else if (kb <= 10) leader_reduce_all<2, 10>(part, G, nv, s_tot);
  else if (kb <= 20) {
    // full grids (592 CTAs): two values x 20 loads per round measured best (16 us per barrier; one value x 20 loads:
    // 17 us; staging whole values through shared memory with cp.async: 22 us)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 2 * warp; q < nv; q += 2 * kEmWarps) {
      const double *s0 = part + (size_t)q * G;
      const bool two = (q + 1 < nv);
      const double *s1 = part + (size_t)(two ? q + 1 : q) * G;
      double t0 = 0.0, t1 = 0.0, v0[20], v1[20];
#pragma unroll
      for (int k = 0; k < 20; k++) {
        const int b = lane + 32 * k;
        v0[k] = (b < G) ? ld_cg(s0 + b) : 0.0;
        v1[k] = (b < G) ? ld_cg(s1 + b) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 20; k++) {
        t0 += v0[k];
        t1 += v1[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      }
      if (lane == 0) {
        s_tot[q] = t0;
        if (two) s_tot[q + 1] = t1;
      }
    }
  }
  else {  // more CTAs than any current device holds: same order of additions, loads one by one
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = warp; q < nv; q += kEmWarps) {
      double t = 0.0;
      for (int b = lane; b < G; b += 32) t += ld_cg(part + (size_t)q * G + b);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) s_tot[q] = t;
    }
  }
  __syncthreads();
}

// ---- the leader's section: the sequential CEM^2 logic, run by ALL threads of the last-arriving CTA ------
// The hot part of the state is copied into shared memory; thread 0 takes the scalar decisions (sums in
// the reference's index order), and everything with independent elements -- divisions, logarithms,
// array shifts, copies, the rows of the Cholesky factor -- is spread over the CTA's threads.  Element-wise
// arithmetic and its order are exactly those of the reference, so results do not depend on this split.
enum EmAct { kActDone = 0, kActPlan, kActEndSweep, kActFinishIter };

template <int DMAX>
struct LeaderS {
  int pass, L, c, next, forced_pending, natural, forced, iters, stop, status, best_L;
  int act, keep, drop, savebest;
  long comp_steps;
  double flops, loglik, cost, cost_prev, cost_best, scal;
  int slot[kEmLmax];
  double lam[kEmLmax], colsum[kEmLmax], S1[kEmLmax], tmp[kEmLmax];
  double Bc[DMAX * (DMAX + 1) / 2];
  double S2[DMAX * (DMAX + 1) / 2];  // fused mode: shifted second moments of the component about to be updated
  double dl[DMAX];                   // fused mode: mean shift S1 / S0
  int chol_ok;
};

// in-place Cholesky of a packed d x d matrix in shared memory by one warp: lane r keeps row r in
// registers, column pivots travel by shuffle (element arithmetic as automix.c:1686-1700)
template <int DMAX>
__device__ __forceinline__ bool warp_chol(double *Bs, int d) {
  const int r = threadIdx.x & 31;
  double row[DMAX];
#pragma unroll
  for (int j = 0; j < DMAX; j++) row[j] = (r < d && j <= r) ? Bs[AMX_TRI(r, j)] : 0.0;
  bool ok = true;
#pragma unroll
  for (int c = 0; c < DMAX; c++) {
    if (c < d) {
      double sp = row[c];
#pragma unroll
      for (int j = 0; j < c; j++) sp = fma(-row[j], row[j], sp);
      const double spc = __shfl_sync(0xffffffffu, sp, c);
      if (!(spc > 0.0)) ok = false;
      const double p = sqrt(spc);
      double t = row[c];
#pragma unroll
      for (int j = 0; j < c; j++) {
        const double acj = __shfl_sync(0xffffffffu, row[j], c);
        t = fma(-row[j], acj, t);
      }
      if (r > c) row[c] = t / p;
      else if (r == c) row[c] = p;
    }
  }
#pragma unroll
  for (int j = 0; j < DMAX; j++)
    if (r < d && j <= r) Bs[AMX_TRI(r, j)] = row[j];
  return ok;
}

// family record of component l (proposal flavour, include/amx_layout.h) from the factor in S.Bc
template <int DMAX>
__device__ __forceinline__ void leader_make_rec(EmCtrl *c, LeaderS<DMAX> &S, int d, int l) {
  const int t = threadIdx.x, tri = d * (d + 1) / 2;
  if (t == 0) {
    double prod = 1.0;
    for (int i = 0; i < d; i++) prod *= S.Bc[AMX_TRI(i, i)];
    const double ld = log(prod);
    c->rec[0] = S.lam[l];
    c->rec[1] = 0.0;
    c->rec[2] = ld;
    c->rec[3] = -(d / 2.0) * log(2.0 * 3.14159265358979323846) - ld;
  }
  if (t < d) {
    c->rec[AMX_REC_HEAD + t] = ld_cg(&c->mu[l][t]);
    c->rec[AMX_REC_HEAD + d + t] = 1.0 / S.Bc[AMX_TRI(t, t)];
  }
  for (int q = t; q < tri; q += blockDim.x) c->rec[AMX_REC_HEAD + 2 * d + q] = S.Bc[q];
}

// remove component `gone` (:823-836, :908-921): thread q moves element q of every later component
template <int DMAX>
__device__ __forceinline__ void leader_drop(EmCtrl *c, LeaderS<DMAX> &S, int d, int gone) {
  const int t = threadIdx.x, tri = d * (d + 1) / 2, L = S.L;
  for (int q = t; q < tri; q += blockDim.x)
    for (int l = gone; l < L - 1; l++) c->B[l][q] = ld_cg(&c->B[l + 1][q]);
  if (t < d)
    for (int l = gone; l < L - 1; l++) c->mu[l][t] = ld_cg(&c->mu[l + 1][t]);
  __syncthreads();
  if (t == 0) {
    for (int l = gone; l < L - 1; l++) {
      S.lam[l] = S.lam[l + 1];
      S.slot[l] = S.slot[l + 1];
    }
    S.L = L - 1;
  }
  __syncthreads();
}

// lam /= sum(lam): the sum in index order by thread 0, the divisions in parallel
template <int DMAX>
__device__ __forceinline__ void leader_renorm(LeaderS<DMAX> &S) {
  const int t = threadIdx.x;
  if (t == 0) {
    double sum = 0.0;
    for (int l = 0; l < S.L; l++) sum += S.lam[l];
    S.scal = sum;
  }
  __syncthreads();
  if (t < S.L) S.lam[t] /= S.scal;
  __syncthreads();
}

// MML cost (:870-876): logarithms in parallel, their sum in index order
template <int DMAX>
__device__ __forceinline__ void leader_cost(LeaderS<DMAX> &S, long n, int nparams) {
  const int t = threadIdx.x;
  if (t < S.L) S.tmp[t] = log((double)n * S.lam[t] / 12.0);
  __syncthreads();
  if (t == 0) {
    double sum = 0.0;
    for (int l = 0; l < S.L; l++) sum += S.tmp[l];
    S.cost = (nparams / 2.0) * sum + (S.L / 2.0) * log((double)n / 12.0) + S.L * (nparams + 1) / 2.0 - S.loglik;
  }
  __syncthreads();
}

template <int DMAX>
__device__ void em_leader_block(EmCtrl *c, const EmArgs &a, int pass, const double *s_tot, LeaderS<DMAX> &S) {
  const int t = threadIdx.x, d = a.d, tri = d * (d + 1) / 2, nparams = d + tri;
  // ---- load the hot state
  if (t == 0) {
    S.pass = ld_cg(&c->pass); S.L = ld_cg(&c->L); S.c = ld_cg(&c->c); S.next = ld_cg(&c->next);
    S.forced_pending = ld_cg(&c->forced_pending); S.natural = ld_cg(&c->natural); S.forced = ld_cg(&c->forced);
    S.iters = ld_cg(&c->iters); S.stop = ld_cg(&c->stop); S.status = ld_cg(&c->status); S.best_L = ld_cg(&c->best_L);
    S.comp_steps = ld_cg(&c->comp_steps); S.flops = ld_cg(&c->flops); S.loglik = ld_cg(&c->loglik);
    S.cost = ld_cg(&c->cost); S.cost_prev = ld_cg(&c->cost_prev); S.cost_best = ld_cg(&c->cost_best);
    S.act = kActDone; S.keep = 0; S.drop = -1; S.savebest = 0; S.chol_ok = 1;
  }
  if (t < kEmLmax) {
    S.slot[t] = ld_cg(&c->slot[t]);
    S.lam[t] = ld_cg(&c->lam[t]);
    S.colsum[t] = ld_cg(&c->colsum[t]);
    S.S1[t] = ld_cg(&c->S1[t]);
  }
  __syncthreads();

  if (pass == kPassInitStats) {
    // :700-723 common isotropic start; s_tot = [sum x_j (d) | sum x_j^2 (d)]
    if (t == 0) {
      double s2 = 0.0;
      const double len = (double)a.n_total;
      for (int j = 0; j < d; j++) s2 += (s_tot[d + j] - s_tot[j] * s_tot[j] / len) / len;
      s2 /= (10.0 * d);
      c->s2 = s2;
      S.scal = sqrt(s2);  // chol of s2 * I (:716-721): sqrt on the diagonal, zeros below
      if (!(s2 > 0.0)) S.status = AMX_ENUMERIC;
      S.L = a.Lmax;
      S.c = 0;
      S.pass = kPassInitDens;
    }
    __syncthreads();
    for (int q = t; q < a.Lmax * d; q += blockDim.x) {
      const int l = q / d, j = q % d;
      c->mu[l][j] = ld_cg(a.init_rows + (size_t)l * d + j);
    }
    for (int q = t; q < a.Lmax * tri; q += blockDim.x) c->B[q / tri][q % tri] = 0.0;
    __syncthreads();
    for (int q = t; q < a.Lmax * d; q += blockDim.x) c->B[q / d][AMX_TRI(q % d, q % d)] = S.scal;
    if (t < a.Lmax) {
      S.slot[t] = t;
      S.lam[t] = 1.0 / a.Lmax;
    }
    for (int q = t; q < tri; q += blockDim.x) S.Bc[q] = 0.0;
    __syncthreads();
    if (t < d) S.Bc[AMX_TRI(t, t)] = S.scal;
    __syncthreads();
    leader_make_rec<DMAX>(c, S, d, 0);
  } else if (pass == kPassInitDens) {  // one component's start densities are in place
    if (t == 0) {
      S.c++;
      if (S.c < S.L) {
        S.pass = kPassInitDens;
      } else {
        S.next = 0;
        S.iters = 0;  // the refresh that follows is the initial E-step (:733-744)
        S.pass = kPassRefresh0;
      }
    }
    for (int q = t; q < tri; q += blockDim.x) S.Bc[q] = ld_cg(&c->B[0][q]);  // all start factors are equal
    __syncthreads();
    if (S.c < S.L) leader_make_rec<DMAX>(c, S, d, S.c);
  } else if (pass == kPassScatter) {
    // :803-813 centred scatter / sum of weights, then Cholesky
    const int cc = S.c;
    for (int q = t; q < tri; q += blockDim.x) S.Bc[q] = s_tot[q] / S.colsum[cc];
    __syncthreads();
    if (t < 32) {
      const bool ok = warp_chol<DMAX>(S.Bc, d);
      if (t == 0) S.chol_ok = ok ? 1 : 0;
    }
    __syncthreads();
    for (int q = t; q < tri; q += blockDim.x) c->B[cc][q] = S.Bc[q];
    leader_make_rec<DMAX>(c, S, d, cc);
    if (t == 0) {
      if (!S.chol_ok) {
        S.status = AMX_ENUMERIC;
        S.stop = 1;
        S.pass = kPassStop;
      } else {
        S.next = (cc + 1 < S.L) ? cc + 1 : 0;
        S.pass = kPassDensRefresh;
      }
    }
  } else {
    // a refresh finished: s_tot = [colsum (Lmax) | loglik | fallbacks | S1 (d)]
    if (t < S.L) S.colsum[t] = s_tot[t];
    if (t < d) S.S1[t] = s_tot[kEmLmax + 2 + t];
    if (a.fused)
      for (int q = t; q < tri; q += blockDim.x) S.S2[q] = s_tot[kEmLmax + 2 + d + q];
    if (t == 0) {
      S.loglik = s_tot[kEmLmax] - 500.0 * s_tot[kEmLmax + 1];
      if (S.iters == 0) {  // initial E-step done: start outer iteration 1
        S.iters = 1;
        S.natural = S.forced = 0;
        S.c = 0;
        S.act = kActPlan;
      } else if (S.forced_pending) {  // refresh after a forced annihilation (:931-958)
        S.act = kActFinishIter;
      } else {
        if (pass == kPassDensRefresh) S.c++;  // component kept: move on (:819)
        S.act = (S.c < S.L) ? kActPlan : kActEndSweep;
      }
    }
    __syncthreads();
    if (S.forced_pending) {  // uniform: the cost after the forced annihilation
      leader_cost<DMAX>(S, a.n_total, nparams);
      if (t == 0) S.forced_pending = 0;
      __syncthreads();
    }
    while (S.act != kActDone) {
      const int act = S.act;
      __syncthreads();
      if (act == kActPlan) {
        // start the update of component S.c from the column sums and the first moment (:773-801)
        const int cc = S.c;
        if (t == 0) {
          double tot = 0.0, wkeep = 0.0;
          for (int l = 0; l < S.L; l++) {
            const double wl = max_m(0.0, (S.colsum[l] - nparams / 2.0));
            if (l == cc) wkeep = wl;
            tot += wl;
          }
          S.lam[cc] = wkeep / tot;
          S.comp_steps++;
          S.flops += (double)a.n_total * (2.0 * d * d + 8.0 * d + 4.0 * S.L + 7.0);
        }
        __syncthreads();
        leader_renorm<DMAX>(S);
        if (S.lam[cc] > 0.005 && !a.fused) {  // uniform branch (shared value)
          if (t < d) {
            const double m = S.S1[t] / S.colsum[cc];
            c->mu[cc][t] = m;
            c->rec[AMX_REC_HEAD + t] = m;  // the scatter pass centres on the NEW mean (:803-809)
          }
          if (t == 0) {
            S.pass = kPassScatter;
            S.act = kActDone;
          }
        } else if (S.lam[cc] > 0.005) {
          // fused: S1, S2 are moments of (x - pivot), pivot = the component's mean before this update.
          //   mean = pivot + S1/S0,   cov = S2/S0 - (S1/S0)(S1/S0)^T   (:797-810 in shifted form)
          const double S0 = S.colsum[cc];
          if (t < d) {
            S.dl[t] = S.S1[t] / S0;
            c->mu[cc][t] = ld_cg(&c->mu[cc][t]) + S.dl[t];
          }
          __syncthreads();
          for (int q = t; q < tri; q += blockDim.x) {
            int j = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
            while ((j + 1) * (j + 2) / 2 <= q) j++;
            while (j * (j + 1) / 2 > q) j--;
            const int k = q - j * (j + 1) / 2;
            S.Bc[q] = (S.S2[q] - S.S1[j] * S.dl[k]) / S0;
          }
          __syncthreads();
          if (t < 32) {
            const bool ok = warp_chol<DMAX>(S.Bc, d);
            if (t == 0) S.chol_ok = ok ? 1 : 0;
          }
          __threadfence();
          __syncthreads();
          for (int q = t; q < tri; q += blockDim.x) c->B[cc][q] = S.Bc[q];
          leader_make_rec<DMAX>(c, S, d, cc);
          if (t == 0) {
            if (!S.chol_ok) {
              S.status = AMX_ENUMERIC;
              S.stop = 1;
              S.pass = kPassStop;
            } else {
              S.next = (cc + 1 < S.L) ? cc + 1 : 0;
              S.pass = kPassDensRefresh;
            }
            S.act = kActDone;
          }
        } else {  // natural annihilation (:821-845): refresh before the next component is looked at
          leader_drop<DMAX>(c, S, d, cc);
          leader_renorm<DMAX>(S);
          if (t == 0) {
            S.natural = 1;
            S.next = (cc < S.L) ? cc : 0;
            S.pass = kPassRefresh;
            S.act = kActDone;
          }
        }
      } else if (act == kActEndSweep) {
        leader_cost<DMAX>(S, a.n_total, nparams);
        if (t == 0) {
          if (S.iters == 1) S.cost_prev = S.cost;
          S.savebest = (S.iters == 1 || S.cost < S.cost_best) ? 1 : 0;  // :881-893
          if (S.savebest) {
            S.best_L = S.L;
            S.cost_best = S.cost;
          }
          S.drop = -1;
          if (fabs(S.cost_prev - S.cost) < min_m(1E-5 * fabs(S.cost_prev), 0.01) && S.iters > 1) {  // :894
            if (S.L == 1) {
              S.stop = 1;
            } else {
              S.forced = 2;
              double lo = S.lam[0];
              int gone = 0;
              for (int l = 1; l < S.L; l++)
                if (lo > S.lam[l]) {
                  lo = S.lam[l];
                  gone = l;
                }
              S.drop = gone;
            }
          }
        }
        __syncthreads();
        if (S.savebest) {
          if (t < S.L) c->best_lam[t] = S.lam[t];
          for (int q = t; q < S.L * d; q += blockDim.x) c->best_mu[q / d][q % d] = ld_cg(&c->mu[q / d][q % d]);
          for (int q = t; q < S.L * tri; q += blockDim.x) c->best_B[q / tri][q % tri] = ld_cg(&c->B[q / tri][q % tri]);
          __syncthreads();
        }
        if (S.drop >= 0) {
          leader_drop<DMAX>(c, S, d, S.drop);
          leader_renorm<DMAX>(S);
          if (t == 0) {
            S.forced_pending = 1;
            S.next = 0;
            S.pass = kPassRefresh;
            S.act = kActDone;
          }
        } else if (t == 0) {
          S.act = kActFinishIter;
        }
      } else {  // kActFinishIter (:961-970)
        if (t == 0) {
          if (S.iters > a.maxit) S.stop = 1;
          S.cost_prev = S.cost;
          const int it = S.iters - 1;
          if (a.trace_ann) a.trace_ann[it] = S.natural + S.forced;
          if (a.trace_cost) a.trace_cost[it] = S.cost;
          if (a.trace_loglik) a.trace_loglik[it] = S.loglik;
          if (a.trace_L) a.trace_L[it] = S.L;
          if (S.stop) {
            S.pass = kPassStop;
            S.act = kActDone;
          } else {
            S.iters++;
            S.natural = S.forced = 0;
            S.c = 0;
            S.act = kActPlan;
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- store the hot state
  if (t == 0) {
    c->pass = S.pass; c->L = S.L; c->c = S.c; c->next = S.next; c->forced_pending = S.forced_pending;
    c->natural = S.natural; c->forced = S.forced; c->iters = S.iters; c->stop = S.stop; c->status = S.status;
    c->best_L = S.best_L; c->comp_steps = S.comp_steps; c->flops = S.flops; c->loglik = S.loglik;
    c->cost = S.cost; c->cost_prev = S.cost_prev; c->cost_best = S.cost_best;
  }
  if (t < kEmLmax) {
    c->slot[t] = S.slot[t];
    c->lam[t] = S.lam[t];
    c->colsum[t] = S.colsum[t];
    c->S1[t] = S.S1[t];
  }
}

// ---- TMA (bulk asynchronous copy) + mbarrier -----------------------------------------------------------
// A tile row (128 consecutive samples of one SoA array = 1 KB) is fetched by one cp.async.bulk issued by
// one thread; completion is counted in bytes on an mbarrier in shared memory.  No registers are staged
// and the fetch overlaps the other resident CTAs' arithmetic.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kEmTS = kEmThreads + 4;  // tile row stride in doubles: spreads the column reduction over the banks

// accumulate rows [R0, R1) of the packed lower triangle of w dx dx^T into acc[0 .. N)
template <int DMAX, int R0, int R1, int NACC>
__device__ __forceinline__ void scatter_rows(double (&acc)[NACC], const double *xs, const double *mu, double w, int tt,
                                             int d) {
  double dx[DMAX];
#pragma unroll
  for (int j = 0; j < DMAX; j++) dx[j] = (j < d) ? xs[j * kEmTS + tt] - mu[j] : 0.0;
  int q = 0;
#pragma unroll
  for (int j = R0; j < R1; j++) {
    const double wd = w * dx[j];
#pragma unroll
    for (int k = 0; k <= j; k++, q++) acc[q] = fma(wd, dx[k], acc[q]);
  }
}

// Log-density of one tile column under component l, operation for operation the reference's lnormprob
// (automix.c:1727-1750: forward substitution with divisions, log of the product of the diagonal), from the
// component's parameters in the control block.  Used only for the rare samples that every component puts below
// exp(-667): there the cached exp(lpd) is (nearly) subnormal and lam * exp(lpd) no longer tracks the reference's
// exp(log(lam) + lpd).
__device__ __noinline__ double lnormprob_slow(const EmCtrl *c, int l, int d, const double *xcol) {
  double r[kEmDmax];
  double det = 1.0;
  for (int i = 0; i < d; i++) r[i] = xcol[i * (kEmThreads + 4)] - ld_cg(&c->mu[l][i]);
  for (int i = 0; i < d; i++) {
    for (int j = 0; j < i; j++) r[i] -= ld_cg(&c->B[l][AMX_TRI(i, j)]) * r[j];
    const double bii = ld_cg(&c->B[l][AMX_TRI(i, i)]);
    r[i] /= bii;
    det *= bii;
  }
  double q = 0.0;
  for (int i = 0; i < d; i++) q += r[i] * r[i];
  return -0.5 * q - (d / 2.0) * log(2.0 * 3.14159265358979323846) - log(det);
}

// ---- the fit kernel -------------------------------------------------------------------------------------
// 128 threads per CTA, one tile = 128 samples.  Shared-memory tile rows (stride kEmTS doubles):
//   xs[d]  sample coordinates     Es[Lmax]  density cache rows of the live components (component order)
//   ws[1]  responsibilities of the "next" component
template <int DMAX>
__global__ void __launch_bounds__(kEmThreads, 4) em_fit_kernel(EmArgs a) {
  constexpr int TRI = DMAX * (DMAX + 1) / 2;
  constexpr int NRED = TRI > 2 * DMAX ? TRI : 2 * DMAX;
  constexpr int RSPLIT = DMAX > 8 ? 8 : DMAX;  // DMAX > 8: two row groups, [0,8) and [8,DMAX)
  extern __shared__ __align__(16) double tile[];
  __shared__ __align__(8) uint64_t s_bar[4];  // one mbarrier per tile buffer / pipeline stage
  // The reduction scratch, the leader's totals and the leader's state are only live while the tile is idle
  // (end of a pass, barrier, leader section), so they alias the tile memory instead of adding ~7 KB of static
  // shared memory per CTA (which would cost a resident CTA per SM).
  constexpr int NTOT = kEmLmax + 2 + DMAX + TRI;
  double *s_red = tile;                         // [kEmWarps * NRED]
  double *s_tot = tile + kEmWarps * NRED;       // [NTOT]
  double *s_chunk = tile + kEmWarps * NRED + NTOT + (NTOT & 1);  // staging area of the partial reduce (idle tile memory)
  const int chunk_cap = a.smem_doubles - (kEmWarps * NRED + NTOT + (NTOT & 1));
  LeaderS<DMAX> &s_lead = *reinterpret_cast<LeaderS<DMAX> *>(tile + kEmWarps * NRED + NTOT + (NTOT & 1));
  __shared__ double s_pivot[DMAX];
  __shared__ double s_rec[AMX_REC_HEAD + 2 * DMAX + TRI];
  __shared__ double s_lam[kEmLmax];
  __shared__ int s_slot[kEmLmax];
  __shared__ int s_pass, s_L, s_c, s_next;

  EmCtrl *ctrl = a.ctrl;
  const int d = a.d;
  const long n = a.n, np = a.npad;
  const long ntiles = np / kEmThreads;
  const long stride = (long)gridDim.x * blockDim.x;
  const long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tile_doubles = (d + a.Lmax + 1 + (a.fused ? d : 0)) * kEmTS;  // rows: x | E | w_next | (w_next * dx)
  double *part_col = a.dev[a.rank].part + blockIdx.x;  // value q of this CTA lives at part_col[q * gridDim.x]
  const size_t pstride = gridDim.x;
  unsigned epoch = 0;
  uint32_t tma_phase[4] = {0u, 0u, 0u, 0u};
  int pass = kPassInitStats;
  if (t == 0) {
    for (int q = 0; q < 4; q++) mbar_init(&s_bar[q], 1);
  }
  __syncthreads();

  for (;;) {
    // ------------------------------------------------------------------ the data pass
    const long long tk0 = clock64();
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
      // warp -> CTA reduction, packed as [sum (d) | sumsq (d)]
#pragma unroll
      for (int j = 0; j < 2 * DMAX; j++) {
        const double r = warp_sum(v[j]);
        if (lane == 0) s_red[warp * NRED + j] = r;
      }
      __syncthreads();
      if (t < 2 * d) {
        const int src = t < d ? t : DMAX + (t - d);
        double tot = 0.0;
        for (int w = 0; w < kEmWarps; w++) tot += s_red[w * NRED + src];
        part_col[(size_t)t * pstride] = tot;
      }
      nv = 2 * d;
    } else if (pass == kPassInitDens) {
      const int slot = s_slot[s_c];
      for (long i = i0; i < n; i += stride) {
        double xv[DMAX], r[DMAX];
#pragma unroll
        for (int j = 0; j < DMAX; j++) xv[j] = (j < d) ? __ldcg(a.xT + (size_t)j * np + i) : 0.0;
        const double lpd = fma(-0.5, solve_lower<DMAX, (DMAX <= 12)>(s_rec, d, xv, r), s_rec[3]);
        __stcg(a.E + (size_t)slot * np + i, exp(lpd));
      }
      nv = 0;
    } else if (pass == kPassScatter) {
      // ---- S2 = sum_i wnxt_i (x_i - mu)(x_i - mu)^T, rows split over two thread groups when DMAX > 8
      constexpr bool kBigD = DMAX > 12;  // entry-parallel scatter from the shared tile (below)
      constexpr int NLO = RSPLIT * (RSPLIT + 1) / 2, NHI = TRI - NLO;
      constexpr int NACC = kBigD ? (TRI + kEmThreads - 1) / kEmThreads : (NLO > NHI ? NLO : NHI);
      double acc[NACC];  // one row group per thread: [0,RSPLIT) for threads 0..63, [RSPLIT,DMAX) for 64..127
#pragma unroll
      for (int q = 0; q < NACC; q++) acc[q] = 0.0;
      const double *mu = s_rec + AMX_REC_HEAD;
      // The scatter pass needs only d+1 rows per tile, so the tile memory holds NST of them: a NST-deep
      // TMA pipeline that keeps NST-1 fetches in flight per CTA and hides the HBM latency.
      const int stage_doubles = (d + 1) * kEmTS;
      int NST = (a.nbuf * tile_doubles) / stage_doubles;
      NST = NST > 4 ? 4 : NST;
      auto fetch = [&](long tl, int st) {  // warp 0: one bulk copy per tile row, lanes in parallel
        double *xs = tile + st * stage_doubles;
        fence_proxy_async();
        if (lane == 0) mbar_expect_tx(&s_bar[st], (uint32_t)((d + 1) * kEmThreads * 8));
        __syncwarp();
        for (int j = lane; j <= d; j += 32)
          tma_load_row(xs + j * kEmTS, (j < d ? a.xT + (size_t)j * np : a.wnxt) + tl * kEmThreads, kEmThreads * 8,
                       &s_bar[st]);
      };
      if (!a.use_tma && !kBigD) {
        // coalesced loads: thread t reads sample tl*128+t of every row (one 1 KB request per warp-row);
        // with DMAX > 8 the two row groups each read samples tt and tt+64
        for (long tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
          const long i = tl * kEmThreads;
          if constexpr (DMAX <= 8) {
            double dx[DMAX];
#pragma unroll
            for (int j = 0; j < DMAX; j++) dx[j] = (j < d) ? __ldcg(a.xT + (size_t)j * np + i + t) - mu[j] : 0.0;
            const double w = __ldcg(a.wnxt + i + t);
            int q = 0;
#pragma unroll
            for (int j = 0; j < DMAX; j++) {
              const double wd = w * dx[j];
#pragma unroll
              for (int k = 0; k <= j; k++, q++) acc[q] = fma(wd, dx[k], acc[q]);
            }
          } else {
            const int tt = t & 63;
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const long ii = i + tt + 64 * h;
              double dx[DMAX];
#pragma unroll
              for (int j = 0; j < DMAX; j++) dx[j] = (j < d) ? __ldcg(a.xT + (size_t)j * np + ii) - mu[j] : 0.0;
              const double w = __ldcg(a.wnxt + ii);
              int q = 0;
              if (t < 64) {
#pragma unroll
                for (int j = 0; j < RSPLIT; j++) {
                  const double wd = w * dx[j];
#pragma unroll
                  for (int k = 0; k <= j; k++, q++) acc[q] = fma(wd, dx[k], acc[q]);
                }
              } else {
#pragma unroll
                for (int j = RSPLIT; j < DMAX; j++) {
                  const double wd = w * dx[j];
#pragma unroll
                  for (int k = 0; k <= j; k++, q++) acc[q] = fma(wd, dx[k], acc[q]);
                }
              }
            }
          }
        }
      }
      const bool tma_sc = a.use_tma || kBigD;
      if (tma_sc && warp == 0)
        for (int q = 0; q < NST - 1; q++)
          if ((long)blockIdx.x + (long)q * gridDim.x < ntiles) fetch(blockIdx.x + (long)q * gridDim.x, q);
      int st = 0;
      for (long tl = blockIdx.x; tma_sc && tl < ntiles; tl += gridDim.x) {
        const long ahead = tl + (long)(NST - 1) * gridDim.x;
        if (warp == 0 && ahead < ntiles) fetch(ahead, (st + NST - 1) % NST);  // that stage was drained last iteration
        mbar_wait(&s_bar[st], tma_phase[st]);
        tma_phase[st] ^= 1u;
        const double *xs = tile + st * stage_doubles, *ws = xs + d * kEmTS;
        if constexpr (kBigD) {
          // d > 12: too many triangle entries for registers per sample.  Thread e owns entries e, e+128, ...
          // of the packed triangle and walks the 128 samples of the tile in shared memory.
          const int tri_d = d * (d + 1) / 2;
#pragma unroll
          for (int q = 0; q < NACC; q++) {
            const int e = t + q * kEmThreads;
            if (e < tri_d) {
              int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
              while ((j + 1) * (j + 2) / 2 <= e) j++;
              while (j * (j + 1) / 2 > e) j--;
              const int k = e - j * (j + 1) / 2;
              const double *xj = xs + j * kEmTS, *xk = xs + k * kEmTS;
              const double mj = mu[j], mk = mu[k];
              double sacc = 0.0;
#pragma unroll 4
              for (int i2 = 0; i2 < kEmThreads; i2++) sacc = fma(ws[i2] * (xj[i2] - mj), xk[i2] - mk, sacc);
              acc[q] += sacc;
            }
          }
        } else if constexpr (DMAX <= 8) {
          scatter_rows<DMAX, 0, DMAX>(acc, xs, mu, ws[t], t, d);  // padding samples carry w = 0
        } else {
          const int tt = t & 63;
          if (t < 64) {
            scatter_rows<DMAX, 0, RSPLIT>(acc, xs, mu, ws[tt], tt, d);
            scatter_rows<DMAX, 0, RSPLIT>(acc, xs, mu, ws[tt + 64], tt + 64, d);
          } else {
            scatter_rows<DMAX, RSPLIT, DMAX>(acc, xs, mu, ws[tt], tt, d);
            scatter_rows<DMAX, RSPLIT, DMAX>(acc, xs, mu, ws[tt + 64], tt + 64, d);
          }
        }
        __syncthreads();  // everyone is done with this stage before it is refilled
        st = (st + 1 == NST) ? 0 : st + 1;
      }
      nv = d * (d + 1) / 2;  // rows j<d of the packed triangle are its first tri(d) entries
      if constexpr (kBigD) {
#pragma unroll
        for (int q = 0; q < NACC; q++) {
          const int e = t + q * kEmThreads;
          if (e < nv) part_col[(size_t)e * pstride] = acc[q];
        }
      } else {
      // warps 0,1 hold rows [0,RSPLIT) at s_red[..][q]; warps 2,3 hold the rest at s_red[..][NLO + q]
#pragma unroll
      for (int q = 0; q < NACC; q++) {
        const double r = warp_sum(acc[q]);
        const int dst = (DMAX <= 8 || warp < 2) ? q : NLO + q;
        if (lane == 0 && dst < NRED) s_red[warp * NRED + dst] = r;
      }
      __syncthreads();
      for (int q = t; q < nv; q += blockDim.x) {
        double tot = 0.0;
        if (DMAX <= 8) {
          for (int w = 0; w < kEmWarps; w++) tot += s_red[w * NRED + q];
        } else if (q < NLO) {
          tot = s_red[0 * NRED + q] + s_red[1 * NRED + q];
        } else {
          tot = s_red[2 * NRED + q] + s_red[3 * NRED + q];
        }
        part_col[(size_t)q * pstride] = tot;
      }
      }
    } else if (pass == kPassDensRefresh || pass == kPassRefresh || pass == kPassRefresh0) {
      // ---- (new density column,) responsibilities, column sums, log-likelihood, next first moment
      // When every density underflows the fit annihilates all of its components, and the reference then keeps
      // "annihilating" into negative component counts until the iteration cap (:893-923 with Lkk == 0; the minimum
      // so far is what it returns).  The leader mirrors that, the data pass sees an empty mixture.
      const int L = s_L < 0 ? 0 : s_L, nx = s_next, cc = s_c;
      const bool dens = (pass == kPassDensRefresh);
      const bool unguarded = (pass == kPassRefresh0);
      const int cslot = dens ? s_slot[cc] : 0;
      const int col = t >> 2, prt = t & 3;  // reduction role: 32 columns x 4 partial sums
      double acc_col = 0.0, acc_s1 = 0.0, ll = 0.0, nfb = 0.0;
      // fused mode: thread (half, e) owns triangle entries e, e+64 over the samples of its half of the tile
      const int tri_d = d * (d + 1) / 2, half = t >> 6;
      double acc2[2] = {0.0, 0.0};
      int ej[2] = {0, 0}, ek[2] = {0, 0};
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int e = (t & 63) + 64 * q;
        int j = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
        while ((j + 1) * (j + 2) / 2 <= e) j++;
        while (j * (j + 1) / 2 > e) j--;
        ej[q] = j;
        ek[q] = e - j * (j + 1) / 2;
      }
      auto fetch = [&](long tl, int bf) {  // warp 0: one bulk copy per tile row, lanes in parallel
        double *xs = tile + bf * tile_doubles, *Es = xs + d * kEmTS;
        fence_proxy_async();
        const int rows = d + L - (dens ? 1 : 0);
        if (lane == 0) mbar_expect_tx(&s_bar[bf], (uint32_t)(rows * kEmThreads * 8));
        __syncwarp();
        for (int q = lane; q < d + L; q += 32) {
          if (q < d) {
            tma_load_row(xs + q * kEmTS, a.xT + (size_t)q * np + tl * kEmThreads, kEmThreads * 8, &s_bar[bf]);
          } else {
            const int l = q - d;
            if (!(dens && l == cc))
              tma_load_row(Es + l * kEmTS, a.E + (size_t)s_slot[l] * np + tl * kEmThreads, kEmThreads * 8, &s_bar[bf]);
          }
        }
      };
      int bf = 0;
      const bool dbl = a.nbuf == 2;
      if (a.use_tma && dbl && warp == 0 && (long)blockIdx.x < ntiles) fetch(blockIdx.x, 0);
      for (long tl = blockIdx.x; tl < ntiles; tl += gridDim.x, bf = dbl ? bf ^ 1 : 0) {
        double *xs = tile + bf * tile_doubles, *Es = xs + d * kEmTS, *ws = Es + a.Lmax * kEmTS;
        // -- per-sample phase: thread t owns sample tl*128 + t
        const long i = tl * kEmThreads + t;
        const bool valid = i < n;
        if (a.use_tma) {
          if (warp == 0) {
            if (!dbl) fetch(tl, 0);
            else if (tl + gridDim.x < ntiles) fetch(tl + gridDim.x, bf ^ 1);
          }
          mbar_wait(&s_bar[bf], tma_phase[bf]);
          tma_phase[bf] ^= 1u;
        } else {
          // coalesced loads: this thread's column of the tile, eight rows in flight at a time
#pragma unroll
          for (int j = 0; j < DMAX; j++)
            if (j < d) xs[j * kEmTS + t] = __ldcg(a.xT + (size_t)j * np + i);
          for (int l0 = 0; l0 < L; l0 += 8) {
            double e[8];
#pragma unroll
            for (int k = 0; k < 8; k++)
              e[k] = (l0 + k < L && !(dens && l0 + k == cc)) ? __ldcg(a.E + (size_t)s_slot[l0 + k] * np + i) : 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++)
              if (l0 + k < L) Es[(l0 + k) * kEmTS + t] = e[k];
          }
        }
        if (dens) {
          double xv[DMAX], r[DMAX];
#pragma unroll
          for (int j = 0; j < DMAX; j++) xv[j] = (j < d) ? xs[j * kEmTS + t] : 0.0;
          const double enew = exp(fma(-0.5, solve_lower<DMAX, (DMAX <= 12)>(s_rec, d, xv, r), s_rec[3]));
          Es[cc * kEmTS + t] = enew;
          if (valid) __stcg(a.E + (size_t)cslot * np + i, enew);
        }
        // sum_l lam_l E_il in component order (as the reference adds them); the loads of four
        // components are in flight at a time, the additions stay sequential
        double *__restrict__ Ecol = Es + t;
        const double *__restrict__ lam = s_lam;
        double sum = 0.0, emax = 0.0;
        int l = 0;
        for (; l + 4 <= L; l += 4) {
          const double e0 = Ecol[(l + 0) * kEmTS], e1 = Ecol[(l + 1) * kEmTS], e2 = Ecol[(l + 2) * kEmTS],
                       e3 = Ecol[(l + 3) * kEmTS];
          const double a0 = lam[l], a1 = lam[l + 1], a2 = lam[l + 2], a3 = lam[l + 3];
          sum = fma(a0, e0, sum);
          sum = fma(a1, e1, sum);
          sum = fma(a2, e2, sum);
          sum = fma(a3, e3, sum);
          emax = fmax(fmax(emax, fmax(e0, e1)), fmax(e2, e3));
        }
        for (; l < L; l++) {
          const double e = Ecol[l * kEmTS];
          sum = fma(lam[l], e, sum);
          emax = fmax(emax, e);
        }
        if (valid && unguarded) {
          // First E-step: w = lam * pdf / sum with no look at the sum, as the reference does (:737-745).  A sample
          // that every start component misses gives 0/0 = NaN, which poisons the column sums; components are then
          // annihilated until the (guarded) refresh after an annihilation clears it -- the reference's own path.
          for (l = 0; l < L; l++) Ecol[l * kEmTS] = (lam[l] * Ecol[l * kEmTS]) / sum;
        } else if (valid) {
          if (emax < 1e-290 && L > 0) {
            // Every component puts this sample below exp(-667): the cached densities are subnormal or zero, and
            // lam * exp(lpd) no longer tracks the reference's exp(log(lam) + lpd) (nor can 1/sum be formed).
            // Redo the sample the reference's way from the components' parameters (:849-866).
            double s2 = 0.0;
            for (l = 0; l < L; l++) {
              const double wl = exp(log(lam[l]) + lnormprob_slow(ctrl, l, d, xs + t));
              Ecol[l * kEmTS] = wl;
              s2 += wl;
            }
            if (s2 > 0) {
              for (l = 0; l < L; l++) Ecol[l * kEmTS] /= s2;
              ll += log(s2);
            } else {
              nfb += 1.0;
              const double w = 1.0 / L;
              for (l = 0; l < L; l++) Ecol[l * kEmTS] = w;
            }
          } else if (sum > 0) {  // the reference's guard (:855-866)
            const double inv = 1.0 / sum;
            ll += log(sum);
            l = 0;
            for (; l + 4 <= L; l += 4) {
              const double e0 = Ecol[(l + 0) * kEmTS], e1 = Ecol[(l + 1) * kEmTS], e2 = Ecol[(l + 2) * kEmTS],
                           e3 = Ecol[(l + 3) * kEmTS];
              Ecol[(l + 0) * kEmTS] = (lam[l] * e0) * inv;
              Ecol[(l + 1) * kEmTS] = (lam[l + 1] * e1) * inv;
              Ecol[(l + 2) * kEmTS] = (lam[l + 2] * e2) * inv;
              Ecol[(l + 3) * kEmTS] = (lam[l + 3] * e3) * inv;
            }
            for (; l < L; l++) Ecol[l * kEmTS] = (lam[l] * Ecol[l * kEmTS]) * inv;
          } else {
            nfb += 1.0;
            const double w = 1.0 / L;
            for (l = 0; l < L; l++) Ecol[l * kEmTS] = w;
          }
        } else {
          for (l = 0; l < L; l++) Ecol[l * kEmTS] = 0.0;
        }
        const double wn = Es[nx * kEmTS + t];
        ws[t] = wn;
        if (a.fused) {  // shift by the pivot: the rows become dx = x - pivot and w_next * dx
          double *wd = ws + kEmTS;
#pragma unroll
          for (int j = 0; j < DMAX; j++)
            if (j < d) {
              const double dxj = xs[j * kEmTS + t] - s_pivot[j];
              xs[j * kEmTS + t] = dxj;
              wd[j * kEmTS + t] = wn * dxj;
            }
        } else if (valid) {
          __stcg(a.wnxt + i, wn);
        }
        __syncthreads();
        // -- reduction phase: thread (col, prt) sums every 4th sample of column col
        if (col < L) {
          double s4 = 0.0;
#pragma unroll 8
          for (int k = 0; k < kEmThreads / 4; k++) s4 += Es[col * kEmTS + prt + 4 * k];
          acc_col += s4;
        }
        if (col < d) {
          double s4 = 0.0;
#pragma unroll 8
          for (int k = 0; k < kEmThreads / 4; k++) s4 = fma(ws[prt + 4 * k], xs[col * kEmTS + prt + 4 * k], s4);
          acc_s1 += s4;
        }
        if (a.fused) {
          const double *wd = ws + kEmTS;
#pragma unroll
          for (int q = 0; q < 2; q++) {
            if ((t & 63) + 64 * q < tri_d) {
              const double *pa = wd + ej[q] * kEmTS + 64 * half, *pb = xs + ek[q] * kEmTS + 64 * half;
              double s2 = 0.0;
#pragma unroll 8
              for (int k = 0; k < 64; k++) s2 = fma(pa[k], pb[k], s2);
              acc2[q] += s2;
            }
          }
        }
        __syncthreads();
      }
      // partial row: [colsum (Lmax) | loglik | fallbacks | S1 (d) | S2 (tri, fused mode)]
      acc_col += __shfl_xor_sync(0xffffffffu, acc_col, 1);
      acc_col += __shfl_xor_sync(0xffffffffu, acc_col, 2);
      acc_s1 += __shfl_xor_sync(0xffffffffu, acc_s1, 1);
      acc_s1 += __shfl_xor_sync(0xffffffffu, acc_s1, 2);
      if (prt == 0) {
        part_col[(size_t)col * pstride] = (col < L) ? acc_col : 0.0;
        if (col < d) part_col[(size_t)(kEmLmax + 2 + col) * pstride] = acc_s1;
      }
      ll = warp_sum(ll);
      nfb = warp_sum(nfb);
      if (lane == 0) {
        s_red[warp * NRED + 0] = ll;
        s_red[warp * NRED + 1] = nfb;
      }
      __syncthreads();
      if (t < 2) {
        double tot = 0.0;
        for (int w = 0; w < kEmWarps; w++) tot += s_red[w * NRED + t];
        part_col[(size_t)(kEmLmax + t) * pstride] = tot;
      }
      nv = kEmLmax + 2 + d;
      if (a.fused) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int e = (t & 63) + 64 * q;
          if (e < tri_d) s_red[half * NRED + e] = acc2[q];
        }
        __syncthreads();
        for (int e = t; e < tri_d; e += blockDim.x)
          part_col[(size_t)(kEmLmax + 2 + d + e) * pstride] = s_red[e] + s_red[NRED + e];
        nv += tri_d;
      }
    }

    // ------------------------------------------------------------------ barrier + leader
    const long long tk1 = clock64();
    const int role = barrier_arrive(a, epoch, nv, s_tot, s_chunk, chunk_cap);
    bool alive = true;
    if (role == 2) {
      const long long tl1 = clock64();
      em_leader_block<DMAX>(ctrl, a, pass, s_tot, s_lead);
      const long long tl2 = clock64();
      barrier_release(a, d, epoch);
      __syncthreads();
      if (threadIdx.x == 0 && a.rank == 0) {
        atomicAdd((unsigned long long *)&ctrl->dbg[3], (unsigned long long)(tl2 - tl1));
        atomicAdd((unsigned long long *)&ctrl->dbg[1], (unsigned long long)(tl1 - tk1));
      }
    } else {
      alive = barrier_wait(a, epoch);
      if (threadIdx.x == 0 && blockIdx.x == 0 && a.rank == 0) atomicAdd((unsigned long long *)&ctrl->dbg[4], (unsigned long long)(clock64() - tk1));
    }
    if (!alive) {  // a partner never arrived: report and leave instead of hanging
      if (threadIdx.x == 0) atomicExch(&ctrl->status, AMX_ECUDA);
      return;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0 && a.rank == 0) {
      atomicAdd((unsigned long long *)&ctrl->dbg[0], (unsigned long long)(tk1 - tk0));
      if (pass == kPassScatter) atomicAdd((unsigned long long *)&ctrl->dbg[6], (unsigned long long)(tk1 - tk0));
      if (pass == kPassDensRefresh || pass == kPassRefresh || pass == kPassRefresh0) atomicAdd((unsigned long long *)&ctrl->dbg[7], (unsigned long long)(tk1 - tk0));
    }
    const long long tk2 = clock64();
    // ------------------------------------------------------------------ reload the control state
    const EmPublic *pub = a.dev[a.rank].pub;  // this GPU's mirror, written by the leader before the release
    if (threadIdx.x == 0) {
      s_pass = ld_cg(&pub->pass);
      s_L = ld_cg(&pub->L);
      s_c = ld_cg(&pub->c);
      s_next = ld_cg(&pub->next);
    }
    if (threadIdx.x < kEmLmax) {
      s_lam[threadIdx.x] = ld_cg(&pub->lam[threadIdx.x]);
      s_slot[threadIdx.x] = ld_cg(&pub->slot[threadIdx.x]);
    }
    const int reclen = AMX_REC_HEAD + 2 * d + d * (d + 1) / 2;
    for (int q = threadIdx.x; q < reclen; q += blockDim.x) s_rec[q] = ld_cg(&pub->rec[q]);
    if (threadIdx.x < DMAX) s_pivot[threadIdx.x] = (threadIdx.x < d) ? ld_cg(&pub->pivot[threadIdx.x]) : 0.0;
    __syncthreads();
    pass = s_pass;
    if (threadIdx.x == 0 && blockIdx.x == 0 && a.rank == 0) atomicAdd((unsigned long long *)&ctrl->dbg[5], (unsigned long long)(clock64() - tk2));
    if (pass == kPassStop) break;
  }

  // optional dump of the responsibilities of the working state (step-parity tests)
  if (a.w_out != nullptr) {
    const int L = s_L;
    for (long i = i0; i < n; i += stride) {
      double sum = 0.0;
      for (int l = 0; l < L; l++) sum += s_lam[l] * __ldcg(a.E + (size_t)s_slot[l] * np + i);
      for (int l = 0; l < L; l++) {
        const double e = __ldcg(a.E + (size_t)s_slot[l] * np + i);
        a.w_out[(size_t)i * a.Lmax + l] = (sum > 0) ? s_lam[l] * e / sum : 1.0 / L;
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

// d > 12: one thread per entry of the mean / packed triangle, samples in order (n = 1000 d in the product)
__global__ void __launch_bounds__(256) autorj_generic_kernel(int d, long n, const double *x, const double *mu,
                                                             double *out) {
  const int nv = mu ? d * (d + 1) / 2 : d;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += gridDim.x * blockDim.x) {
    double acc = 0.0;
    if (mu == nullptr) {
      for (long i = 0; i < n; i++) acc += x[i * d + e];
    } else {
      int j = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((j + 1) * (j + 2) / 2 <= e) j++;
      while (j * (j + 1) / 2 > e) j--;
      const int k = e - j * (j + 1) / 2;
      const double mj = mu[j], mk = mu[k];
      for (long i = 0; i < n; i++) acc = fma(x[i * d + j] - mj, x[i * d + k] - mk, acc);
    }
    out[e] = acc;
  }
}

#include "amx_em2.cuh"

}  // namespace amx

using namespace amx;

// bytes of tile memory the aliased scratch of em_fit_kernel<DMAX> needs: s_red | s_tot | pad | LeaderS
template <int DMAX>
constexpr size_t em_scratch_bytes() {
  constexpr int TRI = DMAX * (DMAX + 1) / 2;
  constexpr int NRED = TRI > 2 * DMAX ? TRI : 2 * DMAX;
  constexpr int NTOT = kEmLmax + 2 + DMAX + TRI;
  return sizeof(double) * (size_t)(kEmWarps * NRED + NTOT + 1) + sizeof(LeaderS<DMAX>) + 16;
}

template <int DMAX>
static int em_occupancy(size_t smem, int *per_sm) {
  AMX_CUDA(cudaFuncSetAttribute(em_fit_kernel<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AMX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, em_fit_kernel<DMAX>, kEmThreads, smem));
  return AMX_OK;
}
template <int DMAX>
static int em_launch_one(EmArgs &a, unsigned grid, size_t smem, cudaStream_t st) {
  void *args[] = {&a};
  AMX_CUDA(cudaLaunchCooperativeKernel((void *)em_fit_kernel<DMAX>, dim3(grid), dim3(kEmThreads), args, smem, st));
  count_launch();
  return AMX_OK;
}
static int em_occupancy_d(int d, size_t smem, int *per_sm) {
  if (d <= 4) return em_occupancy<4>(smem, per_sm);
  if (d <= 8) return em_occupancy<8>(smem, per_sm);
  if (d <= 12) return em_occupancy<12>(smem, per_sm);
  if (d <= 20) return em_occupancy<20>(smem, per_sm);
  return em_occupancy<32>(smem, per_sm);
}
static int em_launch_d(int d, EmArgs &a, unsigned grid, size_t smem, cudaStream_t st) {
  if (d <= 4) return em_launch_one<4>(a, grid, smem, st);
  if (d <= 8) return em_launch_one<8>(a, grid, smem, st);
  if (d <= 12) return em_launch_one<12>(a, grid, smem, st);
  if (d <= 20) return em_launch_one<20>(a, grid, smem, st);
  return em_launch_one<32>(a, grid, smem, st);
}


// ---- second-generation kernel (amx_em2.cuh): d <= 12, one persistent CTA per SM ---------------------------------
struct V2Plan {
  int nteam, ns;
  size_t smem;
  int region0_doubles;
};
template <int DMAX, int NTEAM>
static int em2_plan_one(int d, int Lmax, size_t max_dyn, V2Plan *p) {
  const size_t fixed = v2_fixed_doubles<DMAX>(d, Lmax), scratch = v2_scratch_doubles<DMAX, NTEAM>();
  const size_t stage = (size_t)(d + Lmax + 1) * kV2TS;
  cudaFuncAttributes fa;
  AMX_CUDA(cudaFuncGetAttributes(&fa, em_fit_v2_kernel<DMAX, NTEAM>));
  const size_t avail = (max_dyn - fa.sharedSizeBytes - 64) / sizeof(double);
  if (avail < fixed + scratch + 2 * stage) return fail(AMX_ECUDA, "EM kernel does not fit in shared memory (d=%d, Lmax=%d)", d, Lmax);
  int ns = (int)((avail - fixed) / stage);
  if (ns > 8) ns = 8;
  if (ns < 2) return fail(AMX_ECUDA, "EM kernel: fewer than two ring stages fit (d=%d, Lmax=%d)", d, Lmax);
  size_t region0 = (size_t)ns * stage;
  if (region0 < scratch) region0 = scratch;
  p->nteam = NTEAM;
  p->ns = ns;
  p->region0_doubles = (int)region0;
  p->smem = sizeof(double) * (region0 + fixed);
  AMX_CUDA(cudaFuncSetAttribute(em_fit_v2_kernel<DMAX, NTEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
  return AMX_OK;
}
template <int DMAX, int NTEAM>
static int em2_launch_one(EmArgs &a, V2Args &v, unsigned grid, size_t smem, cudaStream_t st) {
  void *args[] = {&a, &v};
  AMX_CUDA(cudaLaunchCooperativeKernel((void *)em_fit_v2_kernel<DMAX, NTEAM>, dim3(grid), dim3(NTEAM * 128), args, smem, st));
  count_launch();
  return AMX_OK;
}
#define AMX_V2_DISPATCH(FN, ...)                                              \
  do {                                                                        \
    switch (d) {                                                              \
      case 1: return FN<1, 3>(__VA_ARGS__);                                   \
      case 2: return FN<2, 3>(__VA_ARGS__);                                   \
      case 3: return FN<3, 3>(__VA_ARGS__);                                   \
      case 4: return FN<4, 3>(__VA_ARGS__);                                   \
      case 5: return FN<5, 3>(__VA_ARGS__);                                   \
      case 6: return FN<6, 3>(__VA_ARGS__);                                   \
      case 7: return FN<7, 3>(__VA_ARGS__);                                   \
      case 8: return FN<8, 3>(__VA_ARGS__);                                   \
      case 9: return FN<9, 3>(__VA_ARGS__);                                   \
      case 10: return FN<10, 3>(__VA_ARGS__);                                 \
      case 11: return FN<11, 3>(__VA_ARGS__);                                 \
      default: return FN<12, 3>(__VA_ARGS__);                                 \
    }                                                                         \
  } while (0)
static int em2_plan(int d, int nteam, int Lmax, size_t max_dyn, V2Plan *p) { AMX_V2_DISPATCH(em2_plan_one, d, Lmax, max_dyn, p); }
static int em2_launch(int d, int nteam, EmArgs &a, V2Args &v, unsigned grid, size_t smem, cudaStream_t st) {
  AMX_V2_DISPATCH(em2_launch_one, a, v, grid, smem, st);
}

__global__ void em_gather_rows_kernel(const double *x, const int *idx, int L, int d, double *out) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < L * d; q += gridDim.x * blockDim.x)
    out[q] = x[(size_t)idx[q / d] * d + (q % d)];
}

// One fit over ndev GPUs (ndev = 1: the ordinary fit).  The samples are split into contiguous shards, one
// per GPU; every GPU runs the same persistent kernel on its shard and they meet at the cross-GPU barrier of
// the kernel, exchanging only the per-pass partial sums (<= 2 KB per GPU) through peer memory.
// x_host: n x d row-major host samples, or NULL with x_dev0 = device samples (ndev must be 1).
// ---- workspace pool of the fit --------------------------------------------------------------------------------
// A fit needs ~n (2d + Lmax + 2) doubles of device memory (400 MB for n = 1e6, d = 10, Lmax = 30); cudaMalloc and
// cudaFree of blocks that size cost milliseconds each and synchronise the device, which is what an end-to-end
// fit (host samples in, mixture out) was spending most of its non-kernel time on.  Blocks are therefore kept
// after a fit and handed out again when a later fit asks for the same size on the same device; everything a
// kernel reads before writing is initialised explicitly, as it was with fresh memory.  amx_release_workspace()
// returns the idle blocks to the driver; at most 4 GB stay idle.
namespace {
struct WsBlock {
  int dev;
  size_t bytes;
  void *p;
};
std::vector<WsBlock> g_ws_idle, g_ws_live;
std::mutex g_ws_mu;
constexpr size_t kWsIdleCap = (size_t)4 << 30;

template <class T>
cudaError_t ws_malloc(T **p, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_ws_mu);
  for (size_t i = 0; i < g_ws_idle.size(); i++)
    if (g_ws_idle[i].dev == dev && g_ws_idle[i].bytes == bytes) {
      *p = static_cast<T *>(g_ws_idle[i].p);
      g_ws_live.push_back(g_ws_idle[i]);
      g_ws_idle.erase(g_ws_idle.begin() + i);
      return cudaSuccess;
    }
  void *q = nullptr;
  e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return e;
  *p = static_cast<T *>(q);
  g_ws_live.push_back({dev, bytes, q});
  return cudaSuccess;
}

void ws_trim_locked(size_t cap) {
  size_t idle = 0;
  for (const WsBlock &b : g_ws_idle) idle += b.bytes;
  int cur = 0;
  cudaGetDevice(&cur);
  while (idle > cap && !g_ws_idle.empty()) {  // oldest first
    cudaSetDevice(g_ws_idle.front().dev);
    cudaFree(g_ws_idle.front().p);
    idle -= g_ws_idle.front().bytes;
    g_ws_idle.erase(g_ws_idle.begin());
  }
  cudaSetDevice(cur);
}

void ws_free(const void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_ws_mu);
  for (size_t i = 0; i < g_ws_live.size(); i++)
    if (g_ws_live[i].p == p) {
      g_ws_idle.push_back(g_ws_live[i]);
      g_ws_live.erase(g_ws_live.begin() + i);
      ws_trim_locked(kWsIdleCap);
      return;
    }
  cudaFree(const_cast<void *>(p));  // not ours
}
}  // namespace

extern "C" int amx_release_workspace(void) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  ws_trim_locked(0);
  return AMX_OK;
}

static int em_fit_general(int ndev, const int *devices, int d, long n, const double *x_host, const double *x_dev0,
                          int Lmax, int maxit, const int *init_idx, double *wt, double *mean, double *tri, int *trace_L,
                          double *trace_loglik, double *trace_cost, int *trace_ann, double *cur_wt, double *cur_mean,
                          double *cur_tri, int *cur_L, double *cur_w, amx_em_result *res) {
  if (d < 1 || d > kEmDmax || Lmax < 1 || Lmax > kEmLmax || n < 1 || maxit < 0 || !init_idx || !wt || !mean || !tri)
    return fail(AMX_EINVAL, "amx_em_fit: need 1<=d<=%d, 1<=Lmax<=%d, n>=1, maxit>=0 (got d=%d Lmax=%d n=%ld maxit=%d)",
                kEmDmax, kEmLmax, d, Lmax, n, maxit);
  if (n < Lmax) return fail(AMX_EINVAL, "amx_em_fit: fewer samples (%ld) than start components (%d)", n, Lmax);
  if (ndev < 1 || ndev > kEmMaxDev || (ndev > 1 && !x_host)) return fail(AMX_EINVAL, "amx_em_fit: bad device list");
  if (n < (long)ndev * kEmThreads) ndev = 1;  // nothing to share
  for (int l = 0; l < Lmax; l++) {
    if (init_idx[l] < 0 || init_idx[l] >= n) return fail(AMX_EINVAL, "amx_em_fit: init_idx[%d]=%d out of range", l, init_idx[l]);
    for (int m = 0; m < l; m++)
      if (init_idx[m] == init_idx[l]) return fail(AMX_EINVAL, "amx_em_fit: start rows must be distinct");
  }
  int home = 0;
  AMX_CUDA(cudaGetDevice(&home));
  int devs[kEmMaxDev];
  for (int g = 0; g < ndev; g++) devs[g] = devices ? devices[g] : home;
  int sms = 0, coop = 0;
  AMX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devs[0]));
  AMX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, devs[0]));
  if (!coop) return fail(AMX_ECUDA, "device lacks cooperative launch");
  const int tlen = d * (d + 1) / 2, cap = maxit + 1;
  const char *nb = getenv("AMX_EM_NBUF"), *tm = getenv("AMX_EM_TMA");
  const int use_tma = tm ? (atoi(tm) != 0) : 1;  // measured on B200 (n=1e6, d=10, L=30): TMA 183 us/step, plain loads 244
  const int nbuf = (use_tma && nb && atoi(nb) == 2) ? 2 : 1;
  const char *fu = getenv("AMX_EM_FUSED");
  const int fused = (d <= 12 && use_tma && nbuf == 1) ? (fu ? (atoi(fu) != 0) : 1) : 0;  // entries e, e+64 cover tri(12) = 78
  // second-generation kernel (amx_em2.cuh) wherever the fused step applies; AMX_EM_V2=0 keeps the first generation
  const char *v2e = getenv("AMX_EM_V2"), *v2t = getenv("AMX_EM_TEAMS");
  const bool use_v2 = fused && d <= kV2Dmax && (v2e ? atoi(v2e) != 0 : true);
  (void)v2t;
  const int nteam = 3;  // 12 warps: a whole number of register-file allocation units at 168 registers (4 teams at 128
                        // registers measured slower: 107.6 vs 97.4 us per component step; 2 teams at 254: 125)
  V2Plan plan;
  memset(&plan, 0, sizeof(plan));
  if (use_v2) {
    int max_dyn = 0;
    AMX_CUDA(cudaDeviceGetAttribute(&max_dyn, cudaDevAttrMaxSharedMemoryPerBlockOptin, devs[0]));
    for (int g = 0; g < ndev; g++) {  // the attribute is per device
      AMX_CUDA(cudaSetDevice(devs[g]));
      if (int prc = em2_plan(d, nteam, Lmax, (size_t)max_dyn, &plan)) {
        cudaSetDevice(home);
        return prc;
      }
    }
    if (const char *nse = getenv("AMX_EM_STAGES")) {
      const int want_ns = atoi(nse);
      if (want_ns >= 2 && want_ns < plan.ns) plan.ns = want_ns;  // (region0 keeps its size)
    }
    AMX_CUDA(cudaSetDevice(home));
  }
  size_t smem = nbuf * sizeof(double) * (size_t)(d + Lmax + 1 + (fused ? d : 0)) * kEmTS;
  {  // room for the reduction scratch and the leader's state, which alias the tile (see em_fit_kernel)
    const size_t need = d <= 4 ? em_scratch_bytes<4>() : d <= 8 ? em_scratch_bytes<8>() : d <= 12 ? em_scratch_bytes<12>()
                        : d <= 20 ? em_scratch_bytes<20>() : em_scratch_bytes<32>();
    if (smem < need) smem = need;
  }

  EmArgs A[kEmMaxDev];
  V2Args V[kEmMaxDev];
  V2Sync *v2sync[kEmMaxDev];
  memset(V, 0, sizeof(V));
  memset(v2sync, 0, sizeof(v2sync));
  cudaStream_t st[kEmMaxDev];
  int *idx_dev = nullptr;
  double *init_rows = nullptr;
  EmCtrl *ctrl = nullptr;
  int *tr_L = nullptr, *tr_ann = nullptr;
  double *tr_ll = nullptr, *tr_cost = nullptr;
  long off[kEmMaxDev + 1];
  off[0] = 0;
  for (int g = 0; g < ndev; g++) off[g + 1] = off[g] + n / ndev + (g < n % ndev ? 1 : 0);
  memset(A, 0, sizeof(A));
  int rc = AMX_OK;

  // peer access between all participants
  // `devices` naming ONE GPU several times: the ranks are emulated by slices of a single grid on that GPU (the sharded
  // exchange -- per-rank reduction, rows posted to every rank, sums in rank order -- exercised on a single-GPU box)
  bool virt = ndev > 1;
  for (int g = 1; g < ndev; g++) virt = virt && devs[g] == devs[0];
  if (ndev > 1 && !virt)
    for (int g = 0; g < ndev; g++) {
      AMX_CUDA(cudaSetDevice(devs[g]));
      for (int h = 0; h < ndev; h++)
        if (h != g) {
          int can = 0;
          AMX_CUDA(cudaDeviceCanAccessPeer(&can, devs[g], devs[h]));
          if (!can) {
            cudaSetDevice(home);
            return fail(AMX_ECUDA, "GPU %d cannot access GPU %d's memory", devs[g], devs[h]);
          }
          cudaError_t e = cudaDeviceEnablePeerAccess(devs[h], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            cudaSetDevice(home);
            return fail(AMX_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
          }
          cudaGetLastError();
        }
    }

  // one stream per GPU: everything a kernel reads is enqueued on the stream that launches it (never on the legacy
  // default stream, which nothing orders against non-blocking streams)
  for (int g = 0; g < ndev; g++) {
    AMX_CUDA(cudaSetDevice(devs[g]));
    if (g == 0 && ndev == 1) st[g] = stream();
    else AMX_CUDA(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
  }
  // GPU 0: the control block, the traces and the start rows
  AMX_CUDA(cudaSetDevice(devs[0]));
  AMX_CUDA(ws_malloc(&ctrl, sizeof(EmCtrl)));
  AMX_CUDA(cudaMemsetAsync(ctrl, 0, sizeof(EmCtrl), st[0]));
  AMX_CUDA(ws_malloc(&init_rows, sizeof(double) * (size_t)Lmax * d));
  AMX_CUDA(ws_malloc(&tr_L, sizeof(int) * cap));
  AMX_CUDA(ws_malloc(&tr_ann, sizeof(int) * cap));
  AMX_CUDA(ws_malloc(&tr_ll, sizeof(double) * cap));
  AMX_CUDA(ws_malloc(&tr_cost, sizeof(double) * cap));
  if (x_host) {  // the Lmax start rows, gathered on the host: one small copy
    std::vector<double> rows_h((size_t)Lmax * d);
    for (int l = 0; l < Lmax; l++) memcpy(&rows_h[(size_t)l * d], x_host + (size_t)init_idx[l] * d, sizeof(double) * d);
    AMX_CUDA(cudaMemcpyAsync(init_rows, rows_h.data(), sizeof(double) * rows_h.size(), cudaMemcpyHostToDevice, st[0]));
    AMX_CUDA(cudaStreamSynchronize(st[0]));  // rows_h leaves scope
  } else {       // samples already on the device: gather with a kernel, ordered after whatever filled them on this stream
    AMX_CUDA(ws_malloc(&idx_dev, sizeof(int) * kEmLmax));
    AMX_CUDA(cudaMemcpyAsync(idx_dev, init_idx, sizeof(int) * Lmax, cudaMemcpyHostToDevice, st[0]));
    em_gather_rows_kernel<<<1, 256, 0, st[0]>>>(x_dev0, idx_dev, Lmax, d, init_rows);
    count_launch();
    AMX_CUDA(cudaGetLastError());
  }

  // per-GPU shards
  for (int g = 0; g < ndev; g++) {
    AMX_CUDA(cudaSetDevice(devs[g]));
    EmArgs &a = A[g];
    a.ndev = ndev;
    a.rank = g;
    a.n_total = n;
    a.init_rows = init_rows;
    a.d = d;
    a.Lmax = Lmax;
    a.maxit = maxit;
    a.use_tma = use_tma;
    a.nbuf = nbuf;
    a.smem_doubles = (int)(smem / sizeof(double));
    a.fused = fused;
    a.n = off[g + 1] - off[g];
    a.npad = (a.n + kEmThreads - 1) / kEmThreads * kEmThreads;  // whole tiles; the padding is zero and carries no weight
    a.ctrl = ctrl;
    a.trace_L = tr_L;
    a.trace_ann = tr_ann;
    a.trace_loglik = tr_ll;
    a.trace_cost = tr_cost;
    if (x_host) {
      double *xs = nullptr;
      AMX_CUDA(ws_malloc(&xs, sizeof(double) * (size_t)a.n * d));
      AMX_CUDA(cudaMemcpyAsync(xs, x_host + (size_t)off[g] * d, sizeof(double) * (size_t)a.n * d, cudaMemcpyHostToDevice, st[g]));
      a.x = xs;
    } else {
      a.x = x_dev0;
    }
    AMX_CUDA(ws_malloc(&a.xT, sizeof(double) * (size_t)d * a.npad));
    AMX_CUDA(ws_malloc(&a.E, sizeof(double) * (size_t)Lmax * a.npad));
    AMX_CUDA(ws_malloc(&a.wnxt, sizeof(double) * (size_t)a.npad));
    if (!use_v2) {  // the second generation never reads a padding sample's cache, and has no w_next array
      AMX_CUDA(cudaMemsetAsync(a.xT, 0, sizeof(double) * (size_t)d * a.npad, st[g]));
      AMX_CUDA(cudaMemsetAsync(a.E, 0, sizeof(double) * (size_t)Lmax * a.npad, st[g]));
      AMX_CUDA(cudaMemsetAsync(a.wnxt, 0, sizeof(double) * (size_t)a.npad, st[g]));
    }  // (the second generation ignores whatever the padding of the last tile holds)
    if (cur_w) AMX_CUDA(ws_malloc(&a.w_out, sizeof(double) * (size_t)a.n * Lmax));
    int per_sm = 0;
    if (use_v2) per_sm = 1;
    else if ((rc = em_occupancy_d(d, smem, &per_sm))) return rc;
    if (per_sm < 1) return fail(AMX_ECUDA, "EM kernel does not fit on an SM (%zu B of shared memory)", smem);
    long want = a.npad / kEmThreads, capb = (long)sms * per_sm;
    if (virt) {  // the ranks share the GPU: equal slices of its SMs, the same number of CTAs for every rank
      capb = sms / ndev;
      for (int h = 0; h < ndev; h++) {
        const long nh = off[h + 1] - off[h], th = (nh + kEmThreads - 1) / kEmThreads;
        if (th < capb) capb = th;
      }
      if (capb < 1) return fail(AMX_EINVAL, "amx_em_fit_multi: too few samples for %d emulated ranks", ndev);
    }
    unsigned grid = (unsigned)(want < capb ? want : capb);
    if (use_v2) {
      AMX_CUDA(ws_malloc(&v2sync[g], sizeof(V2Sync)));
      AMX_CUDA(cudaMemsetAsync(v2sync[g], 0, sizeof(V2Sync), st[g]));
    }
    EmDev dv;
    memset(&dv, 0, sizeof(dv));
    dv.grid = (int)grid;
    AMX_CUDA(ws_malloc(&dv.part, sizeof(double) * (size_t)grid * kEmNV));
    AMX_CUDA(ws_malloc(&dv.devrow, sizeof(double) * kEmNV));
    AMX_CUDA(ws_malloc(&dv.flags, sizeof(unsigned) * 8 * (size_t)grid));
    AMX_CUDA(ws_malloc(&dv.arrive, sizeof(unsigned) * 8));
    AMX_CUDA(ws_malloc(&dv.pub, sizeof(EmPublic)));
    AMX_CUDA(cudaMemsetAsync(dv.flags, 0, sizeof(unsigned) * 8 * (size_t)grid, st[g]));
    AMX_CUDA(cudaMemsetAsync(dv.arrive, 0, sizeof(unsigned) * 8, st[g]));
    AMX_CUDA(cudaMemsetAsync(dv.pub, 0, sizeof(EmPublic), st[g]));
    for (int h = 0; h < ndev; h++) A[h].dev[g] = dv;
    V[g].part = dv.part;
    V[g].ns = plan.ns;
    V[g].region0_doubles = plan.region0_doubles;
    V[g].debug = getenv("AMX_EM_DEBUG") ? 1 : 0;
  }
  for (int g = 0; g < ndev; g++)
    for (int h = 0; h < ndev; h++) V[g].sync[h] = v2sync[h];
  for (int g = 0; g < ndev; g++) {  // everything above must have landed before any kernel starts
    AMX_CUDA(cudaSetDevice(devs[g]));
    AMX_CUDA(cudaStreamSynchronize(st[g]));
  }

  // launch: one persistent cooperative kernel per GPU; they wait for one another inside
  cudaEvent_t e0, e1;
  AMX_CUDA(cudaSetDevice(devs[0]));
  AMX_CUDA(cudaEventCreate(&e0));
  AMX_CUDA(cudaEventCreate(&e1));
  AMX_CUDA(cudaEventRecord(e0, st[0]));
  int launched = 0;
  EmArgs *va_dev = nullptr;
  V2Args *vv_dev = nullptr;
  if (virt) {
    if (!use_v2) return fail(AMX_EINVAL, "emulated ranks on one GPU need the second-generation kernel (d <= %d)", kV2Dmax);
    AMX_CUDA(ws_malloc(&va_dev, sizeof(EmArgs) * kEmMaxDev));
    AMX_CUDA(ws_malloc(&vv_dev, sizeof(V2Args) * kEmMaxDev));
    AMX_CUDA(cudaMemcpyAsync(va_dev, A, sizeof(EmArgs) * ndev, cudaMemcpyHostToDevice, st[0]));
    AMX_CUDA(cudaMemcpyAsync(vv_dev, V, sizeof(V2Args) * ndev, cudaMemcpyHostToDevice, st[0]));
    V2Args v0 = V[0];
    v0.vranks = ndev;
    v0.va = va_dev;
    v0.vv = vv_dev;
    rc = em2_launch(d, nteam, A[0], v0, (unsigned)(A[0].dev[0].grid * ndev), plan.smem, st[0]);
    if (rc == AMX_OK) launched = 1;
  }
  for (int g = 0; g < ndev && rc == AMX_OK && !virt; g++) {
    cudaSetDevice(devs[g]);
    rc = use_v2 ? em2_launch(d, nteam, A[g], V[g], (unsigned)A[g].dev[g].grid, plan.smem, st[g])
                : em_launch_d(d, A[g], (unsigned)A[g].dev[g].grid, smem, st[g]);
    if (rc == AMX_OK) launched++;
  }
  if (rc != AMX_OK && launched > 0) {
    // a partner failed to start: tell the running kernels to stop (they poll their flags / public state)
    EmPublic stop_pub;
    memset(&stop_pub, 0, sizeof(stop_pub));
    stop_pub.pass = kPassStop;
    cudaStream_t side;
    for (int g = 0; g < launched; g++) {
      cudaSetDevice(devs[g]);
      cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
      cudaMemcpyAsync(A[g].dev[g].pub, &stop_pub, sizeof(stop_pub), cudaMemcpyHostToDevice, side);
      cudaMemsetAsync(A[g].dev[g].flags, 0xFF, sizeof(unsigned) * 8 * (size_t)A[g].dev[g].grid, side);
      // (second generation: the partners' flags never rise; its watchdog ends the kernels that did start)
      cudaStreamSynchronize(side);
      cudaStreamDestroy(side);
    }
  }
  cudaSetDevice(devs[0]);
  cudaEventRecord(e1, st[0]);
  for (int g = 0; g < launched; g++) {
    cudaSetDevice(devs[g]);
    cudaError_t e = cudaStreamSynchronize(st[g]);
    if (e != cudaSuccess && rc == AMX_OK) rc = fail(AMX_ECUDA, "EM kernel on GPU %d: %s", devs[g], cudaGetErrorString(e));
  }
  cudaSetDevice(devs[0]);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);

  int status = 0;
  if (rc == AMX_OK) {
    std::vector<char> hc(sizeof(EmCtrl));
    AMX_CUDA(cudaMemcpy(hc.data(), ctrl, sizeof(EmCtrl), cudaMemcpyDeviceToHost));
    const EmCtrl *c = reinterpret_cast<const EmCtrl *>(hc.data());
    for (int l = 0; l < c->best_L; l++) {
      wt[l] = c->best_lam[l];
      for (int j = 0; j < d; j++) mean[(size_t)l * d + j] = c->best_mu[l][j];
      for (int j = 0; j < tlen; j++) tri[(size_t)l * tlen + j] = c->best_B[l][j];
    }
    const int it = c->iters;
    if (trace_L) AMX_CUDA(cudaMemcpy(trace_L, tr_L, sizeof(int) * it, cudaMemcpyDeviceToHost));
    if (trace_ann) AMX_CUDA(cudaMemcpy(trace_ann, tr_ann, sizeof(int) * it, cudaMemcpyDeviceToHost));
    if (trace_loglik) AMX_CUDA(cudaMemcpy(trace_loglik, tr_ll, sizeof(double) * it, cudaMemcpyDeviceToHost));
    if (trace_cost) AMX_CUDA(cudaMemcpy(trace_cost, tr_cost, sizeof(double) * it, cudaMemcpyDeviceToHost));
    if (cur_L) *cur_L = c->L;
    for (int l = 0; l < c->L; l++) {
      if (cur_wt) cur_wt[l] = c->lam[l];
      if (cur_mean)
        for (int j = 0; j < d; j++) cur_mean[(size_t)l * d + j] = c->mu[l][j];
      if (cur_tri)
        for (int j = 0; j < tlen; j++) cur_tri[(size_t)l * tlen + j] = c->B[l][j];
    }
    if (cur_w)
      for (int g = 0; g < ndev; g++) {
        cudaSetDevice(devs[g]);
        AMX_CUDA(cudaMemcpy(cur_w + (size_t)off[g] * Lmax, A[g].w_out, sizeof(double) * (size_t)A[g].n * Lmax, cudaMemcpyDeviceToHost));
      }
    if (res) {
      res->L = c->best_L;
      res->iters = it;
      res->status = c->status;
      res->comp_steps = c->comp_steps;
      res->kernel_ms = ms;
      res->flops = c->flops;
      res->bytes = 8.0 * d * (double)n * (double)c->comp_steps;
      res->bytes_requested = use_v2 ? c->s2 : 0.0;
    }
    if (getenv("AMX_EM_DEBUG") && use_v2) {
      V2Sync hs;
      if (cudaMemcpy(&hs, v2sync[0], sizeof(hs), cudaMemcpyDeviceToHost) == cudaSuccess) {
        long long lo = hs.dbg_cta[0], hi = hs.dbg_cta[0];
        double av = 0;
        const int gg = A[0].dev[0].grid < 160 ? A[0].dev[0].grid : 160;
        for (int b = 0; b < gg; b++) {
          lo = hs.dbg_cta[b] < lo ? hs.dbg_cta[b] : lo;
          hi = hs.dbg_cta[b] > hi ? hs.dbg_cta[b] : hi;
          av += (double)hs.dbg_cta[b] / gg;
        }
        fprintf(stderr, "[em2 dbg] data-pass cycles per CTA: min %lld avg %.0f max %lld\n", lo, av, hi);
      }
    }
    if (getenv("AMX_EM_DEBUG") && use_v2)
      fprintf(stderr, "[em2 dbg] %d GPU(s), %d teams, %d stages, %zu B smem; CTA 0 of GPU 0, cycles: data passes %lld | exchange %lld (post row + arrive %lld, wait for totals %lld) | sequential section %lld (%ld component steps)\n",
              ndev, nteam, plan.ns, plan.smem, c->dbg[0], c->dbg[1], c->dbg[4], c->dbg[5], c->dbg[3], c->comp_steps);
    else if (getenv("AMX_EM_DEBUG"))
      fprintf(stderr, "[em dbg] %d GPU(s), cycles on GPU 0: block0 data pass %lld | leader: arrive-skew+reduce %lld logic %lld | block0 wait %lld reload %lld (phases ~%ld) | scatter passes %lld refresh passes %lld\n",
              ndev, c->dbg[0], c->dbg[1], c->dbg[3], c->dbg[4], c->dbg[5], 2 * c->comp_steps, c->dbg[6], c->dbg[7]);
    status = c->status;
  }
  for (int g = 0; g < ndev; g++) {
    cudaSetDevice(devs[g]);
    EmArgs &a = A[g];
    if (x_host) ws_free(const_cast<double *>(a.x));
    ws_free(a.xT); ws_free(a.E); ws_free(a.wnxt); ws_free(a.w_out);
    ws_free(a.dev[g].part); ws_free(a.dev[g].devrow); ws_free(a.dev[g].flags); ws_free(a.dev[g].arrive); ws_free(a.dev[g].pub);
    ws_free(v2sync[g]);
    if (!(g == 0 && ndev == 1)) cudaStreamDestroy(st[g]);
  }
  cudaSetDevice(devs[0]);
  ws_free(idx_dev);
  ws_free(va_dev);
  ws_free(vv_dev);
  ws_free(ctrl); ws_free(init_rows); ws_free(tr_L); ws_free(tr_ann); ws_free(tr_ll); ws_free(tr_cost);
  cudaSetDevice(home);
  if (rc != AMX_OK) return rc;
  if (status == AMX_ECUDA) return fail(status, "EM fit: a GPU stopped answering at the cross-GPU barrier");
  if (status && !use_v2) return fail(status, "EM fit: scatter matrix not positive definite");
  // (second generation: a non-positive-definite scatter is carried as the reference carries it -- see v2_leader_warp --
  // and reported in res->status only)
  return AMX_OK;
}

static int em_fit_impl(int d, long n, const double *x_dev, int Lmax, int maxit, const int *init_idx, double *wt,
                       double *mean, double *tri, int *trace_L, double *trace_loglik, double *trace_cost,
                       int *trace_ann, double *cur_wt, double *cur_mean, double *cur_tri, int *cur_L, double *cur_w,
                       amx_em_result *res) {
  return em_fit_general(1, nullptr, d, n, nullptr, x_dev, Lmax, maxit, init_idx, wt, mean, tri, trace_L, trace_loglik,
                        trace_cost, trace_ann, cur_wt, cur_mean, cur_tri, cur_L, cur_w, res);
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
  AMX_CUDA(ws_malloc(&x_dev, sizeof(double) * (size_t)n * d));  // from the workspace pool, like the rest of the fit
  cudaError_t ce = cudaMemcpyAsync(x_dev, x, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, stream());
  int rc = ce == cudaSuccess ? em_fit_impl(d, n, x_dev, Lmax, maxit, init_idx, wt, mean, tri, trace_L, trace_loglik, trace_cost,
                                           trace_ann, cur_wt, cur_mean, cur_tri, cur_L, cur_w, res)
                             : fail(AMX_ECUDA, "amx_em_fit: sample upload: %s", cudaGetErrorString(ce));
  ws_free(x_dev);
  return rc;
}

int amx_em_fit_multi(int ndev, const int *devices, int d, long n, const double *x, int Lmax, int maxit,
                     const int *init_idx, double *wt, double *mean, double *tri, int *trace_L, double *trace_loglik,
                     double *trace_cost, int *trace_ann, double *cur_wt, double *cur_mean, double *cur_tri, int *cur_L,
                     double *cur_w, amx_em_result *res) {
  if (int rc = require_device()) return rc;
  if (!x || !devices) return fail(AMX_EINVAL, "amx_em_fit_multi: null samples or device list");
  return em_fit_general(ndev, devices, d, n, x, nullptr, Lmax, maxit, init_idx, wt, mean, tri, trace_L, trace_loglik,
                        trace_cost, trace_ann, cur_wt, cur_mean, cur_tri, cur_L, cur_w, res);
}

int amx_autorj_fit(int d, long n, const double *x, double *wt, double *mean, double *tri) {
  if (int rc = require_device()) return rc;
  if (d < 1 || d > kEmDmax || n < 2 || !x) return fail(AMX_EINVAL, "amx_autorj_fit: need 1<=d<=%d, n>=2", kEmDmax);
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
    const int nv = phase ? tlen : d;
    if (d > 12) {  // entry-parallel, samples in order
      autorj_generic_kernel<<<(nv + 255) / 256, 256, 0, stream()>>>(d, n, x_dev, mu_arg, part);
      count_launch();
      AMX_CUDA(cudaGetLastError());
      AMX_CUDA(cudaMemcpyAsync(hp.data(), part, sizeof(double) * nv, cudaMemcpyDeviceToHost, stream()));
      AMX_CUDA(cudaStreamSynchronize(stream()));
      for (int q = 0; q < nv; q++) {
        if (phase == 0) mean[q] = hp[q] / (double)n;
        else tri[q] = hp[q] / (double)(n - 1);
      }
    } else {
      if (d <= 4) autorj_moment_kernel<4><<<grid, kEmThreads, 0, stream()>>>(d, n, x_dev, mu_arg, part);
      else autorj_moment_kernel<12><<<grid, kEmThreads, 0, stream()>>>(d, n, x_dev, mu_arg, part);
      count_launch();
      AMX_CUDA(cudaGetLastError());
      AMX_CUDA(cudaMemcpyAsync(hp.data(), part, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost, stream()));
      AMX_CUDA(cudaStreamSynchronize(stream()));
      for (int q = 0; q < nv; q++) {
        double t = 0.0;
        for (unsigned b = 0; b < grid; b++) t += hp[(size_t)b * kEmNV + q];
        if (phase == 0) mean[q] = t / (double)n;
        else tri[q] = t / (double)(n - 1);
      }
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
