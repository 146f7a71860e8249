// amx_rwm_kernels.cuh -- K1 kernel templates and their launcher, in a header so that a plug-in translation unit
// (amx_plugin_tu.cu) instantiates them for a user's __device__ log-posterior.  Host driver: amx_rwm.cu.
#pragma once

#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

#include "amx_internal.cuh"
#include "amx_mailbox.cuh"
#include "amx_targets.cuh"

namespace amx {

constexpr int kRwmThreads = 128;

struct RwmArgs {
  const void *tgt_blob;
  int tgt_flags;
  int model_k, d;
  int dof;                // Student-t proposals when > 0 (rt(), automix.c:1663-1680)
  int nsweepr, nburn;     // total sweeps (incl. the extra tenth), and the adaptation-only prefix
  long nchains;
  const double *init;     // [d]
  const double *gtab;     // [nsweepr] 10 * (sweep+1)^(-2/3)
  unsigned long long seed;
  const double *tape;
  unsigned long long tape_stride;
  double *sig_out;      // [nchains][d]
  double *samples_out;  // [nchains][1000 d][d]
  double *sig_trace0, *acc_trace0;  // chain 0: [nsweepr/100][d]
  int *status;
};

__global__ void rwm_gamma_kernel(double *g, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // gamma = 10.0 * pow(1.0 / (sweep + 1), 2.0 / 3.0), sweep = i + 1   (automix.c:619)
  if (i < n) g[i] = 10.0 * pow(1.0 / (double)(i + 2), 2.0 / 3.0);
}

template <int DMAX, class TGT, class RNG>
__global__ void __launch_bounds__(kRwmThreads) rwm_adapt_kernel(RwmArgs a) {
  TGT T;
  T.bind(a.tgt_blob, a.tgt_flags);
  const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.nchains) return;
  const int d = a.d, k = a.model_k;
  RNG u;
  if constexpr (std::is_same<RNG, TapeStream>::value) u.open(a.tape, a.tape_stride, (unsigned long long)id, 0ull);
  else u.open(a.seed, (unsigned long long)id, 0ull);

  double cur[DMAX], prop[DMAX], sig[DMAX];
  int nacc[DMAX], ntry[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; i++) {
    cur[i] = prop[i] = (i < d) ? a.init[i] : 0.0;
    sig[i] = 10.0;  // :595
    nacc[i] = ntry[i] = 0;
  }
  double lp = T.template eval<DMAX>(k, cur);
  const long nstore = 1000L * d;
  double *out = a.samples_out + (size_t)id * nstore * d;
  long stored = 0;
  int ntrace = 0;
  const double alphastar = 0.25;

  for (int sweep = 1; sweep <= a.nsweepr; sweep++) {
    const int remain = a.nsweepr - sweep;
    const double uu = u.next();
    if (sweep > a.nburn && uu < 0.1) {  // block move, no adaptation (:606-617)
      int i = 0;
      for (; i + 1 < d; i += 2) {
        double z0, z1;
        gauss_pair(u, z0, z1);
        aset(prop, i, z0);
        aset(prop, i + 1, z1);
      }
      if (d & 1) aset(prop, d - 1, gauss_single(u));
      const double den = a.dof > 0 ? t_divisor(a.dof, u) : 1.0;
      for (int j = 0; j < d; j++) {
        const double z = a.dof > 0 ? aget(prop, j) / den : aget(prop, j);
        aset(prop, j, fma(aget(sig, j), z, aget(cur, j)));
      }
      const double lpn = T.template eval<DMAX>(k, prop);
      if (u.next() < mh_prob(lpn - lp)) {
#pragma unroll
        for (int j = 0; j < DMAX; j++) cur[j] = prop[j];
        lp = lpn;
      } else {
#pragma unroll
        for (int j = 0; j < DMAX; j++) prop[j] = cur[j];
      }
    } else {  // coordinate-wise moves with scale adaptation (:618-640)
      const double gam = a.gtab[sweep - 1];
      for (int i = 0; i < d; i++) {
        double z = gauss_single(u);
        if (a.dof > 0) z /= t_divisor(a.dof, u);
        const double si = aget(sig, i);
        aset(prop, i, fma(si, z, aget(cur, i)));
        const double lpn = T.template eval<DMAX>(k, prop);
        const double acc = min_m(1.0, mh_prob(lpn - lp));
        if (u.next() < acc) {
          if constexpr (DMAX <= kRegArrayMax) {
#pragma unroll
            for (int j = 0; j < DMAX; j++) {
              nacc[j] += (j == i);
              ntry[j] += (j == i);
            }
          } else {
            nacc[i]++;
            ntry[i]++;
          }
          aset(cur, i, aget(prop, i));
          lp = lpn;
          aset(sig, i, max_m(0.0, si - gam * (alphastar - 1.0)));
        } else {
          if constexpr (DMAX <= kRegArrayMax) {
#pragma unroll
            for (int j = 0; j < DMAX; j++) ntry[j] += (j == i);
          } else {
            ntry[i]++;
          }
          aset(prop, i, aget(cur, i));
          aset(sig, i, max_m(0.0, si - gam * alphastar));
        }
      }
    }
    if (remain < 10000 * d && remain % 10 == 0) {  // :642-647
      if (stored < nstore) {
#pragma unroll
        for (int j = 0; j < DMAX; j++)
          if (j < d) out[stored * d + j] = cur[j];
      }
      stored++;
    }
    if (sweep % 100 == 0) {  // :648-655
      if (id == 0 && a.sig_trace0 != nullptr) {
#pragma unroll
        for (int j = 0; j < DMAX; j++)
          if (j < d) {
            a.sig_trace0[(size_t)ntrace * d + j] = sig[j];
            a.acc_trace0[(size_t)ntrace * d + j] = (double)nacc[j] / (double)ntry[j];
          }
      }
      ntrace++;
    }
  }
#pragma unroll
  for (int j = 0; j < DMAX; j++)
    if (j < d) a.sig_out[(size_t)id * d + j] = sig[j];
  int status = 0;
  if (u.overrun()) status |= 1;
  if (lp != lp) status |= 2;
  if (status) atomicOr(a.status, status);
}


// ---- speculative form: one WARP per chain, exact semantics ------------------------------------------
// A single chain is a sequence of accept/reject decisions, each needing one log-posterior evaluation of a
// proposal that depends on the decisions before it -- on a GPU a thread walking that sequence is bound by the
// latency of one evaluation after another (coal-mining, d = 13: 12.5 s for the reference's schedule).  But the
// uniforms a step consumes do not depend on the decisions (with Gaussian proposals every coordinate step takes
// exactly three), so the proposal noise of the next steps is known in advance and the only unknown is the
// accept/reject path.  The warp therefore evaluates the whole decision tree of the next five coordinate steps
// at once: lane (2^t - 1) + p owns the node at depth t reached by the accept pattern p of the steps before it,
// replays that prefix without evaluating anything (the proposals and the adapted scales along it are cheap),
// and evaluates the log-posterior of its own proposal.  Every node then knows the log-posterior of its current
// state (its last accepted ancestor's, by shuffle) and decides; five shuffles walk the true path, and every lane
// commits it.  31 evaluations run in the time of one and five steps retire per round, with the arithmetic of
// the sequential kernel operation for operation, so the chain is the same chain bit for bit (the tests run both
// on the same injected tape).  Block-move sweeps (a tenth of the sweeps after burn-in) are taken one at a time.
constexpr int kSpecDepth = 5;

template <int DMAX, class TGT, class RNG>
__global__ void __launch_bounds__(32) rwm_spec_kernel(RwmArgs a) {
  TGT T;
  T.bind(a.tgt_blob, a.tgt_flags);
  const long id = blockIdx.x;
  const int lane = threadIdx.x;
  const unsigned full = 0xffffffffu;
  const int d = a.d, k = a.model_k;
  RNG u;
  if constexpr (std::is_same<RNG, TapeStream>::value) u.open(a.tape, a.tape_stride, (unsigned long long)id, 0ull);
  else u.open(a.seed, (unsigned long long)id, 0ull);

  double cur[DMAX], sig[DMAX];
  double xc[DMAX], sg[DMAX];  // this lane's speculative copy of the state and the scales (one scratch for every use)
  int nacc[DMAX], ntry[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; i++) {
    cur[i] = xc[i] = (i < d) ? a.init[i] : 0.0;
    sig[i] = sg[i] = 10.0;  // :595
    nacc[i] = ntry[i] = 0;
  }
  double lp = T.template eval<DMAX>(k, xc);
  const long nstore = 1000L * d;
  double *out = a.samples_out + (size_t)id * nstore * d;
  long stored = 0;
  int ntrace = 0;
  const double alphastar = 0.25;
  int status = 0;

  // the node of the decision tree this lane evaluates: depth and the accept pattern of the steps above it
  const int t_me = 31 - __clz(lane + 1);
  const int p_me = lane + 1 - (1 << t_me);

  auto end_of_sweep = [&](int sweep) {  // :642-655
    const int remain = a.nsweepr - sweep;
    if (remain < 10000 * d && remain % 10 == 0) {
      if (stored < nstore && lane == 0)
        for (int q = 0; q < d; q++) out[stored * d + q] = cur[q];
      stored++;
    }
    if (sweep % 100 == 0) {
      if (id == 0 && lane == 0 && a.sig_trace0 != nullptr)
        for (int q = 0; q < d; q++) {
          a.sig_trace0[(size_t)ntrace * d + q] = sig[q];
          a.acc_trace0[(size_t)ntrace * d + q] = (double)nacc[q] / (double)ntry[q];
        }
      ntrace++;
    }
  };

  unsigned long long n = 0;  // index of the next uniform of the chain's stream
  int sweep = 1, j = 0;      // next step: coordinate j of `sweep`; j == 0 <=> the sweep's first uniform is not drawn yet
  while (sweep <= a.nsweepr) {
    if (j == 0) {
      const double uu = u.at(n);
      n++;
      if (sweep > a.nburn && uu < 0.1) {  // block move (:606-617): every lane does the same work
        const int npairs = d >> 1;
        double za = 0.0, zb = 0.0;
        if (lane < npairs) {  // gauss_pair
          const double ua = u.at(n + 2 * lane), ub = u.at(n + 2 * lane + 1);
          const double r = sqrt(-2.0 * log(ua));
          double sn, cs;
          sincos(6.283185307179586476925 * ub, &sn, &cs);
          za = r * sn;
          zb = r * cs;
        } else if (lane == npairs && (d & 1)) {  // gauss_single
          const double ua = u.at(n + 2 * lane), ub = u.at(n + 2 * lane + 1);
          za = sqrt(-2.0 * log(ua)) * sin(6.283185307179586476925 * ub);
        }
#pragma unroll
        for (int i = 0; i < DMAX; i++) xc[i] = cur[i];
        for (int i = 0; i < d; i++) {
          const double z = __shfl_sync(full, (i & 1) ? zb : za, i >> 1);
          xc[i] = fma(sig[i], z, cur[i]);
        }
        n += 2ull * (unsigned long long)((d + 1) >> 1);
        const double lpn = T.template eval<DMAX>(k, xc);
        const double uacc = u.at(n);
        n++;
        if (uacc < mh_prob(lpn - lp)) {
#pragma unroll
          for (int i = 0; i < DMAX; i++) cur[i] = xc[i];
          lp = lpn;
        } else {
#pragma unroll
          for (int i = 0; i < DMAX; i++) xc[i] = cur[i];
        }
        end_of_sweep(sweep);
        sweep++;
        continue;
      }
    }
    // The window: up to five coordinate steps from (sweep, j); it may run into the following sweeps, and stops in
    // front of a block-move sweep (whose first uniform is then read again at the top) and at the end of the
    // schedule.  Per step: coordinate, sweep, adaptation gain and position in the uniform stream.  The loops over
    // the window are fully unrolled so that these small arrays are registers.
    int cq[kSpecDepth], sq[kSpecDepth];
    unsigned long long nq[kSpecDepth];
    double gq[kSpecDepth];
    int m = 0;
    {
      int s2 = sweep, j2 = j;
      unsigned long long n2 = n;
      bool open = true;
#pragma unroll
      for (int t = 0; t < kSpecDepth; t++) {
        if (open && j2 == 0 && t > 0) {
          if (s2 > a.nsweepr) {
            open = false;
          } else {
            const double uu2 = u.at(n2);
            if (s2 > a.nburn && uu2 < 0.1) open = false;
            else n2++;
          }
        }
        cq[t] = j2;
        sq[t] = s2;
        nq[t] = n2;
        gq[t] = 0.0;
        if (open) {
          gq[t] = a.gtab[s2 - 1];
          n2 += 3;
          m = t + 1;
          if (++j2 == d) {
            j2 = 0;
            s2++;
          }
        }
      }
      sweep = s2;  // the chain's position after the window
      j = j2;
      n = n2;
    }
    auto pick_n = [&](int t) {
      unsigned long long r = nq[0];
#pragma unroll
      for (int q = 1; q < kSpecDepth; q++) r = (t == q) ? nq[q] : r;
      return r;
    };

    // proposal noise of the window: lane t draws step t's variate (gauss(), :1639-1661), everyone gets all of them
    double zmine = 0.0;
    if (lane < m) {
      const unsigned long long nt = pick_n(lane);
      const double ua = u.at(nt), ub = u.at(nt + 1);
      zmine = sqrt(-2.0 * log(ua)) * sin(6.283185307179586476925 * ub);
    }
    double z[kSpecDepth];
#pragma unroll
    for (int t = 0; t < kSpecDepth; t++) z[t] = __shfl_sync(full, zmine, t);

    // replay the prefix of this lane's node on its mirror of the state (xc, sg), evaluate its proposal, and put
    // the touched entries back
    const bool node = t_me < m;
    double lpn = 0.0;
    if (node) {
#pragma unroll
      for (int q = 0; q < kSpecDepth; q++) {
        if (q < t_me) {
          const int c = cq[q];
          const double si = aget(sg, c);
          if ((p_me >> q) & 1) {
            aset(xc, c, fma(si, z[q], aget(xc, c)));
            aset(sg, c, max_m(0.0, si - gq[q] * (alphastar - 1.0)));
          } else {
            aset(sg, c, max_m(0.0, si - gq[q] * alphastar));
          }
        } else if (q == t_me) {
          const int c = cq[q];
          aset(xc, c, fma(aget(sg, c), z[q], aget(xc, c)));
        }
      }
      lpn = T.template eval<DMAX>(k, xc);
#pragma unroll
      for (int q = 0; q < kSpecDepth; q++)
        if (q <= t_me) {
          const int c = cq[q];
          aset(xc, c, aget(cur, c));
          aset(sg, c, aget(sig, c));
        }
    }
    // the log-posterior of the node's current state: its last accepted ancestor's proposal, else the chain's
    int src = lane;
    if (p_me != 0) {
      const int jh = 31 - __clz(p_me);
      src = (1 << jh) - 1 + (p_me & ((1 << jh) - 1));
    }
    const double lpa = __shfl_sync(full, lpn, src);
    const double lpc = (p_me != 0) ? lpa : lp;
    int dec = 0;
    if (node) {
      const double uacc = u.at(pick_n(t_me) + 2);
      const double acc = min_m(1.0, mh_prob(lpn - lpc));  // :627
      dec = (uacc < acc) ? 1 : 0;
    }
    // walk the true path
    int path = 0;
#pragma unroll
    for (int t = 0; t < kSpecDepth; t++)
      if (t < m) path |= __shfl_sync(full, dec, (1 << t) - 1 + path) << t;
    // commit it: every lane applies the same m steps to the state and to its mirror (:628-640)
#pragma unroll
    for (int t = 0; t < kSpecDepth; t++) {
      if (t < m) {
        const int c = cq[t];
        const double si = aget(sig, c);
        const double lpt = __shfl_sync(full, lpn, (1 << t) - 1 + (path & ((1 << t) - 1)));
        double sn;
        if ((path >> t) & 1) {
          nacc[c]++;
          ntry[c]++;
          const double xn = fma(si, z[t], aget(cur, c));
          aset(cur, c, xn);
          aset(xc, c, xn);
          lp = lpt;
          sn = max_m(0.0, si - gq[t] * (alphastar - 1.0));
        } else {
          ntry[c]++;
          sn = max_m(0.0, si - gq[t] * alphastar);
        }
        aset(sig, c, sn);
        aset(sg, c, sn);
        if (c == d - 1) end_of_sweep(sq[t]);
      }
    }
  }
  if (lane == 0)
    for (int q = 0; q < d; q++) a.sig_out[(size_t)id * d + q] = sig[q];
  if (u.overrun()) status |= 1;
  if (lp != lp) status |= 2;
  if (status) atomicOr(a.status, status);
}


// ---- split form for HOST log-posterior callbacks ---------------------------------------------------
// Same chain, cut at every log-posterior evaluation: kernel j of a sweep finishes proposal j-1 with
// the value the host returned and makes proposal j.  State lives in global memory between kernels.
struct RwmSplit {
  double *cur, *prop, *sig;  // [C][d] chain-major (prop is what the host callback reads)
  int *nacc, *ntry;          // [C][d]
  double *lp, *lpn;          // [C]
  int *mode, *keval;         // [C] mode: 1 = block move this sweep
  long *stored;              // [C]
  unsigned long long *draws; // [C]
};

template <class RNG>
__global__ void __launch_bounds__(kRwmThreads) rwm_split_kernel(RwmArgs a, RwmSplit sp, int sweep, int j) {
  const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.nchains) return;
  const int d = a.d, k = a.model_k;
  RNG u;
  if constexpr (std::is_same<RNG, TapeStream>::value) u.open(a.tape, a.tape_stride, (unsigned long long)id, sp.draws[id]);
  else u.open(a.seed, (unsigned long long)id, sp.draws[id]);
  double *cur = sp.cur + id * d, *prop = sp.prop + id * d, *sig = sp.sig + id * d;
  int *nacc = sp.nacc + id * d, *ntry = sp.ntry + id * d;
  const double alphastar = 0.25;
  int keval = -1;
  if (sweep == 0) {  // chain start: the host evaluates the start point (:599)
    if (j == 0) {
      for (int i = 0; i < d; i++) {
        cur[i] = prop[i] = a.init[i];
        sig[i] = 10.0;
        nacc[i] = ntry[i] = 0;
      }
      sp.stored[id] = 0;
      sp.mode[id] = 0;
      keval = k;
    } else {
      sp.lp[id] = sp.lpn[id];
    }
    sp.keval[id] = keval;
    return;
  }
  double lp = sp.lp[id];
  const double lpn = sp.lpn[id];
  int mode = sp.mode[id];
  // ---- finish the proposal made by the previous kernel of this sweep
  if (j > 0) {
    if (mode == 1) {
      if (j == 1) {
        if (u.next() < mh_prob(lpn - lp)) {
          for (int i = 0; i < d; i++) cur[i] = prop[i];
          lp = lpn;
        } else {
          for (int i = 0; i < d; i++) prop[i] = cur[i];
        }
      }
    } else {
      const int i = j - 1;
      const double gam = a.gtab[sweep - 1];
      const double acc = min_m(1.0, mh_prob(lpn - lp));
      if (u.next() < acc) {
        nacc[i]++;
        ntry[i]++;
        cur[i] = prop[i];
        lp = lpn;
        sig[i] = max_m(0.0, sig[i] - gam * (alphastar - 1.0));
      } else {
        ntry[i]++;
        prop[i] = cur[i];
        sig[i] = max_m(0.0, sig[i] - gam * alphastar);
      }
    }
  }
  // ---- make the next proposal, or close the sweep
  if (j == 0) {
    const double uu = u.next();
    mode = (sweep > a.nburn && uu < 0.1) ? 1 : 0;
    if (mode == 1) {
      int i = 0;
      for (; i + 1 < d; i += 2) {
        double z0, z1;
        gauss_pair(u, z0, z1);
        prop[i] = z0;
        prop[i + 1] = z1;
      }
      if (d & 1) prop[d - 1] = gauss_single(u);
      const double den = a.dof > 0 ? t_divisor(a.dof, u) : 1.0;
      for (int q = 0; q < d; q++) prop[q] = fma(sig[q], a.dof > 0 ? prop[q] / den : prop[q], cur[q]);
    } else {
      double z = gauss_single(u);
      if (a.dof > 0) z /= t_divisor(a.dof, u);
      prop[0] = fma(sig[0], z, cur[0]);
    }
    keval = k;
  } else if (j < d) {
    if (mode == 0) {
      double z = gauss_single(u);
      if (a.dof > 0) z /= t_divisor(a.dof, u);
      prop[j] = fma(sig[j], z, cur[j]);
      keval = k;
    }
  } else {  // j == d: end of sweep (:642-655)
    const int remain = a.nsweepr - sweep;
    if (remain < 10000 * d && remain % 10 == 0) {
      const long st = sp.stored[id];
      if (st < 1000L * d)
        for (int i = 0; i < d; i++) a.samples_out[((size_t)id * 1000 * d + st) * d + i] = cur[i];
      sp.stored[id] = st + 1;
    }
    if (sweep % 100 == 0 && id == 0 && a.sig_trace0 != nullptr) {
      const int row = sweep / 100 - 1;
      for (int i = 0; i < d; i++) {
        a.sig_trace0[(size_t)row * d + i] = sig[i];
        a.acc_trace0[(size_t)row * d + i] = (double)nacc[i] / (double)ntry[i];
      }
    }
    if (sweep == a.nsweepr)
      for (int i = 0; i < d; i++) a.sig_out[(size_t)id * d + i] = sig[i];
  }
  sp.lp[id] = lp;
  sp.mode[id] = mode;
  sp.keval[id] = keval;
  sp.draws[id] = u.n;
  int status = (u.overrun() ? 1 : 0) | ((lp != lp) ? 2 : 0);
  if (status) atomicOr(a.status, status);
}

// ---- persistent form for HOST callbacks (amx_mailbox.cuh) -------------------------------------------------------------
// The chain of rwm_split_kernel kept in one thread for its whole schedule; every log-posterior value comes from the
// host through the CTA's mailbox.  Every sweep makes exactly d exchanges (a block-move sweep uses the first and sits
// the others out), plus one for the start point, so that the host knows each CTA's count.
template <class RNG>
__global__ void __launch_bounds__(kMbThreads) rwm_mailbox_kernel(RwmArgs a, void *mb_base, int ldx) {
  constexpr int DM = AMX_MAX_DIM;
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = gid < a.nchains;
  const long id = active ? gid : a.nchains - 1;
  const int d = a.d, k = active ? a.model_k : -1;
  Mailbox *mb = mailbox_at(mb_base, blockIdx.x, ldx);
  unsigned seq = 0;
  RNG u;
  if constexpr (std::is_same<RNG, TapeStream>::value) u.open(a.tape, a.tape_stride, (unsigned long long)id, 0ull);
  else u.open(a.seed, (unsigned long long)id, 0ull);
  double cur[DM], prop[DM], sig[DM];
  int nacc[DM], ntry[DM];
  for (int i = 0; i < DM; i++) {
    cur[i] = prop[i] = (i < d) ? a.init[i] : 0.0;
    sig[i] = 10.0;  // :595
    nacc[i] = ntry[i] = 0;
  }
  const double alphastar = 0.25;
  double lp = mailbox_exchange<DM>(mb, ldx, ++seq, k, prop, d);  // :599
  long stored = 0;
  for (int sweep = 1; sweep <= a.nsweepr; sweep++) {
    const double uu = u.next();
    if (sweep > a.nburn && uu < 0.1) {  // block move, no adaptation (:606-617)
      int i = 0;
      for (; i + 1 < d; i += 2) {
        double z0, z1;
        gauss_pair(u, z0, z1);
        prop[i] = z0;
        prop[i + 1] = z1;
      }
      if (d & 1) prop[d - 1] = gauss_single(u);
      const double den = a.dof > 0 ? t_divisor(a.dof, u) : 1.0;
      for (int q = 0; q < d; q++) prop[q] = fma(sig[q], a.dof > 0 ? prop[q] / den : prop[q], cur[q]);
      const double lpn = mailbox_exchange<DM>(mb, ldx, ++seq, k, prop, d);
      if (u.next() < mh_prob(lpn - lp)) {
        for (int q = 0; q < d; q++) cur[q] = prop[q];
        lp = lpn;
      } else {
        for (int q = 0; q < d; q++) prop[q] = cur[q];
      }
      for (int j = 1; j < d; j++) mailbox_exchange<DM>(mb, ldx, ++seq, -1, prop, d);
    } else {  // one coordinate after the other, scales adapted towards 25 % acceptance (:619-640)
      const double gam = a.gtab[sweep - 1];
      for (int i = 0; i < d; i++) {
        double z = gauss_single(u);
        if (a.dof > 0) z /= t_divisor(a.dof, u);
        prop[i] = fma(sig[i], z, cur[i]);
        const double lpn = mailbox_exchange<DM>(mb, ldx, ++seq, k, prop, d);
        const double acc = min_m(1.0, mh_prob(lpn - lp));
        if (u.next() < acc) {
          nacc[i]++;
          ntry[i]++;
          cur[i] = prop[i];
          lp = lpn;
          sig[i] = max_m(0.0, sig[i] - gam * (alphastar - 1.0));
        } else {
          ntry[i]++;
          prop[i] = cur[i];
          sig[i] = max_m(0.0, sig[i] - gam * alphastar);
        }
      }
    }
    const int remain = a.nsweepr - sweep;  // :642-655
    if (active && remain < 10000 * d && remain % 10 == 0) {
      if (stored < 1000L * d)
        for (int i = 0; i < d; i++) a.samples_out[((size_t)id * 1000 * d + stored) * d + i] = cur[i];
      stored++;
    }
    if (sweep % 100 == 0 && gid == 0 && a.sig_trace0 != nullptr) {
      const int row = sweep / 100 - 1;
      for (int i = 0; i < d; i++) {
        a.sig_trace0[(size_t)row * d + i] = sig[i];
        a.acc_trace0[(size_t)row * d + i] = (double)nacc[i] / (double)ntry[i];
      }
    }
  }
  if (active) {
    for (int i = 0; i < d; i++) a.sig_out[(size_t)id * d + i] = sig[i];
    const int status = (u.overrun() ? 1 : 0) | ((lp != lp) ? 2 : 0);
    if (status) atomicOr(a.status, status);
  }
}

// Few chains: latency is what matters, give each chain a warp (speculative kernel).  Many chains (pooled
// stage-1 populations): throughput matters, one thread per chain.  Student-t proposals draw a data-dependent
// number of uniforms per step (rgamma's rejection loop), which the look-ahead cannot index: sequential kernel.
inline bool rwm_use_spec(const RwmArgs &a) {
  const char *e = getenv("AMX_RWM_SPEC");
  if (e) return atoi(e) != 0 && a.dof == 0;
  return a.dof == 0 && a.nchains <= 1024;
}

template <class TGT, class RNG>
inline int rwm_launch_d(const RwmArgs &a) {
  if (rwm_use_spec(a)) {
    const unsigned g = (unsigned)a.nchains;
    if constexpr (TargetIsWide<TGT>::value) {
      rwm_spec_kernel<AMX_MAX_DIM, TGT, RNG><<<g, 32, 0, stream()>>>(a);
    } else {
      if (a.d <= 2) rwm_spec_kernel<2, TGT, RNG><<<g, 32, 0, stream()>>>(a);
      else if (a.d <= 8) rwm_spec_kernel<8, TGT, RNG><<<g, 32, 0, stream()>>>(a);
      else rwm_spec_kernel<AMX_MAX_DIM, TGT, RNG><<<g, 32, 0, stream()>>>(a);
    }
    count_launch();
    AMX_CUDA(cudaGetLastError());
    return AMX_OK;
  }
  const unsigned grid = (unsigned)((a.nchains + kRwmThreads - 1) / kRwmThreads);
  if constexpr (TargetIsWide<TGT>::value) {
    rwm_adapt_kernel<AMX_MAX_DIM, TGT, RNG><<<grid, kRwmThreads, 0, stream()>>>(a);
  } else {
    if (a.d <= 2) rwm_adapt_kernel<2, TGT, RNG><<<grid, kRwmThreads, 0, stream()>>>(a);
    else if (a.d <= 8) rwm_adapt_kernel<8, TGT, RNG><<<grid, kRwmThreads, 0, stream()>>>(a);
    else rwm_adapt_kernel<AMX_MAX_DIM, TGT, RNG><<<grid, kRwmThreads, 0, stream()>>>(a);
  }
  count_launch();
  AMX_CUDA(cudaGetLastError());
  return AMX_OK;
}

}  // namespace amx
