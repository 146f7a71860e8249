// amx_rwm.cu -- K1: stage-1 adaptive random-walk Metropolis within one model.
// Replaces rwm_within_model (automix.c:575-662).
//
// One thread per chain; every chain runs the reference's full adaptive schedule
// (1.1 * max(nsweep2, 10000 d) sweeps, Robbins-Monro scale adaptation towards 25 % acceptance,
// 10 % block moves after the first tenth, every 10th state of the last 10000 d sweeps stored),
// so chain 0 fed an injected tape IS the reference chain, step for step.  The chains differ only
// in their counter-based RNG stream; a whole population advances in the wall time of one chain,
// which is what stage 2 needs when it pools the stored samples of many chains.
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
static bool rwm_use_spec(const RwmArgs &a) {
  const char *e = getenv("AMX_RWM_SPEC");
  if (e) return atoi(e) != 0 && a.dof == 0;
  return a.dof == 0 && a.nchains <= 1024;
}

template <class TGT, class RNG>
static int rwm_launch_d(const RwmArgs &a) {
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

template <class RNG>
static int rwm_launch(const amx_target *t, const RwmArgs &a) {
  switch (t->d.kind) {
    case kTargetGaussMix: return rwm_launch_d<GaussMixTarget, RNG>(a);
    case kTargetQuad: return rwm_launch_d<QuadTarget, RNG>(a);
    case kTargetCoal: return rwm_launch_d<CoalTarget, RNG>(a);
    case kTargetMixNorm: return rwm_launch_d<MixNormTarget, RNG>(a);
  }
  return fail(AMX_EINVAL, "plug-in kind %d has no device RWM kernel", t->d.kind);
}


static int rwm_host_run(const amx_target *t, RwmArgs &a) {
  const size_t C = (size_t)a.nchains;
  const int d = a.d;
  RwmSplit sp;
  memset(&sp, 0, sizeof(sp));
  AMX_CUDA(cudaMalloc(&sp.cur, sizeof(double) * C * d));
  AMX_CUDA(cudaMalloc(&sp.prop, sizeof(double) * C * d));
  AMX_CUDA(cudaMalloc(&sp.sig, sizeof(double) * C * d));
  AMX_CUDA(cudaMalloc(&sp.nacc, sizeof(int) * C * d));
  AMX_CUDA(cudaMalloc(&sp.ntry, sizeof(int) * C * d));
  AMX_CUDA(cudaMalloc(&sp.lp, sizeof(double) * C));
  AMX_CUDA(cudaMalloc(&sp.lpn, sizeof(double) * C));
  AMX_CUDA(cudaMalloc(&sp.mode, sizeof(int) * C));
  AMX_CUDA(cudaMalloc(&sp.keval, sizeof(int) * C));
  AMX_CUDA(cudaMalloc(&sp.stored, sizeof(long) * C));
  AMX_CUDA(cudaMalloc(&sp.draws, sizeof(unsigned long long) * C));
  AMX_CUDA(cudaMemsetAsync(sp.draws, 0, sizeof(unsigned long long) * C, stream()));
  AMX_CUDA(cudaMemsetAsync(sp.lpn, 0, sizeof(double) * C, stream()));
  double *h_prop = nullptr, *h_lpn = nullptr;
  int *h_keval = nullptr;
  AMX_CUDA(cudaMallocHost(&h_prop, sizeof(double) * C * d));
  AMX_CUDA(cudaMallocHost(&h_lpn, sizeof(double) * C));
  AMX_CUDA(cudaMallocHost(&h_keval, sizeof(int) * C));
  std::vector<int> kc;
  std::vector<double> xc, lc;
  const unsigned grid = (unsigned)((C + kRwmThreads - 1) / kRwmThreads);
  auto launch = [&](int sweep, int j) -> int {
    if (a.tape) rwm_split_kernel<TapeStream><<<grid, kRwmThreads, 0, stream()>>>(a, sp, sweep, j);
    else rwm_split_kernel<PhiloxStream><<<grid, kRwmThreads, 0, stream()>>>(a, sp, sweep, j);
    count_launch();
    AMX_CUDA(cudaGetLastError());
    return AMX_OK;
  };
  auto evaluate = [&]() -> int {
    AMX_CUDA(cudaMemcpyAsync(h_prop, sp.prop, sizeof(double) * C * d, cudaMemcpyDeviceToHost, stream()));
    AMX_CUDA(cudaMemcpyAsync(h_keval, sp.keval, sizeof(int) * C, cudaMemcpyDeviceToHost, stream()));
    AMX_CUDA(cudaStreamSynchronize(stream()));
    if (t->d.kind == kTargetHostScalar) {
      for (size_t c = 0; c < C; c++)
        if (h_keval[c] >= 0) h_lpn[c] = t->d.scalar(h_keval[c], h_prop + c * d);
    } else {
      kc.clear();
      xc.clear();
      for (size_t c = 0; c < C; c++)
        if (h_keval[c] >= 0) {
          kc.push_back(h_keval[c]);
          xc.insert(xc.end(), h_prop + c * d, h_prop + (c + 1) * d);
        }
      lc.resize(kc.size());
      if (!kc.empty()) t->d.batched((long)kc.size(), kc.data(), xc.data(), d, lc.data(), t->d.user);
      size_t q = 0;
      for (size_t c = 0; c < C; c++)
        if (h_keval[c] >= 0) h_lpn[c] = lc[q++];
    }
    AMX_CUDA(cudaMemcpyAsync(sp.lpn, h_lpn, sizeof(double) * C, cudaMemcpyHostToDevice, stream()));
    return AMX_OK;
  };
  int rc = 0;
  if ((rc = launch(0, 0)) || (rc = evaluate()) || (rc = launch(0, 1))) return rc;
  for (int sweep = 1; sweep <= a.nsweepr && !rc; sweep++) {
    for (int j = 0; j < d && !rc; j++) {
      rc = launch(sweep, j);
      if (!rc) rc = evaluate();
    }
    if (!rc) rc = launch(sweep, d);
  }
  AMX_CUDA(cudaStreamSynchronize(stream()));
  cudaFree(sp.cur); cudaFree(sp.prop); cudaFree(sp.sig); cudaFree(sp.nacc); cudaFree(sp.ntry); cudaFree(sp.lp);
  cudaFree(sp.lpn); cudaFree(sp.mode); cudaFree(sp.keval); cudaFree(sp.stored); cudaFree(sp.draws);
  cudaFreeHost(h_prop); cudaFreeHost(h_lpn); cudaFreeHost(h_keval);
  return rc;
}

}  // namespace amx

using namespace amx;

// One stage-1 run = allocate + enqueue (start) and wait + read back (finish); splitting the two lets
// the runs of all models be in flight at once on separate streams (amx_rwm_adapt_all).
static int g_rwm_dof = 0;  // process-wide, like the reference's sampler flags (amSampler.student_T_dof)

struct RwmJob {
  RwmArgs a;
  double *init_dev, *gtab, *tape_dev, *sig_dev, *samp_dev, *tr_dev;
  int *status_dev;
  cudaEvent_t e0, e1;
  long nstore;
  int ntr;
  bool host_target;
  // host callbacks through the mailbox: the kernel is in flight on `st` and must be served (mailbox_serve)
  void *mb_host;
  int mb_ncta, mb_nthr;
  long mb_nexch;
  cudaStream_t st;
};
static bool rwm_use_mailbox() {
  const char *e = getenv("AMX_HOST_MAILBOX");
  return !(e && atoi(e) == 0);
}

static int rwm_job_start(const amx_target *t, int model_k, int nsweep2, long nchains, const double *init,
                         uint64_t seed, const double *tape, long tape_stride, RwmJob &J) {
  memset(&J, 0, sizeof(J));
  const int d = t->d.dims[model_k];
  RwmArgs &a = J.a;
  int nsw = nsweep2 > 10000 * d ? nsweep2 : 10000 * d;  // :579-582
  a.nburn = nsw / 10;
  a.nsweepr = nsw + a.nburn;
  a.tgt_blob = t->d.blob_dev;
  a.tgt_flags = t->d.flags;
  a.model_k = model_k;
  a.d = d;
  a.nchains = nchains;
  a.seed = seed;
  a.dof = g_rwm_dof;
  J.nstore = 1000L * d;
  J.ntr = a.nsweepr / 100;
  J.host_target = (t->d.kind == kTargetHostScalar || t->d.kind == kTargetHostBatched);
  AMX_CUDA(cudaMalloc(&J.init_dev, sizeof(double) * d));
  AMX_CUDA(cudaMemcpyAsync(J.init_dev, init, sizeof(double) * d, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaMalloc(&J.gtab, sizeof(double) * a.nsweepr));
  AMX_CUDA(cudaMalloc(&J.sig_dev, sizeof(double) * (size_t)nchains * d));
  AMX_CUDA(cudaMalloc(&J.samp_dev, sizeof(double) * (size_t)nchains * J.nstore * d));
  AMX_CUDA(cudaMalloc(&J.tr_dev, sizeof(double) * (size_t)2 * (J.ntr + 1) * d));
  AMX_CUDA(cudaMalloc(&J.status_dev, sizeof(int)));
  AMX_CUDA(cudaMemsetAsync(J.status_dev, 0, sizeof(int), stream()));
  if (tape) {
    AMX_CUDA(cudaMalloc(&J.tape_dev, sizeof(double) * (size_t)tape_stride * nchains));
    AMX_CUDA(cudaMemcpyAsync(J.tape_dev, tape, sizeof(double) * (size_t)tape_stride * nchains, cudaMemcpyHostToDevice, stream()));
  }
  a.init = J.init_dev;
  a.gtab = J.gtab;
  a.tape = J.tape_dev;
  a.tape_stride = (unsigned long long)tape_stride;
  a.sig_out = J.sig_dev;
  a.samples_out = J.samp_dev;
  a.sig_trace0 = J.tr_dev;
  a.acc_trace0 = J.tr_dev + (size_t)(J.ntr + 1) * d;
  a.status = J.status_dev;
  rwm_gamma_kernel<<<(a.nsweepr + 255) / 256, 256, 0, stream()>>>(J.gtab, a.nsweepr);
  count_launch();
  AMX_CUDA(cudaEventCreate(&J.e0));
  AMX_CUDA(cudaEventCreate(&J.e1));
  AMX_CUDA(cudaEventRecord(J.e0, stream()));
  int rc = AMX_OK;
  J.st = stream();
  if (J.host_target && rwm_use_mailbox()) {
    J.mb_nthr = nchains >= kMbThreads ? kMbThreads : (int)((nchains + 31) / 32 * 32);
    J.mb_ncta = (int)((nchains + J.mb_nthr - 1) / J.mb_nthr);
    J.mb_nexch = 1 + (long)a.nsweepr * d;
    if ((rc = mailbox_alloc(&J.mb_host, J.mb_ncta, d))) return rc;
    if (tape) rwm_mailbox_kernel<TapeStream><<<J.mb_ncta, J.mb_nthr, 0, stream()>>>(a, J.mb_host, d);
    else rwm_mailbox_kernel<PhiloxStream><<<J.mb_ncta, J.mb_nthr, 0, stream()>>>(a, J.mb_host, d);
    count_launch();
    AMX_CUDA(cudaGetLastError());
  } else if (J.host_target) {
    rc = rwm_host_run(t, a);
  } else {
    rc = tape ? rwm_launch<TapeStream>(t, a) : rwm_launch<PhiloxStream>(t, a);
  }
  if (rc) return rc;
  AMX_CUDA(cudaEventRecord(J.e1, stream()));
  return AMX_OK;
}

// answer the mailboxes of the jobs whose kernels wait for host log-posterior values (all of them at once: the models'
// chains advance together, the callbacks run on this thread)
static int rwm_jobs_serve(const amx_target *t, RwmJob *jobs, int n) {
  std::vector<MbJob> mj;
  for (int i = 0; i < n; i++)
    if (jobs[i].mb_host) mj.push_back({jobs[i].mb_host, jobs[i].mb_ncta, jobs[i].a.d, jobs[i].mb_nthr, jobs[i].mb_nexch, jobs[i].st, 0u});
  if (mj.empty()) return AMX_OK;
  return mailbox_serve(mj, t->d);
}

static int rwm_job_finish(RwmJob &J, double *sig_out, double *samples_out, double *sig_trace0, double *acc_trace0,
                          double *kernel_ms) {
  const int d = J.a.d;
  const long nchains = J.a.nchains;
  AMX_CUDA(cudaEventSynchronize(J.e1));
  float ms = 0;
  AMX_CUDA(cudaEventElapsedTime(&ms, J.e0, J.e1));
  if (kernel_ms) *kernel_ms = ms;
  int status = 0;
  AMX_CUDA(cudaMemcpy(&status, J.status_dev, sizeof(int), cudaMemcpyDeviceToHost));
  AMX_CUDA(cudaMemcpy(sig_out, J.sig_dev, sizeof(double) * (size_t)nchains * d, cudaMemcpyDeviceToHost));
  AMX_CUDA(cudaMemcpy(samples_out, J.samp_dev, sizeof(double) * (size_t)nchains * J.nstore * d, cudaMemcpyDeviceToHost));
  if (sig_trace0) AMX_CUDA(cudaMemcpy(sig_trace0, J.a.sig_trace0, sizeof(double) * (size_t)J.ntr * d, cudaMemcpyDeviceToHost));
  if (acc_trace0) AMX_CUDA(cudaMemcpy(acc_trace0, J.a.acc_trace0, sizeof(double) * (size_t)J.ntr * d, cudaMemcpyDeviceToHost));
  cudaEventDestroy(J.e0);
  cudaEventDestroy(J.e1);
  cudaFree(J.init_dev); cudaFree(J.gtab); cudaFree(J.tape_dev); cudaFree(J.sig_dev); cudaFree(J.samp_dev);
  cudaFree(J.tr_dev); cudaFree(J.status_dev);
  if (J.mb_host) cudaFreeHost(J.mb_host);
  if (status & 1) return fail(AMX_ETAPE, "injected uniform tape exhausted");
  if (status & 2) return fail(AMX_ENUMERIC, "a chain reached a NaN log-posterior");
  return AMX_OK;
}

extern "C" int amx_rwm_set_dof(int student_t_dof) {
  if (student_t_dof < 0) return fail(AMX_EINVAL, "negative degrees of freedom");
  g_rwm_dof = student_t_dof;
  return AMX_OK;
}

extern "C" int amx_rwm_adapt(const amx_target *t, int model_k, int nsweep2, long nchains, const double *init,
                             uint64_t seed, const double *tape, long tape_stride, double *sig_out,
                             double *samples_out, double *sig_trace0, double *acc_trace0, double *kernel_ms) {
  if (int rc = require_device()) return rc;
  if (!t || model_k < 0 || model_k >= t->d.nmodels || nchains < 1 || nsweep2 < 1 || !init || !sig_out || !samples_out)
    return fail(AMX_EINVAL, "amx_rwm_adapt: bad arguments");
  RwmJob J;
  if (int rc = rwm_job_start(t, model_k, nsweep2, nchains, init, seed, tape, tape_stride, J)) return rc;
  if (int rc = rwm_jobs_serve(t, &J, 1)) return rc;
  return rwm_job_finish(J, sig_out, samples_out, sig_trace0, acc_trace0, kernel_ms);
}

// Stage 1 for every model at once: the models' chains are independent (the reference runs them one after
// another, automix.c:163-176), so their kernels are enqueued on separate streams and overlap on the GPU.
// init_flat / sig_out / samples_out are the per-model arrays concatenated in model order; sig_trace0 and
// acc_trace0 are arrays of nmodels pointers (entries may be NULL).  kernel_ms: wall time of the whole stage.
extern "C" int amx_rwm_adapt_all(const amx_target *t, int nsweep2, long nchains, const double *init_flat,
                                 uint64_t seed, double *sig_out, double *samples_out, double **sig_trace0,
                                 double **acc_trace0, double *kernel_ms) {
  if (int rc = require_device()) return rc;
  if (!t || nchains < 1 || nsweep2 < 1 || !init_flat || !sig_out || !samples_out)
    return fail(AMX_EINVAL, "amx_rwm_adapt_all: bad arguments");
  const int nm = t->d.nmodels;
  const bool host = (t->d.kind == kTargetHostScalar || t->d.kind == kTargetHostBatched);
  std::vector<RwmJob> jobs(nm);
  std::vector<cudaStream_t> streams(nm, nullptr);
  cudaStream_t saved = stream();
  cudaEvent_t w0, w1;
  AMX_CUDA(cudaEventCreate(&w0));
  AMX_CUDA(cudaEventCreate(&w1));
  AMX_CUDA(cudaEventRecord(w0, saved));
  int rc = AMX_OK;
  size_t off_i = 0, off_s = 0, off_x = 0;
  std::vector<size_t> oi(nm), os(nm), ox(nm);
  for (int k = 0; k < nm; k++) {
    const int d = t->d.dims[k];
    oi[k] = off_i; os[k] = off_s; ox[k] = off_x;
    off_i += d;
    off_s += (size_t)nchains * d;
    off_x += (size_t)nchains * 1000 * d * d;
  }
  if (host && !rwm_use_mailbox()) {  // kernel-per-evaluation path: one model after another
    double tot = 0.0;
    for (int k = 0; k < nm && rc == AMX_OK; k++) {
      double ms = 0.0;
      rc = amx_rwm_adapt(t, k, nsweep2, nchains, init_flat + oi[k], seed + 7919u * (uint64_t)k, nullptr, 0, sig_out + os[k],
                         samples_out + ox[k], sig_trace0 ? sig_trace0[k] : nullptr, acc_trace0 ? acc_trace0[k] : nullptr, &ms);
      tot += ms;
    }
    if (kernel_ms) *kernel_ms = tot;
    cudaEventDestroy(w0);
    cudaEventDestroy(w1);
    return rc;
  }
  int started = 0;
  for (int k = 0; k < nm && rc == AMX_OK; k++) {
    if (cudaStreamCreateWithFlags(&streams[k], cudaStreamNonBlocking) != cudaSuccess) {
      rc = fail(AMX_ECUDA, "stream creation failed");
      break;
    }
    cudaStreamWaitEvent(streams[k], w0, 0);
    amx_set_stream(streams[k]);
    rc = rwm_job_start(t, k, nsweep2, nchains, init_flat + oi[k], seed + 7919u * (uint64_t)k, nullptr, 0, jobs[k]);
    if (rc == AMX_OK) started++;
  }
  amx_set_stream(saved);
  if (rc == AMX_OK && host) rc = rwm_jobs_serve(t, jobs.data(), started);  // the models' kernels wait for values
  for (int k = 0; k < started; k++) {
    int r2 = rwm_job_finish(jobs[k], sig_out + os[k], samples_out + ox[k], sig_trace0 ? sig_trace0[k] : nullptr,
                            acc_trace0 ? acc_trace0[k] : nullptr, nullptr);
    if (rc == AMX_OK) rc = r2;
  }
  for (int k = 0; k < nm; k++)
    if (streams[k]) cudaStreamDestroy(streams[k]);
  AMX_CUDA(cudaEventRecord(w1, saved));
  AMX_CUDA(cudaEventSynchronize(w1));
  float ms = 0;
  cudaEventElapsedTime(&ms, w0, w1);
  if (kernel_ms) *kernel_ms = ms;
  cudaEventDestroy(w0);
  cudaEventDestroy(w1);
  return rc;
}
