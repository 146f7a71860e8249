// amx_rj.cuh -- K3: the reversible-jump sweep, one thread per chain, state on-chip for the
// whole launch.  Device-side restatement of reversible_jump_move (automix.c:1035-1288) and
// of the sweep loops of burn_samples / rjmcmc_samples (:77-155).
//
// The sweep is written as four phase functions so that the same code serves
//   * the fused kernel (device log-posterior plug-in: one launch = nsweeps sweeps), and
//   * the split kernels used with a host callback (propose -> host evaluates -> finish).
#pragma once

#include <type_traits>

#include "amx_common.cuh"
#include "amx_targets.cuh"

namespace amx {

template <int DMAX_, int LMAX_, int NMAX_>
struct RjCfg {
  static constexpr int DMAX = DMAX_, LMAX = LMAX_, NMAX = NMAX_;
};
using RjCfgS = RjCfg<2, 8, 2>;     // toy1-sized: everything in registers
using RjCfgM = RjCfg<8, 8, 8>;     // toy2 / tutorial-sized
using RjCfgL = RjCfg<20, 8, 16>;   // large: coal-mining (d<=13), the scaling workload (d<=20); ~1 KB of local vectors
using RjCfgG = RjCfg<AMX_MAX_DIM, AMX_MAX_COMPS, AMX_MAX_MODELS>;  // general (local memory)

constexpr double kHalfLog2Pi = 0.9189385332046727;  // literal at automix.c:1052
constexpr int kRjThreads = 128;
constexpr int kRjGroups = AMX_RJ_GROUPS;  // disjoint groups of chains whose visit counts are kept apart (Monte-Carlo error)
constexpr int kRjWarps = kRjThreads / 32;

// A per-thread scratch vector kept in shared memory, laid out [element][thread] (conflict-free): a run-time
// index costs one LDS/STS, where a register array needs a select chain per access and a local array an L1 trip.
struct LaneVec {
  double *p;
  __device__ __forceinline__ double get(int i) const { return p[i * kRjThreads]; }
  __device__ __forceinline__ void set(int i, double v) { p[i * kRjThreads] = v; }
};
// The large configurations keep the same vector in local memory: their shared memory holds the staged blobs and
// what is left of the SM's 228 KB must stay L1 for the chains' vectors (measured: every KB of carve-out costs).
template <int N>
struct LocalVec {
  double v[N];
  __device__ __forceinline__ double get(int i) const { return v[i]; }
  __device__ __forceinline__ void set(int i, double x) { v[i] = x; }
};
template <class CFG>
using AllocVec = typename std::conditional<(CFG::DMAX <= 8), LaneVec, LocalVec<CFG::LMAX>>::type;

// optional modes (amSampler.student_T_dof, amSampler.doPerm)
// compile-time "no optional modes": the branches below fold away (the small configuration uses this)
struct NoModes {
  static constexpr int dof = 0, do_perm = 0;
  static constexpr double lt_const = 0.0;
};
struct RjModes {
  int dof;          // 0: Gaussian innovations
  int do_perm;      // random permutation of the standardised vector (automix.c:1184-1194)
  double lt_const;  // constant of ltprob for this dof
};

// Read-only view of the proposal family blob (include/amx_layout.h).
struct ProposalView {
  const amx_fam_hdr *h;
  const double *D;
  bool uniform_dims;  // the lanes of a warp share (k, kn): fixed-dimension quadratic forms (quad_form)
  __device__ __forceinline__ void bind(const void *blob, bool uniform = false) {
    h = reinterpret_cast<const amx_fam_hdr *>(blob);
    D = reinterpret_cast<const double *>(h + 1);
    uniform_dims = uniform;
  }
  __device__ __forceinline__ const double *rec(int k, int l) const { return D + h->off[k] + l * h->stride[k]; }
  __device__ __forceinline__ const double *sig(int k) const { return D + h->ext[k]; }
};

template <class CFG>
struct ChainRegs {
  double th[CFG::DMAX];   // current point
  double thn[CFG::DMAX];  // proposal
  double pk[CFG::NMAX];   // adaptive model-jump probabilities (per chain, automix.c:1258-1282)
  double lp;
  double pkllim;
  int k;
  int nreinit;
  // carried from rj_propose to rj_finish
  int kn;
  double lr_pre, t_alloc, t_wt, t_det, gam;
  // counters (runStats, automix.h:180-185) and F_RJ accounting
  unsigned acc_b, try_b, acc_s, try_s, acc_j, try_j;
  unsigned long long flops;
};

// ---- within-model RWM (:1056-1085) -------------------------------------------------------
template <class CFG, class U, class MD>
__device__ __forceinline__ void rwm_block_propose(ChainRegs<CFG> &c, const ProposalView &P, U &u, const MD &md) {
  const int d = P.h->dims[c.k];
  const double *sg = P.sig(c.k);
  c.try_b++;
  int i = 0;
  if (md.dof == 0) {
    for (; i + 1 < d; i += 2) {
      double z0, z1;
      gauss_pair(u, z0, z1);
      aset(c.thn, i, fma(sg[i], z0, aget(c.th, i)));
      aset(c.thn, i + 1, fma(sg[i + 1], z1, aget(c.th, i + 1)));
    }
    if (d & 1) aset(c.thn, d - 1, fma(sg[d - 1], gauss_single(u), aget(c.th, d - 1)));
  } else {  // rt(): all the normals first, then ONE gamma draw scales them (automix.c:1663-1680)
    for (; i + 1 < d; i += 2) {
      double z0, z1;
      gauss_pair(u, z0, z1);
      aset(c.thn, i, z0);
      aset(c.thn, i + 1, z1);
    }
    if (d & 1) aset(c.thn, d - 1, gauss_single(u));
    const double den = t_divisor(md.dof, u);
    for (int j = 0; j < d; j++) aset(c.thn, j, fma(sg[j], aget(c.thn, j) / den, aget(c.th, j)));
  }
}
template <class CFG, class U>
__device__ __forceinline__ void rwm_block_finish(ChainRegs<CFG> &c, const ProposalView &P, U &u, double lpn) {
  const int d = P.h->dims[c.k];
  if (mh_accept(u.next(), lpn - c.lp)) {
    c.acc_b++;
    if constexpr (CFG::DMAX <= kRegArrayMax) {
#pragma unroll
      for (int i = 0; i < CFG::DMAX; i++) c.th[i] = (i < d) ? c.thn[i] : c.th[i];
    } else {
      for (int i = 0; i < d; i++) c.th[i] = c.thn[i];
    }
    c.lp = lpn;
  }
}
// coordinate j (the caller guarantees j < d for active lanes)
template <class CFG, class U, class MD>
__device__ __forceinline__ void rwm_coord_propose(ChainRegs<CFG> &c, const ProposalView &P, U &u, int j, const MD &md) {
  c.try_s++;
  double z = gauss_single(u);
  if (md.dof > 0) z /= t_divisor(md.dof, u);
  aset(c.thn, j, fma(P.sig(c.k)[j], z, aget(c.th, j)));
}
template <class CFG, class U>
__device__ __forceinline__ void rwm_coord_finish(ChainRegs<CFG> &c, U &u, int j, double lpn) {
  if (mh_accept(u.next(), lpn - c.lp)) {
    c.acc_s++;
    aset(c.th, j, aget(c.thn, j));
    c.lp = lpn;
  } else {
    aset(c.thn, j, aget(c.th, j));
  }
}

// Unnormalised allocation weights lambda_l N(x; mu_l, B_l B_l^T) of point x under model k's mixture and their
// sum (:1094-1108 / :1217-1230).  The callers divide only where the reference's quotient is consumed: the
// forward allocation scans all of palloc (:1111-1123), the reverse one reads palloc[ln] alone (:1234).
template <class CFG, class PA>
__device__ __forceinline__ double alloc_weights(const ProposalView &P, int k, const double (&x)[CFG::DMAX], PA &p) {
  const int d = P.h->dims[k], L = P.h->ncomp[k];
  double s = 0.0;
  for (int l = 0; l < L; l++) {
    const double *rec = P.rec(k, l);
    const double v = exp(rec[1] + (fma(-0.5, quad_form<CFG::DMAX>(rec, d, x, P.uniform_dims), rec[3])));
    p.set(l, v);
    s += v;
  }
  return s;
}

// ---- between-model move, everything up to the log-posterior of the proposal -------------
template <class CFG, class U, class MD, class PA>
__device__ __forceinline__ void rj_propose(ChainRegs<CFG> &c, const ProposalView &P, U &u, double gam,
                                           const MD &md, const int *clp_tab, PA &pa) {
  const int nm = P.h->nmodels;
  const int k = c.k, d = P.h->dims[k], L = P.h->ncomp[k];
  double wk[CFG::DMAX];
  c.try_j++;

  // 9.1 allocate the current point to a component (:1090-1123)
  int l = 0;
  double log_pa = 0.0;
  if (L > 1) {
    const double sw = alloc_weights<CFG>(P, k, c.th, pa);
    const double uu = u.next();
    const double flat = 1.0 / L;  // the reference's fallback when every weight underflows (:1107-1110)
    double t = 0.0, pl = 0.0;
    bool found = false;
    for (int i = 0; i < L; i++) {
      const double pi = (sw > 0) ? pa.get(i) / sw : flat;
      if (i == 0) pl = pi;  // component 0 if the scan never fires
      t += pi;
      if (!found && uu < t) {
        l = i;
        pl = pi;
        found = true;
      }
    }
    log_pa = log(pl);
    c.flops += (unsigned)(L * (d * d + 3 * d + 6));
  }
  // 9.2 standardise through component l (:1127-1135)
  const double *recl = P.rec(k, l);
  solve_lower<CFG::DMAX>(recl, d, c.th, wk);

  // 9.3 proposed model and component (:1138-1169)
  int kn = k;
  double lr = 0.0;
  c.gam = 0.0;
  if (nm > 1) {
    c.gam = gam;
    const double uu = u.next();
    double t = 0.0;
    bool found = false;
    kn = 0;
    for (int i = 0; i < nm; i++) {
      t += aget(c.pk, i);
      if (!found && uu < t) {
        kn = i;
        found = true;
      }
    }
    if (kn != k) lr = log(aget(c.pk, k)) - log(aget(c.pk, kn));
  }
  const int dn = P.h->dims[kn], Ln = P.h->ncomp[kn];
  int ln = 0;
  {
    const double uu = u.next();
    double t = 0.0;
    bool found = false;
    for (int i = 0; i < Ln; i++) {
      t += P.rec(kn, i)[0];
      if (!found && uu < t) {
        ln = i;
        found = true;
      }
    }
  }
  // 9.4 dimension matching (:1173-1204)
  auto permute = [&](int n) {  // perm() (:1703-1715): n-1 uniforms
    for (int i = 0; i < n - 1; i++) {
      const int j = i + (int)((n - i) * u.next());
      if (j != i) {
        const double t = aget(wk, j);
        aset(wk, j, aget(wk, i));
        aset(wk, i, t);
      }
    }
  };
  if (d < dn) {
    int i = d;
    for (; i + 1 < dn; i += 2) {
      double z0, z1;
      gauss_pair(u, z0, z1);
      aset(wk, i, z0);
      aset(wk, i + 1, z1);
    }
    if ((dn - d) & 1) aset(wk, dn - 1, gauss_single(u));
    if (md.dof > 0) {
      const double den = t_divisor(md.dof, u);
      for (int j = d; j < dn; j++) aset(wk, j, aget(wk, j) / den);
      for (int j = d; j < dn; j++) lr -= ltprob_dev(md.dof, md.lt_const, aget(wk, j));
    } else {
      for (int j = d; j < dn; j++) {
        const double w = aget(wk, j);
        lr += 0.5 * (w * w) + kHalfLog2Pi;
      }
    }
    if (md.do_perm) permute(dn);
  } else if (d == dn) {
    if (md.do_perm) permute(d);
  } else {
    if (md.do_perm) permute(d);
    if (md.dof > 0) {
      for (int j = dn; j < d; j++) lr += ltprob_dev(md.dof, md.lt_const, aget(wk, j));
    } else {
      for (int j = dn; j < d; j++) {
        const double w = aget(wk, j);
        lr -= (0.5 * (w * w) + kHalfLog2Pi);
      }
    }
  }
  // map through component ln of model kn (:1206-1211)
  const double *recn = P.rec(kn, ln);
  {
    const double *mun = recn + AMX_REC_HEAD, *Tn = mun + 2 * dn;
    if constexpr (CFG::DMAX <= kRegArrayMax) {
#pragma unroll
      for (int i = 0; i < CFG::DMAX; i++)
        if (i < dn) {
          double v = mun[i];
#pragma unroll
          for (int j = 0; j <= i; j++) v = fma(Tn[AMX_TRI(i, j)], wk[j], v);
          c.thn[i] = v;
        }
    } else {
      for (int i = 0; i < dn; i++) {
        double v = mun[i];
        const double *Ti = Tn + AMX_TRI(i, 0);
        for (int j = 0; j <= i; j++) v = fma(Ti[j], wk[j], v);
        c.thn[i] = v;
      }
    }
  }
  // 9.5 reverse allocation (:1216-1235)
  double log_pan = 0.0;
  if (Ln > 1) {
    const double sw = alloc_weights<CFG>(P, kn, c.thn, pa);
    log_pan = log((sw > 0) ? pa.get(ln) / sw : 1.0 / Ln);
    c.flops += (unsigned)(Ln * (dn * dn + 3 * dn + 6));
  }
  c.kn = kn;
  c.lr_pre = lr;
  c.t_alloc = log_pan - log_pa;
  c.t_wt = recl[1] - recn[1];
  c.t_det = recn[2] - recl[2];
  const int dd = d > dn ? d - dn : dn - d;
  c.flops += (unsigned)(d * d + d + dn * dn + 2 * dn + clp_tab[kn] + 4 * dd + 4 * nm + Ln + 27);
}

// 9.6 accept/reject and pk adaptation (:1238-1282).  Returns the model after the sweep.
template <class CFG, class U>
__device__ __forceinline__ void rj_finish(ChainRegs<CFG> &c, const ProposalView &P, U &u, double lpn,
                                          bool adapt) {
  const int nm = P.h->nmodels;
  double lr = c.lr_pre;
  lr += (lpn - c.lp);
  lr += c.t_alloc;
  lr += c.t_wt;
  lr += c.t_det;
  if (mh_accept(u.next(), lr)) {
    const int dn = P.h->dims[c.kn];
    if constexpr (CFG::DMAX <= kRegArrayMax) {
#pragma unroll
      for (int i = 0; i < CFG::DMAX; i++) c.th[i] = (i < dn) ? c.thn[i] : c.th[i];
    } else {
      for (int i = 0; i < dn; i++) c.th[i] = c.thn[i];
    }
    c.lp = lpn;
    c.k = c.kn;
    c.acc_j++;
  }
  if (adapt) {
    bool low = false;
    if constexpr (CFG::NMAX <= kRegArrayMax) {
#pragma unroll
      for (int j = 0; j < CFG::NMAX; j++)
        if (j < nm) {
          const double e = (j == c.k) ? 1.0 : 0.0;
          c.pk[j] += (c.gam * (e - c.pk[j]));
        }
      // the reference stops at the first pk below the limit; only "any" matters
#pragma unroll
      for (int j = 0; j < CFG::NMAX; j++)
        if (j < nm) low |= (c.pk[j] < c.pkllim);
    } else {
      for (int j = 0; j < nm; j++) {
        const double e = (j == c.k) ? 1.0 : 0.0;
        c.pk[j] += (c.gam * (e - c.pk[j]));
      }
      for (int j = 0; j < nm; j++) low |= (c.pk[j] < c.pkllim);
    }
    if (low) {
      c.nreinit++;
      c.pkllim = 1.0 / (10.0 * c.nreinit);
      const double un = 1.0 / nm;
      if constexpr (CFG::NMAX <= kRegArrayMax) {
#pragma unroll
        for (int j = 0; j < CFG::NMAX; j++) c.pk[j] = un;
      } else {
        for (int j = 0; j < nm; j++) c.pk[j] = un;
      }
    }
  }
}

// After the accept step thn must equal th on the coordinates the next sweep touches: the
// reference re-copies theta into thetan at the start of every component-wise sweep (:1070).
template <class CFG>
__device__ __forceinline__ void sync_proposal(ChainRegs<CFG> &c, int d) {
  if constexpr (CFG::DMAX <= kRegArrayMax) {
#pragma unroll
    for (int i = 0; i < CFG::DMAX; i++) c.thn[i] = c.th[i];
  } else {
    for (int i = 0; i < d; i++) c.thn[i] = c.th[i];
  }
}

// ---- global (SoA) chain state -------------------------------------------------------------
struct RjState {
  double *theta;   // [dmax][C]
  double *pk;      // [nmodels][C]
  double *lp;      // [C]
  double *pkllim;  // [C]
  int *k;          // [C]
  int *nreinit;    // [C]
  unsigned long long *draws;  // [C] uniforms consumed by each chain
  long C;
  int dmax, nmodels;
};

struct RjLaunch {
  RjState st;
  const void *prop_blob;
  int prop_bytes;
  const void *tgt_blob;
  int tgt_bytes;
  int tgt_flags;
  unsigned long long seed;
  unsigned long long chain_base;  // global id of chain 0 of this population
  const double *tape;          // parity mode
  unsigned long long tape_stride;
  const double *pk_shared;     // population pk mode: [nmodels] jump probabilities shared by every chain (else NULL)
  const int *order;            // sorted mode: slot -> chain, chains grouped by (model, proposed model), widest first (else NULL)
  const double *gam;           // [nsweeps] pk-adaptation step sizes (sweep_i+1)^(-2/3)
  unsigned long long sweep0;   // sweep_i of the first sweep of this launch
  int nsweeps;
  int adapt;                   // doAdapt && !isBurning
  RjModes modes;
  // outputs
  unsigned long long *visits;  // [nmodels]
  unsigned long long *visits_grp;  // [kRjGroups][AMX_MAX_MODELS] the same counts by group of chains (CTA index mod kRjGroups)
  unsigned long long *cnt;     // [8]: 6 runStats counters, flops, draws
  int *status;                 // bit 0: tape overrun, bit 1: NaN log-posterior
  // trace chains
  int ntrace;
  long tr_stride, tr_off;      // trace row of (chain, sweep s of this launch) = chain * tr_stride + tr_off + s
  int *tr_k;
  double *tr_lp, *tr_theta, *tr_pk;
};

}  // namespace amx
