// amx_em2.cuh -- K2, second generation: the fused CEM^2 component step as a warp-specialised streaming kernel.
// Included by amx_em.cu (inside namespace amx, after the first-generation kernel whose control structures,
// Cholesky routine and mbarrier/TMA helpers it shares).  Same algorithm, same pass sequence and the same
// arithmetic per sample as em_fit_kernel's fused mode (automix.c:664-1006); what changes is how the work is laid on
// the SM and how the GPUs (and the CTAs of one GPU) agree on the next step.
//
//  * ONE persistent CTA per SM: NTEAM teams of 128 threads over a ring of 128-sample stages (d coordinate rows of xT
//    and the L live rows of the density cache E, 1 KB bulk copies, byte-counted on an mbarrier) that fills the SM's
//    shared memory.  Tiles go to the teams round-robin; a team takes its stage as soon as it has landed and, when
//    done with it, refills the same stage itself with the tile NS places ahead.  No CTA-wide barrier inside a
//    pass, no producer to wait for, no idle SM while a tile is in flight: HBM latency is covered by the ring, not by
//    occupancy.  (12 warps = a whole number of register-file allocation units: 168 registers per thread.)
//  * Every accumulator lives in REGISTERS.  A-phase (thread = sample): new density of the refreshed component,
//    sum_l lam_l E_il, T_l += E_il / sum (column sums are lam_l T_l: one FMA per component instead of
//    multiply-multiply-add), log-likelihood; the thread leaves w_next and dx = x - pivot in the stage.  B-phase: the
//    four warps of the team split the d + d(d+1)/2 entries of S1, S2 of the NEXT component four ways, each warp
//    walking all 128 samples of the stage.  (The first generation reduced columns of the shared tile: three shared
//    loads per FMA, which is what bound it.)
//  * The step barrier has no leader and no release: the last CTA of a GPU to arrive sums that GPU's partial rows
//    (value-major, coalesced, fixed order) and posts the row to EVERY GPU (peer stores over NVLink); every CTA of
//    every GPU then adds the rows in GPU order and runs the sequential section -- weight update, annihilation,
//    Cholesky, MML cost, convergence -- redundantly on its own copy of the mixture in shared memory.  Same inputs,
//    same order of operations: every CTA takes the same branches, bit for bit, and nothing has to be published.
//    One cross-GPU latency per step instead of three (arrive, leader's peer loads, public-state push).
#pragma once

constexpr int kV2TS = 128;                        // samples per stage
constexpr int kV2NV = 128;                        // doubles per partial row (>= kEmLmax + 2 + 12 + 78)
constexpr int kV2Dmax = 12;

struct V2Sync {                                   // one per GPU, in that GPU's memory
  unsigned arrive;                                // CTAs of this GPU that have finished the pass (monotonic)
  unsigned pad0[31];
  unsigned flag[kEmMaxDev][32];                   // flag[g][0] = epoch + 1 once GPU g's row of that epoch is in inbox
  double inbox[2][kEmMaxDev][kV2NV];              // by epoch parity
  long long dbg_cta[160];                         // AMX_EM_DEBUG: cycles each CTA spent in its data passes
};

struct V2Args {
  V2Sync *sync[kEmMaxDev];                        // every GPU's block, peer-mapped
  double *part;                                   // [kV2NV][grid] value-major partial rows of this GPU's CTAs
  int ns;                                         // ring stages
  int region0_doubles;                            // ring / reduction scratch (aliased)
  int debug;
  // Several ranks emulated by slices of ONE grid on one GPU (tests of the sharded exchange on a single-GPU box):
  // CTAs [r gv, (r+1) gv) are rank r and take their argument blocks from these device arrays.
  int vranks;
  const EmArgs *va;
  const V2Args *vv;
};

// Data layout in HBM (second generation): TILE-major.  The d coordinate rows of a 128-sample tile are contiguous
// (xT[tile][j][128]) and so are its Lmax density-cache rows (E[tile][slot][128]): a stage is filled by one bulk copy for
// the coordinates plus one per run of consecutive live slots -- a handful of copies per tile instead of d + L.  (A bulk
// copy costs the SM ~170 cycles whatever its size; at one 1 KB row per copy that alone was 3.5 times the HBM time of
// the tile, and it is what bound the first generation too.)
__device__ __forceinline__ size_t v2_x_at(long i, int j, int d) { return ((size_t)(i / kV2TS) * d + j) * kV2TS + (size_t)(i % kV2TS); }
__device__ __forceinline__ size_t v2_e_at(long i, int slot, int Lmax) {
  return ((size_t)(i / kV2TS) * Lmax + slot) * kV2TS + (size_t)(i % kV2TS);
}

template <int DMAX>
struct V2Cfg {
  static constexpr int TRI = DMAX * (DMAX + 1) / 2;
  static constexpr int NE = TRI + DMAX;                // S2 entries then S1 entries
  static constexpr int NE4 = (NE + 3) / 4 * 4;         // ... padded so that the T entries keep l = e (mod 4)
  static constexpr int NEALL = NE4 + kEmLmax;          // ... then T_l, l = 0 .. kEmLmax-1
  static constexpr int NB = NEALL / 4;                 // accumulators per thread: entries e = wq (mod 4)
  static constexpr int NBM = NE4 / 4;                  // of which moments
};

__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait with a watchdog: a copy that never lands becomes a trap (an error the host sees), not a hang
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 20000000000LL) __trap();
  } while (!ok);
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// lnormprob of sample column `s` of a stage under component l, operation for operation the reference's
// (automix.c:1727-1750), from the CTA's own copy of the mixture (see lnormprob_slow of the first generation).
// xcol points at the x rows of the stage; they hold x itself (the caller runs this before the pivot shift).
__device__ __noinline__ double v2_lnormprob_slow(const double *mu, const double *B, int d, const double *xcol) {
  double r[kV2Dmax];
  double det = 1.0;
  for (int i = 0; i < d; i++) r[i] = xcol[i * kV2TS] - mu[i];
  for (int i = 0; i < d; i++) {
    for (int j = 0; j < i; j++) r[i] -= B[AMX_TRI(i, j)] * r[j];
    const double bii = B[AMX_TRI(i, i)];
    r[i] /= bii;
    det *= bii;
  }
  double q = 0.0;
  for (int i = 0; i < d; i++) q += r[i] * r[i];
  return -0.5 * q - (d / 2.0) * log(2.0 * 3.14159265358979323846) - log(det);
}

// A sample whose total density is below 1e-280 (or NaN), redone the reference's way (:849-866): every component puts
// it below exp(-644), the cached densities are near the subnormal range where lam * exp(lpd) no longer tracks the
// reference's exp(log(lam) + lpd) and 1/sum cannot be formed.  Leaves w_l / lam_l (or w_l itself in a `direct` pass, see
// the kernel) in the sample's column of the stage (the caller accumulates it with multiplier 1), returns the
// responsibility of component nx.
__device__ __noinline__ double v2_slow_sample(const double *lam, const double *s_mu, const double *s_B, int d, int tri, int L,
                                              int nx, double sum, bool direct, const double *xcol, double *Ecol, double *ll,
                                              double *nfb) {
  double wn = 0.0;
  bool uniform = !(sum < 1e-280);  // NaN sum: the reference's `sum > 0` fails -> uniform responsibilities, -500 penalty
  if (!uniform) {
    double s2 = 0.0;
    for (int l = 0; l < L; l++) {
      const double wl = exp(log(lam[l]) + v2_lnormprob_slow(s_mu + l * d, s_B + l * tri, d, xcol));
      Ecol[l * kV2TS] = wl;
      s2 += wl;
    }
    if (s2 > 0) {
      *ll += log(s2);
      for (int l = 0; l < L; l++) {
        const double w = Ecol[l * kV2TS] / s2;
        Ecol[l * kV2TS] = direct ? w : w / lam[l];
        if (l == nx) wn = w;
      }
    } else {
      uniform = true;
    }
  }
  if (uniform) {
    *nfb += 1.0;
    const double w = 1.0 / L;
    for (int l = 0; l < L; l++) Ecol[l * kV2TS] = direct ? w : w / lam[l];
    wn = (L > 0) ? w : 0.0;
  }
  return wn;
}

// |T^-1 (x - mu)|^2 for the lower-triangular factor of a family record (forward substitution of lnormprob,
// automix.c:1735-1747).  Column-oriented: once r_j is known every later row takes its -T_ij r_j term, so row i
// receives its terms in the order j = 0 .. i-1 exactly as in the row-oriented solve_lower (same roundings), but the
// rows are d independent chains instead of one chain of d(d+1)/2 operations.
template <int DMAX>
__device__ __forceinline__ double v2_solve_cols(const double *rec, const double (&x)[DMAX]) {
  const double *mu = rec + AMX_REC_HEAD, *rd = mu + DMAX, *T = rd + DMAX;
  double v[DMAX];
#pragma unroll
  for (int i = 0; i < DMAX; i++) v[i] = x[i] - mu[i];
  double q = 0.0;
#pragma unroll
  for (int j = 0; j < DMAX; j++) {
    const double r = v[j] * rd[j];
    q = fma(r, r, q);
#pragma unroll
    for (int i = j + 1; i < DMAX; i++) v[i] = fma(-T[AMX_TRI(i, j)], r, v[i]);
  }
  return q;
}

// B-phase of one sample for the warp that owns the entries e = WQ (mod 4) of (S2 | S1 | T): S2 is the packed lower
// triangle (row-major) of sum w dx dx^T, S1 = sum w dx, T_l = sum E_il * inv.  dxs: the stage's coordinate rows, already
// shifted by the pivot; Es: its density rows.  (DMAX is the dimension itself.)
template <int DMAX, int WQ>
__device__ __forceinline__ void v2_bsample(double (&acc)[V2Cfg<DMAX>::NB], const double *dxs, const double *Es, double w,
                                           double inv, int s, int d, int L) {
  using CF = V2Cfg<DMAX>;
  double dx[DMAX];
#pragma unroll
  for (int j = 0; j < DMAX; j++) dx[j] = dxs[j * kV2TS + s];
  int e = 0;
#pragma unroll
  for (int j = 0; j < DMAX; j++) {
    const double wd = w * dx[j];
#pragma unroll
    for (int k = 0; k <= j; k++, e++)
      if ((e & 3) == WQ) acc[e >> 2] = fma(wd, dx[k], acc[e >> 2]);
    if (((CF::TRI + j) & 3) == WQ) acc[(CF::TRI + j) >> 2] += wd;
  }
#pragma unroll
  for (int i = 0; i < kEmLmax / 4; i++)
    if (4 * i + WQ < L) acc[CF::NBM + i] = fma(Es[(4 * i + WQ) * kV2TS + s], inv, acc[CF::NBM + i]);
}
template <int DMAX, int WQ>
__device__ __forceinline__ void v2_bphase(double (&acc)[V2Cfg<DMAX>::NB], const double *dxs, const double *Es, const double *Enx,
                                          double lam_nx, const double *winv, int lane, int d, int L) {
#pragma unroll 1
  for (int q = 0; q < kV2TS / 32; q++) {
    const int s = lane + 32 * q;
    const double inv = winv[s];
    v2_bsample<DMAX, WQ>(acc, dxs, Es, (lam_nx * Enx[s]) * inv, inv, s, d, L);
  }
}
template <int DMAX>
__device__ __forceinline__ void v2_bphase_any(double (&acc)[V2Cfg<DMAX>::NB], int wq, const double *dxs, const double *Es,
                                              const double *Enx, double lam_nx, const double *winv, int lane, int d, int L) {
  if (wq == 0) v2_bphase<DMAX, 0>(acc, dxs, Es, Enx, lam_nx, winv, lane, d, L);
  else if (wq == 1) v2_bphase<DMAX, 1>(acc, dxs, Es, Enx, lam_nx, winv, lane, d, L);
  else if (wq == 2) v2_bphase<DMAX, 2>(acc, dxs, Es, Enx, lam_nx, winv, lane, d, L);
  else v2_bphase<DMAX, 3>(acc, dxs, Es, Enx, lam_nx, winv, lane, d, L);
}

// The sequential section, run by every CTA on its own copy of the state (see the header) -- by ONE WARP of it: nothing
// in it is wider than 32 (components) except the d(d+1)/2 <= 78 entries of a factor, and a warp-synchronous section
// has no CTA barriers (the block-cooperative form of the first generation spent most of its ~14k cycles in some
// fourteen of them).  Scalars that every lane needs (sums in the reference's index order) are formed by every lane
// redundantly -- same operations, same result -- instead of being broadcast.  Mirrors em_leader_block branch for
// branch; `writer` (GPU 0, CTA 0) also records what the host reads back: traces and the best mixture.
template <int DMAX>
__device__ __forceinline__ void v2w_renorm(LeaderS<DMAX> &S, int lane) {  // lam /= sum(lam), sum in index order (:785-792)
  double sum = 0.0;
  for (int l = 0; l < S.L; l++) sum += S.lam[l];
  __syncwarp();
  if (lane < S.L) S.lam[lane] /= sum;
  __syncwarp();
}
template <int DMAX>
__device__ __forceinline__ double v2w_cost(const LeaderS<DMAX> &S, int lane, long n, int nparams, double loglik) {  // :870-876
  const double mine = (lane < S.L) ? log((double)n * S.lam[lane] / 12.0) : 0.0;
  double sum = 0.0;
  for (int l = 0; l < S.L; l++) sum += __shfl_sync(0xffffffffu, mine, l);
  return (nparams / 2.0) * sum + (S.L / 2.0) * log((double)n / 12.0) + S.L * (nparams + 1) / 2.0 - loglik;
}
template <int DMAX>
__device__ __forceinline__ void v2w_drop(LeaderS<DMAX> &S, double *s_mu, double *s_B, int lane, int gone) {  // :823-836, :908-921
  constexpr int d = DMAX, tri = d * (d + 1) / 2;
  const int L = S.L;
  for (int q = lane; q < tri; q += 32)
    for (int l = gone; l < L - 1; l++) s_B[l * tri + q] = s_B[(l + 1) * tri + q];
  if (lane < d)
    for (int l = gone; l < L - 1; l++) s_mu[l * d + lane] = s_mu[(l + 1) * d + lane];
  __syncwarp();
  if (lane == 0) {
    for (int l = gone; l < L - 1; l++) {
      S.lam[l] = S.lam[l + 1];
      S.slot[l] = S.slot[l + 1];
    }
    S.L = L - 1;
  }
  __syncwarp();
}
template <int DMAX>
__device__ __forceinline__ void v2w_make_rec(LeaderS<DMAX> &S, const double *s_mu, double *s_rec, int lane, int l) {
  constexpr int d = DMAX, tri = d * (d + 1) / 2;
  if (lane == 0) {
    double prod = 1.0;
    for (int i = 0; i < d; i++) prod *= S.Bc[AMX_TRI(i, i)];
    const double ld = log(prod);
    s_rec[0] = S.lam[l];
    s_rec[1] = 0.0;
    s_rec[2] = ld;
    s_rec[3] = -(d / 2.0) * log(2.0 * 3.14159265358979323846) - ld;
  }
  if (lane < d) {
    s_rec[AMX_REC_HEAD + lane] = s_mu[l * d + lane];
    s_rec[AMX_REC_HEAD + d + lane] = 1.0 / S.Bc[AMX_TRI(lane, lane)];
  }
  for (int q = lane; q < tri; q += 32) s_rec[AMX_REC_HEAD + 2 * d + q] = S.Bc[q];
  __syncwarp();
}

template <int DMAX>
__device__ void v2_leader_warp(const EmArgs &a, EmCtrl *c, bool writer, int pass, const double *s_tot, LeaderS<DMAX> &S,
                               double *s_mu, double *s_B, double *s_rec, double *s_piv) {
  constexpr int d = DMAX, tri = d * (d + 1) / 2, nparams = d + tri;
  const int lane = threadIdx.x & 31;
  if (pass == kPassInitStats) {
    // :700-723 common isotropic start; s_tot = [sum x_j (d) | sum x_j^2 (d)]
    double s2 = 0.0;
    const double len = (double)a.n_total;
    for (int j = 0; j < d; j++) s2 += (s_tot[d + j] - s_tot[j] * s_tot[j] / len) / len;
    s2 /= (10.0 * d);
    const double scal = sqrt(s2);
    for (int q = lane; q < a.Lmax * d; q += 32) s_mu[q] = ld_cg(a.init_rows + q);
    for (int q = lane; q < a.Lmax * tri; q += 32) s_B[q] = 0.0;
    for (int q = lane; q < tri; q += 32) S.Bc[q] = 0.0;
    S.slot[lane] = lane;
    S.lam[lane] = (lane < a.Lmax) ? 1.0 / a.Lmax : 0.0;
    __syncwarp();
    for (int q = lane; q < a.Lmax * d; q += 32) s_B[(q / d) * tri + AMX_TRI(q % d, q % d)] = scal;
    if (lane < d) S.Bc[AMX_TRI(lane, lane)] = scal;
    if (lane == 0) {
      S.scal = scal;
      if (!(s2 > 0.0)) S.status = AMX_ENUMERIC;
      S.L = a.Lmax;
      S.c = a.Lmax;  // every start density is formed by the first E-step itself
      S.next = 0;
      S.iters = 0;
      S.pass = kPassRefresh0;
    }
    __syncwarp();
    v2w_make_rec<DMAX>(S, s_mu, s_rec, lane, 0);
    if (lane < d) s_piv[lane] = s_mu[lane];  // pivot of the first pass: the mean of component 0
    __syncwarp();
    return;
  }
  // bytes this pass asked HBM for (all GPUs): the coordinate rows, the density rows it copied, the row(s) it wrote
  if (lane == 0) {
    const double nn = (double)a.n_total, Lp = (double)(S.L < 0 ? 0 : S.L);
    S.tmp[kEmLmax - 1] += 8.0 * nn * (d + (pass == kPassRefresh0 ? 0.0 : Lp)) +
                          8.0 * nn * (pass == kPassRefresh0 ? Lp : (pass == kPassDensRefresh ? 1.0 : 0.0));
  }
  // a refresh finished: s_tot = [T_l (Lmax) | loglik | fallbacks | S1 (d) | S2 (tri)], column sums are lam_l T_l
  S.colsum[lane] = (lane < S.L) ? (S.keep ? s_tot[lane] : S.lam[lane] * s_tot[lane]) : 0.0;  // S.keep: a `direct` pass
  if (lane < d) S.S1[lane] = s_tot[kEmLmax + 2 + lane];
  for (int q = lane; q < tri; q += 32) S.S2[q] = s_tot[kEmLmax + 2 + d + q];
  const double loglik = s_tot[kEmLmax] - 500.0 * s_tot[kEmLmax + 1];
  __syncwarp();
  // every lane carries the scalar state in registers; lane 0 writes it back at the end
  int L = S.L, cc = S.c, next = S.next, iters = S.iters, natural = S.natural, forced = S.forced, stop = S.stop, status = S.status;
  int best_L = S.best_L, npass = S.pass, forced_pending = S.forced_pending;
  long comp_steps = S.comp_steps;
  double flops = S.flops, cost = S.cost, cost_prev = S.cost_prev, cost_best = S.cost_best;
  int act;
  if (iters == 0) {  // initial E-step done: start outer iteration 1
    iters = 1;
    natural = forced = 0;
    cc = 0;
    act = kActPlan;
  } else if (forced_pending) {  // refresh after a forced annihilation (:931-958): the cost after it
    cost = v2w_cost<DMAX>(S, lane, a.n_total, nparams, loglik);
    forced_pending = 0;
    act = kActFinishIter;
  } else {
    if (pass == kPassDensRefresh) cc++;  // component kept: move on (:819)
    act = (cc < L) ? kActPlan : kActEndSweep;
  }
  while (act != kActDone) {
    if (act == kActPlan) {
      // start the update of component cc from the column sums and the moments (:773-801)
      double tot = 0.0, wkeep = 0.0;
      for (int l = 0; l < L; l++) {
        const double wl = max_m(0.0, (S.colsum[l] - nparams / 2.0));
        if (l == cc) wkeep = wl;
        tot += wl;
      }
      // the weight this component is about to get (:785-792), formed without touching the weights yet: a re-pivot
      // pass (below) must see the weights the moments were formed with
      const double lam_cc = wkeep / tot;
      double lsum = 0.0;
      for (int l = 0; l < L; l++) lsum += (l == cc) ? lam_cc : S.lam[l];
      const bool kept = (lam_cc / lsum) > 0.005;
      const double S0 = S.colsum[cc];
      if (kept) {
        // S1, S2 are moments of (x - pivot), pivot = s_piv (the component's mean before this update):
        //   mean = pivot + S1/S0,   cov = S2/S0 - (S1/S0)(S1/S0)^T   (:797-810 in shifted form)
        __syncwarp();
        if (lane < d) S.dl[lane] = S.S1[lane] / S0;
        __syncwarp();
        for (int q = lane; q < tri; q += 32) {
          int j = 0;
          while ((j + 1) * (j + 2) / 2 <= q) j++;
          const int k = q - j * (j + 1) / 2;
          S.Bc[q] = (S.S2[q] - S.S1[j] * S.dl[k]) / S0;
        }
        __syncwarp();
        // The shifted form cancels (S1/S0)^2 out of S2/S0: harmless while the mean moves by less than a few standard
        // deviations per step (always, in a fit that is converging), but a component collapsing onto a few (duplicate)
        // samples can have a variance many orders below its squared shift.  Then the moments are taken again about
        // the NEW mean (one extra pass, same weights), which is the reference's centred formula (:803-810) to rounding.
        bool cancels = false;
        for (int j = 0; j < d; j++) cancels |= !(S.Bc[AMX_TRI(j, j)] * 1000.0 > S.dl[j] * S.dl[j]);
        if (cancels && !S.drop) {
          if (lane < d) s_piv[lane] = s_piv[lane] + S.dl[lane];
          if (lane == 0) S.drop = 1;  // re-pivot pass in flight
          __syncwarp();
          next = cc;
          npass = kPassRefresh;
          act = kActDone;
          continue;
        }
      }
      __syncwarp();
      if (lane == 0) {
        S.lam[cc] = lam_cc;
        S.drop = 0;
      }
      comp_steps++;
      flops += (double)a.n_total * (2.0 * d * d + 8.0 * d + 4.0 * L + 7.0);
      __syncwarp();
      v2w_renorm<DMAX>(S, lane);
      if (kept) {
        if (lane < d) s_mu[cc * d + lane] = s_piv[lane] + S.dl[lane];
        __syncwarp();
        const bool ok = warp_chol<DMAX>(S.Bc, d);
        __syncwarp();
        const bool chol_ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
        for (int q = lane; q < tri; q += 32) s_B[cc * tri + q] = S.Bc[q];
        v2w_make_rec<DMAX>(S, s_mu, s_rec, lane, cc);
        // A scatter matrix that is not positive definite even about its own mean (a component sitting on identical
        // samples): the reference takes sqrt of a non-positive pivot (:1691) and carries on with NaN / zero entries --
        // samples whose density turns NaN get uniform responsibilities and the -500 penalty (:855-866), the iteration
        // cap ends a fit that never recovers, and the minimum-cost mixture seen is returned.  Same here; the status
        // word records that it happened.
        if (!chol_ok) status = AMX_ENUMERIC;
        next = (cc + 1 < L) ? cc + 1 : 0;
        npass = kPassDensRefresh;
        act = kActDone;
      } else {  // natural annihilation (:821-845): refresh before the next component is looked at
        v2w_drop<DMAX>(S, s_mu, s_B, lane, cc);
        L = S.L;
        v2w_renorm<DMAX>(S, lane);
        natural = 1;
        next = (cc < L) ? cc : 0;
        npass = kPassRefresh;
        act = kActDone;
      }
    } else if (act == kActEndSweep) {
      cost = v2w_cost<DMAX>(S, lane, a.n_total, nparams, loglik);
      if (iters == 1) cost_prev = cost;
      const bool savebest = (iters == 1 || cost < cost_best);  // :881-893
      if (savebest) {
        best_L = L;
        cost_best = cost;
      }
      int drop = -1;
      if (fabs(cost_prev - cost) < min_m(1E-5 * fabs(cost_prev), 0.01) && iters > 1) {  // :894
        if (L == 1) {
          stop = 1;
        } else {
          forced = 2;
          double lo = S.lam[0];
          int gone = 0;
          for (int l = 1; l < L; l++)
            if (lo > S.lam[l]) {
              lo = S.lam[l];
              gone = l;
            }
          drop = gone;
        }
      }
      if (savebest && writer) {
        if (lane < L) c->best_lam[lane] = S.lam[lane];
        for (int q = lane; q < L * d; q += 32) c->best_mu[q / d][q % d] = s_mu[q];
        for (int q = lane; q < L * tri; q += 32) c->best_B[q / tri][q % tri] = s_B[q];
      }
      if (drop >= 0) {
        v2w_drop<DMAX>(S, s_mu, s_B, lane, drop);
        L = S.L;
        v2w_renorm<DMAX>(S, lane);
        forced_pending = 1;
        next = 0;
        npass = kPassRefresh;
        act = kActDone;
      } else {
        act = kActFinishIter;
      }
    } else {  // kActFinishIter (:961-970)
      if (iters > a.maxit) stop = 1;
      cost_prev = cost;
      const int it = iters - 1;
      if (writer && lane == 0) {
        if (a.trace_ann) a.trace_ann[it] = natural + forced;
        if (a.trace_cost) a.trace_cost[it] = cost;
        if (a.trace_loglik) a.trace_loglik[it] = loglik;
        if (a.trace_L) a.trace_L[it] = L;
      }
      if (stop) {
        npass = kPassStop;
        act = kActDone;
      } else {
        iters++;
        natural = forced = 0;
        cc = 0;
        act = kActPlan;
      }
    }
  }
  __syncwarp();
  // pivot of the next pass: the current mean of the component whose moments it forms (unless a re-pivot set it)
  if (!S.drop && lane < d && next < kEmLmax) s_piv[lane] = s_mu[next * d + lane];
  if (lane == 0) {
    S.L = L; S.c = cc; S.next = next; S.iters = iters; S.natural = natural; S.forced = forced; S.stop = stop; S.status = status;
    S.best_L = best_L; S.pass = npass; S.forced_pending = forced_pending; S.comp_steps = comp_steps; S.flops = flops;
    S.loglik = loglik; S.cost = cost; S.cost_prev = cost_prev; S.cost_best = cost_best;
  }
  __syncwarp();
}

// ---- the kernel ----------------------------------------------------------------------------------------------
template <int DMAX, int NTEAM>
__global__ void __launch_bounds__(NTEAM * 128, 1) em_fit_v2_kernel(EmArgs a_in, V2Args v_in) {
  // (the argument blocks stay in parameter space -- constant-bank operands; only what differs between emulated ranks is
  // taken from the device arrays, into registers)
  const EmArgs &a = a_in;
  const V2Args &v = v_in;
  const int G = v_in.vranks > 1 ? (int)gridDim.x / v_in.vranks : (int)gridDim.x;  // CTAs of this rank
  const int bid = (int)blockIdx.x % G;                                            // this CTA's index within its rank
  int rk_rank = a_in.rank;
  long rk_n = a_in.n, rk_npad = a_in.npad;
  const double *rk_x = a_in.x;
  double *rk_xT = a_in.xT, *rk_E = a_in.E, *rk_wout = a_in.w_out, *rk_part = v_in.part;
  if (v_in.vranks > 1) {
    const EmArgs *ra = v_in.va + blockIdx.x / G;
    rk_rank = ra->rank; rk_n = ra->n; rk_npad = ra->npad; rk_x = ra->x; rk_xT = ra->xT; rk_E = ra->E; rk_wout = ra->w_out;
    rk_part = v_in.vv[blockIdx.x / G].part;
  }
  using CF = V2Cfg<DMAX>;
  constexpr int TRI = CF::TRI, NB = CF::NB, NE4 = CF::NE4;
  constexpr int NCONS = NTEAM * 128;
  constexpr int kMaxStages = 8;
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages];
  __shared__ volatile unsigned s_gen[kMaxStages];
  __shared__ int s_flag, s_nrun;
  __shared__ int s_run[2 * kEmLmax];

  constexpr int d = DMAX;  // one instantiation per dimension: no per-coordinate predicates anywhere
  const int Lmax = a.Lmax;
  constexpr int tri = d * (d + 1) / 2;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int team = warp >> 2, wq = warp & 3, tt = t & 127;
  const long n = rk_n, np = rk_npad;
  const long ntiles = np / kV2TS;
  const bool writer = (rk_rank == 0 && bid == 0);
  EmCtrl *ctrl = a.ctrl;

  // shared memory: [region0: ring, aliased by the end-of-pass reduction scratch][mu][B][rec][tot][part][LeaderS]
  const int stage_doubles = (d + Lmax + 1) * kV2TS;  // rows: x (d) | E (Lmax, component order) | inv (1/sum, or 1)
  double *ring = smem;
  double *s_mu = smem + v.region0_doubles;           // [Lmax][d]
  double *s_B = s_mu + Lmax * d;                     // [Lmax][tri]
  double *s_rec = s_B + Lmax * tri;                  // family record of the component in progress
  double *s_piv = s_rec + (AMX_REC_HEAD + 2 * DMAX + TRI);  // [DMAX] pivot of the moments of this pass
  double *s_tot = s_piv + DMAX + (DMAX & 1);          // [kV2NV] totals over all CTAs and GPUs
  double *s_part = s_tot + kV2NV;                    // [kV2NV] this CTA's partial row
  LeaderS<DMAX> &S = *reinterpret_cast<LeaderS<DMAX> *>(s_part + kV2NV);
  const int NS = v.ns;

  if (t == 0) {
    for (int q = 0; q < NS; q++) {
      mbar_init(&s_full[q], 1);
      s_gen[q] = 0u;
    }
    memset(&S, 0, sizeof(S));
  }
  for (int q = t; q < kV2NV; q += blockDim.x) s_part[q] = 0.0;
  __syncthreads();

  unsigned epoch = 0;
  unsigned long long seq_base = 0;  // stages used by the passes so far (ring position and mbarrier parity)
  int pass = kPassInitStats;
  long long dbg_pass = 0, dbg_bar = 0, dbg_lead = 0, dbg_x1 = 0, dbg_x2 = 0;

  for (;;) {
    const long long tk0 = clock64();
    int nv = 0;
    if (pass == kPassInitStats) {
      // transpose x -> xT and accumulate sum x_j, sum x_j^2 (:700-711)
      double acc[2 * DMAX];
#pragma unroll
      for (int j = 0; j < 2 * DMAX; j++) acc[j] = 0.0;
      {
        for (long i = (long)bid * NCONS + t; i < n; i += (long)G * NCONS) {
#pragma unroll
          for (int j = 0; j < DMAX; j++) {
            const double xv = rk_x[i * d + j];
            __stcg(rk_xT + v2_x_at(i, j, d), xv);
            acc[j] += xv;
            acc[DMAX + j] = fma(xv, xv, acc[DMAX + j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 2 * DMAX; j++) {
          const double r = warp_sum(acc[j]);
          if (lane == 0) ring[warp * 2 * DMAX + j] = r;
        }
      }
      __syncthreads();
      if (t < 2 * d) {
        const int src = t < d ? t : DMAX + (t - d);
        double tot = 0.0;
        for (int w = 0; w < 4 * NTEAM; w++) tot += ring[w * 2 * DMAX + src];
        s_part[t] = tot;
      }
      fence_proxy_async();
      nv = 2 * d;
    } else {
      // ------------------------------------------------------------ the streaming pass
      const int L = S.L < 0 ? 0 : S.L, nx = S.next, cc = S.c;
      const bool dens_all = (pass == kPassRefresh0);  // first E-step: every density is formed here, unguarded weights
      const bool dens = (pass == kPassDensRefresh);
      const int ntile_cta = (int)((ntiles - (long)bid + G - 1) / G);  // tiles b, b+G, ... of this CTA
      // Runs of consecutive live slots (the slot list is increasing: annihilation removes entries, never reorders):
      // run r covers components [s_run[2r], s_run[2r] + s_run[2r+1]).  The refreshed component's stale row is copied
      // with its run and overwritten in the stage (1 KB per tile, cheaper than splitting the run).
      if (t == 0) {
        int nr = 0;
        for (int l = 0; l < L && !dens_all; l++) {
          if (l == 0 || S.slot[l] != S.slot[l - 1] + 1) {
            s_run[2 * nr] = l;
            s_run[2 * nr + 1] = 0;
            nr++;
          }
          s_run[2 * (nr - 1) + 1]++;
        }
        s_nrun = nr;
      }
      __syncthreads();
      const int nrun = s_nrun;
      const int rows = dens_all ? d : d + L;
      // lane 0 copies the coordinates of tile `it` of this CTA into stage (seq_base + it) % NS, lane r + 1 run r
      auto fetch = [&](int it) {
        const unsigned long long seq = seq_base + (unsigned long long)it;
        const int st = (int)(seq % (unsigned)NS);
        double *xs = ring + st * stage_doubles, *Es = xs + d * kV2TS;
        const long tl = (long)bid + (long)it * G;
        if (lane == 0) {
          s_gen[st] = (unsigned)(seq / (unsigned)NS) + 1u;  // fills issued for this stage
          mbar_expect_tx(&s_full[st], (uint32_t)(rows * kV2TS * 8));
          tma_load_row(xs, rk_xT + (size_t)tl * d * kV2TS, (uint32_t)(d * kV2TS * 8), &s_full[st]);
        }
        __syncwarp();
        if (lane >= 1 && lane <= nrun) {
          const int l0 = s_run[2 * (lane - 1)], len = s_run[2 * (lane - 1) + 1];
          tma_load_row(Es + l0 * kV2TS, rk_E + ((size_t)tl * Lmax + S.slot[l0]) * kV2TS, (uint32_t)(len * kV2TS * 8), &s_full[st]);
        }
      };
      // prologue: the ring is idle (end-of-pass barrier): warp 0 fills it.  The density rows it copies were written
      // with ordinary stores during the previous pass: order those before the bulk copies (async proxy) too.
      fence_proxy_async();
      asm volatile("fence.proxy.async.global;" ::: "memory");
      if (warp == 0)
        for (int it = 0; it < NS && it < ntile_cta; it++) fetch(it);
      {
        double acc[NB];  // this thread's share of (S2 | S1 | T): entries e = wq (mod 4), over the samples it walks
#pragma unroll
        for (int q = 0; q < NB; q++) acc[q] = 0.0;
        double ll = 0.0, nfb = 0.0;
        const int cslot = dens ? S.slot[cc] : 0;
        const double *lam = S.lam;
        // T_l = sum_i E_il inv_i gives the column sums as lam_l T_l only while every weight is a positive number.  Once a
        // weight is NaN (the unguarded first E-step can poison them, :737-745) the reference's guarded refresh falls
        // back to uniform responsibilities and RECOVERS; lam_l T_l could not.  Such a pass accumulates the
        // responsibilities themselves (`direct`: the column holds w_il, the leader takes the sums as they are).
        bool direct = false;
        for (int l = 0; l < L; l++) direct |= !(lam[l] > 0.0);
        if (t == 0) S.keep = direct ? 1 : 0;
        for (int it = team; it < ntile_cta; it += NTEAM) {
          const unsigned long long seq = seq_base + (unsigned long long)it;
          const int st = (int)(seq % (unsigned)NS);
          double *xs = ring + st * stage_doubles, *Es = xs + d * kV2TS, *winv = Es + Lmax * kV2TS;
          const long tl = (long)bid + (long)it * G;
          const long i = tl * kV2TS + tt;
          const bool valid = i < n;
          // try_wait.parity can only tell the barrier's current phase from the one before it, so a team that runs
          // ahead must not look at the barrier before this fill has been issued (the stage's previous fill is then
          // complete and consumed: the barrier is in this fill's phase or just past it)
          {
            const unsigned fill = (unsigned)(seq / (unsigned)NS);
            if (lane == 0) {
              const long long t0 = clock64();
              while (s_gen[st] <= fill) {
                __nanosleep(40);
                if (clock64() - t0 > 20000000000LL) __trap();
              }
            }
            __syncwarp();
            mbar_wait_wd(&s_full[st], fill & 1u);
          }
          // ---- A-phase: thread tt owns sample tt of the stage
          double *Ecol = Es + tt;
          double xv[DMAX];
#pragma unroll
          for (int j = 0; j < DMAX; j++) xv[j] = xs[j * kV2TS + tt];
          if (dens_all) {
            const double rd = s_rec[AMX_REC_HEAD + d], c1 = s_rec[3];  // all start factors are sqrt(s2) I
            for (int l = 0; l < L; l++) {
              const double *mu = s_mu + l * d;
              double q = 0.0;
#pragma unroll
              for (int j = 0; j < DMAX; j++) {
                const double r = (xv[j] - mu[j]) * rd;
                q = fma(r, r, q);
              }
              const double e = exp(fma(-0.5, q, c1));
              Ecol[l * kV2TS] = e;
              if (valid) __stcg(rk_E + v2_e_at(i, S.slot[l], Lmax), e);
            }
          } else if (dens) {
            const double enew = exp(fma(-0.5, v2_solve_cols<DMAX>(s_rec, xv), s_rec[3]));
            Ecol[cc * kV2TS] = enew;
            if (valid) __stcg(rk_E + v2_e_at(i, cslot, Lmax), enew);
          }
          // sum_l lam_l E_il in component order (as the reference adds them)
          double sum;
          {
            double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;  // four chains: the sum moves by rounding only
            int l = 0;
            for (; l + 4 <= L; l += 4) {
              p0 = fma(lam[l], Ecol[l * kV2TS], p0);
              p1 = fma(lam[l + 1], Ecol[(l + 1) * kV2TS], p1);
              p2 = fma(lam[l + 2], Ecol[(l + 2) * kV2TS], p2);
              p3 = fma(lam[l + 3], Ecol[(l + 3) * kV2TS], p3);
            }
            for (; l < L; l++) p0 = fma(lam[l], Ecol[l * kV2TS], p0);
            sum = (p0 + p1) + (p2 + p3);
          }
          // The responsibilities enter the column sums as T_l += E_il * inv (B-phase).  The common case sets
          // inv = 1 / sum; every other case (first E-step, densities near the subnormal range, NaN) first rewrites the
          // sample's column of the stage to w_il / lam_l and uses inv = 1.
          double wn = 0.0, inv = 1.0;
          if (!valid) {  // padding sample: no weight anywhere (its cache column may hold anything)
            for (int l = 0; l < L; l++) Ecol[l * kV2TS] = 0.0;
          } else if (dens_all) {
            // First E-step: w = lam * pdf / sum with no look at the sum, as the reference does (:737-745); 0/0 = NaN
            // poisons the column sums and components are annihilated until a guarded refresh clears it.
            for (int l = 0; l < L; l++) {
              const double w = (lam[l] * Ecol[l * kV2TS]) / sum;
              Ecol[l * kV2TS] = w / lam[l];
              if (l == nx) wn = w;
            }
          } else if (sum >= 1e-280 && !direct) {  // the reference's guard (:855-866) holds, and 1/sum is safe
            inv = 1.0 / sum;
            ll += log(sum);
            wn = (lam[nx] * Ecol[nx * kV2TS]) * inv;
          } else if (sum >= 1e-280) {  // (unreachable while a weight is NaN -- the sum is NaN then -- kept for weights <= 0)
            const double is = 1.0 / sum;
            ll += log(sum);
            for (int l = 0; l < L; l++) Ecol[l * kV2TS] = (lam[l] * Ecol[l * kV2TS]) * is;
          } else {
            wn = v2_slow_sample(lam, s_mu, s_B, d, tri, L, nx, sum, direct, xs + tt, Ecol, &ll, &nfb);
          }
          winv[tt] = inv;
          (void)wn;  // the B-phase forms w_next = lam_nx E_i,nx inv itself (in the rewritten cases that is w to an ulp)
          // shift by the pivot (current mean of the next component): the rows become dx = x - pivot
          {
            const double *piv = s_piv;
#pragma unroll
            for (int j = 0; j < DMAX; j++) {
              const double v0 = xv[j] - piv[j];
              xs[j * kV2TS + tt] = valid ? v0 : 0.0;
            }
          }
          named_bar(1 + team, 128);
          // ---- B-phase: warp wq owns the entries e = wq (mod 4) of (S2 | S1) over all 128 samples
          {
            const double lam_nx = (nx < L) ? (direct ? 1.0 : lam[nx]) : 0.0;  // empty mixture: no weight
            v2_bphase_any<DMAX>(acc, wq, xs, Es, Es + nx * kV2TS, lam_nx, winv, lane, d, L);
          }
          // the stage is free once all four warps are through; the team refills it with the tile NS places ahead
          fence_proxy_async();
          named_bar(1 + team, 128);
          if (wq == 0 && it + NS < ntile_cta) fetch(it + NS);
        }
        // ---- end of pass: the teams' register accumulators -> this CTA's partial row, in a fixed order.
        // Lanes are first folded four to one by shuffles; the 8 x 4 NTEAM group sums per value go through the
        // (now idle) ring memory.
        named_bar(8, NCONS);  // every stage has been consumed
        constexpr int NW = NTEAM * 8;              // writers per entry: 8 four-lane groups of the owning warp of each team
        double *scr = ring;                        // [4 NB][NW]: entry e = 4 q + wq
        double *scr2 = ring + 4 * NB * NW;         // [2][4 NTEAM] log-likelihood and fallback count per warp
#pragma unroll
        for (int q = 0; q < NB; q++) {
          double r = acc[q];
          r += __shfl_xor_sync(0xffffffffu, r, 1);
          r += __shfl_xor_sync(0xffffffffu, r, 2);
          if ((lane & 3) == 0) scr[(4 * q + wq) * NW + team * 8 + (lane >> 2)] = r;
        }
        {
          const double r0 = warp_sum(ll), r1 = warp_sum(nfb);
          if (lane == 0) {
            scr2[warp] = r0;
            scr2[4 * NTEAM + warp] = r1;
          }
        }
        named_bar(8, NCONS);
        for (int e = warp; e < CF::NEALL; e += 4 * NTEAM) {
          double r = (lane < NW) ? scr[e * NW + lane] : 0.0;
          r = warp_sum(r);
          // entry e of the DMAX-packed (S2 | S1 | pad | T) -> position in the partial row [T | ll | nfb | S1 | S2 d-packed]
          // (rows j < d of the packed triangle are its first tri(d) entries in either packing)
          if (lane == 0) {
            if (e >= NE4) s_part[e - NE4] = r;
            else if (e >= TRI) {
              if (e - TRI < d) s_part[kEmLmax + 2 + (e - TRI)] = r;
            } else if (e < tri) {
              s_part[kEmLmax + 2 + d + e] = r;
            }
          }
        }
        if (warp == 0) {
          double r0 = (lane < 4 * NTEAM) ? scr2[lane] : 0.0, r1 = (lane < 4 * NTEAM) ? scr2[4 * NTEAM + lane] : 0.0;
          r0 = warp_sum(r0);
          r1 = warp_sum(r1);
          if (lane == 0) {
            s_part[kEmLmax] = r0;
            s_part[kEmLmax + 1] = r1;
          }
        }
      }
      seq_base += (unsigned long long)ntile_cta;
      nv = kEmLmax + 2 + d + tri;
      fence_proxy_async();  // the ring was used as plain scratch: order that before the next pass's bulk copies
      __syncthreads();
    }

    // ------------------------------------------------------------------ exchange: partial rows -> totals everywhere
    const long long tk1 = clock64();
    V2Sync *me = v.sync[rk_rank];
    for (int q = t; q < nv; q += blockDim.x) __stcg(rk_part + (size_t)q * G + bid, s_part[q]);
    __syncthreads();
    if (t == 0) {
      __threadfence();
      const unsigned prev = atomicAdd(&me->arrive, 1u);
      s_flag = (prev == (unsigned)G * (epoch + 1u) - 1u) ? 1 : 0;
      if (s_flag) __threadfence();
    }
    __syncthreads();
    const int par = (int)(epoch & 1u);
    const long long tx1 = clock64();
    if (s_flag) {
      // last CTA of this GPU: sum the GPU's partial rows (lanes over consecutive CTAs of one value: coalesced;
      // fixed order), then post the row to every GPU's inbox and raise this GPU's flag there
      {  // every load of this warp's values is issued before the first add: one L2 round trip, not one per value
        constexpr int NWARP = 4 * NTEAM, KV = (kV2NV + NWARP - 1) / NWARP, KB = 5;  // KB * 32 >= 148 CTAs
        double vv[KV][KB];
#pragma unroll
        for (int k = 0; k < KV; k++) {
          const int q = warp + NWARP * k;
#pragma unroll
          for (int m = 0; m < KB; m++) {
            const int b = lane + 32 * m;
            vv[k][m] = (q < nv && b < G) ? ld_cg(rk_part + (size_t)q * G + b) : 0.0;
          }
        }
#pragma unroll
        for (int k = 0; k < KV; k++) {
          const int q = warp + NWARP * k;
          double r = 0.0;
#pragma unroll
          for (int m = 0; m < KB; m++) r += vv[k][m];
          for (int b = lane + 32 * KB; b < G; b += 32) r += ld_cg(rk_part + (size_t)(q < nv ? q : 0) * G + b);  // larger grids
          r = warp_sum(r);
          if (lane == 0 && q < nv) s_tot[q] = r;
        }
      }
      __syncthreads();
      for (int q = t; q < nv * a.ndev; q += blockDim.x) {
        const int g = q / nv, k = q - g * nv;
        v.sync[g]->inbox[par][rk_rank][k] = s_tot[k];
      }
      if (a.ndev > 1) __threadfence_system();
      else __threadfence();
      __syncthreads();
      if (t < a.ndev) st_release_sys(&v.sync[t]->flag[rk_rank][0], epoch + 1u);
    }
    if (t < a.ndev) {  // every CTA: wait until every GPU's row of this epoch is in the local inbox
      const unsigned *f = &me->flag[t][0];
      const long long t0 = clock64();
      unsigned ns = 20;
      while (ld_acquire_sys(f) < epoch + 1u) {
        __nanosleep(ns);
        if (ns < 160) ns *= 2;
        if (clock64() - t0 > 40000000000LL) {  // ~20 s: a partner is gone
          S.status = AMX_ECUDA;
          S.stop = 2;
          break;
        }
      }
    }
    __syncthreads();
    const long long tx2 = clock64();
    dbg_x1 += tx1 - tk1;
    dbg_x2 += tx2 - tx1;
    if (S.stop == 2) {
      if (writer && t == 0) atomicExch(&ctrl->status, AMX_ECUDA);
      return;
    }
    for (int q = t; q < nv; q += blockDim.x) {  // rows in GPU order: the same sum on every CTA of every GPU
      double r = 0.0;
      for (int g = 0; g < a.ndev; g++) r += ld_cg(&me->inbox[par][g][q]);
      s_tot[q] = r;
    }
    epoch++;
    __syncthreads();
    const long long tk2 = clock64();
    if (warp == 0) v2_leader_warp<DMAX>(a, ctrl, writer, pass, s_tot, S, s_mu, s_B, s_rec, s_piv);
    __syncthreads();
    const long long tk3 = clock64();
    dbg_pass += tk1 - tk0;
    if (v.debug && t == 0) me->dbg_cta[bid < 160 ? bid : 159] = dbg_pass;
    dbg_bar += tk2 - tk1;
    dbg_lead += tk3 - tk2;
    pass = S.pass;
    if (pass == kPassStop) break;
  }

  // what the host reads back (GPU 0, CTA 0)
  if (writer) {
    if (t == 0) {
      ctrl->pass = S.pass; ctrl->L = S.L; ctrl->c = S.c; ctrl->next = S.next; ctrl->iters = S.iters; ctrl->stop = S.stop;
      ctrl->status = S.status; ctrl->best_L = S.best_L; ctrl->comp_steps = S.comp_steps; ctrl->flops = S.flops;
      ctrl->loglik = S.loglik; ctrl->cost = S.cost; ctrl->cost_prev = S.cost_prev; ctrl->cost_best = S.cost_best;
      ctrl->s2 = S.tmp[kEmLmax - 1];  // bytes requested (S.tmp's last entry is free: v2w_cost keeps its logs in registers)
      ctrl->dbg[0] = dbg_pass; ctrl->dbg[1] = dbg_bar; ctrl->dbg[3] = dbg_lead; ctrl->dbg[4] = dbg_x1; ctrl->dbg[5] = dbg_x2;
    }
    const int Lw = S.L < 0 ? 0 : S.L;
    if (t < kEmLmax) {
      ctrl->lam[t] = S.lam[t];
      ctrl->slot[t] = S.slot[t];
    }
    for (int q = t; q < Lw * d; q += blockDim.x) ctrl->mu[q / d][q % d] = s_mu[q];
    for (int q = t; q < Lw * tri; q += blockDim.x) ctrl->B[q / tri][q % tri] = s_B[q];
  }
  // optional dump of the responsibilities of the working state (step-parity tests)
  if (rk_wout != nullptr) {
    const int L = S.L;
    for (long i = (long)bid * blockDim.x + t; i < n; i += (long)G * blockDim.x) {
      double sum = 0.0;
      for (int l = 0; l < L; l++) sum += S.lam[l] * __ldcg(rk_E + v2_e_at(i, S.slot[l], Lmax));
      for (int l = 0; l < L; l++) {
        const double e = __ldcg(rk_E + v2_e_at(i, S.slot[l], Lmax));
        rk_wout[(size_t)i * a.Lmax + l] = (sum > 0) ? S.lam[l] * e / sum : 1.0 / L;
      }
    }
  }
}

template <int DMAX>
constexpr size_t v2_fixed_doubles(int d, int Lmax) {
  return (size_t)Lmax * d + (size_t)Lmax * (d * (d + 1) / 2) + (AMX_REC_HEAD + 2 * DMAX + V2Cfg<DMAX>::TRI) + DMAX + 1 + 2 * kV2NV +
         (sizeof(LeaderS<DMAX>) + 7) / 8 + 2;
}
template <int DMAX, int NTEAM>
constexpr size_t v2_scratch_doubles() {
  return (size_t)4 * V2Cfg<DMAX>::NB * (NTEAM * 8) + 8 * NTEAM + 64;
}
