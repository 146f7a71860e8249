/*
 * amx_oracle.c -- CPU restatement of LibAutoMix's sampling hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the parity checker for the CUDA
 * path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).  The
 * product library never links, loads or calls it.
 *
 * What it follows (all citations are to /root/reference/src/libautomix/automix.c):
 *   sdrand / sdrni .............. :1297-1316
 *   gauss / rt / rgamma ......... :1639-1661 / :1663-1680 / :1585-1637
 *   chol / det / lnormprob ...... :1682-1701 / :1752-1761 / :1727-1750
 *   perm / ltprob ............... :1703-1715 / :1717-1725
 *   rwm_within_model ............ :575-662
 *   fit_mixture_from_samples .... :664-1006
 *   fit_autorj .................. :1008-1033
 *   reversible_jump_move ........ :1035-1288
 *   initChain + sweep loops ..... :423-449, :77-155
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function
 * below bit-for-bit (same libm, same operation order, no FMA) against the
 * reference itself compiled from its own sources into oracle/_ref/ (see
 * oracle/Makefile), and tests/test_oracle_golden.py checks it against the
 * committed vectors in tests/golden/ that were produced from that build by
 * oracle/gen_golden.py.  The reference's own test-suite holds no per-step
 * vectors, so those are the pins.
 *
 * Layout conventions (differ from the reference on purpose -- flat, no nested
 * pointers): a lower-triangular factor B of order d is stored packed row-major,
 * entry (i,j), j<=i, at i*(i+1)/2 + j.  Mixture components of one model are
 * contiguous.  Samples are row-major n x d.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  -O2 without
 * -ffast-math keeps IEEE evaluation order, which is what makes "bit-for-bit" a
 * meaningful statement.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

typedef double (*orc_target_fn)(int model_k, double *x);

/* ------------------------------------------------------------------------- */
/* Uniform source: either an injected tape or the SuperDuper generator.       */
/* ------------------------------------------------------------------------- */

static const double *g_tape = NULL;
static long g_tape_len = 0, g_tape_pos = 0;
static int g_tape_overrun = 0;
static unsigned long g_cong, g_taus; /* reference's JC, JT (:1297) */

void orc_tape_set(const double *tape, long n) {
  g_tape = tape;
  g_tape_len = n;
  g_tape_pos = 0;
  g_tape_overrun = 0;
}
long orc_tape_used(void) { return g_tape_pos; }
int orc_tape_overrun(void) { return g_tape_overrun; }

/* :1307-1316 */
void orc_sdrni(unsigned long *seed) {
  unsigned long s = *seed;
  if (s == 0) s = (unsigned long)time(0);
  g_taus = s / 65536;
  g_cong = s - 65536 * g_taus;
  g_taus = 65536 * g_taus + 1;
  g_cong = 32768 * g_cong + 1;
  *seed = s;
}

/* :1300-1305 */
double orc_sdrand(void) {
  g_cong = (g_cong * 69069UL) & 0xFFFFFFFFUL;
  g_taus ^= g_taus >> 15;
  g_taus ^= (g_taus << 17) & 0xFFFFFFFFUL;
  return (double)((g_taus ^ g_cong) >> 1) * 4.656612873E-10;
}

static double unif(void) {
  if (g_tape != NULL) {
    if (g_tape_pos < g_tape_len) return g_tape[g_tape_pos++];
    g_tape_overrun = 1;
    g_tape_pos++;
    return 0.5;
  }
  return orc_sdrand();
}

/* ------------------------------------------------------------------------- */
/* Numeric helpers                                                            */
/* ------------------------------------------------------------------------- */

/* :1639-1661 -- Box-Muller, radius uniform first, angle uniform second; the
 * odd tail spends two uniforms for one normal (sine branch only). */
void orc_gauss(double *z, int n) {
  int i = 0;
  for (; i + 1 < n; i += 2) {
    double r = sqrt(-2.0 * log(unif()));
    double a = 2.0 * M_PI * unif();
    z[i] = r * sin(a);
    z[i + 1] = r * cos(a);
  }
  if (n % 2 == 1) {
    double r = sqrt(-2.0 * log(unif()));
    double a = 2.0 * M_PI * unif();
    z[n - 1] = r * sin(a);
  }
}

/* :1585-1637 -- Gamma(s,1) by rejection; three regimes. */
double orc_rgamma(double s) {
  const double e1 = exp(1.0);
  double out;
  if (s < 1.0) {
    double b = (s + e1) / e1, inv = 1.0 / s;
    for (;;) {
      double bu = b * unif();
      if (bu <= 1.0) {
        double t = inv * log(bu);
        out = exp(t < -30.0 ? -30.0 : t);
        if (unif() < exp(-out)) break;
      } else {
        out = -log((b - bu) / s);
        if (unif() < pow(out, s - 1.0)) break;
      }
    }
  } else if (s == 1.0) {
    out = -log(unif());
  } else {
    double c1 = s - 1.0;
    double c2 = (s - 1.0 / (6.0 * s)) / c1;
    double c3 = 2.0 / c1;
    double c4 = c3 + 2.0;
    double c5 = 1.0 / sqrt(s);
    double w;
    for (;;) {
      double u1 = unif();
      double u2 = unif();
      if (s > 2.5) u1 = u2 + c5 * (1.0 - 1.86 * u1);
      if (u1 <= 0.0 || u1 >= 1.0) continue;
      w = c2 * u2 / u1;
      if ((c3 * u1 + w + 1.0 / w) <= c4) break;
      if ((c3 * log(u1) - log(w) + w) >= 1.0) continue;
      break;
    }
    out = c1 * w;
  }
  return out;
}

/* :1663-1680 */
void orc_rt(double *z, int n, int dof) {
  orc_gauss(z, n);
  if (dof > 0) {
    double s = 0.5 * dof;
    double den = sqrt(orc_rgamma(s) / s);
    for (int i = 0; i < n; i++) z[i] /= den;
  }
}

/* :1703-1715 */
void orc_perm(double *v, int n) {
  for (int i = 0; i < n - 1; i++) {
    int j = i + (int)((n - i) * unif());
    if (j != i) {
      double t = v[j];
      v[j] = v[i];
      v[i] = t;
    }
  }
}

/* The reference carries its own Cody-Hillstrom log-gamma (:1323-1579).  It is
 * only reached through the optional Student-t mode (ltprob) and user targets;
 * the restatement uses libm's lgamma, which agrees with it to ~1e-15 relative
 * on the arguments that occur (checked in tests/test_oracle_vs_ref.py). */
double orc_loggamma(double x) { return lgamma(x); }

/* :1717-1725 */
double orc_ltprob(int dof, double z) {
  double c = orc_loggamma(0.5 * (dof + 1)) - orc_loggamma(0.5 * dof) -
             0.5 * log(dof * M_PI);
  return c - 0.5 * (dof + 1) * log(1.0 + pow(z, 2.0) / dof);
}

/* :1682-1701 -- in-place, column by column, packed storage. */
void orc_chol(int d, double *A) {
  for (int c = 0; c < d; c++) {
    double s = A[TRI(c, c)];
    for (int j = 0; j < c; j++) s -= pow(A[TRI(c, j)], 2);
    A[TRI(c, c)] = sqrt(s);
    for (int r = c + 1; r < d; r++) {
      s = A[TRI(r, c)];
      for (int j = 0; j < c; j++) s -= A[TRI(r, j)] * A[TRI(c, j)];
      A[TRI(r, c)] = s / A[TRI(c, c)];
    }
  }
}

/* :1752-1761 */
double orc_det(int d, const double *B) {
  double p = 1.0;
  for (int i = 0; i < d; i++) p *= B[TRI(i, i)];
  return p;
}

/* :1727-1750 */
double orc_lnormprob(int d, const double *mu, const double *B, const double *x) {
  double r[d > 0 ? d : 1];
  for (int i = 0; i < d; i++) r[i] = x[i] - mu[i];
  for (int i = 0; i < d; i++) {
    for (int j = 0; j < i; j++) r[i] -= B[TRI(i, j)] * r[j];
    r[i] /= B[TRI(i, i)];
  }
  double q = 0.0;
  for (int i = 0; i < d; i++) q += r[i] * r[i];
  return -0.5 * q - (d / 2.0) * log(2.0 * M_PI) - log(orc_det(d, B));
}

static double dmax(double a, double b) { return a > b ? a : b; } /* macro :10 */
static double dmin(double a, double b) { return a < b ? a : b; } /* macro :11 */

/* ------------------------------------------------------------------------- */
/* Stage 1: adaptive random-walk Metropolis within one model  (:575-662)      */
/* ------------------------------------------------------------------------- */
/*
 * samples_out: (1000*d) x d row-major.  sig_trace / acc_trace: one row of d per
 * 100 sweeps (may be NULL).  Returns the number of sweeps performed.
 */
int orc_rwm_within_model(int model_k, int d, int nsweep2, int dof,
                         orc_target_fn f, const double *init, double *sig,
                         double *samples_out, double *sig_trace,
                         double *acc_trace, double *final_state,
                         double *final_lp) {
  int nsweepr = nsweep2 > 10000 * d ? nsweep2 : 10000 * d;
  int nburn = nsweepr / 10;
  const double target_acc = 0.25;
  nsweepr += nburn;
  double *cur = malloc(sizeof(double) * d), *prop = malloc(sizeof(double) * d);
  double *z = malloc(sizeof(double) * d);
  int *nacc = calloc(d, sizeof(int)), *ntry = calloc(d, sizeof(int));
  for (int i = 0; i < d; i++) {
    cur[i] = prop[i] = init[i];
    sig[i] = 10.0;
  }
  double lp = f(model_k, cur);
  int nstored = 0, ntrace = 0, remain = nsweepr;
  for (int sweep = 1; sweep <= nsweepr; sweep++) {
    remain--;
    double u = unif();
    if (sweep > nburn && u < 0.1) {
      orc_rt(z, d, dof);
      for (int i = 0; i < d; i++) prop[i] = cur[i] + sig[i] * z[i];
      double lpn = f(model_k, prop);
      if (unif() < exp(dmax(-30.0, dmin(0.0, lpn - lp)))) {
        for (int i = 0; i < d; i++) cur[i] = prop[i];
        lp = lpn;
      }
    } else {
      double gam = 10.0 * pow(1.0 / (sweep + 1), 2.0 / 3.0);
      for (int i = 0; i < d; i++) prop[i] = cur[i];
      for (int i = 0; i < d; i++) {
        double zz;
        orc_rt(&zz, 1, dof);
        prop[i] = cur[i] + sig[i] * zz;
        double lpn = f(model_k, prop);
        double acc = dmin(1, exp(dmax(-30.0, dmin(0.0, lpn - lp))));
        if (unif() < acc) {
          nacc[i]++;
          ntry[i]++;
          cur[i] = prop[i];
          lp = lpn;
          sig[i] = dmax(0, sig[i] - gam * (target_acc - 1));
        } else {
          ntry[i]++;
          prop[i] = cur[i];
          sig[i] = dmax(0, sig[i] - gam * (target_acc));
        }
      }
    }
    if (remain < 10000 * d && remain % 10 == 0) {
      for (int i = 0; i < d; i++) samples_out[(long)nstored * d + i] = cur[i];
      nstored++;
    }
    if (sweep % 100 == 0) {
      if (sig_trace)
        for (int i = 0; i < d; i++) sig_trace[(long)ntrace * d + i] = sig[i];
      if (acc_trace)
        for (int i = 0; i < d; i++)
          acc_trace[(long)ntrace * d + i] = (double)nacc[i] / (double)ntry[i];
      ntrace++;
    }
  }
  if (final_state)
    for (int i = 0; i < d; i++) final_state[i] = cur[i];
  if (final_lp) *final_lp = lp;
  free(cur);
  free(prop);
  free(z);
  free(nacc);
  free(ntry);
  return nsweepr;
}

/* ------------------------------------------------------------------------- */
/* Stage 2: Figueiredo-Jain component-wise EM with annihilation (:664-1006)   */
/* ------------------------------------------------------------------------- */

/* responsibilities + log-likelihood refresh, :847-867 (and :931-951) */
static double em_refresh(int n, int L, int Lmax, const double *lam,
                         const double *lpd, double *w) {
  double loglik = 0.0;
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int l = 0; l < L; l++) {
      double t = log(lam[l]) + lpd[(long)i * Lmax + l];
      w[(long)i * Lmax + l] = exp(t);
      s += w[(long)i * Lmax + l];
    }
    if (s > 0) {
      for (int l = 0; l < L; l++) w[(long)i * Lmax + l] /= s;
      loglik += log(s);
    } else {
      for (int l = 0; l < L; l++) w[(long)i * Lmax + l] = 1.0 / L;
      loglik -= 500.0;
    }
  }
  return loglik;
}

/* drop component `gone`, shifting the later ones down (:823-836, :908-921) */
static void em_drop(int n, int d, int L, int Lmax, int gone, double *lam,
                    double *mu, double *B, double *lpd) {
  int tri = d * (d + 1) / 2;
  for (int l = gone; l < L - 1; l++) {
    lam[l] = lam[l + 1];
    memcpy(mu + (long)l * d, mu + (long)(l + 1) * d, sizeof(double) * d);
    memcpy(B + (long)l * tri, B + (long)(l + 1) * tri, sizeof(double) * tri);
    for (int i = 0; i < n; i++)
      lpd[(long)i * Lmax + l] = lpd[(long)i * Lmax + l + 1];
  }
}

static void renorm(int L, double *lam) {
  double s = 0.0;
  for (int l = 0; l < L; l++) s += lam[l];
  for (int l = 0; l < L; l++) lam[l] /= s;
}

/* MML cost, :870-876 */
static double em_cost(int n, int L, int nparams, const double *lam,
                      double loglik) {
  double s = 0.0;
  for (int l = 0; l < L; l++) s += log(n * lam[l] / 12.0);
  return (nparams / 2.0) * s + (L / 2.0) * log(n / 12.0) +
         L * (nparams + 1) / 2.0 - loglik;
}

/*
 * In/out: lam[Lmax], mu[Lmax*d], B[Lmax*tri] receive the minimum-cost mixture
 * (first *L_out components valid).  init_idx (optional, Lmax ints): if not NULL
 * the indices of the data rows used as initial means are returned there.
 * Traces need capacity maxit+1 (the reference runs maxit+1 outer iterations
 * when the cap binds: `count > NUM_FITMIX_MAX`, :961).
 * cur_* (optional): the *working* state when the loop stopped (not the best).
 */
int orc_fit_mixture(int d, int n, const double *x, int Lmax, int maxit,
                    double *lam, double *mu, double *B, int *L_out,
                    int *trace_L, double *trace_loglik, double *trace_cost,
                    int *trace_ann, int *init_idx, double *cur_lam,
                    double *cur_mu, double *cur_B, int *cur_L, double *cur_w,
                    double *cur_lpd) {
  int tri = d * (d + 1) / 2;
  int L = Lmax;
  int *start = malloc(sizeof(int) * Lmax);

  /* :682-697 distinct random rows */
  for (int l = 0; l < L;) {
    start[l] = (int)floor(n * unif());
    int dup = 0;
    for (int m = 0; m < l; m++)
      if (start[m] == start[l]) {
        dup = 1;
        break;
      }
    if (!dup) l++;
  }
  if (init_idx) memcpy(init_idx, start, sizeof(int) * Lmax);

  /* :700-711 common isotropic start: mean per-dimension variance / 10 */
  double s2 = 0.0;
  for (int j = 0; j < d; j++) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < n; i++) {
      a += x[(long)i * d + j];
      b += x[(long)i * d + j] * x[(long)i * d + j];
    }
    double len = (double)n;
    s2 += (b - a * a / len) / len;
  }
  s2 /= (10.0 * d);

  /* :713-723 */
  for (int l = 0; l < L; l++) {
    double *Bl = B + (long)l * tri;
    for (int j = 0; j < d; j++) {
      mu[(long)l * d + j] = x[(long)start[l] * d + j];
      Bl[TRI(j, j)] = s2;
      for (int k = 0; k < j; k++) Bl[TRI(j, k)] = 0.0;
    }
    orc_chol(d, Bl);
    lam[l] = 1.0 / L;
  }

  double *w = malloc(sizeof(double) * (size_t)n * Lmax);
  double *lpd = malloc(sizeof(double) * (size_t)n * Lmax);
  double *colsum = malloc(sizeof(double) * Lmax);

  /* :733-744 initial E-step (no underflow guard in the reference) */
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int l = 0; l < L; l++) {
      lpd[(long)i * Lmax + l] =
          orc_lnormprob(d, mu + (long)l * d, B + (long)l * tri, x + (long)i * d);
      double t = log(lam[l]) + lpd[(long)i * Lmax + l];
      w[(long)i * Lmax + l] = exp(t);
      s += w[(long)i * Lmax + l];
    }
    for (int l = 0; l < L; l++) w[(long)i * Lmax + l] /= s;
  }

  int nparams = d + (d * (d + 1)) / 2;
  double *best_lam = malloc(sizeof(double) * Lmax);
  double *best_mu = malloc(sizeof(double) * (size_t)Lmax * d);
  double *best_B = malloc(sizeof(double) * (size_t)Lmax * tri);
  int best_L = 0;
  double cost_prev = 0.0, cost_best = 0.0, loglik = 0.0, wkeep = 0.0;
  int iters = 0, stop = 0;

  while (!stop) {
    iters++;
    int natural = 0, forced = 0;
    int c = 0;
    while (c < L) {
      /* :773-792 weight of component c from all column sums, then renormalise */
      double tot = 0.0;
      for (int l = 0; l < L; l++) {
        colsum[l] = 0.0;
        for (int i = 0; i < n; i++) colsum[l] += w[(long)i * Lmax + l];
        double wl = dmax(0.0, (colsum[l] - nparams / 2.0));
        if (l == c) wkeep = wl;
        tot += wl;
      }
      lam[c] = wkeep / tot;
      renorm(L, lam);

      if (lam[c] > 0.005) {
        /* :796-811 mean, then centred scatter row by row */
        double *m = mu + (long)c * d, *Bc = B + (long)c * tri;
        for (int j = 0; j < d; j++) {
          m[j] = 0.0;
          for (int i = 0; i < n; i++)
            m[j] += x[(long)i * d + j] * w[(long)i * Lmax + c];
          m[j] /= colsum[c];
          for (int k = 0; k <= j; k++) {
            double t = 0.0;
            for (int i = 0; i < n; i++)
              t += (x[(long)i * d + j] - m[j]) * (x[(long)i * d + k] - m[k]) *
                   w[(long)i * Lmax + c];
            Bc[TRI(j, k)] = t / colsum[c];
          }
        }
        orc_chol(d, Bc);
        for (int i = 0; i < n; i++)
          lpd[(long)i * Lmax + c] = orc_lnormprob(d, m, Bc, x + (long)i * d);
        c++;
      } else {
        /* :821-845 natural annihilation */
        natural = 1;
        if (c < L - 1) em_drop(n, d, L, Lmax, c, lam, mu, B, lpd);
        L--;
        renorm(L, lam);
      }
      loglik = em_refresh(n, L, Lmax, lam, lpd, w);
    }

    double cost = em_cost(n, L, nparams, lam, loglik);
    if (iters == 1) cost_prev = cost;
    if (iters == 1 || cost < cost_best) { /* :881-893 */
      best_L = L;
      cost_best = cost;
      memcpy(best_lam, lam, sizeof(double) * L);
      memcpy(best_mu, mu, sizeof(double) * (size_t)L * d);
      memcpy(best_B, B, sizeof(double) * (size_t)L * tri);
    }
    if (fabs(cost_prev - cost) < dmin(1E-5 * fabs(cost_prev), 0.01) &&
        iters > 1) { /* :894-960 */
      if (L == 1) {
        stop = 1;
      } else {
        forced = 2;
        double lo = lam[0];
        int gone = 0;
        for (int l = 1; l < L; l++)
          if (lo > lam[l]) {
            lo = lam[l];
            gone = l;
          }
        if (gone < L - 1) em_drop(n, d, L, Lmax, gone, lam, mu, B, lpd);
        L--;
        renorm(L, lam);
        loglik = em_refresh(n, L, Lmax, lam, lpd, w);
        cost = em_cost(n, L, nparams, lam, loglik);
      }
    }
    if (iters > maxit) stop = 1; /* :961-963 */
    cost_prev = cost;
    trace_ann[iters - 1] = natural + forced;
    trace_cost[iters - 1] = cost;
    trace_loglik[iters - 1] = loglik;
    trace_L[iters - 1] = L;
  }

  if (cur_lam) memcpy(cur_lam, lam, sizeof(double) * Lmax);
  if (cur_mu) memcpy(cur_mu, mu, sizeof(double) * (size_t)Lmax * d);
  if (cur_B) memcpy(cur_B, B, sizeof(double) * (size_t)Lmax * tri);
  if (cur_L) *cur_L = L;
  if (cur_w) memcpy(cur_w, w, sizeof(double) * (size_t)n * Lmax);
  if (cur_lpd) memcpy(cur_lpd, lpd, sizeof(double) * (size_t)n * Lmax);

  /* :982-993 */
  *L_out = best_L;
  memcpy(lam, best_lam, sizeof(double) * best_L);
  memcpy(mu, best_mu, sizeof(double) * (size_t)best_L * d);
  memcpy(B, best_B, sizeof(double) * (size_t)best_L * tri);

  free(start);
  free(w);
  free(lpd);
  free(colsum);
  free(best_lam);
  free(best_mu);
  free(best_B);
  return iters;
}

/* :1008-1033 single Gaussian ("AutoRJ") */
void orc_fit_autorj(int d, int n, const double *x, double *lam, double *mu,
                    double *B) {
  lam[0] = 1.0;
  for (int j = 0; j < d; j++) {
    mu[j] = 0.0;
    for (int i = 0; i < n; i++) mu[j] += x[(long)i * d + j];
    mu[j] /= ((double)n);
  }
  for (int r = 0; r < d; r++)
    for (int c = 0; c <= r; c++) {
      double t = 0.0;
      for (int i = 0; i < n; i++)
        t += (x[(long)i * d + r] - mu[r]) * (x[(long)i * d + c] - mu[c]);
      B[TRI(r, c)] = t / ((double)(n - 1));
    }
  orc_chol(d, B);
}

/* ------------------------------------------------------------------------- */
/* Stage 3: reversible-jump sweeps (:1035-1288, loops :77-155, init :423-449) */
/* ------------------------------------------------------------------------- */

typedef struct {
  int nmodels;
  const int *dims, *ncomp;
  const double *lam, *mu, *B, *sig;
  /* offsets per model into the flat arrays */
  long olam[64], omu[64], oB[64], osig[64];
  int dmax, Lmax;
} flatmix;

static int flatmix_bind(flatmix *m, int nmodels, const int *dims,
                        const int *ncomp, const double *lam, const double *mu,
                        const double *B, const double *sig) {
  if (nmodels > 64) return -1;
  m->nmodels = nmodels;
  m->dims = dims;
  m->ncomp = ncomp;
  m->lam = lam;
  m->mu = mu;
  m->B = B;
  m->sig = sig;
  long a = 0, b = 0, c = 0, e = 0;
  m->dmax = 0;
  m->Lmax = 0;
  for (int k = 0; k < nmodels; k++) {
    m->olam[k] = a;
    m->omu[k] = b;
    m->oB[k] = c;
    m->osig[k] = e;
    a += ncomp[k];
    b += (long)ncomp[k] * dims[k];
    c += (long)ncomp[k] * (dims[k] * (dims[k] + 1) / 2);
    e += dims[k];
    if (dims[k] > m->dmax) m->dmax = dims[k];
    if (ncomp[k] > m->Lmax) m->Lmax = ncomp[k];
  }
  return 0;
}

/* chain state, the flat mirror of chainState (automix.h:113-127) */
typedef struct {
  double *theta; /* dmax */
  double *pk;    /* nmodels */
  double lp;
  int k;
  int nreinit;
  double pkllim;
  unsigned long sweep_i;
} chain;

typedef struct {
  unsigned long acc_block, try_block, acc_single, try_single, acc_jump,
      try_jump;
} counters;

/* allocation probabilities, :1094-1110 / :1217-1232 */
static void alloc_probs(const flatmix *m, int k, const double *pt, double *p) {
  int d = m->dims[k], L = m->ncomp[k], tri = d * (d + 1) / 2;
  double s = 0.0;
  for (int l = 0; l < L; l++) {
    p[l] = log(m->lam[m->olam[k] + l]) +
           orc_lnormprob(d, m->mu + m->omu[k] + (long)l * d,
                         m->B + m->oB[k] + (long)l * tri, pt);
    p[l] = exp(p[l]);
    s += p[l];
  }
  if (s > 0) {
    for (int l = 0; l < L; l++) p[l] /= s;
  } else {
    for (int l = 0; l < L; l++) p[l] = 1.0 / L;
  }
}

static int pick(const double *p, int n) { /* cumulative search, default 0 */
  double u = unif(), t = 0.0;
  for (int i = 0; i < n; i++) {
    t += p[i];
    if (u < t) return i;
  }
  return 0;
}

static void rj_move(const flatmix *m, chain *ch, int block, int do_perm,
                    int do_adapt, int burning, int dof, orc_target_fn f,
                    counters *ct, double *thn, double *wk, double *pa,
                    double *pan, double *z) {
  const double half_log_2pi = 0.9189385332046727; /* literal at :1052 */
  int k = ch->k, d = m->dims[k], L = m->ncomp[k];
  double *th = ch->theta;
  const double *sg = m->sig + m->osig[k];

  /* within-model RWM, :1056-1085 */
  if (block) {
    ct->try_block++;
    orc_rt(z, d, dof);
    for (int i = 0; i < d; i++) thn[i] = th[i] + sg[i] * z[i];
    double lpn = f(k, thn);
    if (unif() < exp(dmax(-30.0, dmin(0.0, lpn - ch->lp)))) {
      ct->acc_block++;
      memcpy(th, thn, sizeof(double) * d);
      ch->lp = lpn;
    }
  } else {
    memcpy(thn, th, sizeof(double) * d);
    for (int j = 0; j < d; j++) {
      ct->try_single++;
      double zz;
      orc_rt(&zz, 1, dof);
      thn[j] = th[j] + sg[j] * zz;
      double lpn = f(k, thn);
      if (unif() < exp(dmax(-30.0, dmin(0.0, lpn - ch->lp)))) {
        ct->acc_single++;
        th[j] = thn[j];
        ch->lp = lpn;
      } else {
        thn[j] = th[j];
      }
    }
  }

  /* allocate the current point to a component, :1090-1123 */
  int l = 0, ln = 0;
  ct->try_jump++;
  if (L > 1) {
    alloc_probs(m, k, th, pa);
    l = pick(pa, L);
  } else {
    pa[0] = 1.0;
  }

  /* standardise, :1127-1135 */
  {
    int tri = d * (d + 1) / 2;
    const double *mu = m->mu + m->omu[k] + (long)l * d;
    const double *B = m->B + m->oB[k] + (long)l * tri;
    for (int i = 0; i < d; i++) wk[i] = th[i] - mu[i];
    for (int i = 0; i < d; i++) {
      for (int j = 0; j < i; j++) wk[i] = wk[i] - B[TRI(i, j)] * wk[j];
      wk[i] = wk[i] / B[TRI(i, i)];
    }
  }

  /* target model and component, :1138-1169 */
  int kn = 0;
  double gam = 0.0, lr;
  if (m->nmodels == 1) {
    kn = k;
    lr = 0.0;
  } else {
    gam = pow(1.0 / (ch->sweep_i + 1), (2.0 / 3.0));
    kn = pick(ch->pk, m->nmodels);
    lr = log(ch->pk[k]) - log(ch->pk[kn]);
  }
  int dn = m->dims[kn], Ln = m->ncomp[kn];
  ln = pick(m->lam + m->olam[kn], Ln);

  /* dimension matching, :1173-1204 */
  if (d < dn) {
    orc_rt(wk + d, dn - d, dof);
    if (dof > 0) {
      for (int i = d; i < dn; i++) lr -= orc_ltprob(dof, wk[i]);
    } else {
      for (int i = d; i < dn; i++) lr += 0.5 * pow(wk[i], 2.0) + half_log_2pi;
    }
    if (do_perm) orc_perm(wk, dn);
  } else if (d == dn) {
    if (do_perm) orc_perm(wk, d);
  } else {
    if (do_perm) orc_perm(wk, d);
    if (dof > 0) {
      for (int i = dn; i < d; i++) lr += orc_ltprob(dof, wk[i]);
    } else {
      for (int i = dn; i < d; i++) lr -= (0.5 * pow(wk[i], 2.0) + half_log_2pi);
    }
  }

  /* map through the target component, :1206-1211 */
  int trin = dn * (dn + 1) / 2;
  const double *mun = m->mu + m->omu[kn] + (long)ln * dn;
  const double *Bn = m->B + m->oB[kn] + (long)ln * trin;
  for (int i = 0; i < dn; i++) {
    thn[i] = mun[i];
    for (int j = 0; j <= i; j++) thn[i] += Bn[TRI(i, j)] * wk[j];
  }

  /* reverse allocation, :1216-1235 */
  if (Ln > 1) {
    alloc_probs(m, kn, thn, pan);
  } else {
    pan[ln] = 1.0;
  }

  /* acceptance, :1238-1256 */
  double lpn = f(kn, thn);
  {
    int tri = d * (d + 1) / 2;
    lr += (lpn - ch->lp);
    lr += (log(pan[ln]) - log(pa[l]));
    lr += (log(m->lam[m->olam[k] + l]) - log(m->lam[m->olam[kn] + ln]));
    lr += (log(orc_det(dn, Bn)) -
           log(orc_det(d, m->B + m->oB[k] + (long)l * tri)));
  }
  if (unif() < exp(dmax(-30.0, dmin(0.0, lr)))) {
    for (int i = 0; i < dn; i++) th[i] = thn[i];
    ch->lp = lpn;
    ch->k = kn;
    ct->acc_jump++;
  }

  /* jump-probability adaptation, :1258-1282 */
  if (do_adapt && !burning) {
    for (int j = 0; j < m->nmodels; j++) {
      double e = (j == ch->k) ? 1.0 : 0.0;
      ch->pk[j] += (gam * (e - ch->pk[j]));
    }
    int reset = 0;
    for (int j = 0; j < m->nmodels; j++)
      if (ch->pk[j] < ch->pkllim) {
        reset = 1;
        break;
      }
    if (reset) {
      ch->nreinit++;
      ch->pkllim = 1.0 / (10.0 * ch->nreinit);
      for (int j = 0; j < m->nmodels; j++) ch->pk[j] = 1.0 / m->nmodels;
    }
  }
}

/*
 * Chain start exactly as initChain (:423-449): one uniform picks the model.
 * init_flat is the concatenation of the per-model start vectors.
 */
int orc_chain_init(int nmodels, const int *dims, const double *init_flat,
                   orc_target_fn f, double *theta, double *pk, double *lp,
                   int *k, int *nreinit, double *pkllim,
                   unsigned long *sweep_i) {
  int k0 = (int)floor(nmodels * unif());
  long off = 0;
  for (int j = 0; j < k0; j++) off += dims[j];
  for (int i = 0; i < dims[k0]; i++) theta[i] = init_flat[off + i];
  *lp = f(k0, theta);
  for (int j = 0; j < nmodels; j++) pk[j] = 1.0 / nmodels;
  *k = k0;
  *nreinit = 1;
  *pkllim = 1.0 / 10.0;
  *sweep_i = 1;
  return 0;
}

/*
 * Run nsweeps sweeps of one chain (the loops at :90-125 / :145-152).
 * State is in/out.  Traces (any may be NULL), one row per sweep:
 *   tr_k[n], tr_lp[n], tr_theta[n*dmax], tr_pk[n*nmodels].
 * cnt6: acc_block, try_block, acc_single, try_single, acc_jump, try_jump (added to).
 * visits[nmodels] is incremented with the model index after each sweep.
 */
int orc_rj_sweeps(int nmodels, const int *dims, const int *ncomp,
                  const double *lam, const double *mu, const double *B,
                  const double *sig, orc_target_fn f, long nsweeps,
                  int burning, int do_adapt, int do_perm, int dof,
                  double *theta, double *pk, double *lp, int *k, int *nreinit,
                  double *pkllim, unsigned long *sweep_i, int *tr_k,
                  double *tr_lp, double *tr_theta, double *tr_pk,
                  unsigned long *cnt6, long *visits) {
  flatmix m;
  if (flatmix_bind(&m, nmodels, dims, ncomp, lam, mu, B, sig)) return -1;
  chain ch = {theta, pk, *lp, *k, *nreinit, *pkllim, *sweep_i};
  counters ct = {0, 0, 0, 0, 0, 0};
  double *thn = malloc(sizeof(double) * m.dmax);
  double *wk = malloc(sizeof(double) * m.dmax);
  double *z = malloc(sizeof(double) * m.dmax);
  double *pa = malloc(sizeof(double) * m.Lmax);
  double *pan = malloc(sizeof(double) * m.Lmax);
  for (long s = 0; s < nsweeps; s++, ch.sweep_i++) {
    int block = (ch.sweep_i % 10 == 0);
    rj_move(&m, &ch, block, do_perm, do_adapt, burning, dof, f, &ct, thn, wk,
            pa, pan, z);
    if (visits) visits[ch.k]++;
    if (tr_k) tr_k[s] = ch.k;
    if (tr_lp) tr_lp[s] = ch.lp;
    if (tr_theta) {
      int d = dims[ch.k];
      for (int i = 0; i < m.dmax; i++)
        tr_theta[s * m.dmax + i] = i < d ? ch.theta[i] : 0.0;
    }
    if (tr_pk)
      for (int j = 0; j < nmodels; j++) tr_pk[s * nmodels + j] = ch.pk[j];
  }
  *lp = ch.lp;
  *k = ch.k;
  *nreinit = ch.nreinit;
  *pkllim = ch.pkllim;
  *sweep_i = ch.sweep_i;
  if (cnt6) {
    cnt6[0] += ct.acc_block;
    cnt6[1] += ct.try_block;
    cnt6[2] += ct.acc_single;
    cnt6[3] += ct.try_single;
    cnt6[4] += ct.acc_jump;
    cnt6[5] += ct.try_jump;
  }
  free(thn);
  free(wk);
  free(z);
  free(pa);
  free(pan);
  return 0;
}

/* batched mixture log-density helper for the K4 parity tests:
 * comp[n*L] (may be NULL) and mix[n] = log sum_l lam_l N(x; mu_l, B_l B_l^T),
 * the latter computed the way the reference forms it (plain exp/sum/log,
 * :847-859). */
void orc_mix_logpdf(int d, int L, const double *lam, const double *mu,
                    const double *B, long n, const double *x, double *comp,
                    double *mix) {
  int tri = d * (d + 1) / 2;
  for (long i = 0; i < n; i++) {
    double s = 0.0;
    for (int l = 0; l < L; l++) {
      double v = orc_lnormprob(d, mu + (long)l * d, B + (long)l * tri,
                               x + i * d);
      if (comp) comp[i * L + l] = v;
      s += exp(log(lam[l]) + v);
    }
    if (mix) mix[i] = log(s);
  }
}

/* ---- posterior summaries -------------------------------------------------
 * Integrated autocorrelation time by Sokal's adaptive truncated periodogram, as the reference's
 * report writer computes it for the model-index series (user_examples/logwrite.c:354-403):
 *   1. forward DFT of x, power spectrum, DC term set to zero (= remove the mean) (:369-377);
 *   2. a second forward DFT turns the spectrum into n * circular autocovariance (:379);
 *      var = acov[0] / (n (n-1)) (:380), rho[t] = acov[t] / acov[0] (:381-385);
 *   3. window: sum = -1/3; for i = 0..n-1: sum += rho[i] - 1/6, stop at the first sum < 0;
 *      tau = 2 (sum + i/6), m = i + 1 (:390-401) -- twice Sokal's definition, and with i = n,
 *      m = n + 1 when the sum never turns negative (a constant series gives NaN and lands there).
 * The reference's transform is a radix-4 routine (:405-651); a DFT is a DFT, so this restatement
 * uses a plain iterative radix-2 one.  Results agree to rounding (1e-10 relative in the tests,
 * m exactly).  n must be a power of two >= 4 (the reference prints a message and returns
 * otherwise, :424-439). */
static void orc_fft_pow2(long n, double *re, double *im) {
  for (long i = 1, j = 0; i < n; i++) {
    long bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  for (long len = 2; len <= n; len <<= 1) {
    long half = len >> 1;
    for (long k = 0; k < half; k++) {
      double ang = -6.283185307179586476925 * (double)k / (double)len;
      double wr = cos(ang), wi = sin(ang);
      for (long s = k; s < n; s += len) {
        long q = s + half;
        double tr = re[q] * wr - im[q] * wi, ti = re[q] * wi + im[q] * wr;
        re[q] = re[s] - tr; im[q] = im[s] - ti;
        re[s] += tr; im[s] += ti;
      }
    }
  }
}

int orc_sokal(long n, const double *x, double *var, double *tau, int *m) {
  if (n < 4 || (n & (n - 1)) || n > (1L << 20)) return -1;
  double *re = (double *)malloc(sizeof(double) * n), *im = (double *)calloc(n, sizeof(double));
  for (long i = 0; i < n; i++) re[i] = x[i];
  orc_fft_pow2(n, re, im);
  for (long i = 0; i < n; i++) {
    re[i] = re[i] * re[i] + im[i] * im[i];
    im[i] = 0.0;
  }
  re[0] = 0.0;
  orc_fft_pow2(n, re, im);
  *var = re[0] / ((double)n * (double)(n - 1));
  double c = 1.0 / re[0], sum = -(1.0 / 3.0);
  long i;
  for (i = 0; i < n; i++) {
    sum += re[i] * c - (1.0 / 6.0);
    if (sum < 0.0) break;
  }
  *tau = 2.0 * (sum + (double)i / 6.0);
  *m = (int)i + 1;
  free(re);
  free(im);
  return 0;
}

/* Per-model posterior moments of a set of draws (what a user forms from runStats.theta_summary,
 * automix.c:105-120): count, mean and unbiased covariance of the rows with k[i] == model. */
long orc_model_moments(long n, int dmax, const int *k, const double *theta, int model, int d,
                       double *mean, double *cov) {
  long cnt = 0;
  for (int j = 0; j < d; j++) mean[j] = 0.0;
  for (long i = 0; i < n; i++)
    if (k[i] == model) {
      cnt++;
      for (int j = 0; j < d; j++) mean[j] += theta[i * dmax + j];
    }
  if (cnt == 0) return 0;
  for (int j = 0; j < d; j++) mean[j] /= (double)cnt;
  for (int a = 0; a < d * d; a++) cov[a] = 0.0;
  for (long i = 0; i < n; i++)
    if (k[i] == model)
      for (int a = 0; a < d; a++)
        for (int b = 0; b < d; b++)
          cov[a * d + b] += (theta[i * dmax + a] - mean[a]) * (theta[i * dmax + b] - mean[b]);
  for (int a = 0; a < d * d; a++) cov[a] /= (double)(cnt > 1 ? cnt - 1 : 1);
  return cnt;
}
