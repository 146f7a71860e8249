// amx_rj_kernels.cuh -- K3 kernel templates and their launchers, in a header so that a plug-in translation unit
// (automix_b200/csrc/amx_plugin_tu.cu: a user's __device__ log-posterior compiled against this library) instantiates
// exactly the kernels the built-in families use.  The host driver (amx_rj_* of include/amx.h) is amx_rj.cu.
#pragma once

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <utility>
#include <vector>

#include "amx_internal.cuh"
#include "amx_mailbox.cuh"
#include "amx_rj.cuh"

namespace amx {


__global__ void rj_gamma_kernel(double *g, unsigned long long sweep0, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // gamma = pow(1.0 / (sweep_i + 1), 2.0 / 3.0)   (automix.c:1145)
  if (i < n) g[i] = pow(1.0 / (double)(sweep0 + (unsigned long long)i + 1ull), (2.0 / 3.0));
}

// ---- population pk adaptation ------------------------------------------------------------------------------
// The reference adapts the model-jump probabilities of ITS one chain by pk += gamma_t (1[k_t] - pk) after every
// sweep (automix.c:1258-1282).  Run as thousands of short chains that rule is biased at finite time -- a chain's
// pk is correlated with the chain's own recent path (measured with the reference itself, oracle/ref_population.c:
// 500 chains x (2000 + 2000) sweeps give P(k=5) = 0.102 +- 0.002 on the coal-mining posterior where its single
// long chain gives 0.115 and the same 500 chains without adaptation 0.118).  In the population mode every chain
// proposes from ONE shared pk, and the indicator of the reference's rule is replaced by the population's occupancy:
// the sweep kernels already accumulate the model-visit histogram with warp-aggregated atomics; after a segment of m
// sweeps this kernel applies the m per-sweep updates with that segment's visit fractions f in place of 1[k],
//     pk <- prod(1 - gamma_t) pk + (1 - prod(1 - gamma_t)) f,
// followed by the reference's re-initialisation rule (any pk < pkllim -> uniform, pkllim = 1/(10 nreinit)).
// Within a segment pk is constant, so every sweep is a Metropolis-Hastings kernel that leaves the posterior
// invariant: the estimator is unbiased whatever the segment's statistics are.
struct RjPkShared {
  double pk[AMX_MAX_MODELS];
  double pkllim;
  int nreinit;
  unsigned long long prev[AMX_MAX_MODELS];  // visit histogram at the previous update
};
__global__ void rj_pk_reset_kernel(RjPkShared *ps, int nm) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int j = 0; j < AMX_MAX_MODELS; j++) ps->pk[j] = (j < nm) ? 1.0 / nm : 0.0;
    ps->pkllim = 1.0 / 10.0;
    ps->nreinit = 1;
  }
}
__global__ void rj_pk_population_kernel(RjPkShared *ps, const unsigned long long *visits, const double *gam, int m, int nm,
                                        int adapt) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long dv[AMX_MAX_MODELS], tot = 0;
  for (int j = 0; j < nm; j++) {
    dv[j] = visits[j] - ps->prev[j];
    ps->prev[j] = visits[j];
    tot += dv[j];
  }
  if (!adapt || tot == 0ull) return;
  double keep = 1.0;
  for (int s = 0; s < m; s++) keep *= (1.0 - gam[s]);
  const double G = 1.0 - keep;
  bool low = false;
  for (int j = 0; j < nm; j++) {
    const double f = (double)dv[j] / (double)tot;
    ps->pk[j] += G * (f - ps->pk[j]);
    low |= (ps->pk[j] < ps->pkllim);
  }
  if (low) {
    ps->nreinit++;
    ps->pkllim = 1.0 / (10.0 * ps->nreinit);
    for (int j = 0; j < nm; j++) ps->pk[j] = 1.0 / nm;
  }
}

// stage [prop blob | target blob] into shared memory (8-byte words), or bind to global
__device__ __forceinline__ void stage_blobs(const RjLaunch &a, double *smem, bool staged, const void *&pb,
                                            const void *&tb) {
  if (staged) {
    const int n0 = a.prop_bytes / 8, n1 = a.tgt_bytes / 8;
    const double *s0 = reinterpret_cast<const double *>(a.prop_blob);
    const double *s1 = reinterpret_cast<const double *>(a.tgt_blob);
    for (int i = threadIdx.x; i < n0; i += blockDim.x) smem[i] = s0[i];
    for (int i = threadIdx.x; i < n1; i += blockDim.x) smem[n0 + i] = s1[i];
    pb = smem;
    tb = smem + n0;
  } else {
    pb = a.prop_blob;
    tb = a.tgt_blob;
  }
}

// dlo: coordinates to load.  A chain in model k never reads its coordinates beyond dims[k] (they are rewritten in full
// when a jump to a wider model is accepted), so the fused kernel loads dims[k] of them and stores up to the widest
// model the chain visited during the launch; what lies beyond keeps its old value in memory either way.
template <class CFG>
__device__ __forceinline__ void load_chain(ChainRegs<CFG> &c, const RjState &s, long id, const double *pk_shared = nullptr,
                                           int dlo = AMX_MAX_DIM) {
  c.k = s.k[id];
  c.lp = s.lp[id];
  c.pkllim = s.pkllim[id];
  c.nreinit = s.nreinit[id];
  if (dlo > s.dmax) dlo = s.dmax;
#pragma unroll
  for (int i = 0; i < CFG::DMAX; i++) c.th[i] = (i < dlo) ? s.theta[(long)i * s.C + id] : 0.0;
#pragma unroll
  for (int i = 0; i < CFG::DMAX; i++) c.thn[i] = c.th[i];
#pragma unroll
  for (int j = 0; j < CFG::NMAX; j++)
    c.pk[j] = (j < s.nmodels) ? (pk_shared ? pk_shared[j] : s.pk[(long)j * s.C + id]) : 0.0;
  c.acc_b = c.try_b = c.acc_s = c.try_s = c.acc_j = c.try_j = 0;
  c.flops = 0;
  c.kn = 0;
  c.lr_pre = c.t_alloc = c.t_wt = c.t_det = c.gam = 0.0;
}
template <class CFG>
__device__ __forceinline__ void store_chain(const ChainRegs<CFG> &c, const RjState &s, long id, int dhi = AMX_MAX_DIM,
                                            bool with_pk = true) {
  s.k[id] = c.k;
  s.lp[id] = c.lp;
  s.pkllim[id] = c.pkllim;
  s.nreinit[id] = c.nreinit;
  if (dhi > s.dmax) dhi = s.dmax;
#pragma unroll
  for (int i = 0; i < CFG::DMAX; i++)
    if (i < dhi) s.theta[(long)i * s.C + id] = c.th[i];
  if (with_pk) {
#pragma unroll
    for (int j = 0; j < CFG::NMAX; j++)
      if (j < s.nmodels) s.pk[(long)j * s.C + id] = c.pk[j];
  }
}

template <class RNG>
__device__ __forceinline__ void open_stream(RNG &u, const RjLaunch &a, long id, unsigned long long consumed);
template <>
__device__ __forceinline__ void open_stream<PhiloxStream>(PhiloxStream &u, const RjLaunch &a, long id,
                                                          unsigned long long consumed) {
  u.open(a.seed, a.chain_base + (unsigned long long)id, consumed);
}
template <>
__device__ __forceinline__ void open_stream<PhiloxStreamOL>(PhiloxStreamOL &u, const RjLaunch &a, long id,
                                                            unsigned long long consumed) {
  u.open(a.seed, a.chain_base + (unsigned long long)id, consumed);
}
// the generator type the wide configurations run (same stream, out-of-line block and Box-Muller)
template <class RNG>
struct WideRng {
  using type = RNG;
};
#if AMX_RJ_WIDE_OUTLINE
template <>
struct WideRng<PhiloxStream> {
  using type = PhiloxStreamOL;
};
#endif
template <>
__device__ __forceinline__ void open_stream<TapeStream>(TapeStream &u, const RjLaunch &a, long id,
                                                        unsigned long long consumed) {
  u.open(a.tape, a.tape_stride, (unsigned long long)id, consumed);
}

// One out-of-line copy of the plug-in evaluation per kernel: the sweep calls it from three places, and
// inlining all three made the kernel ~130 KB of SASS (instruction-fetch stalls were as frequent as issues).
#ifndef AMX_RJ_EVAL_INLINE
#define AMX_RJ_EVAL_INLINE 1  // measured on B200: out-of-line 3.9e9 chain-sweeps/s, inline 4.6e9 (toy1)
#endif
#ifndef AMX_RJ_MIN_BLOCKS
#define AMX_RJ_MIN_BLOCKS 4
#endif
// Wide configurations (d > 2), measured on B200 with the sorted population (C5-RJ / coal-mining, 2^18 chains,
// chain-sweeps/s): straight-line sweep with three inlined plug-in evaluations 3.72e8 / 2.77e8, phase loop (one copy)
// 3.70e8 / 2.92e8 -- instruction fetch is the largest stall of these kernels (ncu: 1.9 / 6.6 "no instruction" stalls
// per issue), so the smaller loop wins where the plug-in is large.  Register cap for 3 CTAs per SM: 166 registers,
// no spills (255 uncapped; 128 for 4 CTAs spills and is not faster).
#ifndef AMX_RJ_PHASE_LOOP_WIDE
#define AMX_RJ_PHASE_LOOP_WIDE 1
#endif
#ifndef AMX_RJ_MIN_BLOCKS_WIDE
#define AMX_RJ_MIN_BLOCKS_WIDE 3
#endif
template <class CFG, class TGT>
#if AMX_RJ_EVAL_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
    double
    eval_target(const TGT &T, int k, const double (&x)[CFG::DMAX]) {
  return T.template eval<CFG::DMAX>(k, x);
}

// ---- the fused sweep kernel ------------------------------------------------------------------
enum { kPhaseBlock = 0, kPhaseCoord, kPhaseJump, kPhaseIdle };
template <class CFG, class TGT, class RNG>
__global__ void __launch_bounds__(kRjThreads, (CFG::DMAX <= 2 ? AMX_RJ_MIN_BLOCKS : AMX_RJ_MIN_BLOCKS_WIDE)) rj_sweep_kernel(RjLaunch a, int staged) {
  extern __shared__ double smem[];
  __shared__ unsigned s_hist[kRjWarps][CFG::NMAX];
  __shared__ int s_clp[AMX_MAX_MODELS];
  __shared__ unsigned long long s_cnt[8];
  __shared__ int s_status;
  // allocation weights of the jump: [component][thread] in shared memory for the register-resident configurations
  __shared__ double s_pa[(CFG::DMAX <= 8) ? CFG::LMAX * kRjThreads : 1];
  // sorted mode: a CTA's chains come from anywhere in the population, so the per-group visit counts (Monte-Carlo
  // error, amx_rj_visit_se) are kept per chain id -- group = (id / 128) mod kRjGroups, a fixed partition of the chains
  __shared__ unsigned s_grp[kRjGroups * CFG::NMAX];
  AllocVec<CFG> pa;
  if constexpr (CFG::DMAX <= 8) pa.p = s_pa + threadIdx.x;

  const void *pb, *tb;
  stage_blobs(a, smem, staged != 0, pb, tb);
  for (int i = threadIdx.x; i < kRjWarps * CFG::NMAX; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  if (a.order)
    for (int i = threadIdx.x; i < kRjGroups * CFG::NMAX; i += blockDim.x) s_grp[i] = 0;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_status = 0;
  __syncthreads();
  ProposalView P;
  P.bind(pb, a.order != nullptr);
  TGT T;
  if constexpr (std::is_same<TGT, GaussMixTarget>::value) T.bind(tb, a.tgt_flags | (a.order ? kTargetFlagUniformDims : 0));
  else T.bind(tb, a.tgt_flags);
  const int nm = P.h->nmodels;
  if (threadIdx.x < nm) s_clp[threadIdx.x] = T.flops(threadIdx.x);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int status = 0;
  // optional modes are compiled out of the small configuration (a run that asks for them takes the medium one)
  typename std::conditional<(CFG::DMAX <= 2), NoModes, RjModes>::type md;
  if constexpr (CFG::DMAX > 2) md = a.modes;
  // The grid may be smaller than the population (large configurations keep few threads resident so that a
  // chain's scratch vectors stay in L1): every thread then walks several chains, one after the other.
  for (long base = (long)blockIdx.x * blockDim.x; base < a.st.C; base += (long)gridDim.x * blockDim.x) {
  const long gid = base + threadIdx.x;
  const bool active = gid < a.st.C;
  const long slot = active ? gid : a.st.C - 1;  // tail lanes shadow the last chain, never write
  const long id = a.order ? (long)a.order[slot] : slot;

  ChainRegs<CFG> c;
  int dhi = P.h->dims[a.st.k[id]];  // widest model this chain visits during the launch
  load_chain(c, a.st, id, a.pk_shared, dhi);
  RNG u;
  const unsigned long long draws0 = a.st.draws[id];
  open_stream(u, a, id, draws0);
  const bool traced = active && id < a.ntrace;
  unsigned *my_grp = s_grp + (int)((id >> 7) % kRjGroups) * CFG::NMAX;

  for (int s = 0; s < a.nsweeps; s++) {
    const unsigned long long sweep_i = a.sweep0 + (unsigned long long)s;
    const int d = P.h->dims[c.k];
    // A sweep is a sequence of phases -- propose | evaluate the log-posterior | accept -- and the loop below holds
    // ONE copy of the plug-in evaluation and of each proposal kind (three inlined copies made the kernel 130 KB of
    // SASS and instruction fetch its second largest stall).  Every 10th sweep is a block move (:95, :148), uniform
    // over the grid; otherwise coordinate phase j runs on the lanes whose model has more than j coordinates and the
    // jump waits for the widest model in the warp, so the lanes of a warp always take it together.
    const bool blockmove = (sweep_i % 10ull == 0ull);
    if constexpr (CFG::DMAX <= 8 || AMX_RJ_PHASE_LOOP_WIDE) {
      int last = 1;
      if (blockmove) {
        c.flops += (unsigned)(s_clp[c.k] + 3 * d + 10);
      } else {
        sync_proposal(c, d);
        last = __reduce_max_sync(0xffffffffu, d);
        c.flops += (unsigned)(d * (s_clp[c.k] + 12));
      }
      for (int ph = 0; ph <= last; ph++) {
        const int kind = (ph == last) ? kPhaseJump : (blockmove ? kPhaseBlock : (ph < d ? kPhaseCoord : kPhaseIdle));
        if (kind == kPhaseBlock) rwm_block_propose(c, P, u, md);
        else if (kind == kPhaseCoord) rwm_coord_propose(c, P, u, ph, md);
        else if (kind == kPhaseJump) rj_propose(c, P, u, a.gam[s], md, s_clp, pa);
        double lpn = 0.0;
        if (kind != kPhaseIdle) lpn = eval_target<CFG, TGT>(T, kind == kPhaseJump ? c.kn : c.k, c.thn);
        if (kind == kPhaseBlock) rwm_block_finish(c, P, u, lpn);
        else if (kind == kPhaseCoord) rwm_coord_finish(c, u, ph, lpn);
        else if (kind == kPhaseJump) rj_finish(c, P, u, lpn, a.adapt != 0);
      }
    } else {
      // straight-line form (AMX_RJ_PHASE_LOOP_WIDE=0): three inlined copies of the plug-in evaluation
      if (blockmove) {
        rwm_block_propose(c, P, u, md);
        const double lpn = eval_target<CFG, TGT>(T, c.k, c.thn);
        rwm_block_finish(c, P, u, lpn);
        c.flops += (unsigned)(s_clp[c.k] + 3 * d + 10);
      } else {
        sync_proposal(c, d);
        for (int j = 0; j < d; j++) {
          rwm_coord_propose(c, P, u, j, md);
          const double lpn = eval_target<CFG, TGT>(T, c.k, c.thn);
          rwm_coord_finish(c, u, j, lpn);
        }
        c.flops += (unsigned)(d * (s_clp[c.k] + 12));
      }
      rj_propose(c, P, u, a.gam[s], md, s_clp, pa);
      const double lpn = eval_target<CFG, TGT>(T, c.kn, c.thn);
      rj_finish(c, P, u, lpn, a.adapt != 0);
    }
    if (c.lp != c.lp) status |= 2;
    dhi = max(dhi, P.h->dims[c.k]);

    // model-visit histogram: one ballot per model, lane 0 adds the population count
    __syncwarp();
    for (int m = 0; m < nm; m++) {
      const unsigned b = __ballot_sync(0xffffffffu, active && c.k == m);
      if (lane == 0) s_hist[warp][m] += __popc(b);
    }
    if (a.order && active) atomicAdd(&my_grp[c.k], 1u);
    if (traced) {
      const long row = (long)id * a.tr_stride + a.tr_off + s;
      a.tr_k[row] = c.k;
      a.tr_lp[row] = c.lp;
      const int dk = P.h->dims[c.k];
      for (int i = 0; i < a.st.dmax; i++) a.tr_theta[row * a.st.dmax + i] = (i < dk) ? aget(c.th, i) : 0.0;
      for (int j = 0; j < nm; j++) a.tr_pk[row * nm + j] = aget(c.pk, j);
    }
  }
  if (u.overrun()) status |= 1;
  // every chain tries one jump per sweep and one block move per 10th sweep: grid-uniform counts, set here so
  // that the per-sweep increments inside the phase functions (needed by the split kernels) are dead code
  c.try_j = (unsigned)a.nsweeps;
  c.try_b = (unsigned)((a.sweep0 + (unsigned long long)a.nsweeps + 9ull) / 10ull - (a.sweep0 + 9ull) / 10ull);

  if (active) {
    // pk moves only while adapting per chain; the shared pk of the population mode is stored as before
    store_chain(c, a.st, id, dhi, a.adapt != 0 || a.pk_shared != nullptr);
    a.st.draws[id] = u.n;
  }
  // counters: warp reduce, one shared atomic per warp, one global atomic per CTA
  unsigned long long v[8] = {c.acc_b, c.try_b, c.acc_s, c.try_s, c.acc_j, c.try_j, c.flops, u.n - draws0};
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const unsigned long long r = warp_sum_u64(active ? v[q] : 0ull);
    if (lane == 0) atomicAdd(&s_cnt[q], r);
  }
  }  // chains of this thread
  if (status) atomicOr(&s_status, status);
  __syncthreads();
  if (threadIdx.x < 8) atomicAdd(&a.cnt[threadIdx.x], s_cnt[threadIdx.x]);
  if (threadIdx.x < nm) {
    unsigned long long t = 0;
    for (int w = 0; w < kRjWarps; w++) t += s_hist[w][threadIdx.x];
    atomicAdd(&a.visits[threadIdx.x], t);
    if (!a.order) atomicAdd(&a.visits_grp[(blockIdx.x % kRjGroups) * AMX_MAX_MODELS + threadIdx.x], t);
  }
  if (a.order)
    for (int i = threadIdx.x; i < kRjGroups * CFG::NMAX; i += blockDim.x)
      if (s_grp[i]) atomicAdd(&a.visits_grp[(i / CFG::NMAX) * AMX_MAX_MODELS + (i % CFG::NMAX)], (unsigned long long)s_grp[i]);
  if (threadIdx.x == 0 && s_status) atomicOr(a.status, s_status);
}

// ---- sorted mode: chains grouped by (model, proposed model) before every launch -------------------------------------
// One thread per chain means a warp pays for its widest chain: with models of 2 ... 20 coordinates mixed in a warp,
// every coordinate loop, triangular solve and plug-in evaluation runs at d = 20 for all 32 lanes (C5-RJ: 7.6 times the
// arithmetic the chains need).  Here the population is counting-sorted before each launch of `sort_seg` sweeps by the
// pair (current model k, model kn the coming jump will propose), widest first, and the sweep kernel walks the chains in
// that order, so the lanes of a warp run the same trip counts through the whole sweep.  kn is known in advance: with
// Gaussian innovations a sweep consumes a fixed number of uniforms before the model draw (3 per coordinate, or
// 2 ceil(d/2) + 1 for a block move, plus one for the allocation when the mixture has several components), the
// generators have random access, and pk only changes after the jump -- the same scan of pk as rj_propose (:1138-1169).
// Per-chain arithmetic and random streams are untouched: results are bit-identical to the unsorted kernel.
struct RjSort {
  int *keys;    // [C]
  int *hist;    // [nb] bucket sizes (zero between sorts)
  int *start;   // [nb] exclusive prefix
  int *cursor;  // [nb] (zero between sorts)
  int *order;   // [C]
  int nb;
  unsigned char rank[AMX_MAX_MODELS];  // position of a model in the order "widest first"
};
constexpr int kSortThreads = 256;
constexpr int kSortBuckets = AMX_MAX_MODELS * AMX_MAX_MODELS;

template <class RNG>
__global__ void __launch_bounds__(kSortThreads) rj_sort_key_kernel(RjLaunch a, RjSort so) {
  __shared__ int s_h[kSortBuckets];
  for (int i = threadIdx.x; i < so.nb; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  ProposalView P;
  P.bind(a.prop_blob);
  const int nm = P.h->nmodels;
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.st.C) {
    const int k = a.st.k[c];
    int kn = k;
    if (nm > 1 && a.modes.dof == 0) {
      const int d = P.h->dims[k], L = P.h->ncomp[k];
      const bool blockmove = (a.sweep0 % 10ull == 0ull);
      const unsigned long long pos =
          a.st.draws[c] + (unsigned long long)((blockmove ? 2 * ((d + 1) / 2) + 1 : 3 * d) + (L > 1 ? 1 : 0));
      RNG u;
      open_stream(u, a, c, 0ull);
      const double uu = u.at(pos);
      double t = 0.0;
      bool found = false;
      kn = 0;
      for (int i = 0; i < nm; i++) {
        t += a.pk_shared ? a.pk_shared[i] : a.st.pk[(long)i * a.st.C + c];
        if (!found && uu < t) {
          kn = i;
          found = true;
        }
      }
    }
    const int key = (int)so.rank[k] * nm + (int)so.rank[kn];
    so.keys[c] = key;
    atomicAdd(&s_h[key], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < so.nb; i += blockDim.x)
    if (s_h[i]) atomicAdd(&so.hist[i], s_h[i]);
}
// exclusive prefix of the bucket sizes (nb <= 1024: one CTA); leaves hist and cursor zero for the next sort
__global__ void __launch_bounds__(kSortBuckets) rj_sort_scan_kernel(RjSort so) {
  __shared__ int s[kSortBuckets];
  const int t = threadIdx.x;
  const int v = (t < so.nb) ? so.hist[t] : 0;
  s[t] = v;
  __syncthreads();
  for (int o = 1; o < kSortBuckets; o <<= 1) {
    const int x = (t >= o) ? s[t - o] : 0;
    __syncthreads();
    s[t] += x;
    __syncthreads();
  }
  if (t < so.nb) {
    so.start[t] = s[t] - v;
    so.hist[t] = 0;
    so.cursor[t] = 0;
  }
}
__global__ void __launch_bounds__(kSortThreads) rj_sort_scatter_kernel(RjSort so, long C) {
  __shared__ int s_cnt[kSortBuckets], s_base[kSortBuckets];
  for (int i = threadIdx.x; i < so.nb; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int key = 0, r = 0;
  if (c < C) {
    key = so.keys[c];
    r = atomicAdd(&s_cnt[key], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < so.nb; i += blockDim.x)
    if (s_cnt[i]) s_base[i] = so.start[i] + atomicAdd(&so.cursor[i], s_cnt[i]);
  __syncthreads();
  if (c < C) so.order[s_base[key] + r] = (int)c;
}


// ---- split sweep for HOST log-posterior callbacks -------------------------------------------------
// The scalar contract `double f(int model_k, double *x)` (automix.h:46) can only run on the host.
// A sweep then becomes a sequence of small kernels -- propose | host evaluates | finish -- built from
// the same phase functions as the fused kernel; chain state, proposals and the values carried from
// rj_propose to rj_finish live in global memory between them.  Compatibility path: (d+1) PCIe round
// trips per sweep, regardless of the number of chains.
enum RjPhase { kPhBlockPropose = 0, kPhBlockFinish, kPhCoordPropose, kPhCoordFinish, kPhJumpPropose, kPhJumpFinish };

struct RjSplit {
  double *thn;    // [C][dmax] chain-major: what the host callback reads
  int *keval;     // [C] model index to evaluate, -1 = this chain sits the phase out
  double *lpn;    // [C] values returned by the callback
  int *kn;        // [C]
  double *carry;  // [5][C]: lr_pre, t_alloc, t_wt, t_det, gam
};

template <class RNG>
__global__ void __launch_bounds__(kRjThreads) rj_split_kernel(RjLaunch a, RjSplit sp, int phase, int j, int s) {
  using CFG = RjCfgG;
  __shared__ int s_clp[AMX_MAX_MODELS];
  __shared__ unsigned s_hist[kRjWarps][CFG::NMAX];
  __shared__ unsigned long long s_cnt[8];
  AllocVec<CFG> pa;
  ProposalView P;
  P.bind(a.prop_blob);
  const int nm = P.h->nmodels;
  if (threadIdx.x < AMX_MAX_MODELS) s_clp[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < kRjWarps * CFG::NMAX; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = gid < a.st.C;
  const long id = active ? gid : a.st.C - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ChainRegs<CFG> c;
  load_chain(c, a.st, id, a.pk_shared);
  const int dmax = a.st.dmax;
  for (int i = 0; i < dmax; i++) c.thn[i] = sp.thn[id * dmax + i];
  c.kn = sp.kn[id];
  c.lr_pre = sp.carry[0 * a.st.C + id];
  c.t_alloc = sp.carry[1 * a.st.C + id];
  c.t_wt = sp.carry[2 * a.st.C + id];
  c.t_det = sp.carry[3 * a.st.C + id];
  c.gam = sp.carry[4 * a.st.C + id];
  RNG u;
  const unsigned long long draws0 = a.st.draws[id];
  open_stream(u, a, id, draws0);
  const int d = P.h->dims[c.k];
  const double lpn = sp.lpn[id];
  int keval = -1;
  switch (phase) {
    case kPhBlockPropose:
      rwm_block_propose(c, P, u, a.modes);
      keval = c.k;
      break;
    case kPhBlockFinish:
      rwm_block_finish(c, P, u, lpn);
      break;
    case kPhCoordPropose:
      if (j == 0) sync_proposal(c, d);
      if (j < d) {
        rwm_coord_propose(c, P, u, j, a.modes);
        keval = c.k;
      }
      break;
    case kPhCoordFinish:
      if (j < d) rwm_coord_finish(c, u, j, lpn);
      break;
    case kPhJumpPropose:
      rj_propose(c, P, u, a.gam[s], a.modes, s_clp, pa);
      keval = c.kn;
      break;
    case kPhJumpFinish:
      rj_finish(c, P, u, lpn, a.adapt != 0);
      break;
  }
  int status = (u.overrun() ? 1 : 0) | ((c.lp != c.lp) ? 2 : 0);
  if (phase == kPhJumpFinish) {
    __syncwarp();
    for (int m = 0; m < nm; m++) {
      const unsigned b = __ballot_sync(0xffffffffu, active && c.k == m);
      if (lane == 0) s_hist[warp][m] += __popc(b);
    }
    if (active && gid < a.ntrace) {
      const long row = (long)gid * a.tr_stride + a.tr_off + s;
      a.tr_k[row] = c.k;
      a.tr_lp[row] = c.lp;
      const int dk = P.h->dims[c.k];
      for (int i = 0; i < dmax; i++) a.tr_theta[row * dmax + i] = (i < dk) ? c.th[i] : 0.0;
      for (int q = 0; q < nm; q++) a.tr_pk[row * nm + q] = c.pk[q];
    }
  }
  if (active) {
    store_chain(c, a.st, id);
    a.st.draws[id] = u.n;
    for (int i = 0; i < dmax; i++) sp.thn[id * dmax + i] = c.thn[i];
    sp.keval[id] = keval;
    sp.kn[id] = c.kn;
    sp.carry[0 * a.st.C + id] = c.lr_pre;
    sp.carry[1 * a.st.C + id] = c.t_alloc;
    sp.carry[2 * a.st.C + id] = c.t_wt;
    sp.carry[3 * a.st.C + id] = c.t_det;
    sp.carry[4 * a.st.C + id] = c.gam;
  }
  unsigned long long v[8] = {c.acc_b, c.try_b, c.acc_s, c.try_s, c.acc_j, c.try_j, 0ull, u.n - draws0};
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const unsigned long long r = warp_sum_u64(active ? v[q] : 0ull);
    if (lane == 0) atomicAdd(&s_cnt[q], r);
  }
  __syncthreads();
  if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&a.cnt[threadIdx.x], s_cnt[threadIdx.x]);
  if (phase == kPhJumpFinish && threadIdx.x < nm) {
    unsigned long long t = 0;
    for (int w = 0; w < kRjWarps; w++) t += s_hist[w][threadIdx.x];
    atomicAdd(&a.visits[threadIdx.x], t);
    atomicAdd(&a.visits_grp[(blockIdx.x % kRjGroups) * AMX_MAX_MODELS + threadIdx.x], t);
  }
  if (status && active) atomicOr(a.status, status);
}


// ---- persistent sweep kernel for HOST log-posterior callbacks (amx_mailbox.cuh) -------------------------------------
// The whole sweep loop of the fused kernel with the plug-in evaluation replaced by a mailbox exchange with the host:
// chain state stays in the thread for all nsweeps sweeps.  Every CTA makes the same number of exchanges: a block-move
// sweep has one, any other sweep dmax (chains of smaller models sit the extra coordinates out), then one for the jump.
template <class RNG>
__global__ void __launch_bounds__(kMbThreads) rj_mailbox_kernel(RjLaunch a, void *mb_base, int ldx, unsigned seq0, int cpc) {
  using CFG = RjCfgG;
  __shared__ unsigned s_hist[kMbThreads / 32][CFG::NMAX];  // (the CTA may be launched with fewer than kMbThreads threads)
  __shared__ int s_clp[AMX_MAX_MODELS];
  __shared__ unsigned long long s_cnt[8];
  AllocVec<CFG> pa;
  ProposalView P;
  P.bind(a.prop_blob);
  const int nm = P.h->nmodels, dmax = a.st.dmax;
  for (int i = threadIdx.x; i < (kMbThreads / 32) * CFG::NMAX; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  if (threadIdx.x < AMX_MAX_MODELS) s_clp[threadIdx.x] = 0;
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  Mailbox *mb = mailbox_at(mb_base, blockIdx.x, ldx);
  unsigned seq = seq0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // cpc chains per CTA (the first cpc threads): few chains per mailbox keep the host's share of an exchange short, and
  // the CTAs' exchanges overlap
  const long gid = (long)blockIdx.x * cpc + threadIdx.x;
  const bool active = (int)threadIdx.x < cpc && gid < a.st.C;
  const long id = active ? gid : a.st.C - 1;  // the other threads shadow the last chain: they ask for nothing and never write
  ChainRegs<CFG> c;
  load_chain(c, a.st, id, a.pk_shared);
  RNG u;
  const unsigned long long draws0 = a.st.draws[id];
  open_stream(u, a, id, draws0);
  const bool traced = active && gid < a.ntrace;
  int status = 0;
  for (int s = 0; s < a.nsweeps; s++) {
    const unsigned long long sweep_i = a.sweep0 + (unsigned long long)s;
    const int d = P.h->dims[c.k];
    if (sweep_i % 10ull == 0ull) {
      rwm_block_propose(c, P, u, a.modes);
      const double lpn = mailbox_exchange<CFG::DMAX>(mb, ldx, ++seq, active ? c.k : -1, c.thn, d);
      rwm_block_finish(c, P, u, lpn);
    } else {
      sync_proposal(c, d);
      for (int j = 0; j < dmax; j++) {
        const bool on = j < d;
        if (on) rwm_coord_propose(c, P, u, j, a.modes);
        const double lpn = mailbox_exchange<CFG::DMAX>(mb, ldx, ++seq, (active && on) ? c.k : -1, c.thn, d);
        if (on) rwm_coord_finish(c, u, j, lpn);
      }
    }
    rj_propose(c, P, u, a.gam[s], a.modes, s_clp, pa);
    {
      const double lpn = mailbox_exchange<CFG::DMAX>(mb, ldx, ++seq, active ? c.kn : -1, c.thn, P.h->dims[c.kn]);
      rj_finish(c, P, u, lpn, a.adapt != 0);
    }
    if (active && c.lp != c.lp) status |= 2;
    __syncwarp();
    for (int m = 0; m < nm; m++) {
      const unsigned b = __ballot_sync(0xffffffffu, active && c.k == m);
      if (lane == 0) s_hist[warp][m] += __popc(b);
    }
    if (traced) {
      const long row = (long)gid * a.tr_stride + a.tr_off + s;
      a.tr_k[row] = c.k;
      a.tr_lp[row] = c.lp;
      const int dk = P.h->dims[c.k];
      for (int i = 0; i < dmax; i++) a.tr_theta[row * dmax + i] = (i < dk) ? c.th[i] : 0.0;
      for (int q = 0; q < nm; q++) a.tr_pk[row * nm + q] = c.pk[q];
    }
  }
  if (active && u.overrun()) status |= 1;
  if (active) {
    store_chain(c, a.st, id);
    a.st.draws[id] = u.n;
  }
  unsigned long long v[8] = {c.acc_b, c.try_b, c.acc_s, c.try_s, c.acc_j, c.try_j, 0ull, u.n - draws0};
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const unsigned long long r = warp_sum_u64(active ? v[q] : 0ull);
    if (lane == 0) atomicAdd(&s_cnt[q], r);
  }
  __syncthreads();
  if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&a.cnt[threadIdx.x], s_cnt[threadIdx.x]);
  if (threadIdx.x < nm) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_hist[w][threadIdx.x];
    atomicAdd(&a.visits[threadIdx.x], t);
    atomicAdd(&a.visits_grp[(blockIdx.x % kRjGroups) * AMX_MAX_MODELS + threadIdx.x], t);
  }
  if (status && active) atomicOr(a.status, status);
}

// chain start for host callbacks: pick the model and copy the start vector; lp comes from the host
template <class RNG>
__global__ void __launch_bounds__(kRjThreads) rj_init_split_kernel(RjLaunch a, RjSplit sp, const double *init_flat,
                                                                   int finish) {
  ProposalView P;
  P.bind(a.prop_blob);
  const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.st.C) return;
  const int nm = P.h->nmodels, dmax = a.st.dmax;
  if (finish) {
    a.st.lp[id] = sp.lpn[id];
    return;
  }
  RNG u;
  open_stream(u, a, id, 0ull);
  int k0 = (int)floor(nm * u.next());
  if (k0 >= nm) k0 = nm - 1;
  int off = 0;
  for (int q = 0; q < k0; q++) off += P.h->dims[q];
  const int d = P.h->dims[k0];
  for (int i = 0; i < dmax; i++) {
    const double v = (i < d) ? init_flat[off + i] : 0.0;
    a.st.theta[(long)i * a.st.C + id] = v;
    sp.thn[id * dmax + i] = v;
  }
  for (int q = 0; q < nm; q++) a.st.pk[(long)q * a.st.C + id] = 1.0 / nm;
  a.st.k[id] = k0;
  a.st.nreinit[id] = 1;
  a.st.pkllim[id] = 1.0 / 10.0;
  a.st.draws[id] = u.n;
  sp.keval[id] = k0;
  if (u.overrun()) atomicOr(a.status, 1);
}

// ---- chain start (initChain, automix.c:423-449) -------------------------------------------------
template <class TGT, class RNG>
__global__ void __launch_bounds__(kRjThreads) rj_init_kernel(RjLaunch a, const double *init_flat) {
  using CFG = RjCfgG;
  ProposalView P;
  P.bind(a.prop_blob);
  TGT T;
  T.bind(a.tgt_blob, a.tgt_flags);
  const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.st.C) return;
  RNG u;
  open_stream(u, a, id, 0ull);
  const int nm = P.h->nmodels;
  int k0 = (int)floor(nm * u.next());
  if (k0 >= nm) k0 = nm - 1;
  int off = 0;
  for (int j = 0; j < k0; j++) off += P.h->dims[j];
  double th[CFG::DMAX];
  const int d = P.h->dims[k0];
  for (int i = 0; i < CFG::DMAX; i++) th[i] = (i < d) ? init_flat[off + i] : 0.0;
  const double lp = T.template eval<CFG::DMAX>(k0, th);
  for (int i = 0; i < a.st.dmax; i++) a.st.theta[(long)i * a.st.C + id] = th[i];
  for (int j = 0; j < nm; j++) a.st.pk[(long)j * a.st.C + id] = 1.0 / nm;
  a.st.lp[id] = lp;
  a.st.k[id] = k0;
  a.st.nreinit[id] = 1;
  a.st.pkllim[id] = 1.0 / 10.0;
  a.st.draws[id] = u.n;
  if (u.overrun()) atomicOr(a.status, 1);
}

// ---- chain-major (host API) <-> coordinate-major (device) state transposition ----------------------------
__global__ void rj_state_scatter_kernel(RjState st, long first, long count, const double *theta_cm, const double *pk_cm) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= count) return;
  for (int i = 0; i < st.dmax; i++) st.theta[(long)i * st.C + first + c] = theta_cm[c * st.dmax + i];
  for (int j = 0; j < st.nmodels; j++) st.pk[(long)j * st.C + first + c] = pk_cm[c * st.nmodels + j];
}
__global__ void rj_state_gather_kernel(RjState st, long first, long count, double *theta_cm, double *pk_cm) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= count) return;
  if (theta_cm)
    for (int i = 0; i < st.dmax; i++) theta_cm[c * st.dmax + i] = st.theta[(long)i * st.C + first + c];
  if (pk_cm)
    for (int j = 0; j < st.nmodels; j++) pk_cm[c * st.nmodels + j] = st.pk[(long)j * st.C + first + c];
}

// ---- batched evaluation of a plug-in (amx_target_eval; also the target parity tests) --------------
struct EvalDims {
  int dims[AMX_MAX_MODELS];
};
template <class TGT>
__global__ void __launch_bounds__(kRjThreads) target_eval_kernel(const void *blob, int flags, EvalDims ed, long n,
                                                                 const int *k, const double *x, long ldx,
                                                                 double *out) {
  TGT T;
  T.bind(blob, flags);
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[AMX_MAX_DIM];
  const int d = ed.dims[k[i]];
  for (int j = 0; j < AMX_MAX_DIM; j++) v[j] = (j < d) ? x[i * ldx + j] : 0.0;
  out[i] = T.template eval<AMX_MAX_DIM>(k[i], v);
}


// ---- launchers ---------------------------------------------------------------------------------------------
template <class CFG, class TGT, class RNG>
inline int launch_sweeps(const RjLaunch &a) {
  const size_t need = (size_t)a.prop_bytes + (size_t)a.tgt_bytes;
  int staged = need <= 160 * 1024 ? 1 : 0;
  // Sorted mode launches once per sweep: copying ~60 KB of blobs into every CTA's shared memory per launch costs more
  // than it saves (C5-RJ 3.17e8 -> 3.54e8 chain-sweeps/s without), and the lanes of a sorted warp read the same
  // records, so the L1-cached global loads are broadcasts.
  if (a.order) staged = 0;
  if (const char *e = getenv("AMX_RJ_STAGE")) staged = staged && atoi(e);
  const size_t smem = staged ? need : 0;
  auto kern = rj_sweep_kernel<CFG, TGT, RNG>;
  if (smem > 48 * 1024) AMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned grid = (unsigned)((a.st.C + kRjThreads - 1) / kRjThreads);
  if (CFG::DMAX > 8) {
    // Large configurations keep a chain's vectors (~1-1.8 KB per thread) in local memory.  At full
    // occupancy that is > 1 MB per SM and spills past L1 and L2; with a couple of resident CTAs per SM it stays
    // in L1, so cap the grid and let the kernel's chain loop cover the population.
    int dev = 0, sms = 0;
    AMX_CUDA(cudaGetDevice(&dev));
    AMX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const char *e = getenv("AMX_RJ_BLOCKS_PER_SM");
    const unsigned per_sm = e ? (unsigned)atoi(e) : (a.order ? 3u : (CFG::DMAX <= 20 ? 2u : 1u));
    const unsigned cap = (unsigned)sms * (per_sm ? per_sm : 1u);
    if (grid > cap) grid = cap;
    // shared memory per SM: per resident CTA the staged blobs, the static arrays (allocation weights, histograms)
    // and the driver's 1 KB; whatever is left stays L1 for the chains' local vectors
    cudaFuncAttributes fa;
    AMX_CUDA(cudaFuncGetAttributes(&fa, kern));
    const size_t per_cta = smem + fa.sharedSizeBytes + 1024;
    int pct = (int)((per_cta * per_sm * 100) / (228 * 1024) + 3);
    AMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct));
  }
  kern<<<grid, kRjThreads, smem, stream()>>>(a, staged);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  return AMX_OK;
}

template <class TGT, class RNG>
inline int launch_cfg(const RjLaunch &a, int dmax, int Lmax, int nm) {
  using WRNG = typename WideRng<RNG>::type;
  if constexpr (TargetIsWide<TGT>::value) {
    if (dmax <= RjCfgL::DMAX && Lmax <= RjCfgL::LMAX && nm <= RjCfgL::NMAX) return launch_sweeps<RjCfgL, TGT, WRNG>(a);
    return launch_sweeps<RjCfgG, TGT, WRNG>(a);
  } else {
    const bool plain = a.modes.dof == 0 && a.modes.do_perm == 0;
    if (plain && dmax <= RjCfgS::DMAX && Lmax <= RjCfgS::LMAX && nm <= RjCfgS::NMAX) return launch_sweeps<RjCfgS, TGT, RNG>(a);
    if (dmax <= RjCfgM::DMAX && Lmax <= RjCfgM::LMAX && nm <= RjCfgM::NMAX) return launch_sweeps<RjCfgM, TGT, RNG>(a);
    if (dmax <= RjCfgL::DMAX && Lmax <= RjCfgL::LMAX && nm <= RjCfgL::NMAX) return launch_sweeps<RjCfgL, TGT, WRNG>(a);
    return launch_sweeps<RjCfgG, TGT, WRNG>(a);
  }
}


}  // namespace amx
