// amx_rj.cu -- K3 kernels and their host driver (amx_rj_* of include/amx.h).
//
// Mapping: one thread per chain, 128 threads per CTA.  The proposal mixtures and the
// plug-in parameters are staged once per CTA in shared memory; a chain's state is loaded
// once, lives in registers (small configurations) or L1-resident local memory (general
// configuration) for all nsweeps sweeps of the launch, and is written back once.  There
// is no per-sweep global traffic except the optional trace chains.  Model-visit counts are
// accumulated with warp-aggregated ballots into a per-warp shared histogram and flushed with
// one 64-bit atomic per model per CTA.
#include "amx_rj_kernels.cuh"

using namespace amx;

struct amx_rj {
  const amx_proposal *prop;
  const amx_target *tgt;
  long C;
  int dmax, nm, ntrace;
  unsigned long long seed, sweep_i, chain_base;
  RjModes modes;
  RjState st;
  double *init_dev;
  double *tape_dev;
  unsigned long long tape_stride;
  unsigned long long *visits_dev, *cnt_dev, *grp_dev;
  unsigned long long grp_host[AMX_RJ_GROUPS * AMX_MAX_MODELS];  // group counts as of the last amx_rj_collect
  int *status_dev;
  double *gam_dev;
  long gam_cap;
  int pk_mode, pk_seg;      // AMX_PK_PER_CHAIN / AMX_PK_POPULATION; sweeps per population update
  RjPkShared *pk_shared;    // device
  int sort_seg;             // sorted mode: -1 automatic, 0 off, n > 0 re-sort the chains before every n sweeps
  RjSort so;                // its device buffers (allocated on first use)
  int *tr_k;
  double *tr_lp, *tr_theta, *tr_pk;
  long tr_cap, last_nsweeps;
  cudaEvent_t e0, e1;
  double kernel_ms;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *pending;
  MomentsBuf *mom;    // posterior-moment accumulators (amx_summary.cu), created on first use
  double *stage_dev;  // chain-major staging for set/get_state
  long stage_cap;
  // host-callback mode
  void *mb_host;  // mailboxes (mapped pinned host memory), one per CTA of rj_mailbox_kernel
  unsigned mb_seq;  // exchanges they have carried so far
  RjSplit sp;
  double *h_thn, *h_lpn;  // pinned mirrors
  int *h_keval;
  std::vector<int> *h_kc;
  std::vector<double> *h_xc, *h_lc;
};

template <class RNG>
static int launch_tgt(const amx_rj *rj, const RjLaunch &a) {
  const amx_fam_hdr &h = rj->prop->hdr;
  switch (rj->tgt->d.kind) {
    case kTargetGaussMix: return launch_cfg<GaussMixTarget, RNG>(a, h.dmax, h.Lmax, h.nmodels);
    case kTargetQuad: return launch_cfg<QuadTarget, RNG>(a, h.dmax, h.Lmax, h.nmodels);
    case kTargetCoal: return launch_cfg<CoalTarget, RNG>(a, h.dmax, h.Lmax, h.nmodels);
    case kTargetMixNorm: return launch_cfg<MixNormTarget, RNG>(a, h.dmax, h.Lmax, h.nmodels);
    case kTargetPlugin:
      return rj->tgt->d.plugin->rj_sweeps(&a, h.dmax, h.Lmax, h.nmodels, std::is_same<RNG, TapeStream>::value ? 1 : 0);
  }
  return fail(AMX_EINVAL, "plug-in kind %d has no fused sweep kernel", rj->tgt->d.kind);
}

static RjLaunch base_launch(const amx_rj *rj) {
  RjLaunch a;
  memset(&a, 0, sizeof(a));
  a.st = rj->st;
  a.prop_blob = rj->prop->blob_dev;
  a.prop_bytes = rj->prop->blob_bytes;
  a.tgt_blob = rj->tgt->d.blob_dev;
  a.tgt_bytes = rj->tgt->d.blob_bytes;
  a.tgt_flags = rj->tgt->d.flags;
  a.seed = rj->seed;
  a.chain_base = rj->chain_base;
  a.modes = rj->modes;
  a.tape = rj->tape_dev;
  a.tape_stride = rj->tape_stride;
  a.visits = rj->visits_dev;
  a.visits_grp = rj->grp_dev;
  a.cnt = rj->cnt_dev;
  a.status = rj->status_dev;
  return a;
}


static bool is_host_target(const amx_rj *rj) {
  return rj->tgt->d.kind == kTargetHostScalar || rj->tgt->d.kind == kTargetHostBatched;
}

// ---- sorted mode (rj_sort_*_kernel, amx_rj_kernels.cuh) ----------------------------------------------------------
// Sweeps per sort for this call: 0 = unsorted.  Automatic: populations of >= 65536 chains in the wide configurations
// (a model with more than 8 coordinates) on a device plug-in, where a warp of mixed models pays for its widest chain.
// A sorted sweep costs a fixed ~0.1-0.3 ms (three sort kernels, a launch with cold caches, the latency of the widest
// tile): smaller populations -- the drop-in's default 16384 chains: 0.16 ms per sweep unsorted, 0.42 sorted -- stay on
// one launch for all sweeps.  The small register-resident configurations (d <= 8) run hundreds of sweeps per launch
// with their state in registers and lose more to the per-sort state traffic than they gain.
static int sort_segment(const amx_rj *rj) {
  if (is_host_target(rj) || rj->nm < 2) return 0;
  if (rj->sort_seg >= 0) return rj->sort_seg;
  if (const char *e = getenv("AMX_RJ_SORT")) return atoi(e);
  return (rj->dmax > 8 && rj->C >= 65536) ? 1 : 0;
}
static int sort_alloc(amx_rj *rj) {
  if (rj->so.order) return AMX_OK;
  RjSort &so = rj->so;
  const amx_fam_hdr &h = rj->prop->hdr;
  so.nb = rj->nm * rj->nm;
  // rank of a model in the order "widest first" (ties by index)
  for (int k = 0; k < rj->nm; k++) {
    int r = 0;
    for (int q = 0; q < rj->nm; q++)
      if (h.dims[q] > h.dims[k] || (h.dims[q] == h.dims[k] && q < k)) r++;
    so.rank[k] = (unsigned char)r;
  }
  AMX_CUDA(cudaMalloc(&so.keys, sizeof(int) * (size_t)rj->C));
  AMX_CUDA(cudaMalloc(&so.order, sizeof(int) * (size_t)rj->C));
  AMX_CUDA(cudaMalloc(&so.hist, sizeof(int) * 3 * kSortBuckets));
  so.start = so.hist + kSortBuckets;
  so.cursor = so.start + kSortBuckets;
  AMX_CUDA(cudaMemsetAsync(so.hist, 0, sizeof(int) * 3 * kSortBuckets, stream()));
  return AMX_OK;
}
static int sort_chains(amx_rj *rj, const RjLaunch &a) {
  const unsigned grid = (unsigned)((rj->C + kSortThreads - 1) / kSortThreads);
  if (rj->tape_dev) rj_sort_key_kernel<TapeStream><<<grid, kSortThreads, 0, stream()>>>(a, rj->so);
  else rj_sort_key_kernel<PhiloxStream><<<grid, kSortThreads, 0, stream()>>>(a, rj->so);
  rj_sort_scan_kernel<<<1, kSortBuckets, 0, stream()>>>(rj->so);
  rj_sort_scatter_kernel<<<grid, kSortThreads, 0, stream()>>>(rj->so, rj->C);
  count_launch(3);
  AMX_CUDA(cudaGetLastError());
  return AMX_OK;
}

static int split_alloc(amx_rj *rj) {
  if (rj->sp.thn) return AMX_OK;
  const size_t C = (size_t)rj->C;
  AMX_CUDA(cudaMalloc(&rj->sp.thn, sizeof(double) * C * rj->dmax));
  AMX_CUDA(cudaMalloc(&rj->sp.keval, sizeof(int) * C));
  AMX_CUDA(cudaMalloc(&rj->sp.lpn, sizeof(double) * C));
  AMX_CUDA(cudaMalloc(&rj->sp.kn, sizeof(int) * C));
  AMX_CUDA(cudaMalloc(&rj->sp.carry, sizeof(double) * 5 * C));
  AMX_CUDA(cudaMemsetAsync(rj->sp.thn, 0, sizeof(double) * C * rj->dmax, stream()));
  AMX_CUDA(cudaMemsetAsync(rj->sp.lpn, 0, sizeof(double) * C, stream()));
  AMX_CUDA(cudaMemsetAsync(rj->sp.kn, 0, sizeof(int) * C, stream()));
  AMX_CUDA(cudaMemsetAsync(rj->sp.carry, 0, sizeof(double) * 5 * C, stream()));
  AMX_CUDA(cudaMallocHost(&rj->h_thn, sizeof(double) * C * rj->dmax));
  AMX_CUDA(cudaMallocHost(&rj->h_lpn, sizeof(double) * C));
  AMX_CUDA(cudaMallocHost(&rj->h_keval, sizeof(int) * C));
  rj->h_kc = new std::vector<int>();
  rj->h_xc = new std::vector<double>();
  rj->h_lc = new std::vector<double>();
  return AMX_OK;
}

// proposals -> host, callback on the chains that take part in this phase, values -> device
static int split_evaluate(amx_rj *rj) {
  const size_t C = (size_t)rj->C;
  const int dmax = rj->dmax;
  AMX_CUDA(cudaMemcpyAsync(rj->h_thn, rj->sp.thn, sizeof(double) * C * dmax, cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaMemcpyAsync(rj->h_keval, rj->sp.keval, sizeof(int) * C, cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  const TargetDesc &t = rj->tgt->d;
  if (t.kind == kTargetHostScalar) {
    for (size_t c = 0; c < C; c++)
      if (rj->h_keval[c] >= 0) rj->h_lpn[c] = t.scalar(rj->h_keval[c], rj->h_thn + c * dmax);
  } else {
    std::vector<int> &kc = *rj->h_kc;
    std::vector<double> &xc = *rj->h_xc, &lc = *rj->h_lc;
    kc.clear();
    xc.clear();
    for (size_t c = 0; c < C; c++)
      if (rj->h_keval[c] >= 0) {
        kc.push_back(rj->h_keval[c]);
        xc.insert(xc.end(), rj->h_thn + c * dmax, rj->h_thn + (c + 1) * dmax);
      }
    lc.resize(kc.size());
    if (!kc.empty()) t.batched((long)kc.size(), kc.data(), xc.data(), dmax, lc.data(), t.user);
    size_t q = 0;
    for (size_t c = 0; c < C; c++)
      if (rj->h_keval[c] >= 0) rj->h_lpn[c] = lc[q++];
  }
  AMX_CUDA(cudaMemcpyAsync(rj->sp.lpn, rj->h_lpn, sizeof(double) * C, cudaMemcpyHostToDevice, stream()));
  return AMX_OK;
}

static int split_phase(amx_rj *rj, const RjLaunch &a, int phase, int j, int s) {
  const unsigned grid = (unsigned)((rj->C + kRjThreads - 1) / kRjThreads);
  if (rj->tape_dev) rj_split_kernel<TapeStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->sp, phase, j, s);
  else rj_split_kernel<PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->sp, phase, j, s);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  return AMX_OK;
}

// host callbacks through the mailbox: launch the persistent kernel, then serve it from this thread until every CTA
// has had all its values (AMX_HOST_MAILBOX=0 keeps the kernel-per-evaluation path below)
static int mailbox_sweeps(amx_rj *rj, RjLaunch &a) {
  const char *ce = getenv("AMX_MAILBOX_CPC");
  int cpc = ce ? atoi(ce) : 32;
  cpc = cpc < 1 ? 1 : (cpc > kMbThreads ? kMbThreads : cpc);
  const int ncta = (int)((rj->C + cpc - 1) / cpc), ldx = rj->dmax;
  if (!rj->mb_host)
    if (int rc = mailbox_alloc(&rj->mb_host, ncta, ldx)) return rc;
  long nexch = 0;
  for (int s = 0; s < a.nsweeps; s++) nexch += (((a.sweep0 + (unsigned long long)s) % 10ull == 0ull) ? 1 : rj->dmax) + 1;
  if (rj->mb_seq + (unsigned long long)nexch > 0xFFFFFF00ull) {  // 32-bit request numbers: start a fresh set of mailboxes
    AMX_CUDA(cudaStreamSynchronize(stream()));
    cudaFreeHost(rj->mb_host);
    rj->mb_host = nullptr;
    rj->mb_seq = 0;
    if (int rc = mailbox_alloc(&rj->mb_host, ncta, ldx)) return rc;
  }
  const int nthr = (cpc + 31) / 32 * 32;  // no more threads than chains: every thread of the CTA writes its slot each exchange
  if (rj->tape_dev) rj_mailbox_kernel<TapeStream><<<ncta, nthr, 0, stream()>>>(a, rj->mb_host, ldx, rj->mb_seq, cpc);
  else rj_mailbox_kernel<PhiloxStream><<<ncta, nthr, 0, stream()>>>(a, rj->mb_host, ldx, rj->mb_seq, cpc);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  std::vector<MbJob> jobs(1);
  jobs[0] = {rj->mb_host, ncta, ldx, nthr, nexch, stream(), rj->mb_seq};
  rj->mb_seq += (unsigned)nexch;
  if (int rc = mailbox_serve(jobs, rj->tgt->d)) return rc;
  return AMX_OK;
}
static bool use_mailbox() {
  const char *e = getenv("AMX_HOST_MAILBOX");
  return !(e && atoi(e) == 0);
}

static int split_sweeps(amx_rj *rj, RjLaunch &a) {
  if (int rc = split_alloc(rj)) return rc;
  for (int s = 0; s < a.nsweeps; s++) {
    const unsigned long long sweep_i = a.sweep0 + (unsigned long long)s;
    int rc = 0;
    if (sweep_i % 10ull == 0ull) {
      if ((rc = split_phase(rj, a, kPhBlockPropose, 0, s)) || (rc = split_evaluate(rj)) ||
          (rc = split_phase(rj, a, kPhBlockFinish, 0, s)))
        return rc;
    } else {
      for (int j = 0; j < rj->dmax; j++)
        if ((rc = split_phase(rj, a, kPhCoordPropose, j, s)) || (rc = split_evaluate(rj)) ||
            (rc = split_phase(rj, a, kPhCoordFinish, j, s)))
          return rc;
    }
    if ((rc = split_phase(rj, a, kPhJumpPropose, 0, s)) || (rc = split_evaluate(rj)) ||
        (rc = split_phase(rj, a, kPhJumpFinish, 0, s)))
      return rc;
  }
  return AMX_OK;
}

extern "C" {

amx_rj *amx_rj_create(const amx_proposal *p, const amx_target *t, long nchains, const double *init_flat,
                      uint64_t seed, int n_trace) {
  if (require_device()) return nullptr;
  if (!p || !t || nchains < 1) {
    fail(AMX_EINVAL, "amx_rj_create: null proposal/target or nchains<1");
    return nullptr;
  }
  if (t->d.nmodels != p->hdr.nmodels) {
    fail(AMX_EINVAL, "proposal has %d models, plug-in %d", p->hdr.nmodels, t->d.nmodels);
    return nullptr;
  }
  for (int k = 0; k < p->hdr.nmodels; k++)
    if (t->d.dims[k] != p->hdr.dims[k]) {
      fail(AMX_EINVAL, "model %d: proposal dimension %d, plug-in %d", k, p->hdr.dims[k], t->d.dims[k]);
      return nullptr;
    }
  amx_rj *rj = new amx_rj();
  memset(rj, 0, sizeof(*rj));
  rj->prop = p;
  rj->tgt = t;
  rj->C = nchains;
  rj->dmax = p->hdr.dmax;
  rj->nm = p->hdr.nmodels;
  rj->seed = seed;
  rj->sweep_i = 1;
  rj->ntrace = n_trace < 0 ? 0 : (n_trace > nchains ? (int)nchains : n_trace);
  RjState &s = rj->st;
  s.C = nchains;
  s.dmax = rj->dmax;
  s.nmodels = rj->nm;
  int total_d = 0;
  for (int k = 0; k < rj->nm; k++) total_d += p->hdr.dims[k];
  AMX_CUDA_PTR(cudaMalloc(&s.theta, sizeof(double) * (size_t)rj->dmax * nchains));
  AMX_CUDA_PTR(cudaMalloc(&s.pk, sizeof(double) * (size_t)rj->nm * nchains));
  AMX_CUDA_PTR(cudaMalloc(&s.lp, sizeof(double) * nchains));
  AMX_CUDA_PTR(cudaMalloc(&s.pkllim, sizeof(double) * nchains));
  AMX_CUDA_PTR(cudaMalloc(&s.k, sizeof(int) * nchains));
  AMX_CUDA_PTR(cudaMalloc(&s.nreinit, sizeof(int) * nchains));
  AMX_CUDA_PTR(cudaMalloc(&s.draws, sizeof(unsigned long long) * nchains));
  AMX_CUDA_PTR(cudaMemsetAsync(s.draws, 0, sizeof(unsigned long long) * nchains, stream()));
  AMX_CUDA_PTR(cudaMalloc(&rj->init_dev, sizeof(double) * total_d));
  AMX_CUDA_PTR(cudaMemcpyAsync(rj->init_dev, init_flat, sizeof(double) * total_d, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA_PTR(cudaMalloc(&rj->visits_dev, sizeof(unsigned long long) * AMX_MAX_MODELS));
  AMX_CUDA_PTR(cudaMalloc(&rj->cnt_dev, sizeof(unsigned long long) * 8));
  AMX_CUDA_PTR(cudaMalloc(&rj->grp_dev, sizeof(rj->grp_host)));
  AMX_CUDA_PTR(cudaMemsetAsync(rj->grp_dev, 0, sizeof(rj->grp_host), stream()));
  AMX_CUDA_PTR(cudaMalloc(&rj->status_dev, sizeof(int)));
  AMX_CUDA_PTR(cudaMemsetAsync(rj->visits_dev, 0, sizeof(unsigned long long) * AMX_MAX_MODELS, stream()));
  AMX_CUDA_PTR(cudaMemsetAsync(rj->cnt_dev, 0, sizeof(unsigned long long) * 8, stream()));
  AMX_CUDA_PTR(cudaMemsetAsync(rj->status_dev, 0, sizeof(int), stream()));
  AMX_CUDA_PTR(cudaMalloc(&rj->pk_shared, sizeof(RjPkShared)));
  AMX_CUDA_PTR(cudaMemsetAsync(rj->pk_shared, 0, sizeof(RjPkShared), stream()));
  rj_pk_reset_kernel<<<1, 32, 0, stream()>>>(rj->pk_shared, rj->nm);
  rj->pk_mode = AMX_PK_PER_CHAIN;
  rj->pk_seg = 25;
  rj->sort_seg = -1;
  AMX_CUDA_PTR(cudaEventCreate(&rj->e0));
  AMX_CUDA_PTR(cudaEventCreate(&rj->e1));
  rj->pending = new std::vector<std::pair<cudaEvent_t, cudaEvent_t>>();
  AMX_CUDA_PTR(cudaStreamSynchronize(stream()));
  return rj;
}

void amx_rj_destroy(amx_rj *rj) {
  if (!rj) return;
  RjState &s = rj->st;
  cudaFree(s.theta); cudaFree(s.pk); cudaFree(s.lp); cudaFree(s.pkllim); cudaFree(s.k);
  cudaFree(s.nreinit); cudaFree(s.draws); cudaFree(rj->init_dev); cudaFree(rj->tape_dev);
  cudaFree(rj->visits_dev); cudaFree(rj->cnt_dev); cudaFree(rj->status_dev); cudaFree(rj->gam_dev); cudaFree(rj->pk_shared); cudaFree(rj->grp_dev);
  cudaFree(rj->tr_k); cudaFree(rj->tr_lp); cudaFree(rj->tr_theta); cudaFree(rj->tr_pk);
  cudaFree(rj->so.keys); cudaFree(rj->so.order); cudaFree(rj->so.hist);
  cudaEventDestroy(rj->e0);
  cudaEventDestroy(rj->e1);
  for (auto &pr : *rj->pending) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  delete rj->pending;
  moments_free(rj->mom);
  cudaFree(rj->stage_dev);
  cudaFree(rj->sp.thn); cudaFree(rj->sp.keval); cudaFree(rj->sp.lpn); cudaFree(rj->sp.kn); cudaFree(rj->sp.carry);
  if (rj->mb_host) cudaFreeHost(rj->mb_host);
  if (rj->h_thn) cudaFreeHost(rj->h_thn);
  if (rj->h_lpn) cudaFreeHost(rj->h_lpn);
  if (rj->h_keval) cudaFreeHost(rj->h_keval);
  delete rj->h_kc;
  delete rj->h_xc;
  delete rj->h_lc;
  delete rj;
}

// ---- posterior summaries (kernels in amx_summary.cu) ------------------------------------------------
int amx_rj_moments_reset(amx_rj *rj) {
  if (!rj) return fail(AMX_EINVAL, "null population");
  return moments_reset(&rj->mom, rj->prop->hdr);
}

int amx_rj_moments_accumulate(amx_rj *rj) {
  if (!rj) return fail(AMX_EINVAL, "null population");
  if (!rj->mom)
    if (int rc = moments_reset(&rj->mom, rj->prop->hdr)) return rc;
  return moments_accumulate(rj->mom, rj->st.k, rj->st.theta, rj->st.lp, rj->C, rj->dmax, rj->prop->blob_dev);
}

int amx_rj_moments_get(const amx_rj *rj, int model, unsigned long long *count, double *mean, double *cov,
                       double *mean_lp) {
  if (!rj || !rj->mom) return fail(AMX_EINVAL, "no moments accumulated");
  return moments_get(rj->mom, rj->prop, model, count, mean, cov, mean_lp);
}

int amx_rj_sokal(const amx_rj *rj, long nkeep, double *var, double *tau, int *m) {
  if (!rj || rj->ntrace < 1 || rj->last_nsweeps < 1) return fail(AMX_EINVAL, "no trace recorded");
  return sokal_ktrace(rj->tr_k, rj->last_nsweeps, nkeep, rj->ntrace, var, tau, m);
}

int amx_rj_set_modes(amx_rj *rj, int student_t_dof, int do_perm) {
  if (!rj || student_t_dof < 0) return fail(AMX_EINVAL, "amx_rj_set_modes: bad arguments");
  rj->modes.dof = student_t_dof;
  rj->modes.do_perm = do_perm ? 1 : 0;
  rj->modes.lt_const = student_t_dof > 0 ? lgamma(0.5 * (student_t_dof + 1)) - lgamma(0.5 * student_t_dof) -
                                               0.5 * log(student_t_dof * 3.14159265358979323846)
                                         : 0.0;
  return AMX_OK;
}

int amx_rj_set_pk_mode(amx_rj *rj, int mode, int segment_sweeps) {
  if (!rj || (mode != AMX_PK_PER_CHAIN && mode != AMX_PK_POPULATION) || segment_sweeps < 0)
    return fail(AMX_EINVAL, "amx_rj_set_pk_mode: bad arguments");
  rj->pk_mode = mode;
  if (segment_sweeps > 0) rj->pk_seg = segment_sweeps;
  return AMX_OK;
}

int amx_rj_get_pk_shared(const amx_rj *rj, double *pk, int *nreinit, double *pkllim) {
  if (!rj) return fail(AMX_EINVAL, "null handle");
  RjPkShared h;
  AMX_CUDA(cudaMemcpyAsync(&h, rj->pk_shared, sizeof(h), cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  if (pk)
    for (int j = 0; j < rj->nm; j++) pk[j] = h.pk[j];
  if (nreinit) *nreinit = h.nreinit;
  if (pkllim) *pkllim = h.pkllim;
  return AMX_OK;
}

int amx_rj_set_chain_base(amx_rj *rj, uint64_t first_chain_id) {
  if (!rj) return fail(AMX_EINVAL, "null handle");
  rj->chain_base = first_chain_id;
  return AMX_OK;
}

int amx_rj_set_tape(amx_rj *rj, const double *tape, long stride) {
  if (!rj || !tape || stride < 1) return fail(AMX_EINVAL, "amx_rj_set_tape: bad arguments");
  cudaFree(rj->tape_dev);
  rj->tape_dev = nullptr;
  const size_t nb = sizeof(double) * (size_t)stride * rj->C;
  AMX_CUDA(cudaMalloc(&rj->tape_dev, nb));
  AMX_CUDA(cudaMemcpyAsync(rj->tape_dev, tape, nb, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  rj->tape_stride = (unsigned long long)stride;
  // a new tape is read from its beginning
  AMX_CUDA(cudaMemsetAsync(rj->st.draws, 0, sizeof(unsigned long long) * rj->C, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  return AMX_OK;
}

int amx_rj_init_chains(amx_rj *rj) {
  if (!rj) return fail(AMX_EINVAL, "null handle");
  rj_pk_reset_kernel<<<1, 32, 0, stream()>>>(rj->pk_shared, rj->nm);  // pk = 1/nmodels, nreinit = 1, pkllim = 0.1 (:441-446)
  count_launch();
  RjLaunch a = base_launch(rj);
  const unsigned grid = (unsigned)((rj->C + kRjThreads - 1) / kRjThreads);
  const bool tape = rj->tape_dev != nullptr;
  if (is_host_target(rj)) {
    if (int rc = split_alloc(rj)) return rc;
    for (int finish = 0; finish < 2; finish++) {
      if (tape) rj_init_split_kernel<TapeStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->sp, rj->init_dev, finish);
      else rj_init_split_kernel<PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->sp, rj->init_dev, finish);
      count_launch();
      AMX_CUDA(cudaGetLastError());
      if (!finish)
        if (int rc = split_evaluate(rj)) return rc;
    }
    rj->sweep_i = 1;
    return AMX_OK;
  }
  switch (rj->tgt->d.kind) {
    case kTargetGaussMix:
      if (tape) rj_init_kernel<GaussMixTarget, TapeStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      else rj_init_kernel<GaussMixTarget, PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      break;
    case kTargetQuad:
      if (tape) rj_init_kernel<QuadTarget, TapeStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      else rj_init_kernel<QuadTarget, PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      break;
    case kTargetCoal:
      if (tape) rj_init_kernel<CoalTarget, TapeStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      else rj_init_kernel<CoalTarget, PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      break;
    case kTargetMixNorm:
      if (tape) rj_init_kernel<MixNormTarget, TapeStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      else rj_init_kernel<MixNormTarget, PhiloxStream><<<grid, kRjThreads, 0, stream()>>>(a, rj->init_dev);
      break;
    case kTargetPlugin:
      if (int rc = rj->tgt->d.plugin->rj_init(&a, rj->init_dev, tape ? 1 : 0)) return rc;
      rj->sweep_i = 1;
      return AMX_OK;
    default:
      return fail(AMX_EINVAL, "plug-in kind %d has no device chain start", rj->tgt->d.kind);
  }
  count_launch();
  AMX_CUDA(cudaGetLastError());
  rj->sweep_i = 1;
  return AMX_OK;
}

int amx_rj_set_state(amx_rj *rj, long first, long count, const double *theta, const double *pk,
                     const double *lp, const int *k, const int *nreinit, const double *pkllim,
                     unsigned long long sweep_i) {
  if (!rj || first < 0 || count < 1 || first + count > rj->C) return fail(AMX_EINVAL, "amx_rj_set_state: range");
  RjState &s = rj->st;
  // host arrays are chain-major; the device layout is coordinate-major: one contiguous copy, then a
  // transposing kernel
  const long need = count * (rj->dmax + rj->nm);
  if (need > rj->stage_cap) {
    cudaFree(rj->stage_dev);
    rj->stage_dev = nullptr;
    AMX_CUDA(cudaMalloc(&rj->stage_dev, sizeof(double) * need));
    rj->stage_cap = need;
  }
  double *th_cm = rj->stage_dev, *pk_cm = rj->stage_dev + count * rj->dmax;
  AMX_CUDA(cudaMemcpyAsync(th_cm, theta, sizeof(double) * count * rj->dmax, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaMemcpyAsync(pk_cm, pk, sizeof(double) * count * rj->nm, cudaMemcpyHostToDevice, stream()));
  rj_state_scatter_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream()>>>(s, first, count, th_cm, pk_cm);
  count_launch();
  AMX_CUDA(cudaGetLastError());
  AMX_CUDA(cudaMemcpyAsync(s.lp + first, lp, sizeof(double) * count, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaMemcpyAsync(s.k + first, k, sizeof(int) * count, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaMemcpyAsync(s.nreinit + first, nreinit, sizeof(int) * count, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaMemcpyAsync(s.pkllim + first, pkllim, sizeof(double) * count, cudaMemcpyHostToDevice, stream()));
  if (!defer_sync()) AMX_CUDA(cudaStreamSynchronize(stream()));
  rj->sweep_i = sweep_i;
  return AMX_OK;
}

int amx_rj_get_state(const amx_rj *rj, long first, long count, double *theta, double *pk, double *lp, int *k,
                     int *nreinit, double *pkllim, unsigned long long *sweep_i) {
  if (!rj || first < 0 || count < 1 || first + count > rj->C) return fail(AMX_EINVAL, "amx_rj_get_state: range");
  const RjState &s = rj->st;
  amx_rj *w = const_cast<amx_rj *>(rj);
  if (theta || pk) {
    const long need = count * (rj->dmax + rj->nm);
    if (need > w->stage_cap) {
      cudaFree(w->stage_dev);
      w->stage_dev = nullptr;
      AMX_CUDA(cudaMalloc(&w->stage_dev, sizeof(double) * need));
      w->stage_cap = need;
    }
    double *th_cm = w->stage_dev, *pk_cm = w->stage_dev + count * rj->dmax;
    rj_state_gather_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream()>>>(s, first, count, theta ? th_cm : nullptr,
                                                                                  pk ? pk_cm : nullptr);
    count_launch();
    AMX_CUDA(cudaGetLastError());
    if (theta) AMX_CUDA(cudaMemcpyAsync(theta, th_cm, sizeof(double) * count * rj->dmax, cudaMemcpyDeviceToHost, stream()));
    if (pk) AMX_CUDA(cudaMemcpyAsync(pk, pk_cm, sizeof(double) * count * rj->nm, cudaMemcpyDeviceToHost, stream()));
  }
  if (lp) AMX_CUDA(cudaMemcpyAsync(lp, s.lp + first, sizeof(double) * count, cudaMemcpyDeviceToHost, stream()));
  if (k) AMX_CUDA(cudaMemcpyAsync(k, s.k + first, sizeof(int) * count, cudaMemcpyDeviceToHost, stream()));
  if (nreinit) AMX_CUDA(cudaMemcpyAsync(nreinit, s.nreinit + first, sizeof(int) * count, cudaMemcpyDeviceToHost, stream()));
  if (pkllim) AMX_CUDA(cudaMemcpyAsync(pkllim, s.pkllim + first, sizeof(double) * count, cudaMemcpyDeviceToHost, stream()));
  if (!defer_sync()) AMX_CUDA(cudaStreamSynchronize(stream()));  // deferred: the caller owns pinned buffers and syncs
  if (sweep_i) *sweep_i = rj->sweep_i;
  return AMX_OK;
}

int amx_rj_sweeps(amx_rj *rj, long nsweeps, int burning, int do_adapt) {
  if (!rj || nsweeps < 1 || nsweeps > 2000000000L) return fail(AMX_EINVAL, "amx_rj_sweeps: bad arguments");
  if (nsweeps > rj->gam_cap) {
    cudaFree(rj->gam_dev);
    rj->gam_dev = nullptr;
    rj->gam_cap = 0;
    AMX_CUDA(cudaMalloc(&rj->gam_dev, sizeof(double) * nsweeps));
    rj->gam_cap = nsweeps;
  }
  if (rj->ntrace > 0 && nsweeps * rj->ntrace > rj->tr_cap) {
    cudaFree(rj->tr_k); cudaFree(rj->tr_lp); cudaFree(rj->tr_theta); cudaFree(rj->tr_pk);
    rj->tr_k = nullptr;
    rj->tr_lp = rj->tr_theta = rj->tr_pk = nullptr;
    rj->tr_cap = 0;
    const size_t rows = (size_t)nsweeps * rj->ntrace;
    AMX_CUDA(cudaMalloc(&rj->tr_k, sizeof(int) * rows));
    AMX_CUDA(cudaMalloc(&rj->tr_lp, sizeof(double) * rows));
    AMX_CUDA(cudaMalloc(&rj->tr_theta, sizeof(double) * rows * rj->dmax));
    AMX_CUDA(cudaMalloc(&rj->tr_pk, sizeof(double) * rows * rj->nm));
    rj->tr_cap = (long)rows;
  }
  RjLaunch a = base_launch(rj);
  const int adapt = (do_adapt && !burning) ? 1 : 0;
  const bool pop = rj->pk_mode == AMX_PK_POPULATION;
  a.ntrace = rj->ntrace;
  a.tr_stride = nsweeps;
  a.tr_k = rj->tr_k;
  a.tr_lp = rj->tr_lp;
  a.tr_theta = rj->tr_theta;
  a.tr_pk = rj->tr_pk;
  a.pk_shared = pop ? rj->pk_shared->pk : nullptr;
  a.adapt = pop ? 0 : adapt;  // population mode: the shared pk moves between segments, not inside the sweep
  cudaEvent_t e0, e1;
  AMX_CUDA(cudaEventCreate(&e0));
  AMX_CUDA(cudaEventCreate(&e1));
  rj_gamma_kernel<<<(unsigned)((nsweeps + 255) / 256), 256, 0, stream()>>>(rj->gam_dev, rj->sweep_i, (int)nsweeps);
  count_launch();
  AMX_CUDA(cudaEventRecord(e0, stream()));
  // Population pk mode while adapting: segments of pk_seg sweeps, each followed by the shared update.  Sorted mode:
  // launches of sort_seg sweeps, each preceded by the sort.  Otherwise the whole call is one launch.
  const long seg = (pop && adapt) ? (long)rj->pk_seg : nsweeps;
  const long sseg = sort_segment(rj);
  int rc = AMX_OK;
  if (sseg > 0) {
    rc = sort_alloc(rj);
    a.order = rj->so.order;
  }
  for (long off = 0; off < nsweeps && rc == AMX_OK; off += seg) {
    const long m = (nsweeps - off < seg) ? nsweeps - off : seg;
    const long step = sseg > 0 ? sseg : m;
    for (long o2 = 0; o2 < m && rc == AMX_OK; o2 += step) {
      a.gam = rj->gam_dev + off + o2;
      a.sweep0 = rj->sweep_i + (unsigned long long)(off + o2);
      a.nsweeps = (int)((m - o2 < step) ? m - o2 : step);
      a.tr_off = off + o2;
      if (sseg > 0) rc = sort_chains(rj, a);
      if (rc) break;
      if (is_host_target(rj)) rc = use_mailbox() ? mailbox_sweeps(rj, a) : split_sweeps(rj, a);
      else rc = rj->tape_dev ? launch_tgt<TapeStream>(rj, a) : launch_tgt<PhiloxStream>(rj, a);
    }
    if (rc == AMX_OK && pop) {  // also while burning: the histogram baseline must follow the visits
      rj_pk_population_kernel<<<1, 32, 0, stream()>>>(rj->pk_shared, rj->visits_dev, rj->gam_dev + off, (int)m, rj->nm, adapt);
      count_launch();
    }
  }
  if (rc) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
  }
  AMX_CUDA(cudaEventRecord(e1, stream()));
  rj->pending->push_back({e0, e1});
  rj->sweep_i += (unsigned long long)nsweeps;
  rj->last_nsweeps = nsweeps;
  return AMX_OK;
}

int amx_rj_set_sort(amx_rj *rj, int sweeps_per_sort) {
  if (!rj || sweeps_per_sort < -1) return fail(AMX_EINVAL, "amx_rj_set_sort: bad arguments");
  rj->sort_seg = sweeps_per_sort;
  return AMX_OK;
}

int amx_rj_collect(amx_rj *rj, unsigned long long *visits, amx_rj_stats *st, int reset) {
  if (!rj) return fail(AMX_EINVAL, "null handle");
  AMX_CUDA(cudaStreamSynchronize(stream()));
  for (auto &pr : *rj->pending) {
    float ms = 0;
    AMX_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    rj->kernel_ms += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  rj->pending->clear();
  unsigned long long cnt[8];
  int status = 0;
  AMX_CUDA(cudaMemcpyAsync(cnt, rj->cnt_dev, sizeof(cnt), cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaMemcpyAsync(&status, rj->status_dev, sizeof(int), cudaMemcpyDeviceToHost, stream()));
  if (visits)
    AMX_CUDA(cudaMemcpyAsync(visits, rj->visits_dev, sizeof(unsigned long long) * rj->nm, cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaMemcpyAsync(rj->grp_host, rj->grp_dev, sizeof(rj->grp_host), cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  if (st) {
    st->acc_block = cnt[0]; st->try_block = cnt[1]; st->acc_single = cnt[2]; st->try_single = cnt[3];
    st->acc_jump = cnt[4]; st->try_jump = cnt[5]; st->flops = cnt[6]; st->draws = cnt[7];
    st->kernel_ms = rj->kernel_ms;
  }
  if (reset) {  // on the library stream: the next sweep kernel's atomics are ordered after these
    AMX_CUDA(cudaMemsetAsync(rj->visits_dev, 0, sizeof(unsigned long long) * AMX_MAX_MODELS, stream()));
    AMX_CUDA(cudaMemsetAsync(rj->pk_shared->prev, 0, sizeof(unsigned long long) * AMX_MAX_MODELS, stream()));
    AMX_CUDA(cudaMemsetAsync(rj->cnt_dev, 0, sizeof(unsigned long long) * 8, stream()));
    AMX_CUDA(cudaMemsetAsync(rj->grp_dev, 0, sizeof(rj->grp_host), stream()));
    AMX_CUDA(cudaMemsetAsync(rj->status_dev, 0, sizeof(int), stream()));
    rj->kernel_ms = 0.0;
  }
  if (status & 1) return fail(AMX_ETAPE, "injected uniform tape exhausted");
  if (status & 2) return fail(AMX_ENUMERIC, "a chain reached a NaN log-posterior");
  return AMX_OK;
}

int amx_rj_visit_se(const amx_rj *rj, double *p, double *se, int *ngroups) {
  if (!rj) return fail(AMX_EINVAL, "null handle");
  const int nm = rj->nm;
  double tot = 0.0, gt[AMX_RJ_GROUPS];
  int G = 0;
  for (int g = 0; g < AMX_RJ_GROUPS; g++) {
    gt[g] = 0.0;
    for (int k = 0; k < nm; k++) gt[g] += (double)rj->grp_host[g * AMX_MAX_MODELS + k];
    tot += gt[g];
    if (gt[g] > 0.0) G++;
  }
  if (ngroups) *ngroups = G;
  for (int k = 0; k < nm; k++) {
    double vk = 0.0;
    for (int g = 0; g < AMX_RJ_GROUPS; g++) vk += (double)rj->grp_host[g * AMX_MAX_MODELS + k];
    const double pk = tot > 0.0 ? vk / tot : 0.0;
    double v = 0.0;
    for (int g = 0; g < AMX_RJ_GROUPS; g++)
      if (gt[g] > 0.0) {
        const double w = gt[g] / tot, dg = (double)rj->grp_host[g * AMX_MAX_MODELS + k] / gt[g] - pk;
        v += w * w * dg * dg;
      }
    if (p) p[k] = pk;
    if (se) se[k] = G > 1 ? sqrt(v * (double)G / (double)(G - 1)) : 0.0 / 0.0;
  }
  return AMX_OK;
}

int amx_rj_get_trace(const amx_rj *rj, int *k, double *lp, double *theta, double *pk) {
  if (!rj || rj->ntrace < 1 || rj->last_nsweeps < 1) return fail(AMX_EINVAL, "no trace recorded");
  AMX_CUDA(cudaStreamSynchronize(stream()));
  const size_t rows = (size_t)rj->last_nsweeps * rj->ntrace;
  if (k) AMX_CUDA(cudaMemcpy(k, rj->tr_k, sizeof(int) * rows, cudaMemcpyDeviceToHost));
  if (lp) AMX_CUDA(cudaMemcpy(lp, rj->tr_lp, sizeof(double) * rows, cudaMemcpyDeviceToHost));
  if (theta) AMX_CUDA(cudaMemcpy(theta, rj->tr_theta, sizeof(double) * rows * rj->dmax, cudaMemcpyDeviceToHost));
  if (pk) AMX_CUDA(cudaMemcpy(pk, rj->tr_pk, sizeof(double) * rows * rj->nm, cudaMemcpyDeviceToHost));
  return AMX_OK;
}

void *amx_rj_visits_dev(amx_rj *rj) { return rj ? rj->visits_dev : nullptr; }

int amx_target_eval(const amx_target *t, long n, const int *model_k, const double *x, long ldx, double *lp_out) {
  if (!t || n < 1) return fail(AMX_EINVAL, "amx_target_eval: bad arguments");
  if (t->d.kind == kTargetHostScalar) {
    for (long i = 0; i < n; i++) lp_out[i] = t->d.scalar(model_k[i], const_cast<double *>(x) + i * ldx);
    return AMX_OK;
  }
  if (t->d.kind == kTargetHostBatched) {
    t->d.batched(n, model_k, x, ldx, lp_out, t->d.user);
    return AMX_OK;
  }
  if (int rc = require_device()) return rc;
  int *k_dev = nullptr;
  double *x_dev = nullptr, *o_dev = nullptr;
  AMX_CUDA(cudaMalloc(&k_dev, sizeof(int) * n));
  AMX_CUDA(cudaMalloc(&x_dev, sizeof(double) * n * ldx));
  AMX_CUDA(cudaMalloc(&o_dev, sizeof(double) * n));
  AMX_CUDA(cudaMemcpyAsync(k_dev, model_k, sizeof(int) * n, cudaMemcpyHostToDevice, stream()));
  AMX_CUDA(cudaMemcpyAsync(x_dev, x, sizeof(double) * n * ldx, cudaMemcpyHostToDevice, stream()));
  const unsigned grid = (unsigned)((n + kRjThreads - 1) / kRjThreads);
  EvalDims ed;
  memset(&ed, 0, sizeof(ed));
  for (int q = 0; q < t->d.nmodels && q < AMX_MAX_MODELS; q++) ed.dims[q] = t->d.dims[q];
  switch (t->d.kind) {
    case kTargetGaussMix:
      target_eval_kernel<GaussMixTarget><<<grid, kRjThreads, 0, stream()>>>(t->d.blob_dev, t->d.flags, ed, n, k_dev, x_dev, ldx, o_dev);
      break;
    case kTargetQuad:
      target_eval_kernel<QuadTarget><<<grid, kRjThreads, 0, stream()>>>(t->d.blob_dev, t->d.flags, ed, n, k_dev, x_dev, ldx, o_dev);
      break;
    case kTargetCoal:
      target_eval_kernel<CoalTarget><<<grid, kRjThreads, 0, stream()>>>(t->d.blob_dev, t->d.flags, ed, n, k_dev, x_dev, ldx, o_dev);
      break;
    case kTargetMixNorm:
      target_eval_kernel<MixNormTarget><<<grid, kRjThreads, 0, stream()>>>(t->d.blob_dev, t->d.flags, ed, n, k_dev, x_dev, ldx, o_dev);
      break;
    case kTargetPlugin:
      if (int rc = t->d.plugin->eval(t->d.blob_dev, t->d.flags, t->d.dims, n, k_dev, x_dev, ldx, o_dev)) return rc;
      break;
  }
  count_launch();
  AMX_CUDA(cudaGetLastError());
  AMX_CUDA(cudaMemcpyAsync(lp_out, o_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, stream()));
  AMX_CUDA(cudaStreamSynchronize(stream()));
  cudaFree(k_dev);
  cudaFree(x_dev);
  cudaFree(o_dev);
  return AMX_OK;
}

}  // extern "C"
