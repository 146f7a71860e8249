// amx_rwm.cu -- K1: stage-1 adaptive random-walk Metropolis within one model.
// Replaces rwm_within_model (automix.c:575-662).
//
// One thread per chain; every chain runs the reference's full adaptive schedule
// (1.1 * max(nsweep2, 10000 d) sweeps, Robbins-Monro scale adaptation towards 25 % acceptance,
// 10 % block moves after the first tenth, every 10th state of the last 10000 d sweeps stored),
// so chain 0 fed an injected tape IS the reference chain, step for step.  The chains differ only
// in their counter-based RNG stream; a whole population advances in the wall time of one chain,
// which is what stage 2 needs when it pools the stored samples of many chains.
#include "amx_rwm_kernels.cuh"

namespace amx {

template <class RNG>
static int rwm_launch(const amx_target *t, const RwmArgs &a) {
  switch (t->d.kind) {
    case kTargetGaussMix: return rwm_launch_d<GaussMixTarget, RNG>(a);
    case kTargetQuad: return rwm_launch_d<QuadTarget, RNG>(a);
    case kTargetCoal: return rwm_launch_d<CoalTarget, RNG>(a);
    case kTargetMixNorm: return rwm_launch_d<MixNormTarget, RNG>(a);
    case kTargetPlugin: return t->d.plugin->rwm(&a, std::is_same<RNG, TapeStream>::value ? 1 : 0);
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
