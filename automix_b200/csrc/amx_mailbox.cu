// amx_mailbox.cu -- host side of the mailbox protocol (amx_mailbox.cuh): the calling thread serves the CTAs of one or
// several persistent kernels with values of the user's host log-posterior callback.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "amx_mailbox.cuh"
#include "amx_targets.cuh"

namespace amx {

int mailbox_alloc(void **base, int ncta, int ldx) {
  const size_t bytes = mailbox_bytes(ldx) * (size_t)ncta;
  AMX_CUDA(cudaHostAlloc(base, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(*base, 0, bytes);
  const char *e = getenv("AMX_MAILBOX_WATCHDOG_S");
  for (int b = 0; b < ncta; b++) {
    Mailbox *mb = mailbox_at(*base, b, ldx);
    mb->wd_cycles = (long long)((e && atof(e) > 0 ? atof(e) : 60.0) * 2e9);
    for (int s2 = 0; s2 < kMbThreads; s2++) mb->k[s2] = -1;
  }
  __sync_synchronize();
  return AMX_OK;
}

int mailbox_serve(std::vector<MbJob> &jobs, const TargetDesc &t) {
  struct Cta {
    Mailbox *mb;
    int ldx, job, nslots;
    unsigned next, last;
  };
  std::vector<Cta> ctas;
  long remaining = 0;
  for (size_t j = 0; j < jobs.size(); j++)
    for (int b = 0; b < jobs[j].ncta; b++) {
      ctas.push_back({mailbox_at(jobs[j].base, b, jobs[j].ldx), jobs[j].ldx, (int)j, jobs[j].nslots, jobs[j].seq0 + 1u,
                      jobs[j].seq0 + (unsigned)jobs[j].nexch});
      remaining += jobs[j].nexch;
    }
  std::vector<int> kc, slot;
  std::vector<double> xc, lc;
  unsigned long idle = 0;
  while (remaining > 0) {
    bool any = false;
    for (Cta &c : ctas) {
      if (c.next > c.last) continue;
      if (__atomic_load_n((const unsigned *)&c.mb->req_seq, __ATOMIC_ACQUIRE) != c.next) continue;
      any = true;
      double *x = mailbox_x(c.mb);
      if (t.kind == kTargetHostScalar) {
        for (int s = 0; s < c.nslots; s++)
          if (c.mb->k[s] >= 0) c.mb->lp[s] = t.scalar(c.mb->k[s], x + (size_t)s * c.ldx);
      } else {
        kc.clear();
        slot.clear();
        xc.clear();
        for (int s = 0; s < c.nslots; s++)
          if (c.mb->k[s] >= 0) {
            kc.push_back(c.mb->k[s]);
            slot.push_back(s);
            xc.insert(xc.end(), x + (size_t)s * c.ldx, x + (size_t)(s + 1) * c.ldx);
          }
        lc.resize(kc.size());
        if (!kc.empty()) t.batched((long)kc.size(), kc.data(), xc.data(), c.ldx, lc.data(), t.user);
        for (size_t q = 0; q < kc.size(); q++) c.mb->lp[slot[q]] = lc[q];
      }
      __atomic_store_n((unsigned *)&c.mb->resp_seq, c.next, __ATOMIC_RELEASE);
      c.next++;
      remaining--;
    }
    if (any) {
      idle = 0;
    } else if ((++idle & 0xFFFF) == 0) {  // nobody asked for a while: make sure the kernels are still there
      bool alive = false;
      for (MbJob &j : jobs) {
        const cudaError_t e = cudaStreamQuery(j.st);
        if (e == cudaErrorNotReady) alive = true;
        else if (e != cudaSuccess) {
          char buf[400];
          int o = 0;
          for (size_t q = 0; q < ctas.size() && q < 6 && o < 330; q++)
            o += snprintf(buf + o, sizeof(buf) - o, " [cta %zu: req %u next %u last %u]", q, ctas[q].mb->req_seq, ctas[q].next,
                          ctas[q].last);
          return fail(AMX_ECUDA, "host-callback kernel: %s;%s", cudaGetErrorString(e), buf);
        }
      }
      if (!alive) {  // every stream is idle: re-check once, then the kernels ended early
        bool pending = false;
        for (Cta &c : ctas)
          if (c.next <= c.last && __atomic_load_n((const unsigned *)&c.mb->req_seq, __ATOMIC_ACQUIRE) == c.next) pending = true;
        if (!pending) return fail(AMX_ECUDA, "host-callback kernel ended with %ld exchanges outstanding", remaining);
      }
    }
  }
  return AMX_OK;
}

}  // namespace amx
