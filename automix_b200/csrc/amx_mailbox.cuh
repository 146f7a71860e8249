// amx_mailbox.cuh -- log-posterior values from a HOST callback without leaving the kernel.
//
// The reference's plug-in contract is a C function `double f(int model_k, double *x)` (automix.h:46); it can only run
// on the host.  The first generation of this library cut every sweep into propose / evaluate / finish kernels with two
// copies and a stream synchronisation per evaluation (~35 us each, whatever the number of chains).  Here the chains
// stay in ONE persistent kernel, and each CTA talks to the host through a mailbox in mapped pinned host memory:
//
//   device (all threads of the CTA): write (model, point) of every chain that needs a value; fence; thread 0 raises
//   req_seq; thread 0 spins on resp_seq; fence; every thread reads its value.
//   host (the calling thread, amx::mailbox_serve): spins over the CTAs' req_seq, calls the user's function for the
//   filled slots of a CTA whose request is up, writes the values, raises that CTA's resp_seq.
//   (A variant in which every thread polls a 16-byte {value, request number} slot of its own, saving the second
//   barrier and one PCIe read, was tried and lost answers under three concurrent kernels; not pursued.)
//
// One exchange costs about two PCIe latencies (~3-5 us) plus the callbacks themselves; the chains' state never leaves
// the registers.  Every CTA must make the same, known number of exchanges (the kernels pad their sweeps with idle
// slots), so the host knows when it is done.
#pragma once

#include <vector>

#include "amx_common.cuh"

namespace amx {

constexpr int kMbThreads = 128;  // chains (threads) per CTA = slots per mailbox

struct Mailbox {                  // one per CTA, in cudaHostAllocMapped memory
  volatile unsigned req_seq;      // device -> host: request number `req_seq` is complete
  unsigned pad0[13];
  long long wd_cycles;            // how long the device waits for an answer before it gives up (set by the host)
  volatile unsigned resp_seq;     // host -> device: the values of request `resp_seq` are in lp[]
  unsigned pad1[15];
  int k[kMbThreads];              // model index to evaluate, -1 = this chain sits the exchange out
  double lp[kMbThreads];
  // followed by x[kMbThreads][ldx] doubles (slot-major: what the callback reads in place)
};

__host__ __device__ inline size_t mailbox_bytes(int ldx) { return sizeof(Mailbox) + sizeof(double) * (size_t)kMbThreads * ldx; }
__host__ __device__ inline Mailbox *mailbox_at(void *base, int cta, int ldx) {
  return reinterpret_cast<Mailbox *>(reinterpret_cast<char *>(base) + (size_t)cta * mailbox_bytes(ldx));
}
__host__ __device__ inline double *mailbox_x(Mailbox *m) { return reinterpret_cast<double *>(m + 1); }

// Device side of one exchange; every thread of the CTA calls it (k < 0: no request).  seq counts this CTA's exchanges
// from 1 (and never restarts).  Returns the value for this thread's request (0 for none).
template <int DMAXV>
__device__ __forceinline__ double mailbox_exchange(Mailbox *mb, int ldx, unsigned seq, int k, const double (&x)[DMAXV], int d) {
  __shared__ int s_dead;
  double *xs = mailbox_x(mb) + (size_t)threadIdx.x * ldx;
  mb->k[threadIdx.x] = k;
  if (k >= 0)
    for (int i = 0; i < d; i++) xs[i] = x[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long wd = mb->wd_cycles;
    mb->req_seq = seq;
    const long long t0 = clock64();
    int dead = 0;
    while (mb->resp_seq != seq) {
      if (clock64() - t0 > wd) {  // (a minute by default) without an answer: the host side is gone
        dead = 1;
        break;
      }
    }
    s_dead = dead;
    __threadfence_system();
  }
  __syncthreads();
  if (s_dead) __trap();
  const double v = (k >= 0) ? *reinterpret_cast<volatile double *>(&mb->lp[threadIdx.x]) : 0.0;
  return v;
}

// Host side.  A job = the mailboxes of one kernel (ncta CTAs, nexch exchanges each, running on stream st).
// mailbox_serve returns when every exchange of every job has been answered -- with the plug-in's scalar or batched
// host callback -- or with AMX_ECUDA if the kernels ended (faulted) first.
struct MbJob {
  void *base;
  int ncta, ldx;
  int nslots;  // threads per CTA of the kernel (<= kMbThreads): the slots the host looks at
  long nexch;
  cudaStream_t st;
  unsigned seq0;  // exchanges these mailboxes have already carried: request numbers never restart (a reset could
                  // overtake the device's read of its last answer)
};
int mailbox_alloc(void **base_host, int ncta, int ldx);
int mailbox_serve(std::vector<MbJob> &jobs, const struct TargetDesc &t);

}  // namespace amx
