// amx_internal.cuh -- host-side objects behind the opaque handles of include/amx.h.
#pragma once

#include "amx_common.cuh"

struct amx_proposal {
  amx_fam_hdr hdr;   // host copy (shapes, offsets)
  void *blob_dev;    // [amx_fam_hdr | doubles] on the device
  int blob_bytes;
};

namespace amx {
int upload_blob(const amx_fam_hdr &h, const double *data, void **dev, int *bytes);
}

// posterior summaries (amx_summary.cu)
namespace amx {
struct MomentsBuf {
  int *pairs_dev;       // [P] (model, entry) codes
  double *partial_dev;  // [nblocks][P]
  double *total_dev;    // [P] running totals
  int P, nblocks;
  int off[AMX_MAX_MODELS];  // first entry of model k: count, sum lp, S1[d], S2[tri]
  long snapshots;
};
int moments_reset(MomentsBuf **mb, const amx_fam_hdr &h);
int moments_accumulate(MomentsBuf *mb, const int *k, const double *theta, const double *lp, long C, int dmax,
                       const void *prop_blob);
int moments_get(const MomentsBuf *mb, const amx_proposal *prop, int model, unsigned long long *count, double *mean,
                double *cov, double *mean_lp);
void moments_free(MomentsBuf *mb);
int sokal_ktrace(const int *tr_k, long nsweeps, long nkeep, int ntrace, double *var, double *tau, int *m);
}  // namespace amx
