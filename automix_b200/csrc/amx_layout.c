/*
 * amx_layout.c -- packers for the device "family blob" (include/amx_layout.h).
 * Plain C; part of the product library (and reused by the host-side target
 * callbacks of the test harness).
 */
#include "amx_layout.h"

#include <math.h>

int amx_fam_plan(amx_fam_hdr *h, int nmodels, const int *dims, const int *ncomp,
                 const int *extlen) {
  if (nmodels < 1 || nmodels > AMX_MAX_MODELS) return -1;
  h->nmodels = nmodels;
  h->dmax = 0;
  h->Lmax = 0;
  int pos = 0;
  for (int k = 0; k < AMX_MAX_MODELS; k++) {
    h->dims[k] = h->ncomp[k] = h->off[k] = h->stride[k] = h->ext[k] =
        h->extlen[k] = 0;
  }
  for (int k = 0; k < nmodels; k++) {
    int d = dims[k], L = ncomp[k];
    if (d < 1 || d > AMX_MAX_DIM || L < 1 || L > AMX_MAX_COMPS) return -1;
    h->dims[k] = d;
    h->ncomp[k] = L;
    h->stride[k] = AMX_REC_HEAD + 2 * d + d * (d + 1) / 2;
    h->off[k] = pos;
    pos += L * h->stride[k];
    h->ext[k] = pos;
    h->extlen[k] = extlen ? extlen[k] : 0;
    pos += h->extlen[k];
    if (d > h->dmax) h->dmax = d;
    if (L > h->Lmax) h->Lmax = L;
  }
  h->total = pos;
  return pos;
}

void amx_fam_pack(const amx_fam_hdr *h, int kind, const double *wt,
                  const double *mean, const double *tri, const double *ext,
                  double *data) {
  long iw = 0, im = 0, it = 0, ie = 0;
  for (int k = 0; k < h->nmodels; k++) {
    int d = h->dims[k], L = h->ncomp[k], nt = d * (d + 1) / 2;
    for (int l = 0; l < L; l++) {
      double *r = data + h->off[k] + (long)l * h->stride[k];
      const double *T = tri + it + (long)l * nt;
      double prod = 1.0;
      for (int i = 0; i < d; i++) prod *= T[AMX_TRI(i, i)];
      r[0] = wt[iw + l];
      r[1] = log(wt[iw + l]);
      if (kind == AMX_FAM_PROPOSAL) {
        r[2] = log(prod);
        r[3] = -(d / 2.0) * log(2.0 * M_PI) - r[2];
      } else {
        r[2] = wt[iw + l] * pow(2.0 * M_PI, -d / 2.0) / prod;
        r[3] = log(r[2]);
      }
      for (int i = 0; i < d; i++) {
        r[AMX_REC_HEAD + i] = mean[im + (long)l * d + i];
        r[AMX_REC_HEAD + d + i] = 1.0 / T[AMX_TRI(i, i)];
      }
      for (int i = 0; i < nt; i++) r[AMX_REC_HEAD + 2 * d + i] = T[i];
    }
    for (int i = 0; i < h->extlen[k]; i++) data[h->ext[k] + i] = ext[ie + i];
    iw += L;
    im += (long)L * d;
    it += (long)L * nt;
    ie += h->extlen[k];
  }
}
