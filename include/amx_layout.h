/*
 * amx_layout.h -- flat, pointer-free layouts shared by the C host side and the
 * CUDA side of automix-b200.
 *
 * The reference keeps its proposal distribution as nested pointer arrays
 * (`double ****B`, automix.h:134-153).  A GPU cannot chase those, so everything
 * that crosses the C-ABI is flat, and on the device it is re-packed once into a
 * single "family blob" that a CTA can stage in shared memory with one bulk copy.
 *
 * Flat interchange format (what amx.h entry points take), for nmodels models:
 *   dims[k], ncomp[k]
 *   wt  : concat over k of wt[k][l]                      (sum_k L_k doubles)
 *   mean: concat over k,l of mean[k][l][0..d_k)          (sum_k L_k d_k)
 *   tri : concat over k,l of the lower triangle of the d_k x d_k Cholesky
 *         factor, packed row-major: (i,j), j<=i, at i(i+1)/2+j
 *   ext : concat over k of a per-model vector (RWM scales sig[k][0..d_k) for a
 *         proposal, one model weight for a Gaussian-mixture target)
 *
 * Device family blob: header (amx_fam_hdr) followed by doubles.  Per model k a
 * block at hdr.off[k] holds L_k component records of `stride[k]` doubles each,
 * then the per-model ext vector at hdr.ext[k]:
 *   record = [ wt, logwt, c0, c1, mean[d], rdiag[d], tri[d(d+1)/2] ]
 *   proposal: c0 = log(prod_i B_ii)   (the reference takes log of the product,
 *                                      automix.c:1748, :1244-1245)
 *             c1 = -(d/2.0)*log(2*pi) - c0   (the constant part of lnormprob)
 *   target  : c0 = wt * (2*pi)^(-d/2) / prod_i S_ii,  c1 = log(c0)
 *   rdiag[i] = 1 / B_ii.
 */
#ifndef AMX_LAYOUT_H
#define AMX_LAYOUT_H

#define AMX_MAX_MODELS 32
#define AMX_MAX_DIM 32
#define AMX_MAX_COMPS 32

#define AMX_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
#define AMX_REC_HEAD 4 /* wt, logwt, c0, c1 */

typedef struct amx_fam_hdr {
  int nmodels;
  int dmax;
  int Lmax;
  int total; /* number of doubles following the header */
  int dims[AMX_MAX_MODELS];
  int ncomp[AMX_MAX_MODELS];
  int off[AMX_MAX_MODELS];    /* first record of model k (in doubles) */
  int stride[AMX_MAX_MODELS]; /* record length of model k */
  int ext[AMX_MAX_MODELS];    /* per-model ext vector (in doubles) */
  int extlen[AMX_MAX_MODELS];
} amx_fam_hdr;

#define AMX_FAM_PROPOSAL 0
#define AMX_FAM_TARGET 1

#ifdef __cplusplus
extern "C" {
#endif

/* Number of doubles a blob needs; fills hdr.  Returns <0 on bad arguments. */
int amx_fam_plan(amx_fam_hdr *hdr, int nmodels, const int *dims,
                 const int *ncomp, const int *extlen);

/* Fill `data` (hdr->total doubles) from the flat interchange arrays. */
void amx_fam_pack(const amx_fam_hdr *hdr, int kind, const double *wt,
                  const double *mean, const double *tri, const double *ext,
                  double *data);

#ifdef __cplusplus
}
#endif

#endif /* AMX_LAYOUT_H */
