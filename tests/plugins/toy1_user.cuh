// toy1_user.cuh -- a USER-WRITTEN __device__ log-posterior for the plug-in path (tests/test_gpu_plugin.py): the toy1
// targets of the reference (src/user_examples/usertoy1.c:72-100: model k is modw_k * sum_g c_g exp(-q_g / 2)), written
// against the library's SDK headers the way a user would, and compared bit for bit with the built-in Gaussian-mixture
// family on the same parameter blob.
//
// Parameters: the family blob of include/amx_layout.h (amx_fam_plan + amx_fam_pack with AMX_FAM_TARGET): per
// component a record [wt, logwt, c0 = wt (2 pi)^(-d/2) / prod S_ii, c1, mean, 1/diag, packed lower triangle], one model
// weight per model.
struct AmxUserTarget {
  const amx_fam_hdr *h;
  const double *D;
  __device__ void bind(const void *blob, int /*flags*/) {
    h = reinterpret_cast<const amx_fam_hdr *>(blob);
    D = reinterpret_cast<const double *>(h + 1);
  }
  __device__ int flops(int k) const {
    const int d = h->dims[k];
    return h->ncomp[k] * (d * d + 3 * d + 6) + 2;
  }
  template <int DMAX>
  __device__ double eval(int k, const double (&x)[DMAX]) const {
    const int d = h->dims[k], G = h->ncomp[k], st = h->stride[k];
    const double *rec = D + h->off[k];
    const double modw = D[h->ext[k]];
    double r[DMAX];
    double s = 0.0;
    for (int g = 0; g < G; g++) s = fma(rec[g * st + 2], exp(-0.5 * amx::solve_lower<DMAX>(rec + g * st, d, x, r)), s);
    return log(modw * s);
  }
};
