/* test_proposal_io.c -- CPU-only: the proposal distribution survives a save/load round trip bit for bit,
 * files in the reference's own six-decimal format load too, and the loader's checks match the reference
 * reader's (logwrite.c:27-109).  No GPU is touched: initAMSampler and the file routines are host C. */
#include "automix.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int failures = 0;
#define CHECK(c, msg) do { if (!(c)) { failures++; printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); } } while (0)
static double lp(int k, double *x) { (void)k; return -x[0] * x[0]; }

int main(int argc, char **argv) {
  const char *dir = argc > 1 ? argv[1] : "/tmp";
  char path[512], path2[512];
  snprintf(path, sizeof(path), "%s/prop_mix.data", dir);
  snprintf(path2, sizeof(path2), "%s/ref_style_mix.data", dir);
  int dims[3] = {1, 3, 2};
  amSampler a, b;
  CHECK(initAMSampler(&a, 3, dims, lp, NULL) == EXIT_SUCCESS, "init a");
  CHECK(initAMSampler(&b, 3, dims, lp, NULL) == EXIT_SUCCESS, "init b");
  unsigned long seed = 77;
  sdrni(&seed);
  for (int k = 0; k < 3; k++) {
    const int d = dims[k], L = k + 1;
    a.jd.nMixComps[k] = L;
    double tot = 0;
    for (int l = 0; l < L; l++) tot += (a.jd.lambda[k][l] = 0.1 + sdrand());
    for (int l = 0; l < L; l++) {
      a.jd.lambda[k][l] /= tot;
      for (int i = 0; i < d; i++) {
        a.jd.mu[k][l][i] = 1e3 * (sdrand() - 0.5) / 3.0;
        for (int j = 0; j <= i; j++) a.jd.B[k][l][i][j] = (i == j ? 0.5 : -0.25) + sdrand() * M_PI;
      }
    }
    for (int i = 0; i < d; i++) a.jd.sig[k][i] = sdrand() * 7.0 / 9.0;
  }
  CHECK(amx_sampler_save_proposal(&a, path) == 0, "save");
  CHECK(!b.cpstats.isInitialized, "not estimated before load");
  CHECK(amx_sampler_load_proposal(&b, path) == 0, "load");
  CHECK(b.cpstats.isInitialized, "load marks the proposal as estimated (stages 1-2 are skipped)");
  for (int k = 0; k < 3; k++) {
    const int d = dims[k], L = a.jd.nMixComps[k];
    CHECK(b.jd.nMixComps[k] == L, "component count");
    double s = 0;
    for (int l = 0; l < L; l++) s += a.jd.lambda[k][l];
    for (int l = 0; l < L; l++) {
      /* the loader renormalises when the weights do not sum to exactly one, as the reference does */
      CHECK(fabs(b.jd.lambda[k][l] - a.jd.lambda[k][l] / s) <= 1e-16, "weights");
      for (int i = 0; i < d; i++) {
        CHECK(b.jd.mu[k][l][i] == a.jd.mu[k][l][i], "means are bit-identical");
        for (int j = 0; j <= i; j++) CHECK(b.jd.B[k][l][i][j] == a.jd.B[k][l][i][j], "factors are bit-identical");
      }
    }
    for (int i = 0; i < d; i++) CHECK(b.jd.sig[k][i] == a.jd.sig[k][i], "scales are bit-identical");
  }
  /* a file as the reference writes it (%lf) */
  FILE *f = fopen(path2, "w");
  fprintf(f, "3\n1\n3\n2\n");
  fprintf(f, "%lf\n1\n%lf\n%lf\n%lf\n", 4.909002, 1.0, 0.503322, 1.051363);
  fprintf(f, "1.0\n2.0\n3.0\n2\n0.25\n0\n0\n0\n1\n0\n1\n0\n0\n1\n0.75\n1\n1\n1\n2\n0.5\n2\n0\n0\n2\n");
  fprintf(f, "0.5\n0.5\n1\n1.000000\n0.1\n0.2\n1\n0\n1\n");
  fclose(f);
  amSampler c;
  initAMSampler(&c, 3, dims, lp, NULL);
  CHECK(amx_sampler_load_proposal(&c, path2) == 0, "load a reference-style file");
  CHECK(c.jd.nMixComps[0] == 1 && c.jd.nMixComps[1] == 2 && c.jd.nMixComps[2] == 1, "counts from a reference-style file");
  CHECK(fabs(c.jd.mu[0][0][0] - 0.503322) < 1e-12 && fabs(c.jd.B[1][1][2][0] - 0.0) < 1e-12, "values from a reference-style file");
  /* the reader's checks */
  int dims2[3] = {1, 3, 3};
  amSampler d2;
  initAMSampler(&d2, 3, dims2, lp, NULL);
  CHECK(amx_sampler_load_proposal(&d2, path) != 0, "dimension mismatch is rejected");
  CHECK(!d2.cpstats.isInitialized, "a rejected file leaves the sampler unestimated");
  CHECK(amx_sampler_load_proposal(&d2, "/nonexistent/file") != 0, "missing file is rejected");
  f = fopen(path2, "w");
  fprintf(f, "3\n1\n3\n2\n1.0\n2\n0.5\n0\n1\n0.4\n0\n1\n");
  fclose(f);
  CHECK(amx_sampler_load_proposal(&c, path2) != 0, "weights that do not sum to one are rejected");
  /* weights and Cholesky diagonals feed logarithms on the device: a file that would turn into NaN there is refused here */
  f = fopen(path2, "w");
  fprintf(f, "3\n1\n3\n2\n1.0\n2\n1.5\n0\n1\n-0.5\n0\n1\n");
  fclose(f);
  CHECK(amx_sampler_load_proposal(&c, path2) != 0, "a negative weight is rejected (even when the weights sum to one)");
  f = fopen(path2, "w");
  fprintf(f, "3\n1\n3\n2\n1.0\n1\n1.0\n0\n0.0\n");
  fclose(f);
  CHECK(amx_sampler_load_proposal(&c, path2) != 0, "a zero Cholesky diagonal is rejected");
  /* Either side reads the other's files: the reference's own reader and writer (user_examples/logwrite.c,
   * compiled as it lies into oracle/_ref/libref_logwrite.so; its structs have this header's layout). */
  if (argc > 2) {
    void *h = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    CHECK(h != NULL, "dlopen libref_logwrite.so");
    if (h) {
      int (*ref_read)(char *, amSampler *) = (int (*)(char *, amSampler *))dlsym(h, "read_mixture_params");
      void (*ref_write)(char *, proposalDist) = (void (*)(char *, proposalDist))dlsym(h, "write_mix_to_file");
      CHECK(ref_read && ref_write, "reference reader / writer symbols");
      char stem[512], stem2[512], file2[560];
      snprintf(stem, sizeof(stem), "%s/prop", dir); /* the reference appends _mix.data */
      snprintf(stem2, sizeof(stem2), "%s/byref", dir);
      snprintf(file2, sizeof(file2), "%s_mix.data", stem2);
      amSampler r, w;
      initAMSampler(&r, 3, dims, lp, NULL);
      initAMSampler(&w, 3, dims, lp, NULL);
      CHECK(ref_read(stem, &r) == EXIT_SUCCESS, "the reference's reader accepts our file");
      ref_write(stem2, a.jd);
      CHECK(amx_sampler_load_proposal(&w, file2) == 0, "our loader accepts the reference writer's file");
      for (int k = 0; k < 3; k++) {
        const int d = dims[k], L = a.jd.nMixComps[k];
        CHECK(r.jd.nMixComps[k] == L && w.jd.nMixComps[k] == L, "component counts across implementations");
        for (int l = 0; l < L; l++) {
          CHECK(r.jd.lambda[k][l] == b.jd.lambda[k][l], "weights: their reader == our reader, bit for bit");
          CHECK(fabs(w.jd.lambda[k][l] - a.jd.lambda[k][l]) < 2e-6, "weights through their six-decimal writer");
          for (int i = 0; i < d; i++) {
            CHECK(r.jd.mu[k][l][i] == a.jd.mu[k][l][i], "means through their reader are bit-identical");
            CHECK(fabs(w.jd.mu[k][l][i] - a.jd.mu[k][l][i]) < 5.1e-7, "means through their writer");
            for (int j = 0; j <= i; j++) {
              CHECK(r.jd.B[k][l][i][j] == a.jd.B[k][l][i][j], "factors through their reader are bit-identical");
              CHECK(fabs(w.jd.B[k][l][i][j] - a.jd.B[k][l][i][j]) < 5.1e-7, "factors through their writer");
            }
          }
        }
        for (int i = 0; i < d; i++) CHECK(r.jd.sig[k][i] == a.jd.sig[k][i], "scales through their reader");
      }
      freeAMSampler(&r);
      freeAMSampler(&w);
      printf("cross-read with the reference's logwrite.c: done\n");
    }
  }
  freeAMSampler(&a);
  freeAMSampler(&b);
  freeAMSampler(&c);
  freeAMSampler(&d2);
  printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
  return failures ? 1 : 0;
}
