/*
 * ref_usertargets.c -- TEST INFRASTRUCTURE.  Thin `double f(int, double*)`
 * adapters around the reference's own example log-posteriors, which are
 * compiled from /root/reference/src/user_examples/user{toy1,toy2,cpt}.c where
 * they lie (oracle/Makefile).  Used only to validate that our workload
 * definitions (oracle/host_targets.c) are the same functions.
 */
void toy1_logpost(int k, int nkk, double *theta, double *lp, double *llh);
void toy2_logpost(int k, int nkk, double *theta, double *lp, double *llh);
void cpt_logpost(int k, int nkk, double *theta, double *lp, double *llh);
void cpt_get_rwm_init(int k, int mdim, double *rwm);

double ref_toy1(int k, double *x) {
  double lp = 0, llh = 0;
  toy1_logpost(k, 0, x, &lp, &llh);
  return lp;
}
double ref_toy2(int k, double *x) {
  double lp = 0, llh = 0;
  toy2_logpost(k, 0, x, &lp, &llh);
  return lp;
}
double ref_cpt(int k, double *x) {
  double lp = 0, llh = 0;
  cpt_logpost(k, 0, x, &lp, &llh);
  return lp;
}
void ref_cpt_init(int k, int mdim, double *rwm) { cpt_get_rwm_init(k, mdim, rwm); }
