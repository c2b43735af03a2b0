// Host build of korali_b200/csrc/jacobi_inner.cuh for tests/test_jacobi_inner.py (g++ -O2 -shared -fPIC -ffp-contract=off).
// The MUFU seeds are emulated by float-rounded 1/x and 1/sqrt(x) (the same ~2^-22 accuracy).
#include "../../korali_b200/csrc/jacobi_inner.cuh"

extern "C" {
// gamma: 8x8 symmetric (row-major). mode 0: cross rounds (rows 0-3 x rows 4-7); mode 1: all 28 pairs.
// Outputs: R (8x8 row-major, rows_new = R rows_old), gamma_out (8x8 symmetric), rotations, big.
void jacobi_inner_host(const double* gamma, double tol, int mode, double* R, double* gamma_out, int* rotations, int* big) {
  for (int j = 0; j < 8; j++) {   // "lane" j accumulates column j of R
    kc::Inner8 m;
    for (int a = 0; a < 8; a++)
      for (int b = 0; b < 8; b++) m.g[a][b] = gamma[a * 8 + b];
    for (int a = 0; a < 8; a++) m.rc[a] = (a == j) ? 1.0 : 0.0;
    m.rotations = 0; m.big = 0;
    if (mode == 0) kc::inner_cross(m, tol * tol); else kc::inner_full(m, tol * tol);
    for (int a = 0; a < 8; a++) R[a * 8 + j] = m.rc[a];
    if (j == 0) {
      for (int a = 0; a < 8; a++)
        for (int b = 0; b < 8; b++) gamma_out[a * 8 + b] = m.g[a < b ? a : b][a < b ? b : a];
      *rotations = m.rotations; *big = m.big;
    }
  }
}
void jacobi_cs_host(double alpha, double beta, double gamma, double* c, double* s, int* safe) {
  *safe = kc::jacobi_cs_fast(alpha, beta, gamma, *c, *s) ? 1 : 0;
  if (!*safe) kc::jacobi_cs_scaled(alpha, beta, gamma, *c, *s);
}
// Tournament schedules of the persistent eigensolver: order 0 = round-robin, 1 = ring (np = power of two).
void jacobi_pair_host(int order, int np, int step, int k, int* p, int* q) {
  if (order == 1) kc::ring_pair(np, step, k, *p, *q); else kc::rr_pair(np, step, k, *p, *q);
}
}
