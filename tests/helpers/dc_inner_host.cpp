// Host build of korali_b200/csrc/dc_inner.cuh for tests/test_dc_inner.py (g++; the header is host/device code).
#include "../../korali_b200/csrc/dc_inner.cuh"

extern "C" {

// All K roots of 1/rho + sum w_i^2/(dl_i - lam) = 0: lam[j], delta[j*K + i] = dl_i - lam_j (accurate differences), iters[j].
void dc_secular_host(int K, const double* dl, const double* w, double rho, double* lam, double* delta, int* iters) {
  kc::SerialLanes cx;
  for (int j = 0; j < K; j++) {
    int o; double mu;
    iters[j] = kc::secular_root(cx, j, K, dl, w, rho, o, mu);
    lam[j] = dl[o] + mu;
    for (int i = 0; i < K; i++) delta[(size_t)j * K + i] = (dl[i] - dl[o]) - mu;
  }
}

}
