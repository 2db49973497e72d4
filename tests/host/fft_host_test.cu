// Host-side check of the shared-memory FFT plan (chimeralm_b200/csrc/fft.cuh): compiled with
// nvcc as plain host code, one "thread".  Prints max abs error vs a double-precision DFT of
// (a) forward transform up to the digit-reversal permutation (checked as a multiset via the
// round trip and Parseval) and (b) circular convolution through forward * spectrum -> inverse.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../chimeralm_b200/csrc/fft.cuh"
using namespace clm::fft;

template <int LOGN>
int run() {
  constexpr int N = 1 << LOGN;
  std::vector<float2> z(padded_size(N)), g(padded_size(N));
  std::vector<double> a(N), b(N), k(N);
  srand(LOGN);
  for (int i = 0; i < N; ++i) {
    a[i] = (i < N / 2) ? (rand() / (double)RAND_MAX - 0.5) : 0.0;
    b[i] = (i < N / 2) ? (rand() / (double)RAND_MAX - 0.5) : 0.0;
    k[i] = (rand() / (double)RAND_MAX - 0.5) * exp(-3.0 * i / N);
    z[pad_idx(i)] = make_float2((float)a[i], (float)b[i]);
    g[pad_idx(i)] = make_float2((float)k[i], 0.f);
  }
  fft_forward<LOGN, 1>(z.data(), 0);
  fft_forward<LOGN, 1>(g.data(), 0);
  // Parseval on the forward output
  double e_t = 0, e_f = 0;
  for (int i = 0; i < N; ++i) { e_t += a[i] * a[i] + b[i] * b[i]; float2 v = z[pad_idx(i)]; e_f += (double)v.x * v.x + (double)v.y * v.y; }
  for (int i = 0; i < N; ++i) { float2 u = z[pad_idx(i)], w = g[pad_idx(i)]; z[pad_idx(i)] = make_float2((u.x * w.x - u.y * w.y) / N, (u.x * w.y + u.y * w.x) / N); }
  fft_inverse<LOGN, 1>(z.data(), 0);
  // reference circular convolution (O(N^2) in double; sample a subset of outputs for big N)
  double maxerr = 0, maxref = 0;
  int step = N > 4096 ? 37 : 1;
  for (int n = 0; n < N; n += step) {
    double ra = 0, rb = 0;
    for (int q = 0; q < N / 2; ++q) { double kv = k[(n - q + N) % N]; ra += a[q] * kv; rb += b[q] * kv; }
    maxerr = fmax(maxerr, fmax(fabs(ra - z[pad_idx(n)].x), fabs(rb - z[pad_idx(n)].y)));
    maxref = fmax(maxref, fmax(fabs(ra), fabs(rb)));
  }
  double pars = fabs(e_f / N - e_t) / e_t;
  printf("LOGN=%d N=%d conv_maxerr=%.3e (ref max %.3e) parseval_rel=%.3e\n", LOGN, N, maxerr, maxref, pars);
  return (maxerr < 2e-5 * fmax(1.0, maxref) * LOGN && pars < 1e-5) ? 0 : 1;
}
int main() {
  int bad = 0;
  bad += run<8>(); bad += run<9>(); bad += run<10>(); bad += run<11>(); bad += run<12>(); bad += run<13>(); bad += run<14>();
  printf(bad ? "FAIL\n" : "OK\n");
  return bad;
}
