// Coarse rigid initialisation on the device: the annealed robust Procrustes of `inlier_from_NN`
// (spateo/alignment/methods/utils.py:1220-1280) over the mutual-nearest-neighbour voxel pairs. The host version spends
// ~3 s in numpy temporaries for 3e5 pairs x 100 iterations; here every iteration is two passes over the pairs plus two
// single-thread steps, all fp64, enqueued back-to-back by one C call (no host synchronisation).
#include "common.cuh"

namespace {

struct InlierState {
  double R[9], t[3];
  double sigma2, gamma, alpha;
  double Sp;       // sum of the UNCLAMPED posterior of the last E-step (the reference normalises the means with it)
  double mom[20];  // [0] sum Pc, [1..3] sum Pc x, [4..6] sum Pc y, [7..15] sum Pc y x^T, [16] sum Pc resid  (Pc = clamped P)
  double sumP;     // accumulator of the new unclamped posterior
};

constexpr double kTwoPiC = 6.283185307179586;

// pass A: clamp the posterior of the previous E-step (utils.py:1265) and accumulate the weighted moments
// (utils.py:1248-1252) plus sum Pc * resid for the previous iteration's sigma2 (utils.py:1268)
__global__ void __launch_bounds__(256) inlier_moments_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                             double* __restrict__ P, const double* __restrict__ resid,
                                                             int N, int D, int clamp, InlierState* st) {
  double m[17];
#pragma unroll
  for (int q = 0; q < 17; ++q) m[q] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    double p = P[i];
    if (clamp) {
      p = fmax(p, 1e-6);
      P[i] = p;
    }
    double xv[3] = {0, 0, 0}, yv[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) {
      xv[d] = x[(int64_t)i * 3 + d];
      yv[d] = y[(int64_t)i * 3 + d];
    }
    m[0] += p;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      m[1 + d] += p * xv[d];
      m[4 + d] += p * yv[d];
#pragma unroll
      for (int e = 0; e < 3; ++e) m[7 + d * 3 + e] += p * yv[d] * xv[e];
    }
    m[16] += p * resid[i];
  }
  block_reduce_atomic<17>(m, st->mom);
}

// single thread: sigma2 of the previous iteration (utils.py:1268), this iteration's annealed alpha (utils.py:1269-1270),
// then the weighted Procrustes solve (utils.py:1248-1257)
__global__ void inlier_solve_kernel(InlierState* st, int D, int iter, int do_solve, double alpha_dec) {
  if (threadIdx.x != 0) return;
  double* m = st->mom;
  if (iter > 0) st->sigma2 = m[16] / ((double)D * st->Sp);
  st->alpha = pow(alpha_dec, (double)max(0, iter - 21));
  if (do_solve) {
    const double Sp = st->Sp;
    double mux[3], muy[3];
    for (int d = 0; d < 3; ++d) {
      mux[d] = m[1 + d] / Sp;
      muy[d] = m[4 + d] / Sp;
    }
    double A[9], R[9];
    for (int q = 0; q < 9; ++q) A[q] = R[q] = 0.0;
    for (int d = 0; d < D; ++d)
      for (int e = 0; e < D; ++e)
        A[d * 3 + e] = m[7 + d * 3 + e] - muy[d] * m[1 + e] - m[4 + d] * mux[e] + m[0] * muy[d] * mux[e];
    rotation_from(A, D, R);
    for (int q = 0; q < 9; ++q) st->R[q] = R[q];
    for (int d = 0; d < 3; ++d) {
      double s = 0.0;
      if (d < D) {
        s = muy[d];
        for (int e = 0; e < D; ++e) s -= mux[e] * R[d * 3 + e];
      }
      st->t[d] = s;
    }
  }
  for (int q = 0; q < 20; ++q) m[q] = 0.0;
  st->sumP = 0.0;
}

// pass B: residuals under the new transform (utils.py:1258-1260) and the new unclamped posterior (utils.py:1260-1263);
// final_pass != 0: the closing posterior with fixed sigma2 = 1e-2, gamma = 0.1 on the last residuals (utils.py:1274-1278)
__global__ void __launch_bounds__(256) inlier_posterior_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                               const double* __restrict__ dist, double* __restrict__ P,
                                                               double* __restrict__ resid, int N, int D, double area,
                                                               double dmin, int iter, int final_pass, InlierState* st) {
  // weights in force: exp(-dist) un-normalised through iteration 21, then exp(-dist*alpha)/max (utils.py:1269-1272)
  const double alpha = st->alpha;
  const bool normalised = iter >= 22;
  const double wnorm = normalised ? exp(-dmin * alpha) : 1.0;
  const double maxw = normalised ? 1.0 : exp(-dmin * alpha);
  const double sigma2 = final_pass ? 1e-2 : st->sigma2;
  const double gamma = final_pass ? 0.1 : st->gamma;
  const double outlier = maxw * (1.0 - gamma) * pow(kTwoPiC * sigma2, 0.5 * D) / (gamma * area);
  double v[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    double r;
    if (!final_pass) {
      r = 0.0;
      for (int d = 0; d < D; ++d) {
        double yh = st->t[d];
        for (int e = 0; e < D; ++e) yh += x[(int64_t)i * 3 + e] * st->R[d * 3 + e];
        const double df = y[(int64_t)i * 3 + d] - yh;
        r += df * df;
      }
      resid[i] = r;
    } else {
      r = resid[i];
    }
    const double w = exp(-dist[i] * alpha) / wnorm;
    const double t1 = exp(-r / (2.0 * sigma2)) * w;
    const double p = t1 / (t1 + outlier);
    P[i] = p;
    v[0] += p;
  }
  block_reduce_atomic<1>(v, &st->sumP);
}

__global__ void inlier_gamma_kernel(InlierState* st, int N, double* out16) {
  if (threadIdx.x != 0) return;
  st->Sp = st->sumP;
  st->gamma = fmin(fmax(st->sumP / (double)N, 0.01), 0.99);
  if (out16 != nullptr) {
    for (int q = 0; q < 9; ++q) out16[q] = st->R[q];
    for (int d = 0; d < 3; ++d) out16[9 + d] = st->t[d];
    out16[12] = st->sigma2;
    out16[13] = st->gamma;
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

// x, y: [N][3] doubles (unused dims 0); dist: [N] clamped at 0 and normalised (utils.py:1227-1229); P: [N] in = initial
// weights exp(-dist) (utils.py:1234-1236), out = closing posterior; resid: [N] scratch; state: >= 512 bytes device scratch;
// out16 (device): R[9] row-major 3x3, t[3], sigma2, gamma.
extern "C" int spb_inlier_from_nn(const double* x, const double* y, const double* dist, int64_t N, int32_t D, double area,
                                  double dmin, double sigma2_init, double sumP_init, double* P, double* resid, void* state,
                                  double* out16, void* stream) {
  if (D < 2 || D > 3 || N <= 0) return SPB_EINVAL;
  InlierState h;
  memset(&h, 0, sizeof(h));
  h.R[0] = h.R[4] = h.R[8] = 1.0;
  h.sigma2 = sigma2_init;
  h.gamma = 0.5;
  h.alpha = 1.0;
  h.Sp = sumP_init;
  InlierState* st = reinterpret_cast<InlierState*>(state);
  cudaError_t e = cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ST);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(resid, 0, sizeof(double) * N, ST);
  if (e != cudaSuccess) return (int)e;
  const int64_t want_blocks = (N + 255) / 256;
  const int blocks = (int)(want_blocks < 592 ? want_blocks : 592);
  const double alpha_dec = pow(0.1, 1.0 / 80.0);  // (alpha_end / alpha)^(1 / (max_iter - 20)), utils.py:1239
  for (int it = 0; it < 100; ++it) {
    inlier_moments_kernel<<<blocks, 256, 0, ST>>>(x, y, P, resid, (int)N, D, it > 0, st);
    SPB_CHECK_LAUNCH();
    inlier_solve_kernel<<<1, 32, 0, ST>>>(st, D, it, 1, alpha_dec);
    SPB_CHECK_LAUNCH();
    inlier_posterior_kernel<<<blocks, 256, 0, ST>>>(x, y, dist, P, resid, (int)N, D, area, dmin, it, 0, st);
    SPB_CHECK_LAUNCH();
    inlier_gamma_kernel<<<1, 32, 0, ST>>>(st, (int)N, nullptr);
    SPB_CHECK_LAUNCH();
  }
  // tail of iteration 99: clamp, sigma2, last annealing step; then the closing fixed-variance posterior
  inlier_moments_kernel<<<blocks, 256, 0, ST>>>(x, y, P, resid, (int)N, D, 1, st);
  SPB_CHECK_LAUNCH();
  inlier_solve_kernel<<<1, 32, 0, ST>>>(st, D, 100, 0, alpha_dec);
  SPB_CHECK_LAUNCH();
  inlier_posterior_kernel<<<blocks, 256, 0, ST>>>(x, y, dist, P, resid, (int)N, D, area, dmin, 100, 1, st);
  SPB_CHECK_LAUNCH();
  inlier_gamma_kernel<<<1, 32, 0, ST>>>(st, (int)N, out16);
  SPB_CHECK_LAUNCH();
  return 0;
}
