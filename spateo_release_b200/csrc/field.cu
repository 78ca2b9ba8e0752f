// Gaussian-kernel vector field utilities shared by the alignment and by st.tdr:
//   U^T = exp(-beta |x - z|^2)            (con_K, spateo/alignment/methods/utils.py:1132-1158)
//   field evaluation on query points      (BA_transform, spateo/alignment/transform.py:93-103;
//                                          _gp_velocity, spateo/tdr/morphometrics/morphofield/gaussian_process.py:109-117)
#include "common.cuh"

namespace {

__global__ void rbf_kernel_T_kernel(const float* __restrict__ x, int64_t n, int64_t ldx, const float* __restrict__ z,
                                    int K, float beta, float* __restrict__ UT) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= ldx) return;
  float v = 0.f;
  if (i < n) {
    const float d0 = x[i] - z[k * 3 + 0], d1 = x[ldx + i] - z[k * 3 + 1], d2 = x[2 * ldx + i] - z[k * 3 + 2];
    v = expf(-beta * (d0 * d0 + d1 * d1 + d2 * d2));
  }
  UT[(int64_t)k * ldx + i] = v;
}

// out[i][:] = sum_k exp(-beta |q_i - z_k|^2) Coff[k][:]   (fp64; K*D*2 doubles staged in shared memory)
__global__ void field_eval_kernel(const double* __restrict__ q, int64_t n, int D, const double* __restrict__ z,
                                  const double* __restrict__ Coff, int K, double beta, double* __restrict__ out) {
  extern __shared__ double shf[];
  double* zs = shf;
  double* cs = shf + (size_t)K * D;
  for (int t = threadIdx.x; t < K * D; t += blockDim.x) {
    zs[t] = z[t];
    cs[t] = Coff[t];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double qi[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
  for (int d = 0; d < D; ++d) qi[d] = q[i * D + d];
  for (int k = 0; k < K; ++k) {
    double d2 = 0;
    for (int d = 0; d < D; ++d) {
      const double df = qi[d] - zs[k * D + d];
      d2 += df * df;
    }
    const double w = exp(-beta * d2);
    for (int d = 0; d < D; ++d) acc[d] += w * cs[k * D + d];
  }
  for (int d = 0; d < D; ++d) out[i * D + d] = acc[d];
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int spb_rbf_kernel_T(const float* x, int64_t n, int64_t ldx, const float* z, int32_t K, float beta, float* UT,
                                void* stream) {
  dim3 grid((unsigned)((ldx + 255) / 256), (unsigned)K);
  rbf_kernel_T_kernel<<<grid, 256, 0, ST>>>(x, n, ldx, z, K, beta, UT);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_field_eval(const double* q, int64_t n, int32_t D, const double* z, const double* Coff, int32_t K,
                              double beta, double* out, void* stream) {
  if (n <= 0) return 0;
  if (D < 1 || D > 3) return SPB_EINVAL;
  const size_t smem = sizeof(double) * 2 * (size_t)K * D;
  if (smem > 96 * 1024) return SPB_EUNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(field_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  field_eval_kernel<<<(unsigned)((n + 127) / 128), 128, smem, ST>>>(q, n, D, z, Coff, K, beta, out);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_field_eval_host(const double* q_host, int64_t n, int32_t D, const double* z_host,
                                   const double* Coff_host, int32_t K, double beta, double* out_host) {
  double *q = nullptr, *z = nullptr, *c = nullptr, *o = nullptr;
  cudaError_t e;
  int rc = 0;
  if ((e = cudaMalloc(&q, sizeof(double) * n * D)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&z, sizeof(double) * K * D)) != cudaSuccess) { cudaFree(q); return (int)e; }
  if ((e = cudaMalloc(&c, sizeof(double) * K * D)) != cudaSuccess) { cudaFree(q); cudaFree(z); return (int)e; }
  if ((e = cudaMalloc(&o, sizeof(double) * n * D)) != cudaSuccess) { cudaFree(q); cudaFree(z); cudaFree(c); return (int)e; }
  cudaMemcpy(q, q_host, sizeof(double) * n * D, cudaMemcpyHostToDevice);
  cudaMemcpy(z, z_host, sizeof(double) * K * D, cudaMemcpyHostToDevice);
  cudaMemcpy(c, Coff_host, sizeof(double) * K * D, cudaMemcpyHostToDevice);
  rc = spb_field_eval(q, n, D, z, c, K, beta, o, nullptr);
  if (rc == 0) {
    e = cudaMemcpy(out_host, o, sizeof(double) * n * D, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = (int)e;
  }
  cudaFree(q); cudaFree(z); cudaFree(c); cudaFree(o);
  return rc;
}
