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


// Differential geometry of the Gaussian-process field, one thread per query point, everything in registers (fp64):
//   velocity  (_gp_velocity, gaussian_process.py:102-127), Jacobian (Jacobian_GP_gaussian_kernel, GPVectorField.py:143-190),
//   acceleration / curvature / curl / torsion / divergence (GPVectorField.py:12-125), det J (differential_geometry.py:336).
struct GeomOut {
  double *V, *J, *acc, *acc_mat, *curv, *curv_mat, *curl, *torsion, *div, *det;
};

template <int D>
__global__ void field_geometry_kernel(spb_field_desc f, const double* __restrict__ X, int64_t n,
                                      const double* __restrict__ z, const double* __restrict__ Coff, GeomOut o) {
  extern __shared__ double shf[];
  double* zs = shf;
  double* cs = shf + (size_t)f.K * D;
  for (int t = threadIdx.x; t < f.K * D; t += blockDim.x) {
    zs[t] = z[t];
    cs[t] = Coff[t];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[D], xn[D], vel[D], J[D][D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    x[d] = X[i * D + d];
    xn[d] = (x[d] - f.mean_transformed[d]) / f.scale_transformed;
    vel[d] = 0.0;
#pragma unroll
    for (int e = 0; e < D; ++e) J[d][e] = 0.0;
  }
  for (int k = 0; k < f.K; ++k) {
    double df[D], d2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      df[d] = xn[d] - zs[k * D + d];
      d2 += df[d] * df[d];
    }
    const double w = exp(-f.beta * d2);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double wc = w * cs[k * D + a];
      vel[a] += wc;
#pragma unroll
      for (int b = 0; b < D; ++b) J[a][b] += wc * df[b];
    }
  }
  const double jscale = -2.0 * f.beta * (f.scale_fixed / f.scale_transformed);
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) J[a][b] *= jscale;
  // velocity in raw units / velocity_divisor (10000 for the GP field, 1 for a plain RBF field)
  double v[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (f.nonrigid_only) {
      v[d] = vel[d] * f.scale_fixed + (f.scale_fixed - f.scale_transformed) * xn[d];
    } else {
      double r = f.t[d];
#pragma unroll
      for (int e = 0; e < D; ++e) r += xn[e] * f.R[d * 3 + e];
      v[d] = (vel[d] + r) * f.scale_fixed + f.mean_fixed[d] - x[d];
    }
    v[d] /= f.velocity_divisor;
  }
  double a[D], vv = 0.0, va = 0.0, aa = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < D; ++e) s += J[d][e] * v[e];
    a[d] = s;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    vv += v[d] * v[d];
    va += v[d] * a[d];
    aa += a[d] * a[d];
  }
  if (o.V)
    for (int d = 0; d < D; ++d) o.V[i * D + d] = v[d];
  if (o.J)
    for (int p = 0; p < D; ++p)
      for (int q = 0; q < D; ++q) o.J[(i * D + p) * D + q] = J[p][q];
  if (o.acc) o.acc[i] = sqrt(aa);
  if (o.acc_mat)
    for (int d = 0; d < D; ++d) o.acc_mat[i * D + d] = a[d];
  if (o.curv || o.curv_mat) {
    const double nv = sqrt(vv);
    if (f.curvature_formula == 1) {
      if (o.curv) o.curv[i] = sqrt(vv * aa) / (nv * nv * nv);
    } else {
      double c2 = 0.0;
      const double nv4 = (nv * nv) * (nv * nv);
      for (int d = 0; d < D; ++d) {
        const double c = (a[d] * vv - v[d] * va) / nv4;
        if (o.curv_mat) o.curv_mat[i * D + d] = c;
        c2 += c * c;
      }
      if (o.curv) o.curv[i] = sqrt(c2);
    }
  }
  if (o.curl) {
    if (D == 2) {
      o.curl[i] = J[1][0] - J[0][1];
    } else if (D == 3) {
      o.curl[i * 3 + 0] = J[2 % D][1] - J[1][2 % D];
      o.curl[i * 3 + 1] = J[0][2 % D] - J[2 % D][0];
      o.curl[i * 3 + 2] = J[1][0] - J[0][1];
    }
  }
  if (o.torsion && D == 3) {
    // tau = outer(v, a) . (J a) / |outer(v, a)|_F^2 = v (a . J a) / (|v|^2 |a|^2)
    double Ja[D], aJa = 0.0;
    for (int d = 0; d < D; ++d) {
      double s = 0.0;
      for (int e = 0; e < D; ++e) s += J[d][e] * a[e];
      Ja[d] = s;
    }
    for (int d = 0; d < D; ++d) aJa += a[d] * Ja[d];
    const double nrm = sqrt(vv * aa);
    for (int d = 0; d < D; ++d) o.torsion[i * D + d] = v[d] * aJa / (nrm * nrm);
  }
  if (o.div) {
    double tr = 0.0;
    for (int d = 0; d < D; ++d) tr += J[d][d];
    o.div[i] = tr;
  }
  if (o.det) {
    double dt;
    if (D == 1) dt = J[0][0];
    else if (D == 2) dt = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    else
      dt = J[0][0] * (J[1][1] * J[2 % D][2 % D] - J[1][2 % D] * J[2 % D][1]) -
           J[0][1] * (J[1][0] * J[2 % D][2 % D] - J[1][2 % D] * J[2 % D][0]) +
           J[0][2 % D] * (J[1][0] * J[2 % D][1] - J[1][1] * J[2 % D][0]);
    o.det[i] = dt;
  }
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
  static bool attr_set[SPB_MAX_DEVICES] = {};  // the opt-in is per device (one process may drive several GPUs)
  const int dev_ = spb_current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(field_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev_] = true;
  }
  field_eval_kernel<<<(unsigned)((n + 127) / 128), 128, smem, ST>>>(q, n, D, z, Coff, K, beta, out);
  SPB_CHECK_LAUNCH();
  return 0;
}


extern "C" int spb_field_geometry(const spb_field_desc* f, const double* X, int64_t n, const double* z,
                                  const double* Coff, double* V, double* J, double* acc, double* acc_mat, double* curv,
                                  double* curv_mat, double* curl, double* torsion, double* div, double* det,
                                  void* stream) {
  if (n <= 0) return 0;
  if (f == nullptr || f->D < 2 || f->D > 3 || f->K < 1) return SPB_EINVAL;
  if (torsion != nullptr && f->D != 3) return SPB_EINVAL;
  const size_t smem = sizeof(double) * 2 * (size_t)f->K * f->D;
  if (smem > 96 * 1024) return SPB_EUNSUPPORTED;
  static bool attr_set[SPB_MAX_DEVICES] = {};  // the opt-in is per device (one process may drive several GPUs)
  const int dev_ = spb_current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(field_geometry_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(field_geometry_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev_] = true;
  }
  GeomOut o{V, J, acc, acc_mat, curv, curv_mat, curl, torsion, div, det};
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (f->D == 2)
    field_geometry_kernel<2><<<grid, 128, smem, ST>>>(*f, X, n, z, Coff, o);
  else
    field_geometry_kernel<3><<<grid, 128, smem, ST>>>(*f, X, n, z, Coff, o);
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
