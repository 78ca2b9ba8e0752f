// Voxelisation of the coarse rigid initialisation (voxel_data, spateo/alignment/methods/utils.py:1283-1336).
// A cell belongs to EVERY grid point closer than voxel_size / 2 (overlapping balls, not a partition); a voxel's expression
// is the mean over its members. The reference loops over the (up to 31^3) grid points in Python; here one thread per cell
// enumerates the few grid points that can contain it and applies the reference's own membership test
//   sqrt(sum((x - g)^2)) < radius   evaluated in the coordinates' dtype with individually rounded operations,
// so memberships are identical. Pass 1 counts members, pass 2 adds exp[i][:] / count[v] into the voxel means (fp64).
#include "common.cuh"

namespace {

struct VoxGeom {
  const void* ax[3];
  int n[3];
  int D;
  double radius;
  double lo[3], step[3];  // conservative candidate search only
};

template <typename T>
__device__ __forceinline__ T vox_dist(const T* x, const T* g, int D);
template <>
__device__ __forceinline__ float vox_dist<float>(const float* x, const float* g, int D) {
  float s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float df = __fsub_rn(x[d], g[d]);
    const float sq = __fmul_rn(df, df);
    s = d == 0 ? sq : __fadd_rn(s, sq);
  }
  return __fsqrt_rn(s);
}
template <>
__device__ __forceinline__ double vox_dist<double>(const double* x, const double* g, int D) {
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    const double df = __dsub_rn(x[d], g[d]);
    const double sq = __dmul_rn(df, df);
    s = d == 0 ? sq : __dadd_rn(s, sq);
  }
  return __dsqrt_rn(s);
}

// calls f(v) for every grid point v (flat index in np.meshgrid('xy') + reshape order) whose ball contains cell i
template <typename T, typename F>
__device__ __forceinline__ void for_each_voxel(const VoxGeom& gm, const T* __restrict__ coords, int64_t i, F&& f) {
  T x[3] = {0, 0, 0};
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int d = 0; d < gm.D; ++d) {
    x[d] = coords[i * gm.D + d];
    const double a = ((double)x[d] - gm.radius - gm.lo[d]) / gm.step[d];
    const double b = ((double)x[d] + gm.radius - gm.lo[d]) / gm.step[d];
    lo[d] = max(0, (int)floor(a) - 1);
    hi[d] = min(gm.n[d] - 1, (int)ceil(b) + 1);
  }
  const T* a0 = (const T*)gm.ax[0];
  const T* a1 = (const T*)gm.ax[1];
  const T* a2 = (const T*)gm.ax[2];
  for (int i1 = lo[1]; i1 <= hi[1]; ++i1)
    for (int i0 = lo[0]; i0 <= hi[0]; ++i0)
      for (int i2 = lo[2]; i2 <= hi[2]; ++i2) {
        T g[3];
        g[0] = a0[i0];
        g[1] = a1[i1];
        g[2] = gm.D == 3 ? a2[i2] : (T)0;
        if ((double)vox_dist<T>(x, g, gm.D) < gm.radius) {
          const int64_t v = gm.D == 3 ? ((int64_t)i1 * gm.n[0] + i0) * gm.n[2] + i2 : (int64_t)i1 * gm.n[0] + i0;
          f(v);
        }
      }
}

template <typename T>
__global__ void voxel_count_kernel(VoxGeom gm, const T* __restrict__ coords, int64_t N, int32_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for_each_voxel<T>(gm, coords, i, [&](int64_t v) { atomicAdd(counts + v, 1); });
}

// one CTA per cell: every thread enumerates the member voxels (identical, divergence-free work) and adds its slice of
// the cell's expression row into each of them — no cap on the number of memberships
template <typename T>
__global__ void __launch_bounds__(256)
voxel_accumulate_kernel(VoxGeom gm, const T* __restrict__ coords, int64_t N, const int32_t* __restrict__ counts,
                        const int32_t* __restrict__ new_id, const float* __restrict__ exp, int64_t ldg, int G,
                        double* __restrict__ means, int64_t ldm) {
  const int64_t i = blockIdx.x;
  const float* e = exp + i * ldg;
  for_each_voxel<T>(gm, coords, i, [&](int64_t v) {
    const double w = 1.0 / (double)counts[v];
    double* dst = means + (int64_t)new_id[v] * ldm;
    for (int g = threadIdx.x; g < G; g += blockDim.x) atomicAdd(dst + g, (double)e[g] * w);
  });
}

int fill_geom(VoxGeom& gm, int D, const void* ax0, int n0, const void* ax1, int n1, const void* ax2, int n2, double radius,
              const double* lo, const double* step) {
  if (D < 2 || D > 3 || n0 < 1 || n1 < 1 || (D == 3 && n2 < 1)) return SPB_EINVAL;
  gm.ax[0] = ax0; gm.ax[1] = ax1; gm.ax[2] = D == 3 ? ax2 : ax0;
  gm.n[0] = n0; gm.n[1] = n1; gm.n[2] = D == 3 ? n2 : 1;
  gm.D = D;
  gm.radius = radius;
  for (int d = 0; d < 3; ++d) {
    gm.lo[d] = d < D ? lo[d] : 0.0;
    gm.step[d] = d < D ? step[d] : 1.0;
    if (!(gm.step[d] > 0.0)) return SPB_EINVAL;
  }
  return 0;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int spb_voxel_count(const void* coords, int32_t is_f64, int64_t N, int32_t D, const void* ax0, int32_t n0,
                               const void* ax1, int32_t n1, const void* ax2, int32_t n2, double radius, const double* lo3,
                               const double* step3, int32_t* counts, void* stream) {
  if (N <= 0) return 0;
  VoxGeom gm;
  int rc = fill_geom(gm, D, ax0, n0, ax1, n1, ax2, n2, radius, lo3, step3);
  if (rc) return rc;
  const unsigned grid = (unsigned)((N + 127) / 128);
  if (is_f64) voxel_count_kernel<double><<<grid, 128, 0, ST>>>(gm, (const double*)coords, N, counts);
  else voxel_count_kernel<float><<<grid, 128, 0, ST>>>(gm, (const float*)coords, N, counts);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_voxel_accumulate(const void* coords, int32_t is_f64, int64_t N, int32_t D, const void* ax0, int32_t n0,
                                    const void* ax1, int32_t n1, const void* ax2, int32_t n2, double radius,
                                    const double* lo3, const double* step3, const int32_t* counts, const int32_t* new_id,
                                    const float* exp, int64_t ldg, int32_t G, double* means, int64_t ldm, void* stream) {
  if (N <= 0) return 0;
  VoxGeom gm;
  int rc = fill_geom(gm, D, ax0, n0, ax1, n1, ax2, n2, radius, lo3, step3);
  if (rc) return rc;
  if (is_f64)
    voxel_accumulate_kernel<double><<<(unsigned)N, 256, 0, ST>>>(gm, (const double*)coords, N, counts, new_id, exp, ldg, G,
                                                                 means, ldm);
  else
    voxel_accumulate_kernel<float><<<(unsigned)N, 256, 0, ST>>>(gm, (const float*)coords, N, counts, new_id, exp, ldg, G, means,
                                                                ldm);
  SPB_CHECK_LAUNCH();
  return 0;
}
