// Fused E-step of the morpho-align EM (replaces calc_distance("euc") + get_P_core + every reduction that consumes P:
// spateo/alignment/methods/utils.py:993-1096, morpho_class.py:1147-1200). P is never materialised in the loop.
//
//   sweep 1  streams GT (one contiguous row per fixed cell j) through a 3-stage bulk-async (TMA 1-D) shared-memory
//            ring and produces the four column sums  C1=sum_i s, C2=sum_i s m, C3=sum_i q m, C4=sum_i q m g.
//   col_finalize   turns them into the per-column constants a_j, b_j, c_j (and K_NB_j).
//   sweep 2  streams GT again and accumulates, per moving cell i (thread-owned registers), K_NA_spatial, K_NA_sigma2,
//            sum_j Psigma d, K_NA and the D components of P @ XB.
//   row_finalize   reduces the per-segment partials in fp64 and forms the global sums.
//
// Algorithmic HBM traffic: 4 bytes per cell pair per sweep (the fp32 g_ij), 8 B/pair/iteration in total.
// Thread mapping: 256 consumer threads own 4 consecutive rows each (one float4 of a GT row), one extra warp is the
// bulk-copy producer. Column constants are broadcast from shared memory.
#include "common.cuh"

namespace {

constexpr int kRowTile = SPB_ROW_TILE;   // 1024 rows per CTA
constexpr int kColStage = SPB_COL_STAGE; // 8 columns per stage
constexpr int kStages = SPB_STAGES;
constexpr int kConsumers = SPB_THREADS;  // 256
constexpr int kThreads = kConsumers + 32;

struct __align__(16) SmemLayout {
  float tile[kStages][kColStage][kRowTile];  // 3 x 32 KB
  float4 cols[kStages][kColStage][2];        // per-column constants (sweep 1 uses only [.][.][0])
  float red[2][kConsumers / 32][32];         // sweep-1 cross-warp staging
  uint64_t full[kStages];
  uint64_t empty[kStages];
};

__device__ __forceinline__ void producer_loop(SmemLayout& sm, const float* __restrict__ GT, int64_t ldx,
                                              const int32_t* __restrict__ col_index, const float* __restrict__ colsrc,
                                              int col_floats, int i0, int j_begin, int j_end, int lane) {
  const int nst = (j_end - j_begin + kColStage - 1) / kColStage;
  for (int st = 0; st < nst; ++st) {
    const int s = st % kStages;
    if (st >= kStages) mbar_wait(&sm.empty[s], ((st / kStages) - 1) & 1);
    const int jb = j_begin + st * kColStage;
    const int ncol = min(kColStage, j_end - jb);
    if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)(ncol * kRowTile * 4 + ncol * col_floats * 4));
    __syncwarp();
    if (lane < ncol) {
      const int j = jb + lane;
      const int64_t row = col_index ? (int64_t)col_index[j] : (int64_t)j;
      bulk_g2s(&sm.tile[s][lane][0], GT + row * ldx + i0, kRowTile * 4, &sm.full[s]);
    }
    // per-column constants: sweep 2 reads [ncol][8] floats in one copy; sweep 1 reads one float4 per column into the
    // first slot of cols[][2]
    if (col_floats == 8) {
      if (lane == 0) bulk_g2s(&sm.cols[s][0][0], colsrc + (int64_t)jb * 8, ncol * 32, &sm.full[s]);
    } else {
      if (lane < ncol) bulk_g2s(&sm.cols[s][lane][0], colsrc + (int64_t)(jb + lane) * 4, 16, &sm.full[s]);
    }
  }
}

__device__ __forceinline__ float sqdist(float x0, float x1, float x2, const float4& y) {
  const float d0 = x0 - y.x, d1 = x1 - y.y, d2 = x2 - y.z;
  return fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
}

// ---------------------------------------------------------------------------------------------------------------------
// sweep 1: column sums
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
estep_sweep1_kernel(const float* __restrict__ GT, int64_t ldx, const int32_t* __restrict__ col_index,
                    const float* __restrict__ colgeom, const float* __restrict__ XA, const float* __restrict__ lm,
                    const float* __restrict__ mm, const spb_scalars* __restrict__ sc, float* __restrict__ colpart,
                    int NBb, int nbb_pad, int cols_per_seg) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb = blockIdx.x, seg = blockIdx.y;
  const int i0 = rb * kRowTile;
  const int j_begin = seg * cols_per_seg;
  const int j_end = min(NBb, j_begin + cols_per_seg);
  if (j_begin >= j_end) return;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kConsumers / 32) {
    producer_loop(sm, GT, ldx, col_index, colgeom, 4, i0, j_begin, j_end, lane);
    return;
  }
  // ---- consumers ----
  const float c_q = sc->c_q, c_s = sc->c_s;
  const int r = i0 + tid * 4;
  const float4 X0 = *reinterpret_cast<const float4*>(XA + r);
  const float4 X1 = *reinterpret_cast<const float4*>(XA + ldx + r);
  const float4 X2 = *reinterpret_cast<const float4*>(XA + 2 * ldx + r);
  const float4 LM = *reinterpret_cast<const float4*>(lm + r);
  const float4 MM = *reinterpret_cast<const float4*>(mm + r);
  const int nst = (j_end - j_begin + kColStage - 1) / kColStage;
  for (int st = 0; st < nst; ++st) {
    const int s = st % kStages;
    mbar_wait(&sm.full[s], (st / kStages) & 1);
    const int jb = j_begin + st * kColStage;
    const int ncol = min(kColStage, j_end - jb);
    float acc[32];  // index v * 8 + jj
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.f;
#pragma unroll
    for (int jj = 0; jj < kColStage; ++jj) {
      if (jj < ncol) {
        const float4 y = sm.cols[s][jj][0];
        const float4 g = *reinterpret_cast<const float4*>(&sm.tile[s][jj][tid * 4]);
        float d, sv, qm, c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f;
        d = sqdist(X0.x, X1.x, X2.x, y); sv = ex2f(c_s * d); qm = ex2f(fmaf(c_q, d, LM.x));
        c1 += sv; c2 = fmaf(sv, MM.x, c2); c3 += qm; c4 = fmaf(qm, g.x, c4);
        d = sqdist(X0.y, X1.y, X2.y, y); sv = ex2f(c_s * d); qm = ex2f(fmaf(c_q, d, LM.y));
        c1 += sv; c2 = fmaf(sv, MM.y, c2); c3 += qm; c4 = fmaf(qm, g.y, c4);
        d = sqdist(X0.z, X1.z, X2.z, y); sv = ex2f(c_s * d); qm = ex2f(fmaf(c_q, d, LM.z));
        c1 += sv; c2 = fmaf(sv, MM.z, c2); c3 += qm; c4 = fmaf(qm, g.z, c4);
        d = sqdist(X0.w, X1.w, X2.w, y); sv = ex2f(c_s * d); qm = ex2f(fmaf(c_q, d, LM.w));
        c1 += sv; c2 = fmaf(sv, MM.w, c2); c3 += qm; c4 = fmaf(qm, g.w, c4);
        acc[0 * 8 + jj] = c1; acc[1 * 8 + jj] = c2; acc[2 * 8 + jj] = c3; acc[3 * 8 + jj] = c4;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);  // stage buffer is free again
    // butterfly transpose-reduce: 31 shuffles reduce all 32 values across the warp; lane q ends with value q
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int q = 0; q < n; ++q) {
        const float mine = up ? acc[q + n] : acc[q];
        const float theirs = up ? acc[q] : acc[q + n];
        acc[q] = mine + __shfl_xor_sync(0xffffffffu, theirs, off);
      }
    }
    const int buf = st & 1;
    sm.red[buf][warp][lane] = acc[0];
    named_bar_sync(1, kConsumers);
    if (warp == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kConsumers / 32; ++w) t += sm.red[buf][w][lane];
      const int v = lane >> 3, jj = lane & 7;
      if (jj < ncol) colpart[((int64_t)rb * 4 + v) * nbb_pad + jb + jj] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// column constants
// ---------------------------------------------------------------------------------------------------------------------
__global__ void col_finalize_kernel(const float* __restrict__ colpart, int nrb, int nbb_pad, int NBb,
                                    const float* __restrict__ colgeom, const spb_scalars* __restrict__ sc,
                                    float* __restrict__ colconst, float* __restrict__ K_NB) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= NBb) return;
  double C[4] = {0, 0, 0, 0};
  for (int rb = 0; rb < nrb; ++rb) {
#pragma unroll
    for (int v = 0; v < 4; ++v) C[v] += (double)colpart[((int64_t)rb * 4 + v) * nbb_pad + j];
  }
  const double omega = sc->omega;
  const double inl = 1.0 - omega / (omega + C[0]);          // utils.py:1055
  const double a = 1.0 / (omega + C[1]);                     // utils.py:1059
  const double b = inl / (C[2] + 1e-8);                      // utils.py:1073
  const double c = inl / (C[3] + 1e-8);                      // utils.py:1083
  const float4 y = *reinterpret_cast<const float4*>(colgeom + (int64_t)j * 4);
  float4* out = reinterpret_cast<float4*>(colconst + (int64_t)j * 8);
  out[0] = make_float4(y.x, y.y, y.z, (float)a);
  out[1] = make_float4((float)b, (float)c, 0.f, 0.f);
  K_NB[j] = (float)(c * C[3]);                                // column sum of P (morpho_class.py:1176)
}

// ---------------------------------------------------------------------------------------------------------------------
// sweep 2: row statistics
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
estep_sweep2_kernel(const float* __restrict__ GT, int64_t ldx, const int32_t* __restrict__ col_index,
                    const float* __restrict__ colconst, const float* __restrict__ XA, const float* __restrict__ lm,
                    const spb_scalars* __restrict__ sc, float* __restrict__ rowpart, int NBb, int cols_per_seg) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb = blockIdx.x, seg = blockIdx.y;
  const int i0 = rb * kRowTile;
  const int j_begin = seg * cols_per_seg;
  const int j_end = min(NBb, j_begin + cols_per_seg);
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], kConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == kConsumers / 32) {
    if (j_begin < j_end) producer_loop(sm, GT, ldx, col_index, colconst, 8, i0, j_begin, j_end, lane);
    return;
  }
  const float c_q = sc->c_q, c_s = sc->c_s;
  const int r = i0 + tid * 4;
  const float4 X0 = *reinterpret_cast<const float4*>(XA + r);
  const float4 X1 = *reinterpret_cast<const float4*>(XA + ldx + r);
  const float4 X2 = *reinterpret_cast<const float4*>(XA + 2 * ldx + r);
  const float4 LM = *reinterpret_cast<const float4*>(lm + r);
  float4 a_sp = make_float4(0, 0, 0, 0), a_s2 = a_sp, a_sd = a_sp, a_k = a_sp, px = a_sp, py = a_sp, pz = a_sp;
  const int nst = j_begin < j_end ? (j_end - j_begin + kColStage - 1) / kColStage : 0;
  for (int st = 0; st < nst; ++st) {
    const int s = st % kStages;
    mbar_wait(&sm.full[s], (st / kStages) & 1);
    const int ncol = min(kColStage, j_end - (j_begin + st * kColStage));
#pragma unroll
    for (int jj = 0; jj < kColStage; ++jj) {
      if (jj < ncol) {
        const float4 ya = sm.cols[s][jj][0];  // y0 y1 y2 a_j
        const float4 bc = sm.cols[s][jj][1];  // b_j c_j
        const float4 g = *reinterpret_cast<const float4*>(&sm.tile[s][jj][tid * 4]);
        float d, sv, qm, t, p;
#define SPB_ROW(C)                                                     \
  d = sqdist(X0.C, X1.C, X2.C, ya);                                    \
  sv = ex2f(c_s * d);                                                  \
  qm = ex2f(fmaf(c_q, d, LM.C));                                       \
  a_sp.C = fmaf(sv, ya.w, a_sp.C);                                     \
  t = qm * bc.x;                                                       \
  a_s2.C += t;                                                         \
  a_sd.C = fmaf(t, d, a_sd.C);                                         \
  p = (qm * g.C) * bc.y;                                               \
  a_k.C += p;                                                          \
  px.C = fmaf(p, ya.x, px.C);                                          \
  py.C = fmaf(p, ya.y, py.C);                                          \
  pz.C = fmaf(p, ya.z, pz.C);
        SPB_ROW(x) SPB_ROW(y) SPB_ROW(z) SPB_ROW(w)
#undef SPB_ROW
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
  }
  float* out = rowpart + ((int64_t)seg * 8) * ldx + r;
  *reinterpret_cast<float4*>(out + 0 * ldx) = a_sp;
  *reinterpret_cast<float4*>(out + 1 * ldx) = a_s2;
  *reinterpret_cast<float4*>(out + 2 * ldx) = a_sd;
  *reinterpret_cast<float4*>(out + 3 * ldx) = a_k;
  *reinterpret_cast<float4*>(out + 4 * ldx) = px;
  *reinterpret_cast<float4*>(out + 5 * ldx) = py;
  *reinterpret_cast<float4*>(out + 6 * ldx) = pz;
}

// per row: fold the segment partials (fp64), write the fp32 statistics, accumulate the global sums
__global__ void row_finalize_kernel(const float* __restrict__ rowpart, int nseg, int ldx, int NA,
                                    const float* __restrict__ mm, float* __restrict__ K_NA_spatial,
                                    float* __restrict__ K_NA_sigma2, float* __restrict__ K_NA, float* __restrict__ PXB,
                                    spb_scalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v[4] = {0, 0, 0, 0};  // Sp_spatial, Sp_sigma2, Sp, S2
  if (i < NA) {
    double a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < nseg; ++s) {
#pragma unroll
      for (int q = 0; q < 7; ++q) a[q] += (double)rowpart[((int64_t)s * 8 + q) * ldx + i];
    }
    const double ksp = a[0] * (double)mm[i];
    K_NA_spatial[i] = (float)ksp;
    K_NA_sigma2[i] = (float)a[1];
    K_NA[i] = (float)a[3];
    PXB[i] = (float)a[4];
    PXB[ldx + i] = (float)a[5];
    PXB[2 * ldx + i] = (float)a[6];
    v[0] = ksp; v[1] = a[1]; v[2] = a[3]; v[3] = a[2];
  }
  block_reduce_atomic<4>(v, sc->sums);
}

// gather this iteration's fixed-slice coordinates (SVI batch or all columns) (morpho_class.py:1149)
__global__ void gather_cols_kernel(const float* __restrict__ xb4, const int32_t* __restrict__ idx, int NBb,
                                   float* __restrict__ colgeom) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= NBb) return;
  const int64_t src = idx ? idx[j] : j;
  reinterpret_cast<float4*>(colgeom)[j] = reinterpret_cast<const float4*>(xb4)[src];
}

// dense P for the caller (utils.py:1083): P[i][j] = qm_ij g_ij c_j, transposed through shared memory
__global__ void materialize_P_kernel(const float* __restrict__ GT, int64_t ldx, const int32_t* __restrict__ col_index,
                                     const float* __restrict__ colconst, const float* __restrict__ XA,
                                     const float* __restrict__ lm, const spb_scalars* __restrict__ sc, int NA, int NBb,
                                     float* __restrict__ P, int64_t ldp) {
  __shared__ float tile[32][33];
  const float c_q = sc->c_q;
  const int jb = blockIdx.y * 32, ib = blockIdx.x * 32;
  for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
    const int j = jb + jj, i = ib + threadIdx.x;
    float p = 0.f;
    if (j < NBb && i < NA) {
      const int64_t row = col_index ? col_index[j] : j;
      const float4 ya = *reinterpret_cast<const float4*>(colconst + (int64_t)j * 8);
      const float cj = colconst[(int64_t)j * 8 + 5];
      const float d = sqdist(XA[i], XA[ldx + i], XA[2 * ldx + i], ya);
      p = ex2f(fmaf(c_q, d, lm[i])) * GT[row * ldx + i] * cj;
    }
    tile[jj][threadIdx.x] = p;
  }
  __syncthreads();
  for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
    const int i = ib + ii, j = jb + threadIdx.x;
    if (i < NA && j < NBb) P[(int64_t)i * ldp + j] = tile[threadIdx.x][ii];
  }
}

int cols_per_segment(int NBb, int nseg) {
  int c = (NBb + nseg - 1) / nseg;
  return ((c + kColStage - 1) / kColStage) * kColStage;
}

}  // namespace

static inline const int32_t* batch_ptr(const spb_em_params* p, int iter) {
  return (p->svi && p->batch_idx) ? p->batch_idx + (int64_t)iter * p->NBb : nullptr;
}

extern "C" int spb_gather_cols(const spb_em_params* p, int32_t iter, void* stream) {
  gather_cols_kernel<<<(p->NBb + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p->xb4, batch_ptr(p, iter), p->NBb, p->colgeom);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_estep_sweep1(const spb_em_params* p, int32_t iter, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(estep_sweep1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayout));
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int nrb = p->ldx / kRowTile;
  dim3 grid(nrb, p->seg1);
  estep_sweep1_kernel<<<grid, kThreads, sizeof(SmemLayout), (cudaStream_t)stream>>>(
      p->GT, p->ldx, batch_ptr(p, iter), p->colgeom, p->XAHat, p->lm, p->mm, p->sc, p->colpart, p->NBb, p->nbb_pad,
      cols_per_segment(p->NBb, p->seg1));
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_col_finalize(const spb_em_params* p, void* stream) {
  col_finalize_kernel<<<(p->NBb + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      p->colpart, p->ldx / kRowTile, p->nbb_pad, p->NBb, p->colgeom, p->sc, p->colconst, p->K_NB);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_estep_sweep2(const spb_em_params* p, int32_t iter, void* stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(estep_sweep2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLayout));
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(p->ldx / kRowTile, p->seg2);
  estep_sweep2_kernel<<<grid, kThreads, sizeof(SmemLayout), (cudaStream_t)stream>>>(
      p->GT, p->ldx, batch_ptr(p, iter), p->colconst, p->XAHat, p->lm, p->sc, p->rowpart, p->NBb,
      cols_per_segment(p->NBb, p->seg2));
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_row_finalize(const spb_em_params* p, void* stream) {
  row_finalize_kernel<<<(p->NA + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      p->rowpart, p->seg2, p->ldx, p->NA, p->mm, p->K_NA_spatial, p->K_NA_sigma2, p->K_NA, p->PXB, p->sc);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_materialize_P(const spb_em_params* p, int32_t iter, float* P, int64_t ldp, void* stream) {
  dim3 grid((p->NA + 31) / 32, (p->NBb + 31) / 32), block(32, 8);
  materialize_P_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(p->GT, p->ldx, batch_ptr(p, iter), p->colconst, p->XAHat,
                                                                p->lm, p->sc, p->NA, p->NBb, P, ldp);
  SPB_CHECK_LAUNCH();
  return 0;
}
