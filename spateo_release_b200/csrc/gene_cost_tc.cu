// Expression cost matrix on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate through a 3xTF32
// split:  x = hi + lo  (hi = tf32(x), lo = x - hi),  A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi  accumulated in fp32 in TMEM.
// Same contract as gene_cost_kernel (gene_cost.cu): GT[j][i] (op)= prob(metric(A_i, B_j)).
//
// Per CTA (persistent, one per SM): tile = 128 fixed cells (UMMA M, TMEM lanes) x 256 moving cells (UMMA N, TMEM columns).
//   warp 0   TMA producer: 2-stage ring, per k-block (32 features = one 128-byte swizzle row) four 2-D tensor-map loads
//            (Bfix hi/lo 128x32, Amov hi/lo 256x32) into the canonical K-major SWIZZLE_128B layout
//   warp 1   MMA issuer: 4 k-steps x 3 products of tcgen05.mma.kind::tf32 (M128 N256 K8) per k-block, tcgen05.commit
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns), cost -> probability, store to GT; double-buffered TMEM
//            accumulators (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int TM = 128;   // fixed cells per tile (UMMA M)
constexpr int TN = 256;   // moving cells per tile (UMMA N)
constexpr int TK = 32;    // features per k-block (128 bytes)
constexpr int kTcStages = 2;
constexpr int kTcThreads = 192;  // 6 warps

struct __align__(1024) TcSmem {
  float bfix_hi[kTcStages][TM * TK];  // 16 KB each, SWIZZLE_128B K-major (8-row groups of 1024 B)
  float bfix_lo[kTcStages][TM * TK];
  float amov_hi[kTcStages][TN * TK];  // 32 KB each
  float amov_lo[kTcStages][TN * TK];
  float rowterm_a[2][TN];             // per-tile row terms of the moving cells
  uint64_t full[kTcStages];
  uint64_t empty[kTcStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset = 1024 B (8 rows x 128 B) >> 4 in
// [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, K-major A and B
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4)                    // c_format = F32
         | (2u << 7)                  // a_format = TF32
         | (2u << 10)                 // b_format = TF32
         | ((uint32_t)(N >> 3) << 17) // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
      "%25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float tc_cost_to_prob(float dot, float ta, float tb, int metric, int prob_type, float neg_inv2b) {
  float e;
  if (metric == SPB_METRIC_KL) e = (ta - tb) - dot;  // tb = centring term c_j of the fixed cell
  else if (metric == SPB_METRIC_SYMKL) e = 0.5f * ((ta + tb) - dot);             // utils.py:922-932
  else if (metric == SPB_METRIC_COS) e = fmaf(-0.5f, dot, 0.5f);
  else {
    e = fmaxf(ta + tb - 2.0f * dot, 0.0f);
    if (metric == SPB_METRIC_SQRT_EUC) e = sqrtf(e);
  }
  if (prob_type == SPB_PROB_GAUSS) return __expf(e * neg_inv2b);
  if (prob_type == SPB_PROB_COS) return 1.0f - e;
  return e;
}

// Tile order: bands of kBand fixed-cell tiles, moving-cell tiles fastest inside a band, so the ~148 tiles in flight
// touch ~kBand B-side and ~148/kBand A-side operand panels (tens of MB, L2 resident) instead of 148 distinct A panels.
constexpr int kBand = 16;
__device__ __forceinline__ void tile_coords(int tile, int tiles_i, int tiles_j, int& ti, int& tj) {
  const int per_band = kBand * tiles_i;
  const int band = tile / per_band, r = tile % per_band;
  const int bh = min(kBand, tiles_j - band * kBand);
  ti = r / bh;
  tj = band * kBand + r % bh;
}

__global__ void __launch_bounds__(kTcThreads, 1)
gene_cost_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                    const float* __restrict__ rtA, const float* __restrict__ rtB, int64_t NA, int64_t NB, int nkb,
                    int tiles_i, int tiles_j, int metric, int prob_type, float neg_inv2b, int accumulate,
                    float* __restrict__ GT, int64_t ldx) {
  extern __shared__ uint8_t tc_smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles: align the dynamic shared-memory window by hand
  TcSmem& sm = *reinterpret_cast<TcSmem*>(tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = tiles_i * tiles_j;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.acc_full[b], 1);
      mbar_init(&sm.acc_empty[b], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {  // TMEM allocation by one warp: all 512 columns (two 256-column accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int ti, tj;
        tile_coords(tile, tiles_i, tiles_j, ti, tj);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kTcStages;
          if (it >= kTcStages) mbar_wait(&sm.empty[s], ((it / kTcStages) - 1) & 1);
          mbar_expect_tx(&sm.full[s], (uint32_t)((2 * TM + 2 * TN) * TK * 4));
          tma_load_2d(sm.bfix_hi[s], &map_b_hi, kb * TK, tj * TM, &sm.full[s]);
          tma_load_2d(sm.bfix_lo[s], &map_b_lo, kb * TK, tj * TM, &sm.full[s]);
          tma_load_2d(sm.amov_hi[s], &map_a_hi, kb * TK, ti * TN, &sm.full[s]);
          tma_load_2d(sm.amov_lo[s], &map_a_lo, kb * TK, ti * TN, &sm.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
      int it = 0, t_local = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_local) {
        const int buf = t_local & 1;
        if (t_local >= 2) mbar_wait(&sm.acc_empty[buf], ((t_local >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kTcStages;
          mbar_wait(&sm.full[s], (it / kTcStages) & 1);
          tc_fence_after();
          const uint64_t d_bhi = umma_desc_k_sw128(sm.bfix_hi[s]), d_blo = umma_desc_k_sw128(sm.bfix_lo[s]);
          const uint64_t d_ahi = umma_desc_k_sw128(sm.amov_hi[s]), d_alo = umma_desc_k_sw128(sm.amov_lo[s]);
#pragma unroll
          for (int k = 0; k < TK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per K = 8 step inside the 128-byte swizzle row
            // UMMA "A" (M side) = fixed cells, "B" (N side) = moving cells; small cross terms first
            umma_tf32(tmem_d, d_blo + adv, d_ahi + adv, idesc, (kb | k) != 0);
            umma_tf32(tmem_d, d_bhi + adv, d_alo + adv, idesc, 1);
            umma_tf32(tmem_d, d_bhi + adv, d_ahi + adv, idesc, 1);
          }
          umma_commit(&sm.empty[s]);  // frees the smem stage once these MMAs have read it
        }
        umma_commit(&sm.acc_full[buf]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===== epilogue warps (2..5): TMEM lanes 32 * (warp % 4) =====
    const int q = warp & 3;
    const int et = threadIdx.x - 64;  // 0..127
    int t_local = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t_local) {
      const int buf = t_local & 1;
      int ti, tj;
      tile_coords(tile, tiles_i, tiles_j, ti, tj);
      const int64_t j = (int64_t)tj * TM + q * 32 + lane;
      const int64_t i0 = (int64_t)ti * TN;
      // stage the moving cells' row terms of this tile
      for (int c = et; c < TN; c += 128) sm.rowterm_a[buf][c] = (rtA != nullptr && i0 + c < NA) ? rtA[i0 + c] : 0.f;
      named_bar_sync(2, 128);
      const float tb = (rtB != nullptr && j < NB) ? rtB[j] : 0.f;
      mbar_wait(&sm.acc_full[buf], (t_local >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TN);
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        if (j < NB) {
          float* dst = GT + j * ldx + i0 + c0;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            if (i0 + c0 + c < ldx) {
              float o[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + c0 + c + u;
                o[u] = i < NA ? tc_cost_to_prob(__uint_as_float(r[c + u]), sm.rowterm_a[buf][c0 + c + u], tb, metric,
                                                prob_type, neg_inv2b)
                              : 0.f;
              }
              float4* d4 = reinterpret_cast<float4*>(dst + c);
              if (accumulate) {
                const float4 old = *d4;
                o[0] *= old.x; o[1] *= old.y; o[2] *= old.z; o[3] *= old.w;
              }
              *d4 = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// x -> (hi, lo): hi keeps the 10 explicit mantissa bits a tf32 operand keeps, lo = x - hi (exact in fp32)
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h;
    lo[i] = v - h;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t Gp, int64_t pitch, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return SPB_EUNSUPPORTED;
  const cuuint64_t dims[2] = {(cuuint64_t)Gp, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 700 + (int)r;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int spb_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  if (n <= 0) return 0;
  split_tf32_kernel<<<1184, 256, 0, ST>>>(x, hi, lo, n);
  SPB_CHECK_LAUNCH();
  return 0;
}

extern "C" int spb_gene_cost_tc(const float* A_hi, const float* A_lo, int64_t lda, const float* rowtermA, const float* B_hi,
                                const float* B_lo, int64_t ldb, const float* rowtermB, int64_t NA, int64_t NB, int64_t G,
                                int32_t metric, int32_t prob_type, float prob_param, int32_t accumulate, float* GT,
                                int64_t ldx, void* stream) {
  if (lda % 4 != 0 || ldb % 4 != 0 || ldx % 4 != 0) return SPB_EINVAL;
  const int64_t Gp = ((G + TK - 1) / TK) * TK;
  if (lda < Gp || ldb < Gp) return SPB_EINVAL;  // operands must be zero-padded to a multiple of 32 features
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, A_hi, NA, Gp, lda, TN))) return rc;
  if ((rc = make_map(&ma_lo, A_lo, NA, Gp, lda, TN))) return rc;
  if ((rc = make_map(&mb_hi, B_hi, NB, Gp, ldb, TM))) return rc;
  if ((rc = make_map(&mb_lo, B_lo, NB, Gp, ldb, TM))) return rc;
  const int tiles_i = (int)((ldx + TN - 1) / TN), tiles_j = (int)((NB + TM - 1) / TM);
  const float neg_inv2b = prob_type == SPB_PROB_GAUSS ? -1.0f / (2.0f * prob_param) : 0.f;
  static int n_sm_dev[SPB_MAX_DEVICES] = {};  // SM count + shared-memory opt-in, per device
  const int dev_ = spb_current_device();
  if (n_sm_dev[dev_] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev_);
    cudaError_t e = cudaFuncSetAttribute(gene_cost_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem) + 1024);
    if (e != cudaSuccess) return (int)e;
    n_sm_dev[dev_] = n;
  }
  const int n_sm = n_sm_dev[dev_];
  const int grid = min(n_sm, tiles_i * tiles_j);
  gene_cost_tc_kernel<<<grid, kTcThreads, sizeof(TcSmem) + 1024, ST>>>(ma_hi, ma_lo, mb_hi, mb_lo, rowtermA, rowtermB, NA, NB,
                                                                        (int)(Gp / TK), tiles_i, tiles_j, metric, prob_type,
                                                                        neg_inv2b, accumulate, GT, ldx);
  SPB_CHECK_LAUNCH();
  return 0;
}
