// bf16 GEMM  C[M,N] = A[M,K] * B[N,K]^T (+ R)  on the 5th-gen tensor cores (tcgen05.mma, accumulators
// in TMEM), operands staged by TMA (128B swizzle) through a multi-stage mbarrier ring; persistent
// CTAs, one per SM, warp-specialised:
//   warp 0 : TMA producer (one elected thread)
//   warp 1 : MMA issuer   (one elected thread)
//   warp 2 : TMEM allocator / deallocator
//   warps 4-7 : epilogue (TMEM -> registers -> bf16/fp32 -> global, optional residual add)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// This kernel replaces every frozen nn.Linear of the reference (llama/model.py:89,99-100,128,142,
// 348,354) and — with the load-time transposed weight copies — its dX-only backward.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace fvqa {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_UMMA_K = 16;

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = GEMM_BM * GEMM_BK * 2;
  static constexpr int kStageBytesB = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two: 256 or 512)
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024 manual alignment
};

// Epilogue extras. ROPE (QKV projection only): columns < rope_cols hold q|k heads of width hd; the
// interleaved pairs (2i, 2i+1) of row r are rotated by angle[pos = r % S][i] (llama/model.py:61-67)
// while the fp32 accumulator is still in registers, so attention never sees un-rotated q/k.
struct GemmEpi {
  const void* R;      // residual, same dtype as the output (nullptr = none)
  int ldr;
  const float* cosT;  // [S, hd/2]
  const float* sinT;
  int rope_cols, hd, S;
};

template <int BN, bool OUT_F32, bool ROPE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    void* __restrict__ Cout, const GemmEpi epi, int M, int N, int K, int ldc) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem holder
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t holder = bar_base + 8u * (2 * Cfg::kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kStages * Cfg::kStageBytes + 8 * (2 * Cfg::kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = K / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(holder, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * GEMM_BM;   // M fastest: concurrent CTAs share the weight tile
        const int n0 = (tile / tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kStageBytesA;
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * GEMM_BK, m0);
          tma_load_2d(sb, &tmap_b, full_bar(stage), kb * GEMM_BK, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kStageBytesA;
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle row: +2 in 16-byte units
            umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));               // smem slot reusable once these MMAs finish
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));                   // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % tiles_m) * GEMM_BM;
      const int n0 = (tile / tiles_m) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN + c * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row_ok && col0 < N) {
          if constexpr (OUT_F32) {
            float* crow = reinterpret_cast<float*>(Cout) + static_cast<long>(row) * ldc + col0;
            const float* rrow = epi.R ? reinterpret_cast<const float*>(epi.R) + static_cast<long>(row) * epi.ldr + col0 : nullptr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + j * 4 < N) {
                float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                if (rrow != nullptr) {
                  const float4 rr = *reinterpret_cast<const float4*>(rrow + j * 4);
                  o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                *reinterpret_cast<float4*>(crow + j * 4) = o;
              }
            }
          } else {
            bf16* crow = reinterpret_cast<bf16*>(Cout) + static_cast<long>(row) * ldc + col0;
            const bf16* rrow = epi.R ? reinterpret_cast<const bf16*>(epi.R) + static_cast<long>(row) * epi.ldr + col0 : nullptr;
            if constexpr (ROPE) {
              if (col0 < epi.rope_cols) {
                const int pos = row % epi.S;
                const int i0 = (col0 % epi.hd) >> 1;
                const float4* c4 = reinterpret_cast<const float4*>(epi.cosT + static_cast<long>(pos) * (epi.hd >> 1) + i0);
                const float4* s4 = reinterpret_cast<const float4*>(epi.sinT + static_cast<long>(pos) * (epi.hd >> 1) + i0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float4 cc = __ldg(c4 + q), ss = __ldg(s4 + q);
                  const float cv[4] = {cc.x, cc.y, cc.z, cc.w}, sv[4] = {ss.x, ss.y, ss.z, ss.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float a = __uint_as_float(v[8 * q + 2 * e]), b = __uint_as_float(v[8 * q + 2 * e + 1]);
                    v[8 * q + 2 * e] = __float_as_uint(a * cv[e] - b * sv[e]);
                    v[8 * q + 2 * e + 1] = __float_as_uint(a * sv[e] + b * cv[e]);
                  }
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col0 + j * 8 < N) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
                if (rrow != nullptr) {
                  float r[8];
                  unpack8(*reinterpret_cast<const uint4*>(rrow + j * 8), r);
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] += r[e];
                }
                *reinterpret_cast<uint4*>(crow + j * 8) = pack8(f);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static std::mutex g_mu;

struct MapKey {
  const void* ptr;
  int rows, cols, ld, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.cols) * 0xC2B2AE3D27D4EB4Full) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.ld) * 0x165667B19E3779F9ull) + (h << 6) + (h >> 2);
    h ^= static_cast<size_t>(k.box_rows) + (h << 6) + (h >> 2);
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

int gemm_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_encode != nullptr) return FVQA_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  FVQA_REQUIRE(e == cudaSuccess && fn != nullptr && qres == cudaDriverEntryPointSuccess, FVQA_ERR_CUDA,
               "cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
  int dev = 0;
  e = cudaGetDevice(&dev);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  FVQA_REQUIRE(prop.major == 10, FVQA_ERR_UNSUPPORTED, "this library only runs on sm_100 (found sm_%d%d)", prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
#define FVQA_SET_SMEM(BN, F32, ROPE)                                                                          \
  e = cudaFuncSetAttribute(gemm_bf16_nt_kernel<BN, F32, ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                           GemmCfg<BN>::kSmemBytes);                                                          \
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
  FVQA_SET_SMEM(256, false, false)
  FVQA_SET_SMEM(256, true, false)
  FVQA_SET_SMEM(128, false, false)
  FVQA_SET_SMEM(128, true, false)
  FVQA_SET_SMEM(256, false, true)
  FVQA_SET_SMEM(128, false, true)
#undef FVQA_SET_SMEM
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return FVQA_OK;
}

static int get_tmap(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, rows, cols, ld, box_rows};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return FVQA_OK;
  }
  CUtensorMap m;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FVQA_REQUIRE(r == CUDA_SUCCESS, FVQA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%d cols=%d ld=%d box_rows=%d",
               static_cast<int>(r), ptr, rows, cols, ld, box_rows);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return FVQA_OK;
}

template <int BN, bool OUT_F32, bool ROPE>
static int launch_gemm(const bf16* A, int lda, const bf16* B, int ldb, void* C, int ldc, const GemmEpi& epi, int M,
                       int N, int K, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = get_tmap(A, M, K, lda, GEMM_BM, &ta);
  if (rc) return rc;
  rc = get_tmap(B, N, K, ldb, BN, &tb);
  if (rc) return rc;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_bf16_nt_kernel<BN, OUT_F32, ROPE><<<grid, GEMM_THREADS, GemmCfg<BN>::kSmemBytes, stream>>>(ta, tb, C, epi, M, N, K, ldc);
  return check_launch("gemm_bf16_nt");
}

static int check_gemm_args(const void* A, int lda, const void* B, int ldb, const void* C, int ldc, const void* R, int ldr, int M, int N, int K) {
  FVQA_REQUIRE(g_encode != nullptr, FVQA_ERR_INVALID_ARG, "fvqa_init() has not been called");
  FVQA_REQUIRE(M > 0 && N > 0 && K > 0, FVQA_ERR_INVALID_ARG, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  FVQA_REQUIRE(K % GEMM_BK == 0, FVQA_ERR_UNSUPPORTED, "gemm: K=%d must be a multiple of %d", K, GEMM_BK);
  FVQA_REQUIRE(N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && (R == nullptr || ldr % 8 == 0),
               FVQA_ERR_UNSUPPORTED, "gemm: N/lda/ldb/ldc/ldr must be multiples of 8 (N=%d lda=%d ldb=%d ldc=%d ldr=%d)", N, lda, ldb, ldc, ldr);
  FVQA_REQUIRE(lda >= K && ldb >= K && ldc >= N, FVQA_ERR_INVALID_ARG, "gemm: leading dimensions too small");
  FVQA_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (reinterpret_cast<uintptr_t>(R) & 15) == 0,
               FVQA_ERR_INVALID_ARG, "gemm: pointers must be 16-byte aligned");
  return FVQA_OK;
}

// Tile choice: 128x256 unless N is small or the 128x128 tiling fills the SM waves clearly better.
static bool prefer_bn128(int M, int N) {
  const int tm = (M + GEMM_BM - 1) / GEMM_BM;
  auto eff = [&](int bn) {
    const long tn = (N + bn - 1) / bn;
    const long tiles = static_cast<long>(tm) * tn;
    const long waves = (tiles + g_num_sms - 1) / g_num_sms;
    return (static_cast<double>(tiles) / (static_cast<double>(waves) * g_num_sms)) * (static_cast<double>(N) / (tn * bn));
  };
  return (N <= 128) || (eff(128) > eff(256) + 0.08);
}

}  // namespace fvqa

using namespace fvqa;

extern "C" int fvqa_gemm_bf16_nt(const fvqa_bf16* A, int lda, const fvqa_bf16* B, int ldb, void* C, int ldc,
                                 const void* R, int ldr, int M, int N, int K, int out_fp32, void* stream) {
  int rc = check_gemm_args(A, lda, B, ldb, C, ldc, R, ldr, M, N, K);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16* a = reinterpret_cast<const bf16*>(A);
  const bf16* b = reinterpret_cast<const bf16*>(B);
  GemmEpi epi{R, ldr, nullptr, nullptr, 0, 0, 1};
  if (prefer_bn128(M, N)) {
    return out_fp32 ? launch_gemm<128, true, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s)
                    : launch_gemm<128, false, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
  }
  return out_fp32 ? launch_gemm<256, true, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s)
                  : launch_gemm<256, false, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
}

extern "C" int fvqa_gemm_bf16_nt_rope(const fvqa_bf16* A, int lda, const fvqa_bf16* B, int ldb, fvqa_bf16* C, int ldc, int M, int N,
                                      int K, const float* rope_cos, const float* rope_sin, int rope_cols, int hd, int S,
                                      void* stream) {
  int rc = check_gemm_args(A, lda, B, ldb, C, ldc, nullptr, 0, M, N, K);
  if (rc) return rc;
  FVQA_REQUIRE((hd == 64 || hd == 128) && rope_cols % hd == 0 && rope_cols <= N && S > 0 && rope_cos && rope_sin,
               FVQA_ERR_INVALID_ARG, "gemm_rope: bad rope arguments (hd=%d rope_cols=%d S=%d)", hd, rope_cols, S);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16* a = reinterpret_cast<const bf16*>(A);
  const bf16* b = reinterpret_cast<const bf16*>(B);
  GemmEpi epi{nullptr, 0, rope_cos, rope_sin, rope_cols, hd, S};
  if (prefer_bn128(M, N)) return launch_gemm<128, false, true>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
  return launch_gemm<256, false, true>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
}
