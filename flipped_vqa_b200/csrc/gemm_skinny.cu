// Skinny GEMM  C[M,N] = A[M,K] * B[N,K]^T  for M <= 16: the adapter-prompt projections
//   akv_l      = adapter_l[A=10, d] . [Wk;Wv]_l^T                    (llama/model.py:99-100)
//   d adapter_l = dK_a . Wk + dV_a . Wv  = dakv[10, 2d] . Wkv_t^T     (its backward; SURVEY.md 8(a) addendum)
// These read a 67 MB weight block for 0.7 GFLOP: HBM/L2-bound, and a 128-row tcgen05 tile would waste
// 92 % of every MMA and serialise 64-128 k-blocks per CTA. Here a CTA of 8 warps owns 8*NT output columns;
// warp w streams K-slice w of the weight rows with 16-byte loads straight into mma.sync.m16n8k16
// B fragments (the k index inside each 32-element chunk is permuted identically for A and B, which a
// contraction does not see), and the 8 partial tiles are summed in shared memory in a fixed order.
#include <atomic>

#include "common.cuh"

namespace fvqa {

namespace {

__device__ __forceinline__ void mma16816_sk(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." FVQA_MMA_TYPE "." FVQA_MMA_TYPE ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;

}  // namespace

// Optional epilogue of the single-group launch (the M = bsz rows of a KV-cached decode step, StepEngine.generate):
// fp32 residual add (h = x + attn / out = h + ffn, llama/model.py:185-186) or RoPE of the q|k columns by each row's position
// (llama/model.py:61-67), exactly what the tcgen05 GEMM's epilogues do for the M > 16 problems.
struct SkinnyEpi {
  const float* R = nullptr;          // [M, N] fp32 residual (OUT_F32 only)
  int ldr = 0;
  const float* cosT = nullptr;       // [positions, hd / 2]
  const float* sinT = nullptr;
  int rope_cols = 0, hd = 1, S = 1;
  const int32_t* pos_ids = nullptr;  // position of row r (nullptr -> r % S)
};

// Grouped form (blockIdx.y = group g): A_g = A + g * strideA, C_g = C + g * strideC (elements), B_g = Bptrs[g] when a device
// pointer table is given (the per-layer weight blocks live in separate allocations), else B. One launch then covers the
// adapter projections of ALL layers (2.1 GB of weights at 7B) instead of 32 launches of 67 MB that are mostly ramp and tail.
template <int NT, bool OUT_F32>
__global__ void __launch_bounds__(SK_THREADS) gemm_skinny_kernel(const h16* __restrict__ A, long strideA, int lda,
                                                                 const h16* __restrict__ B, const h16* const* __restrict__ Bptrs, int ldb,
                                                                 void* __restrict__ C, long strideC, int ldc, int M, int N, int K,
                                                                 const SkinnyEpi epi) {
  __shared__ float red[SK_WARPS][16][8 * NT + 1];
  pdl_launch_dependents();
  pdl_wait();
  {
    const int grp = blockIdx.y;
    A += grp * strideA;
    if (Bptrs != nullptr) B = Bptrs[grp];
    if constexpr (OUT_F32) C = reinterpret_cast<float*>(C) + grp * strideC;
    else C = reinterpret_cast<h16*>(C) + grp * strideC;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 8 * NT;
  const int kslice = K / SK_WARPS;                     // multiple of 32 (checked by the launcher)
  const int k_begin = warp * kslice;
  const bool row_lo = g < M, row_hi = g + 8 < M;
  const uint4* a_lo = reinterpret_cast<const uint4*>(A + static_cast<long>(row_lo ? g : 0) * lda + k_begin + 8 * t);
  const uint4* a_hi = reinterpret_cast<const uint4*>(A + static_cast<long>(row_hi ? g + 8 : 0) * lda + k_begin + 8 * t);
  const uint4* b_ptr[NT];
  bool b_ok[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int col = n0 + j * 8 + g;
    b_ok[j] = col < N;
    b_ptr[j] = reinterpret_cast<const uint4*>(B + static_cast<long>(b_ok[j] ? col : 0) * ldb + k_begin + 8 * t);
  }
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const int chunks = kslice / 32;                      // 32 k per chunk = 4 uint4 per row
#pragma unroll 8
  for (int c = 0; c < chunks; ++c) {
    const uint4 al = row_lo ? __ldg(a_lo + 4 * c) : zero;
    const uint4 ah = row_hi ? __ldg(a_hi + 4 * c) : zero;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint4 b = b_ok[j] ? __ldg(b_ptr[j] + 4 * c) : zero;
      mma16816_sk(acc[j], al.x, ah.x, al.y, ah.y, b.x, b.y);
      mma16816_sk(acc[j], al.z, ah.z, al.w, ah.w, b.z, b.w);
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    red[warp][g][j * 8 + 2 * t] = acc[j][0];
    red[warp][g][j * 8 + 2 * t + 1] = acc[j][1];
    red[warp][g + 8][j * 8 + 2 * t] = acc[j][2];
    red[warp][g + 8][j * 8 + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  // one thread per (row, column PAIR): a RoPE pair (2i, 2i + 1) never straddles threads (N is even: checked by the launcher)
  for (int idx = threadIdx.x; idx < 16 * 4 * NT; idx += SK_THREADS) {
    const int r = idx / (4 * NT), cc = 2 * (idx - r * (4 * NT));
    const int col = n0 + cc;
    if (r < M && col < N) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) { s0 += red[w][r][cc]; s1 += red[w][r][cc + 1]; }
      if (epi.cosT != nullptr && col < epi.rope_cols) {
        const int pos = epi.pos_ids != nullptr ? __ldg(epi.pos_ids + r) : r % epi.S;
        const long ti = static_cast<long>(pos) * (epi.hd >> 1) + ((col % epi.hd) >> 1);
        const float cv = __ldg(epi.cosT + ti), sv = __ldg(epi.sinT + ti);
        const float a = s0, b = s1;
        s0 = a * cv - b * sv;
        s1 = a * sv + b * cv;
      }
      if constexpr (OUT_F32) {
        float* o = reinterpret_cast<float*>(C) + static_cast<long>(r) * ldc + col;
        if (epi.R != nullptr) { s0 += epi.R[static_cast<long>(r) * epi.ldr + col]; s1 += epi.R[static_cast<long>(r) * epi.ldr + col + 1]; }
        o[0] = s0; o[1] = s1;
      } else {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<h16*>(C) + static_cast<long>(r) * ldc + col) = pack_h16x2(s0, s1);
      }
    }
  }
}

// R: fp32 residual is handled (out_fp32 launches); an h16 residual is not
bool gemm_skinny_supported(int M, int K, const void* R) { return M <= 16 && R == nullptr && K % (SK_WARPS * 32) == 0; }
bool gemm_skinny_shape_ok(int M, int N, int K) { return M <= 16 && K % (SK_WARPS * 32) == 0 && N % 2 == 0; }

std::atomic<int> g_skinny_force_nt{0};   // tuning hook (fvqa_gemm_debug_skinny_nt): column blocks of 8 * nt per CTA; 0 = heuristic

static int skinny_launch(const h16* A, long strideA, int lda, const h16* B, const h16* const* Bptrs, int ldb, void* C, long strideC,
                         int ldc, int M, int N, int K, int groups, int out_fp32, int num_sms, const SkinnyEpi& epi, cudaStream_t stream) {
  // 16 columns per CTA when that still gives every SM ~2 CTAs, else 8. (32 columns per CTA cost registers / resident CTAs:
  // measured 4.96 vs 5.38 TB/s on the 2.1 GB forward launch and 3.5 vs 4.3 TB/s on a backward chunk, tools/skinny_bench.py.)
  const long blocks16 = static_cast<long>(N / 16) * groups;
  int nt = blocks16 >= 2 * num_sms ? 2 : 1;
  if (g_skinny_force_nt == 1 || g_skinny_force_nt == 2 || g_skinny_force_nt == 4) nt = g_skinny_force_nt.load();
  const int cols = 8 * nt;
  const dim3 grid((N + cols - 1) / cols, groups);
#define FVQA_SK(NT_)                                                                                                                \
  if (out_fp32) launch_k(gemm_skinny_kernel<NT_, true>, grid, dim3(SK_THREADS), 0, stream, A, strideA, lda, B, Bptrs, ldb, C, strideC, ldc, M, N, K, epi); \
  else launch_k(gemm_skinny_kernel<NT_, false>, grid, dim3(SK_THREADS), 0, stream, A, strideA, lda, B, Bptrs, ldb, C, strideC, ldc, M, N, K, epi);
  if (nt == 4) { FVQA_SK(4) } else if (nt == 2) { FVQA_SK(2) } else { FVQA_SK(1) }
#undef FVQA_SK
  return check_launch("gemm_skinny");
}

int gemm_skinny_grouped(const h16* A, long strideA, int lda, const h16* B, const h16* const* Bptrs, int ldb, void* C, long strideC,
                        int ldc, int M, int N, int K, int groups, int out_fp32, int num_sms, cudaStream_t stream) {
  return skinny_launch(A, strideA, lda, B, Bptrs, ldb, C, strideC, ldc, M, N, K, groups, out_fp32, num_sms, SkinnyEpi{}, stream);
}

int gemm_skinny(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, int M, int N, int K, int out_fp32, int num_sms,
                cudaStream_t stream) {
  return skinny_launch(A, 0, lda, B, nullptr, ldb, C, 0, ldc, M, N, K, 1, out_fp32, num_sms, SkinnyEpi{}, stream);
}

// + fp32 residual (out_fp32) or RoPE epilogue: the decode-step projections of the generation evaluator (M = bsz <= 16 rows)
int gemm_skinny_epi(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, int M, int N, int K, int out_fp32, const float* R,
                    int ldr, const float* cosT, const float* sinT, int rope_cols, int hd, int S, const int32_t* pos_ids, int num_sms,
                    cudaStream_t stream) {
  SkinnyEpi epi;
  epi.R = R; epi.ldr = ldr; epi.cosT = cosT; epi.sinT = sinT; epi.rope_cols = rope_cols; epi.hd = hd > 0 ? hd : 1; epi.S = S > 0 ? S : 1;
  epi.pos_ids = pos_ids;
  return skinny_launch(A, 0, lda, B, nullptr, ldb, C, 0, ldc, M, N, K, 1, out_fp32, num_sms, epi, stream);
}

}  // namespace fvqa
