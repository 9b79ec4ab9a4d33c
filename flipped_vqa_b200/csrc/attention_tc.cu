// tcgen05 / TMEM attention for the step's dominant shape: S <= 128 tokens, head_dim = 128
// (7B / 13B NExT-QA: S = 128). One CTA of 128 threads owns one (sequence, head): the whole
// 128 x 128 score tile lives in TMEM, thread t owns query row t (TMEM lane t).
//
//   forward : TMA(Q,K,Ka | V,Va) -> UMMA S = Q K^T, S_a = Q Ka^T -> per-row softmax in registers
//             (causal + gate2 block bias; separate adapter softmax x tanh(gate1), llama/model.py:111-122)
//             -> P, P_a as h16 UMMA operands in shared memory (P overwrites K) -> UMMA O = P V + P_a Va
//             -> h16 store. 2 CTAs per SM (110 KB smem, 256 TMEM columns each) overlap one unit's
//             loads with the other's math: at S = 128 attention is HBM-bound (65 FLOP/B), not tensor-bound.
//
// Shared-memory operand layouts: everything TMA loads is the 128-byte-swizzled [rows][64 elem] box.
// Used K-major when the contraction runs along the 64-element rows (Q, K, Ka, P) and MN-major when it
// runs along the box rows (V, Va as B of P.V: N = head dim contiguous, K = keys).
#include <atomic>

#include "../../include/fvqa_debug.h"
#include "attention_tc.cuh"
#include "tmap.h"

namespace fvqa {

namespace {

// ---- forward shared-memory map (bytes from the 1024-aligned base) ----
constexpr int F_SQ = 0;               // [2][128][64] h16; O staging for the TMA store at the end
constexpr int F_SK = 32768;           // [2][128][64]; overwritten by P after S is complete
constexpr int F_SV = 65536;           // [2][128][64]
constexpr int F_SKA = 98304;          // [2][16][64]
constexpr int F_SVA = 102400;         // [2][16][64]
constexpr int F_SPA = 106496;         // P_a: [16 row groups][2 k-chunks][8 rows][8] h16, no swizzle (4 KB)
constexpr int F_BAR = 110592;
constexpr int F_ROW = F_BAR + 64;     // [4][128] floats: partial row max / sum
constexpr int F_SMEM = F_ROW + 2048 + 1024;
constexpr int FW_THREADS = 256;

}  // namespace

// 256 threads: warps w and w + 4 share TMEM lane quadrant w & 3 (rows) and split the 128 key columns in halves.
__global__ void __launch_bounds__(FW_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_akv,
                   const __grid_constant__ CUtensorMap tm_out, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_qk = sbase + F_BAR, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_o = bar_qk + 24, holder = bar_qk + 32;
  const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, part = tid >> 7, quad = warp & 3;
  const int h = blockIdx.x, n = blockIdx.y;
  const int S = p.S, D = p.H * 128;

  if (tid == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 256);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + F_BAR + 32);
  pdl_wait();                                          // set-up above overlaps the previous kernel's tail

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_akv);
    tma_prefetch_desc(&tm_out);
    const int c = h * 128;
    mbar_arrive_expect_tx(bar_qk, 2 * 32768 + 4096);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + F_SQ + kb * 16384, &tm_qkv, bar_qk, c + kb * 64, 0, n);
      tma_load_3d(sbase + F_SK + kb * 16384, &tm_qkv, bar_qk, D + c + kb * 64, 0, n);
      tma_load_2d(sbase + F_SKA + kb * 2048, &tm_akv, bar_qk, c + kb * 64, 0);
    }
    mbar_arrive_expect_tx(bar_v, 32768 + 4096);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + F_SV + kb * 16384, &tm_qkv, bar_v, 2 * D + c + kb * 64, 0, n);
      tma_load_2d(sbase + F_SVA + kb * 2048, &tm_akv, bar_v, D + c + kb * 64, 0);
    }
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    constexpr uint32_t id_s = idesc_h16(128, 128, 0, 0), id_a = idesc_h16(128, 16, 0, 0);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t a = umma_desc_k_sw128(sbase + F_SQ + (ks >> 2) * 16384) + static_cast<uint64_t>(2 * (ks & 3));
      const uint64_t b = umma_desc_k_sw128(sbase + F_SK + (ks >> 2) * 16384) + static_cast<uint64_t>(2 * (ks & 3));
      const uint64_t ba = umma_desc_k_sw128(sbase + F_SKA + (ks >> 2) * 2048) + static_cast<uint64_t>(2 * (ks & 3));
      umma_h16_ss(tmem, a, b, id_s, ks > 0 ? 1u : 0u);            // S   -> columns [0,128)
      umma_h16_ss(tmem + 128, a, ba, id_a, ks > 0 ? 1u : 0u);     // S_a -> columns [128,144)
    }
    umma_commit(bar_s);
  }
  __syncwarp();
  mbar_wait(bar_s, 0);
  tc_fence_after();

  // ---------------- softmax: thread = (query row r, column half `part`) ----------------
  const uint32_t tlane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const float scale2 = rsqrtf(128.f) * TC_LOG2E;
  const int vs = p.vstart[n];
  const float bias2 = (vs >= 0) ? p.gate2[h] * TC_LOG2E : 0.f;
  const bool row_biased = (vs >= 0) && (r >= vs + p.F);
  const int bias_c0 = vs, bias_c1 = vs + p.F;
  float* s_max = reinterpret_cast<float*>(sgen + F_ROW);          // [2][128] partial row max | [2][128] partial row sum
  float* s_sum = s_max + 256;
  float s[64];
  float mx = -INFINITY;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = 2 * part + cc;
    if (c <= quad) {                                   // warp-uniform: chunks beyond the warp's last row are fully masked
      uint32_t v[32];
      tmem_ld_32x32(tlane + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        float x = __uint_as_float(v[j]) * scale2;
        if (row_biased && col >= bias_c0 && col < bias_c1) x += bias2;
        if (col > r) x = -INFINITY;
        s[cc * 32 + j] = x;
        mx = fmaxf(mx, x);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) s[cc * 32 + j] = -INFINITY;
    }
  }
  s_max[part * 128 + r] = mx;
  // adapter branch (part 1 threads; part 0 always has live text chunks): separate softmax x tanh(gate1)
  uint32_t pa[8];
  if (part == 1) {
    uint32_t v[32];
    tmem_ld_32x16(tlane + 128u, v);
    tmem_ld_wait();
    const float tg = tanhf(p.gate1[h]);
    float sa[16], ma = -INFINITY, la = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sa[j] = (j < p.A) ? __uint_as_float(v[j]) * scale2 : -INFINITY;
      ma = fmaxf(ma, sa[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sa[j] = exp2f(sa[j] - ma);
      la += sa[j];
    }
    const float ia = tg / la;
#pragma unroll
    for (int e = 0; e < 8; ++e) pa[e] = pack_h16x2(sa[2 * e] * ia, sa[2 * e + 1] * ia);
  }
  tc_fence_before();
  __syncthreads();                                     // every S / S_a value is in registers; partial maxima published
  tc_fence_after();
  mx = fmaxf(s_max[r], s_max[128 + r]);
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const float e = exp2f(s[j] - mx);                  // masked entries: exp2(-inf) = 0
    s[j] = e;
    l += e;
  }
  s_sum[part * 128 + r] = l;
  __syncthreads();
  l = s_sum[r] + s_sum[128 + r];
  const float inv = 1.f / l;
  // P (normalised, h16, two keys per 32-bit column) -> TMEM columns [32 part, 32 part + 32) over the consumed S
  {
    uint32_t pk[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) pk[e] = pack_h16x2(s[2 * e] * inv, s[2 * e + 1] * inv);
    tmem_st_32x32(tlane + static_cast<uint32_t>(part * 32), pk);
  }
  if (part == 1) tmem_st_32x8(tlane + 64u, pa);        // P_a -> TMEM columns [64,72)
  tmem_st_wait();
  if (part == 0 && r < S) p.lse[(static_cast<long>(n) * p.H + h) * S + r] = (mx + log2f(l)) * TC_LN2;
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    mbar_wait(bar_v, 0);
    tc_fence_after();
    constexpr uint32_t id_o = idesc_h16(128, 128, 0, 1);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)                                 // O = P V, A = P from TMEM (8 columns = 16 keys per step)
      umma_h16_ts(tmem + 128, tmem + static_cast<uint32_t>(ks * 8), desc_mn_sw128(sbase + F_SV + ks * 2048, 16384), id_o, ks > 0 ? 1u : 0u);
    umma_h16_ts(tmem + 128, tmem + 64u, desc_mn_sw128(sbase + F_SVA, 2048), id_o, 1u);   // += P_a Va
    umma_commit(bar_o);
  }
  __syncwarp();
  mbar_wait(bar_o, 0);
  tc_fence_after();
  // O -> h16 -> swizzled staging (Q's buffer: every MMA that read it has retired) -> TMA store
#pragma unroll
  for (int c = 2 * part; c < 2 * part + 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tlane + 128u + static_cast<uint32_t>(c * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
      *reinterpret_cast<uint4*>(sgen + F_SQ + (c >> 1) * 16384 + sw128_off(r, (c & 1) * 4 + q)) = pack8(f);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tma_store_3d(&tm_out, sbase + F_SQ, h * 128, 0, n);            // rows >= S are out of bounds -> dropped
    tma_store_3d(&tm_out, sbase + F_SQ + 16384, h * 128 + 64, 0, n);
    tma_store_commit();
    tma_store_wait_read();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// backward: one CTA (128 threads, 1 per SM: 210 KB smem, all 512 TMEM columns) per (sequence, head).
//   UMMA  S = Q K^T, S_a = Q Ka^T, dP = dO V^T, dP_a = dO Va^T          (TMEM)
//   rows  pass 1: P = exp2(S - lse) (kept as packed h16 in registers, also written over V as a UMMA operand),
//                 D = sum_k P dP  (= <dO, O> minus the adapter part: no O / dO reads from global memory);
//         pass 2: dS = P (dP - D) / sqrt(hd) (+ gate2 partial); adapter softmax backward (gate1 partial);
//   UMMA  dV = P^T dO, dK = dS^T Q, dQ = dS K + dS_a Ka, dKa^T = Q^T dS_a, dVa^T = dO^T (tanh(g1) P_a)
//   rows  inverse RoPE on dQ / dK (fp16 cos|sin table staged in smem at kernel start), h16 -> swizzled smem
//         -> TMA stores; per-(sequence, head) adapter / gate partials to the workspace (reduced over sequences
//         in a fixed order by attn_bwd_reduce_kernel -> deterministic).
// Every [128][64]-element box is used K-major for one product and MN-major for another (same bytes).
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int B_SQ = 0;                // later: dQ staging
constexpr int B_SK = 32768;            // later: dK staging
constexpr int B_SV = 65536;            // P after dP is complete; later: dV staging
constexpr int B_SDO = 98304;
constexpr int B_SDS = 131072;
constexpr int B_SKA = 163840;          // [2][16][64]
constexpr int B_SVA = 167936;
constexpr int B_SPA = 172032;          // tanh(g1) P_a, no swizzle (4 KB)
constexpr int B_SDSA = 176128;         // dS_a, no swizzle (4 KB)
constexpr int B_ROPE = 180224;         // [128 rows][16 chunks ^ (row & 7)][4 x half2(cos, sin)]  (32 KB)
constexpr int B_BAR = 212992;
constexpr int B_RED = B_BAR + 64;      // 32 floats
constexpr int B_ROW = B_BAR + 256;     // [4][128] floats
constexpr int B_SMEM = B_ROW + 2048 + 1024;
constexpr int BW_THREADS = 512;
}  // namespace

// 512 threads: warps w, w+4, w+8, w+12 share TMEM lane quadrant w & 3 (rows 32 (w & 3) ..); thread = (row r, part):
// part owns the 32-column chunk `part` of every 128-column matrix. With one warp per scheduler the row-wise math
// was pure exposed latency (ncu: 36 K cycles per CTA).
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_akv,
                   const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dqkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_a = sbase + B_BAR, bar_b = bar_a + 8, bar_m1 = bar_a + 16, bar_m2 = bar_a + 24, holder = bar_a + 32;
  float* sred = reinterpret_cast<float*>(sgen + B_RED);           // 32 floats: per-warp gate partials
  float* s_row = reinterpret_cast<float*>(sgen + B_ROW);          // [4][128] per-row partial sums of D
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = tid & 127, part = tid >> 7, quad = warp & 3;
  const int h = blockIdx.x, n = blockIdx.y;
  const int S = p.S, D = p.H * 128;

  if (tid == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_b, 1);
    mbar_init(bar_m1, 1);
    mbar_init(bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(holder, 512);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + B_BAR + 32);
  pdl_wait();
  constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 384, T_SA = 256, T_DPA = 272, T_DQ = 0, T_DKA = 128;

  auto kmaj = [&](int off, int blk, int ks) {
    return umma_desc_k_sw128(sbase + off + (ks >> 2) * blk) + static_cast<uint64_t>(2 * (ks & 3));
  };
  auto mnmaj = [&](int off, int lbo, int ks) { return desc_mn_sw128(sbase + off + ks * 2048, lbo); };

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_akv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqkv);
    const int c = h * 128;
    mbar_arrive_expect_tx(bar_a, 2 * 32768 + 4096);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + B_SQ + kb * 16384, &tm_qkv, bar_a, c + kb * 64, 0, n);
      tma_load_3d(sbase + B_SK + kb * 16384, &tm_qkv, bar_a, D + c + kb * 64, 0, n);
      tma_load_2d(sbase + B_SKA + kb * 2048, &tm_akv, bar_a, c + kb * 64, 0);
    }
    mbar_arrive_expect_tx(bar_b, 2 * 32768 + 4096);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + B_SV + kb * 16384, &tm_qkv, bar_b, 2 * D + c + kb * 64, 0, n);
      tma_load_3d(sbase + B_SDO + kb * 16384, &tm_do, bar_b, c + kb * 64, 0, n);
      tma_load_2d(sbase + B_SVA + kb * 2048, &tm_akv, bar_b, D + c + kb * 64, 0);
    }
  }
  __syncwarp();
  // RoPE table of this sequence's positions -> smem as fp16 (cos, sin) pairs (coalesced; overlaps the TMA loads)
  stage_rope_table(sgen + B_ROPE, p.cosT, p.sinT, 0, S, tid, BW_THREADS);
  if (tid == 0) {
    constexpr uint32_t id_s = idesc_h16(128, 128, 0, 0), id_a = idesc_h16(128, 16, 0, 0);
    mbar_wait(bar_a, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      umma_h16_ss(tmem + T_S, kmaj(B_SQ, 16384, ks), kmaj(B_SK, 16384, ks), id_s, ks > 0 ? 1u : 0u);
      umma_h16_ss(tmem + T_SA, kmaj(B_SQ, 16384, ks), kmaj(B_SKA, 2048, ks), id_a, ks > 0 ? 1u : 0u);
    }
    mbar_wait(bar_b, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      umma_h16_ss(tmem + T_DP, kmaj(B_SDO, 16384, ks), kmaj(B_SV, 16384, ks), id_s, ks > 0 ? 1u : 0u);
      umma_h16_ss(tmem + T_DPA, kmaj(B_SDO, 16384, ks), kmaj(B_SVA, 2048, ks), id_a, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar_m1);
  }
  __syncwarp();

  const bool row_ok = r < S;
  const float nlse = row_ok ? -p.lse[(static_cast<long>(n) * p.H + h) * S + r] * TC_LOG2E : -1e30f;   // rows past S: P = 0
  const float scale = rsqrtf(128.f);
  const float scale2 = scale * TC_LOG2E;
  const int vs = p.vstart[n];
  const float bias2 = (vs >= 0) ? p.gate2[h] * TC_LOG2E : 0.f;
  const bool row_biased = (vs >= 0) && (r >= vs + p.F);
  const int bias_c0 = vs, bias_c1 = vs + p.F;
  const uint32_t tlane = tmem + (static_cast<uint32_t>(quad * 32) << 16);

  mbar_wait(bar_m1, 0);
  tc_fence_after();

  float g1_part = 0.f, g2_part = 0.f;
  // ---------------- adapter branch (one thread per row) ----------------
  if (part == 0) {
    const float tg = tanhf(p.gate1[h]);
    uint32_t v[32], w[32];
    tmem_ld_32x16(tlane + T_SA, v);
    tmem_ld_32x16(tlane + T_DPA, w);
    tmem_ld_wait();
    float sa[16], ma = -INFINITY, la = 0.f, da = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sa[j] = (j < p.A) ? __uint_as_float(v[j]) * scale2 : -INFINITY;
      ma = fmaxf(ma, sa[j]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sa[j] = exp2f(sa[j] - ma);
      la += sa[j];
    }
    const float ia = row_ok ? 1.f / la : 0.f;        // rows past the sequence contribute nothing
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      sa[j] *= ia;
      da += sa[j] * __uint_as_float(w[j]);             // <dO, P_a V_a>
    }
    g1_part = da;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float fp[8], fd[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float pa = sa[c * 8 + e];
        fp[e] = tg * pa;
        fd[e] = tg * pa * (__uint_as_float(w[c * 8 + e]) - da) * scale;
      }
      const int off = (r >> 3) * 256 + c * 128 + (r & 7) * 16;
      *reinterpret_cast<uint4*>(sgen + B_SPA + off) = pack8(fp);
      *reinterpret_cast<uint4*>(sgen + B_SDSA + off) = pack8(fd);
    }
  }
  // ---------------- text keys: this thread's 32-key chunk. P -> smem (over V), D partial -> smem ----------------
  const int ch = part;
  const bool live = ch <= quad;                        // warp-uniform causal skip: chunks beyond the warp's last row are masked
  float pf[32];
  uint32_t w[32];                                      // dP chunk, kept for pass 2
  {
    float dpart = 0.f;
    if (live) {
      uint32_t v[32];
      tmem_ld_32x32(tlane + T_S + static_cast<uint32_t>(ch * 32), v);
      tmem_ld_32x32(tlane + T_DP + static_cast<uint32_t>(ch * 32), w);
      tmem_ld_wait();
      const bool causal = ch == quad;
      const bool bias_any = row_biased && ch * 32 < bias_c1 && ch * 32 + 32 > bias_c0;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = ch * 32 + e;
        float t = fmaf(__uint_as_float(v[e]), scale2, nlse);
        if (bias_any && col >= bias_c0 && col < bias_c1) t += bias2;
        float pe = exp2f(t);
        if (causal && col > r) pe = 0.f;
        pf[e] = pe;
        dpart += pe * __uint_as_float(w[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) pf[e] = 0.f;
    }
    s_row[part * 128 + r] = dpart;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f8[e] = pf[q * 8 + e];
      *reinterpret_cast<uint4*>(sgen + B_SV + (ch >> 1) * 16384 + sw128_off(r, (ch & 1) * 4 + q)) = pack8(f8);
    }
  }
  __syncthreads();
  // ---------------- pass 2: dS = P (dP - D) / sqrt(hd), D = sum_k P dP (= <dO, O> minus the adapter part) ----------------
  {
    const float dxs = ((s_row[r] + s_row[128 + r]) + (s_row[256 + r] + s_row[384 + r])) * scale;
    const bool bias_any = row_biased && ch * 32 < bias_c1 && ch * 32 + 32 > bias_c0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float fd[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = q * 8 + e;
        // P as the UMMA sees it (h16), like the unfused formulation
        const float pb = h2f(f2h(pf[j]));
        const float ds = live ? pb * fmaf(__uint_as_float(w[j]), scale, -dxs) : 0.f;
        if (bias_any && ch * 32 + j >= bias_c0 && ch * 32 + j < bias_c1) g2_part += ds;
        fd[e] = ds;
      }
      *reinterpret_cast<uint4*>(sgen + B_SDS + (ch >> 1) * 16384 + sw128_off(r, (ch & 1) * 4 + q)) = pack8(fd);
    }
    g2_part *= 1.f / scale;                            // partial sums were taken on dS / sqrt(hd)
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t id_tt = idesc_h16(128, 128, 1, 1), id_q = idesc_h16(128, 128, 0, 1), id_at = idesc_h16(128, 16, 1, 1);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)      // dQ[row][d] = sum_keys dS[row][key] K[key][d]
      umma_h16_ss(tmem + T_DQ, kmaj(B_SDS, 16384, ks), mnmaj(B_SK, 16384, ks), id_q, ks > 0 ? 1u : 0u);
    umma_h16_ss(tmem + T_DQ, desc_nosw(sbase + B_SDSA, 128, 256), desc_mn_sw128(sbase + B_SKA, 2048), id_q, 1u);   // += dS_a Ka
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)      // dK[key][d] = sum_rows dS[row][key] Q[row][d]
      umma_h16_ss(tmem + T_DK, mnmaj(B_SDS, 16384, ks), mnmaj(B_SQ, 16384, ks), id_tt, ks > 0 ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)      // dV[key][d] = sum_rows P[row][key] dO[row][d]
      umma_h16_ss(tmem + T_DV, mnmaj(B_SV, 16384, ks), mnmaj(B_SDO, 16384, ks), id_tt, ks > 0 ? 1u : 0u);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {    // dKa^T[d][a] = sum_rows Q[row][d] dS_a[row][a];  dVa^T[d][a] = sum_rows dO[row][d] (tg P_a)[row][a]
      umma_h16_ss(tmem + T_DKA, mnmaj(B_SQ, 16384, ks), desc_nosw(sbase + B_SDSA + ks * 512, 256, 128), id_at, ks > 0 ? 1u : 0u);
      umma_h16_ss(tmem + T_DKA + 16, mnmaj(B_SDO, 16384, ks), desc_nosw(sbase + B_SPA + ks * 512, 256, 128), id_at, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar_m2);
  }
  __syncwarp();

  // gate partial sums of this (sequence, head) in a fixed order while the MMAs run
  g1_part = warp_sum(g1_part);
  g2_part = warp_sum(g2_part);
  if (lane == 0) { sred[warp] = g1_part; sred[16 + warp] = g2_part; }
  __syncthreads();
  if (tid == 0) {
    float* wsg = p.ws_gate + (static_cast<long>(n) * p.H + h) * 2;
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 16; ++w2) { a1 += sred[w2]; a2 += sred[16 + w2]; }
    wsg[0] = a1;
    wsg[1] = a2;
  }

  mbar_wait(bar_m2, 0);
  tc_fence_after();
  // ---------------- epilogue: TMEM -> (inverse RoPE) -> h16 -> swizzled smem -> TMA store ----------------
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {            // 0: dQ (row = query), 1: dK (row = key), 2: dV
    const uint32_t tcol = which == 0 ? T_DQ : (which == 1 ? T_DK : T_DV);
    const int sdst = which == 0 ? B_SQ : (which == 1 ? B_SK : B_SV);
    {
      uint32_t v[32];
      tmem_ld_32x32(tlane + tcol + static_cast<uint32_t>(ch * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
        if (which < 2) inv_rope8(f, sgen + B_ROPE, r, ch * 4 + q);      // rotate pair (2i, 2i+1) by -angle(pos = r, i)
        *reinterpret_cast<uint4*>(sgen + sdst + (ch >> 1) * 16384 + sw128_off(r, (ch & 1) * 4 + q)) = pack8(f);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm_dqkv, sbase + sdst, which * D + h * 128, 0, n);          // rows >= S out of bounds -> dropped
      tma_store_3d(&tm_dqkv, sbase + sdst + 16384, which * D + h * 128 + 64, 0, n);
      tma_store_commit();
    }
  }
  if (part == 0) {
    // adapter partials: thread = head-dim index d; columns [128,144) dKa^T, [144,160) dVa^T
    uint32_t v[32];
    tmem_ld_32x32(tlane + T_DKA, v);
    tmem_ld_wait();
    float* wsa = p.ws_akv + (static_cast<long>(n) * p.H + h) * 2 * AT_AP * 128;
#pragma unroll
    for (int a = 0; a < AT_AP; ++a) {
      wsa[a * 128 + r] = __uint_as_float(v[a]);
      wsa[AT_AP * 128 + a * 128 + r] = __uint_as_float(v[16 + a]);
    }
  }
  if (tid == 0) tma_store_wait_read();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_use_tc{1};   // test hook (fvqa_attn_debug_use_tc), process-wide

bool attn_tc_supported(int S, int hd, int A) { return g_use_tc != 0 && hd == 128 && S <= 128 && A <= AT_AP; }
bool attn_tcl_supported(int S, int hd, int A) { return g_use_tc != 0 && hd == 128 && S > 128 && A <= AT_AP; }

int attn_tc_init() {
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(attn_fwd_tc): %s", cudaGetErrorString(e));
  e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(attn_bwd_tc): %s", cudaGetErrorString(e));
  return FVQA_OK;
}

static int make_maps(const AttnParams& p, CUtensorMap* qkv, CUtensorMap* akv) {
  const int D = p.H * 128;
  int rc = get_tmap_seq(p.qkv, p.n_seq, p.S, 3 * D, 3 * D, 128, qkv);
  if (rc) return rc;
  return get_tmap(p.akv, p.A, 2 * D, p.akv_ld, 16, akv);     // rows >= A are out of bounds -> zero-filled
}

int attn_fwd_tc(const AttnParams& p, cudaStream_t stream) {
  CUtensorMap tq, ta, to;
  int rc = make_maps(p, &tq, &ta);
  if (rc) return rc;
  const int D = p.H * 128;
  rc = get_tmap_seq(p.out, p.n_seq, p.S, D, D, 128, &to);
  if (rc) return rc;
  launch_k(attn_fwd_tc_kernel, dim3(p.H, p.n_seq), dim3(FW_THREADS), F_SMEM, stream, tq, ta, to, p);
  return check_launch("attn_fwd_tc");
}

// Launches the per-(sequence, head) kernel only; the caller runs attn_bwd_reduce_kernel afterwards
// (workspace layout identical to the mma.sync path with qblocks = 1).
int attn_bwd_tc(const AttnParams& p, cudaStream_t stream) {
  CUtensorMap tq, ta, td, tg;
  int rc = make_maps(p, &tq, &ta);
  if (rc) return rc;
  const int D = p.H * 128;
  rc = get_tmap_seq(p.dout, p.n_seq, p.S, D, D, 128, &td);
  if (rc) return rc;
  rc = get_tmap_seq(p.dqkv, p.n_seq, p.S, 3 * D, 3 * D, 128, &tg);
  if (rc) return rc;
  launch_k(attn_bwd_tc_kernel, dim3(p.H, p.n_seq), dim3(BW_THREADS), B_SMEM, stream, tq, ta, td, tg, p);
  return check_launch("attn_bwd_tc");
}

}  // namespace fvqa

/* 1 if fvqa_attn_fwd / fvqa_attn_bwd take the tcgen05 path for this shape (bwd then launches 2 kernels, else 3). */
extern "C" int fvqa_attn_uses_tc(int S, int hd, int A) { return fvqa::attn_tc_supported(S, hd, A) ? 1 : 0; }

/* Test hook: 0 forces the mma.sync attention kernels, 1 (default) lets S <= 128 / hd = 128 shapes take the
 * tcgen05 path. Returns the previous setting. */
extern "C" int fvqa_attn_debug_use_tc(int on) {
  const int prev = fvqa::g_use_tc;
  fvqa::g_use_tc = on;
  return prev;
}
