// tcgen05 / TMEM attention for sequences longer than one 128-row tile (DramaQA-shaped S = 384, TVQA-shaped
// S = 650), head_dim = 128. Same operand layouts and per-row softmax code as attention_tc.cu, tiled:
//
//   forward   : CTA = (128-query tile i, head, sequence), 2 CTAs/SM. Two passes over the key tiles j <= i:
//               pass 1 UMMA S = Q K_j^T -> running row max / sum (no rescaling of O is ever needed);
//               pass 2 UMMA S again -> P = exp2(S - m) / l -> smem -> UMMA O += P V_j (TMEM accumulate).
//               QK^T twice is cheap (attention is latency / HBM-bound), K tiles come back from L2.
//   backward A: CTA = (query tile i, ...): D = <dO, O> from TMA-staged tiles, adapter softmax backward,
//               loop j <= i: S, dP -> dS -> UMMA dQ += dS K_j; writes dQ (inverse RoPE), D, gate partials and the
//               per-tile adapter partials dKa^T, dVa^T.
//   backward B: CTA = (key tile j, ...): loop i >= j: S, dP -> P, dS -> UMMA dV += P^T dO_i, dK += dS^T Q_i
//               (TMEM accumulators live across the loop); writes dK (inverse RoPE), dV.
// Two deterministic passes, no atomics; partials are reduced in a fixed order by attn_bwd_reduce_kernel.
#include "attention_tc.cuh"
#include "tmap.h"

namespace fvqa {

namespace {

// masked / biased log2-domain score of element (row_g, col_g)
struct ScoreCtx {
  float scale2, bias2;
  int bias_row0, bias_c0, bias_c1;   // rows >= bias_row0, cols in [c0, c1) get bias2
};
__device__ __forceinline__ ScoreCtx make_score_ctx(const AttnParams& p, int n, int h) {
  ScoreCtx c;
  c.scale2 = rsqrtf(128.f) * TC_LOG2E;
  const int vs = p.vstart[n];
  c.bias2 = (vs >= 0) ? p.gate2[h] * TC_LOG2E : 0.f;
  c.bias_row0 = (vs >= 0) ? vs + p.F : 0x7fffffff;
  c.bias_c0 = vs;
  c.bias_c1 = vs + p.F;
  return c;
}

// ---- forward smem map: identical to the S <= 128 kernel ----
constexpr int LF_SQ = 0, LF_SK = 32768, LF_SV = 65536, LF_SKA = 98304, LF_SVA = 102400, LF_SPA = 106496, LF_BAR = 110592;
constexpr int LF_SMEM = LF_BAR + 64 + 1024;

}  // namespace

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_fwd_tcl_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_akv,
                    const __grid_constant__ CUtensorMap tm_out, const AttnParams p) {
  // Single pass over the key tiles j <= i with an online softmax: S = Q K_j^T (UMMA) -> per-row p = exp2(s - m_ref)
  // -> P (h16) written back into TMEM over S -> O += P V_j (TS-mode UMMA, accumulator in TMEM). The reference maximum
  // m_ref only moves when a tile's maximum exceeds it by more than 2^8 (then O and l are rescaled through TMEM), so
  // rescaling is rare; the final O / l and the adapter term (tanh(g1) softmax_a, pre-multiplied by l) end the row.
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_q = sbase + LF_BAR, bar_k = bar_q + 8, bar_v = bar_q + 16, bar_s = bar_q + 24, bar_o = bar_q + 32, holder = bar_q + 40;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int qi = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);   // longest key loops are scheduled first
  const int h = blockIdx.y, n = blockIdx.z;
  const int S = p.S, D = p.H * 128;
  const int row0 = qi * 128, row_g = row0 + tid;

  if (tid == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(holder, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + LF_BAR + 40);
  const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int c = h * 128;
  constexpr uint32_t id_s = idesc_h16(128, 128, 0, 0), id_a = idesc_h16(128, 16, 0, 0), id_o = idesc_h16(128, 128, 0, 1);
  constexpr uint32_t T_S = 0, T_O = 128, T_SA = 128;      // S_a sits in O's columns until the first P.V overwrites them
  auto kdesc = [&](int off, int blk, int ks) { return umma_desc_k_sw128(sbase + off + (ks >> 2) * blk) + static_cast<uint64_t>(2 * (ks & 3)); };
  auto load_k = [&](int j) {
    mbar_arrive_expect_tx(bar_k, 32768);
    tma_load_3d(sbase + LF_SK, &tm_qkv, bar_k, D + c, j * 128, n);
    tma_load_3d(sbase + LF_SK + 16384, &tm_qkv, bar_k, D + c + 64, j * 128, n);
  };
  auto load_v = [&](int j) {
    mbar_arrive_expect_tx(bar_v, 32768);
    tma_load_3d(sbase + LF_SV, &tm_qkv, bar_v, 2 * D + c, j * 128, n);
    tma_load_3d(sbase + LF_SV + 16384, &tm_qkv, bar_v, 2 * D + c + 64, j * 128, n);
  };
  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_akv); tma_prefetch_desc(&tm_out);
    mbar_arrive_expect_tx(bar_q, 32768 + 8192);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + LF_SQ + kb * 16384, &tm_qkv, bar_q, c + kb * 64, row0, n);
      tma_load_2d(sbase + LF_SKA + kb * 2048, &tm_akv, bar_q, c + kb * 64, 0);
      tma_load_2d(sbase + LF_SVA + kb * 2048, &tm_akv, bar_q, D + c + kb * 64, 0);
    }
    load_k(0);
    load_v(0);
  }
  const ScoreCtx sc = make_score_ctx(p, n, h);
  const bool row_biased = row_g >= sc.bias_row0;
  uint32_t ph_k = 0, ph_v = 0, ph_s = 0, ph_o = 0;
  float m_ref = -INFINITY, l_run = 0.f;
  float pa_n[16];                                        // tanh(g1) * softmax over the adapter keys (fp32, row-local)

  for (int j = 0; j <= qi; ++j) {
    if (tid == 0) {
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(bar_k, ph_k);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        umma_h16_ss(tmem + T_S, kdesc(LF_SQ, 16384, ks), kdesc(LF_SK, 16384, ks), id_s, ks > 0 ? 1u : 0u);
        if (j == 0) umma_h16_ss(tmem + T_SA, kdesc(LF_SQ, 16384, ks), kdesc(LF_SKA, 2048, ks), id_a, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_s, ph_s);
    tc_fence_after();
    if (tid == 0 && j < qi) load_k(j + 1);                // K's buffer is free as soon as S is complete
    if (j == 0) {
      uint32_t v[32];
      tmem_ld_32x16(tlane + T_SA, v);
      tmem_ld_wait();
      const float tg = tanhf(p.gate1[h]);
      float ma = -INFINITY, la = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        pa_n[e] = (e < p.A) ? __uint_as_float(v[e]) * sc.scale2 : -INFINITY;
        ma = fmaxf(ma, pa_n[e]);
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) { pa_n[e] = exp2f(pa_n[e] - ma); la += pa_n[e]; }
      const float ia = tg / la;
#pragma unroll
      for (int e = 0; e < 16; ++e) pa_n[e] *= ia;
    }
    // scores of this tile (diagonal tile: 32-column chunks beyond the warp's last row are fully masked)
    const int nch = (j == qi) ? warp + 1 : 4;
    float x[128];
    float mx = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch < nch) {
        uint32_t v[32];
        tmem_ld_32x32(tlane + T_S + static_cast<uint32_t>(ch * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col_g = j * 128 + ch * 32 + e;
          float t = __uint_as_float(v[e]) * sc.scale2;
          if (row_biased && col_g >= sc.bias_c0 && col_g < sc.bias_c1) t += sc.bias2;
          if (col_g > row_g) t = -INFINITY;
          x[ch * 32 + e] = t;
          mx = fmaxf(mx, t);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) x[ch * 32 + e] = -INFINITY;
      }
    }
    // lazy rescale of the running state: only when this tile's maximum beats the reference by > 2^8 (warp-uniform
    // decision: tcgen05.ld/st are warp-collective); the previous P.V has retired (S_j was issued after it completed)
    const bool bump = mx > m_ref + 8.f;
    if (__any_sync(0xffffffffu, bump)) {
      const float m_new = bump ? mx : m_ref;
      const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);
      if (j > 0) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tlane + T_O + static_cast<uint32_t>(ch * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * f);
          tmem_st_32x32(tlane + T_O + static_cast<uint32_t>(ch * 32), v);
        }
      }
      l_run *= f;
      m_ref = m_new;
    }
    // P = exp2(x - m_ref) (h16, two keys per column) -> TMEM columns [0,64) over the consumed S
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t pk[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float p0 = exp2f(x[hf * 64 + 2 * e] - m_ref), p1 = exp2f(x[hf * 64 + 2 * e + 1] - m_ref);
        l_run += p0 + p1;
        pk[e] = pack_h16x2(p0, p1);
      }
      tmem_st_32x32(tlane + T_S + static_cast<uint32_t>(hf * 32), pk);
    }
    if (j == qi) {
      // adapter probabilities pre-multiplied by l so that the final O / l leaves tanh(g1) softmax_a . Va
      uint32_t pa[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) pa[e] = pack_h16x2(pa_n[2 * e] * l_run, pa_n[2 * e + 1] * l_run);
      tmem_st_32x8(tlane + T_S + 64u, pa);
    }
    tmem_st_wait();
    ph_k ^= 1u; ph_s ^= 1u;
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, ph_v);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_h16_ts(tmem + T_O, tmem + T_S + static_cast<uint32_t>(ks * 8), desc_mn_sw128(sbase + LF_SV + ks * 2048, 16384), id_o,
                     (j > 0 || ks > 0) ? 1u : 0u);
      if (j == qi) umma_h16_ts(tmem + T_O, tmem + T_S + 64u, desc_mn_sw128(sbase + LF_SVA, 2048), id_o, 1u);
      umma_commit(bar_o);
      // the next S overwrites P's columns and the next V load overwrites V: both wait for this P.V
      mbar_wait(bar_o, ph_o);
      if (j < qi) load_v(j + 1);
    }
    __syncwarp();
    ph_v ^= 1u; ph_o ^= 1u;
  }
  // every thread: the last P.V (completion #qi of bar_o)
  mbar_wait(bar_o, static_cast<uint32_t>(qi) & 1u);
  tc_fence_after();
  const float inv_l = 1.f / l_run;
  if (row_g < S) p.lse[(static_cast<long>(n) * p.H + h) * S + row_g] = (m_ref + log2f(l_run)) * TC_LN2;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(tlane + T_O + static_cast<uint32_t>(ch * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) * inv_l;
      *reinterpret_cast<uint4*>(sgen + LF_SQ + (ch >> 1) * 16384 + sw128_off(tid, (ch & 1) * 4 + q)) = pack8(f);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tma_store_3d(&tm_out, sbase + LF_SQ, c, row0, n);
    tma_store_3d(&tm_out, sbase + LF_SQ + 16384, c + 64, row0, n);
    tma_store_commit();
    tma_store_wait_read();
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// ---------------------------------------------------------------------------------------------
// backward A: query tile owner -> dQ, D, gate partials, adapter partials
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int A_SQ = 0, A_SDO = 32768, A_SK = 65536, A_SV = 98304, A_SDS = 131072;   // A_SDS first holds the O tile
constexpr int A_SKA = 163840, A_SVA = 167936, A_SPA = 172032, A_SDSA = 176128, A_ROPE = 180224, A_BAR = 212992;
constexpr int A_RED = A_BAR + 64;                     // 32 floats: per-warp gate partials
constexpr int A_ROW = A_BAR + 256;                    // [4][128] floats of per-row scratch
constexpr int A_SMEM = A_ROW + 2048 + 1024;
constexpr int LQ_THREADS = 512;
}  // namespace

// 512 threads: the four warps w, w+4, w+8, w+12 share TMEM lane quadrant w & 3 (query rows 32 (w & 3) ..) and take one
// 32-column chunk each (part = tid / 128); with one warp per scheduler the row-wise math is pure exposed latency.
__global__ void __launch_bounds__(LQ_THREADS, 1)
attn_bwd_tcl_dq_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_akv,
                       const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_o,
                       const __grid_constant__ CUtensorMap tm_dqkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_q = sbase + A_BAR, bar_kv0 = bar_q + 8, bar_m1 = bar_q + 16, bar_m2 = bar_q + 24, bar_ma = bar_q + 32, holder = bar_q + 40,
                 bar_kv1 = bar_q + 48;
  float* sred = reinterpret_cast<float*>(sgen + A_RED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = tid & 127, part = tid >> 7, quad = warp & 3;
  const int qi = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);   // longest key loops are scheduled first
  const int h = blockIdx.y, n = blockIdx.z;
  const int S = p.S, D = p.H * 128;
  const int row0 = qi * 128, row_g = row0 + r;
  const bool row_ok = row_g < S;

  if (tid == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_kv0, 1); mbar_init(bar_kv1, 1); mbar_init(bar_m1, 1); mbar_init(bar_m2, 1); mbar_init(bar_ma, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + A_BAR + 40);
  const uint32_t tlane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256, T_SA = 384, T_DPA = 400, T_DKA = 416;   // dVa^T at T_DKA + 16
  const int c = h * 128;
  auto kdesc = [&](int off, int blk, int ks) { return umma_desc_k_sw128(sbase + off + (ks >> 2) * blk) + static_cast<uint64_t>(2 * (ks & 3)); };
  auto mndesc = [&](int off, int lbo, int ks) { return desc_mn_sw128(sbase + off + ks * 2048, lbo); };
  constexpr uint32_t id_s = idesc_h16(128, 128, 0, 0), id_a = idesc_h16(128, 16, 0, 0), id_q = idesc_h16(128, 128, 0, 1),
                     id_at = idesc_h16(128, 16, 1, 1);

  // K/V tile buffers: 0 = (A_SK, A_SV); 1 = (A_SDS once the O tile has been consumed, A_ROPE)
  auto koff = [&](int buf) { return buf ? A_SDS : A_SK; };
  auto voff = [&](int buf) { return buf ? A_ROPE : A_SV; };
  auto load_kv = [&](int j, int buf) {                  // tid 0 only
    const uint32_t bar = buf ? bar_kv1 : bar_kv0;
    mbar_arrive_expect_tx(bar, 65536);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + koff(buf) + kb * 16384, &tm_qkv, bar, D + c + kb * 64, j * 128, n);
      tma_load_3d(sbase + voff(buf) + kb * 16384, &tm_qkv, bar, 2 * D + c + kb * 64, j * 128, n);
    }
  };
  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_akv); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_o); tma_prefetch_desc(&tm_dqkv);
    mbar_arrive_expect_tx(bar_q, 3 * 32768 + 8192);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + A_SQ + kb * 16384, &tm_qkv, bar_q, c + kb * 64, row0, n);
      tma_load_3d(sbase + A_SDO + kb * 16384, &tm_do, bar_q, c + kb * 64, row0, n);
      tma_load_3d(sbase + A_SDS + kb * 16384, &tm_o, bar_q, c + kb * 64, row0, n);
      tma_load_2d(sbase + A_SKA + kb * 2048, &tm_akv, bar_q, c + kb * 64, 0);
      tma_load_2d(sbase + A_SVA + kb * 2048, &tm_akv, bar_q, D + c + kb * 64, 0);
    }
    load_kv(0, 0);
  }
  __syncwarp();
  if (tid == 0) {
    mbar_wait(bar_q, 0);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      umma_h16_ss(tmem + T_SA, kdesc(A_SQ, 16384, ks), kdesc(A_SKA, 2048, ks), id_a, ks > 0 ? 1u : 0u);
      umma_h16_ss(tmem + T_DPA, kdesc(A_SDO, 16384, ks), kdesc(A_SVA, 2048, ks), id_a, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar_ma);
  }
  __syncwarp();
  mbar_wait(bar_q, 0);
  // D_total = <dO, O> of this row from the staged tiles: each of the 4 parts sums 4 of the 16 chunks
  float* s_row = reinterpret_cast<float*>(sgen + A_ROW);          // [4][128] partial D | later [128] -lse, [128] D / sqrt(hd)
  {
    float dpart = 0.f;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c16 = part * 4 + cc;
      float a[8], b[8];
      const uint32_t off = static_cast<uint32_t>((c16 >> 3) * 16384) + sw128_off(r, c16 & 7);
      unpack8(*reinterpret_cast<const uint4*>(sgen + A_SDS + off), a);
      unpack8(*reinterpret_cast<const uint4*>(sgen + A_SDO + off), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) dpart += a[e] * b[e];
    }
    s_row[part * 128 + r] = dpart;
  }
  __syncthreads();
  const ScoreCtx sc = make_score_ctx(p, n, h);
  const bool row_biased = row_g >= sc.bias_row0;
  const float scale = rsqrtf(128.f);
  float g1_part = 0.f, g2_part = 0.f;
  mbar_wait(bar_ma, 0);
  tc_fence_after();
  float nlse_own = 0.f, dxs_own = 0.f;
  if (part == 0) {                                        // adapter branch + row constants: one thread per row
    const float dtot = (s_row[r] + s_row[128 + r]) + (s_row[256 + r] + s_row[384 + r]);
    const float tg = tanhf(p.gate1[h]);
    const float lse2 = row_ok ? p.lse[(static_cast<long>(n) * p.H + h) * S + row_g] * TC_LOG2E : 0.f;
    uint32_t v[32], w[32];
    tmem_ld_32x16(tlane + T_SA, v);
    tmem_ld_32x16(tlane + T_DPA, w);
    tmem_ld_wait();
    float sa[16], ma = -INFINITY, la = 0.f, da = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      sa[e] = (e < p.A) ? __uint_as_float(v[e]) * sc.scale2 : -INFINITY;
      ma = fmaxf(ma, sa[e]);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) { sa[e] = exp2f(sa[e] - ma); la += sa[e]; }
    const float ia = row_ok ? 1.f / la : 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { sa[e] *= ia; da += sa[e] * __uint_as_float(w[e]); }
    g1_part = da;
    const float dx = dtot - tg * da;
    if (row_ok) p.ws_dx[(static_cast<long>(n) * p.H + h) * (p.qblocks * 128) + row_g] = dx;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      float fp[8], fd[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float pa = sa[cc * 8 + e];
        fp[e] = tg * pa;
        fd[e] = tg * pa * (__uint_as_float(w[cc * 8 + e]) - da) * scale;
      }
      const int off = (r >> 3) * 256 + cc * 128 + (r & 7) * 16;
      *reinterpret_cast<uint4*>(sgen + A_SPA + off) = pack8(fp);
      *reinterpret_cast<uint4*>(sgen + A_SDSA + off) = pack8(fd);
    }
    nlse_own = row_ok ? -lse2 : -1e30f;                   // p = exp2(s c + nlse); rows past the sequence give p = 0
    dxs_own = dx * scale;                                 // ds = p (dp / sqrt(hd) - dxs)
  }
  __syncthreads();                                        // partial D consumed
  if (part == 0) { s_row[r] = nlse_own; s_row[128 + r] = dxs_own; }
  tc_fence_before();
  __syncthreads();                                       // O tile fully consumed: its buffer becomes dS
  tc_fence_after();

  const float nlse = s_row[r], dxs = s_row[128 + r];      // (written before the barrier above)
  uint32_t ph_kv[2] = {0, 0}, ph_m1 = 0, ph_m2 = 0;
  for (int j = 0; j <= qi; ++j) {
    const int buf = j & 1;
    if (tid == 0) {
      if (j > 0) mbar_wait(bar_m2, ph_m2 ^ 1u);          // dQ UMMA of tile j-1 retired: its K buffer and the dS columns are free
      if (j < qi) load_kv(j + 1, buf ^ 1);                // prefetch the next key tile
      mbar_wait(buf ? bar_kv1 : bar_kv0, ph_kv[buf]);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        umma_h16_ss(tmem + T_S, kdesc(A_SQ, 16384, ks), kdesc(koff(buf), 16384, ks), id_s, ks > 0 ? 1u : 0u);
        umma_h16_ss(tmem + T_DP, kdesc(A_SDO, 16384, ks), kdesc(voff(buf), 16384, ks), id_s, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_m1);
    }
    __syncwarp();
    mbar_wait(bar_m1, ph_m1);
    tc_fence_after();
    // dS (h16, two keys per 32-bit column) is written back IN PLACE over S: this thread owns chunk `part` (fp32 columns
    // [32 part, +32) -> packed columns [16 part, +16)); all four parts read before anyone writes (barrier in between)
    const int ch = part;
    const bool live = (j < qi) || ch <= quad;              // diagonal tile: chunks beyond the warp's last row are masked
    uint32_t v[32], w[32];
    if (live) {
      tmem_ld_32x32(tlane + T_S + static_cast<uint32_t>(ch * 32), v);
      tmem_ld_32x32(tlane + T_DP + static_cast<uint32_t>(ch * 32), w);
      tmem_ld_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    {
      uint32_t dd[16];
      if (live) {
        const int col0_g = j * 128 + ch * 32;
        const bool causal = (j == qi) && ch == quad;      // only the chunk on the diagonal needs the key > row test
        const bool bias_any = row_biased && col0_g < sc.bias_c1 && col0_g + 32 > sc.bias_c0;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float ds[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col_g = col0_g + e + u;
            float t = fmaf(__uint_as_float(v[e + u]), sc.scale2, nlse);
            const bool biased = bias_any && col_g >= sc.bias_c0 && col_g < sc.bias_c1;
            if (biased) t += sc.bias2;
            float pe = exp2f(t);
            if (causal && col_g > row_g) pe = 0.f;
            const float dse = pe * fmaf(__uint_as_float(w[e + u]), scale, -dxs);     // dS / sqrt(hd)
            if (biased) g2_part += dse;
            ds[u] = dse;
          }
          dd[e >> 1] = pack_h16x2(ds[0], ds[1]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) dd[e] = 0u;
      }
      tmem_st_32x16(tlane + T_S + static_cast<uint32_t>(ch * 16), dd);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)                      // dQ[row][d] += sum_keys dS[row][key] K_j[key][d], A = dS from TMEM
        umma_h16_ts(tmem + T_DQ, tmem + T_S + static_cast<uint32_t>(ks * 8), mndesc(koff(buf), 16384, ks), id_q, (j > 0 || ks > 0) ? 1u : 0u);
      if (j == qi) {
        umma_h16_ss(tmem + T_DQ, desc_nosw(sbase + A_SDSA, 128, 256), desc_mn_sw128(sbase + A_SKA, 2048), id_q, 1u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma_h16_ss(tmem + T_DKA, mndesc(A_SQ, 16384, ks), desc_nosw(sbase + A_SDSA + ks * 512, 256, 128), id_at, ks > 0 ? 1u : 0u);
          umma_h16_ss(tmem + T_DKA + 16, mndesc(A_SDO, 16384, ks), desc_nosw(sbase + A_SPA + ks * 512, 256, 128), id_at, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(bar_m2);
    }
    __syncwarp();
    ph_kv[buf] ^= 1u; ph_m1 ^= 1u; ph_m2 ^= 1u;
  }
  g2_part *= 1.f / scale;                                 // the partial sums were taken on dS / sqrt(hd)
  g1_part = warp_sum(g1_part);
  g2_part = warp_sum(g2_part);
  if (lane == 0) { sred[warp] = g1_part; sred[16 + warp] = g2_part; }
  __syncthreads();
  if (tid == 0) {
    float* wsg = p.ws_gate + ((static_cast<long>(n) * p.H + h) * p.qblocks + qi) * 2;
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 16; ++w2) { a1 += sred[w2]; a2 += sred[16 + w2]; }     // fixed order
    wsg[0] = a1;
    wsg[1] = a2;
  }
  mbar_wait(bar_m2, ph_m2 ^ 1u);                         // last completion: every buffer is free
  tc_fence_after();
  stage_rope_table(sgen + A_SV, p.cosT, p.sinT, row0, S, tid, LQ_THREADS);
  __syncthreads();
  {
    const int ch = part;
    uint32_t v[32];
    tmem_ld_32x32(tlane + T_DQ + static_cast<uint32_t>(ch * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
      inv_rope8(f, sgen + A_SV, r, ch * 4 + q);
      *reinterpret_cast<uint4*>(sgen + A_SK + (ch >> 1) * 16384 + sw128_off(r, (ch & 1) * 4 + q)) = pack8(f);
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tma_store_3d(&tm_dqkv, sbase + A_SK, c, row0, n);
    tma_store_3d(&tm_dqkv, sbase + A_SK + 16384, c + 64, row0, n);
    tma_store_commit();
  }
  if (part == 0) {
    uint32_t v[32];
    tmem_ld_32x32(tlane + T_DKA, v);                     // [416,432) dKa^T, [432,448) dVa^T; thread = head-dim index
    tmem_ld_wait();
    float* wsa = p.ws_akv + ((static_cast<long>(n) * p.qblocks + qi) * p.H + h) * 2 * AT_AP * 128;
#pragma unroll
    for (int a = 0; a < AT_AP; ++a) {
      wsa[a * 128 + r] = __uint_as_float(v[a]);
      wsa[AT_AP * 128 + a * 128 + r] = __uint_as_float(v[16 + a]);
    }
  }
  if (tid == 0) tma_store_wait_read();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward B: key tile owner -> dK, dV. Transposed formulation: S^T = K_j Q_i^T and dP^T = V_j dO_i^T put the KEYS on
// the TMEM lanes, so P^T and dS^T are written back into TMEM (h16, two rows per 32-bit column, in place over S^T /
// dP^T) and feed dV += P^T dO_i, dK += dS^T Q_i as TMEM-resident A operands (TS-mode UMMA): no shared-memory round
// trip, and the 64 KB that P / dS occupied become a second Q / dO buffer -> the TMA loads of tile i+1 overlap tile i.
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int K_SK = 0, K_SV = 32768;                 // resident key tile
constexpr int K_BUF0 = 65536;                         // [2] x { Q_i 32 KB | dO_i 32 KB }; buffer 0 doubles as RoPE table / dK staging at the end
constexpr int K_BUFSZ = 65536;
constexpr int K_LSE = 196608;                         // [2][128] lse (log2 domain) | [2][128] D of the query tile
constexpr int K_BAR = K_LSE + 2048;
constexpr int K_SMEM = K_BAR + 64 + 1024;
constexpr int LB_THREADS = 256;
}  // namespace

// 256 threads: warps w and w + 4 share TMEM lane quadrant w & 3 (keys 32 (w & 3) ..) and split the 128 columns (query rows).
__global__ void __launch_bounds__(LB_THREADS, 1)
attn_bwd_tcl_dkv_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_dqkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar_kv = sbase + K_BAR, bar_q0 = bar_kv + 8, bar_q1 = bar_kv + 16, bar_m1 = bar_kv + 24, bar_m2 = bar_kv + 32, holder = bar_kv + 40;
  float* s_lse = reinterpret_cast<float*>(sgen + K_LSE);          // [2][128]
  float* s_dx = s_lse + 256;                                       // [2][128]
  const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7, quad = warp & 3;
  const int kj = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int S = p.S, D = p.H * 128;
  const int key0 = kj * 128, key_g = key0 + r;
  const int qtiles = p.qblocks;

  if (tid == 0) {
    mbar_init(bar_kv, 1); mbar_init(bar_q0, 1); mbar_init(bar_q1, 1); mbar_init(bar_m1, 1); mbar_init(bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(holder, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + K_BAR + 40);
  const uint32_t tlane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  constexpr uint32_t T_ST = 0, T_DPT = 128, T_DV = 256, T_DK = 384;
  const int c = h * 128;
  auto kdesc = [&](int off, int ks) { return umma_desc_k_sw128(sbase + off + (ks >> 2) * 16384) + static_cast<uint64_t>(2 * (ks & 3)); };
  auto mndesc = [&](int off, int ks) { return desc_mn_sw128(sbase + off + ks * 2048, 16384); };
  constexpr uint32_t id_s = idesc_h16(128, 128, 0, 0), id_t = idesc_h16(128, 128, 0, 1);
  const long nh = static_cast<long>(n) * p.H + h;

  auto load_q_tile = [&](int qi, int buf) {            // tid 0 only
    const uint32_t bar = buf ? bar_q1 : bar_q0;
    const int off = K_BUF0 + buf * K_BUFSZ;
    mbar_arrive_expect_tx(bar, 65536);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + off + kb * 16384, &tm_qkv, bar, c + kb * 64, qi * 128, n);
      tma_load_3d(sbase + off + 32768 + kb * 16384, &tm_do, bar, c + kb * 64, qi * 128, n);
    }
  };
  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_dqkv);
    mbar_arrive_expect_tx(bar_kv, 65536);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_3d(sbase + K_SK + kb * 16384, &tm_qkv, bar_kv, D + c + kb * 64, key0, n);
      tma_load_3d(sbase + K_SV + kb * 16384, &tm_qkv, bar_kv, 2 * D + c + kb * 64, key0, n);
    }
    load_q_tile(kj, 0);
  }
  const ScoreCtx sc = make_score_ctx(p, n, h);
  const float scale = rsqrtf(128.f);
  const bool key_biased = key_g >= sc.bias_c0 && key_g < sc.bias_c1;
  uint32_t ph_q[2] = {0, 0}, ph_m1 = 0, ph_m2 = 0;
  int it = 0;
  for (int qi = kj; qi < qtiles; ++qi, ++it) {
    const int buf = it & 1;
    const int qoff = K_BUF0 + buf * K_BUFSZ, dooff = qoff + 32768;
    // lse / D of this query tile -> smem (threads 0..127: lse, 128..255: D)
    {
      const int row_g = qi * 128 + r;
      // -lse (log2 domain; +inf-like for rows past the sequence so that P = 0 there) and D / sqrt(hd)
      float v = half ? 0.f : -1e30f;
      if (row_g < S) v = half ? p.ws_dx[nh * (qtiles * 128) + row_g] * scale : -p.lse[nh * S + row_g] * TC_LOG2E;
      (half ? s_dx : s_lse)[buf * 128 + r] = v;
    }
    if (tid == 0) {
      if (it > 0) mbar_wait(bar_m2, ph_m2 ^ 1u);         // UMMAs of tile it-1 are done with the other Q / dO buffer
      if (qi + 1 < qtiles) load_q_tile(qi + 1, buf ^ 1);  // prefetch the next query tile
      if (it == 0) mbar_wait(bar_kv, 0);
      mbar_wait(buf ? bar_q1 : bar_q0, ph_q[buf]);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        umma_h16_ss(tmem + T_ST, kdesc(K_SK, ks), kdesc(qoff, ks), id_s, ks > 0 ? 1u : 0u);      // S^T  = K_j Q_i^T
        umma_h16_ss(tmem + T_DPT, kdesc(K_SV, ks), kdesc(dooff, ks), id_s, ks > 0 ? 1u : 0u);    // dP^T = V_j dO_i^T
      }
      umma_commit(bar_m1);
    }
    __syncthreads();                                      // lse / D staged; also orders tid 0's issue before the waits below
    mbar_wait(bar_m1, ph_m1);
    tc_fence_after();
    // this thread: key key_g (lane), query rows [64 half, 64 half + 64) of the tile (columns)
    uint32_t vs_[64], vd_[64];
    const bool diag = (qi == kj);
    // on the diagonal tile rows < 32 quad are all masked for this warp's keys: 32-row chunks c < quad are zero
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int ch = 2 * half + cc;
      uint32_t a[32], b[32];
      tmem_ld_32x32(tlane + T_ST + static_cast<uint32_t>(ch * 32), a);
      tmem_ld_32x32(tlane + T_DPT + static_cast<uint32_t>(ch * 32), b);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) { vs_[cc * 32 + e] = a[e]; vd_[cc * 32 + e] = b[e]; }
    }
    tc_fence_before();
    __syncthreads();                                      // both column halves of every lane are in registers: in-place overwrite is safe
    tc_fence_after();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int ch = 2 * half + cc;
      uint32_t pp[16], dd[16];
      if (!diag || ch >= quad) {
        // gate2 bias: this key is a video column and the row is past the video block (uniform per 32-row chunk
        // except for the one chunk that contains bias_row0)
        const int row0_g = qi * 128 + ch * 32;
        const float badd_all = (key_biased && row0_g >= sc.bias_row0) ? sc.bias2 : 0.f;
        const bool bias_mixed = key_biased && row0_g < sc.bias_row0 && row0_g + 32 > sc.bias_row0;
        const bool causal = diag && ch == quad;           // only the chunk on the diagonal needs the key > row test
        const float4* l4 = reinterpret_cast<const float4*>(s_lse + buf * 128 + ch * 32);
        const float4* d4 = reinterpret_cast<const float4*>(s_dx + buf * 128 + ch * 32);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 nl = l4[q4], dxs = d4[q4];          // broadcast reads: 4 rows at a time
          const float nlv[4] = {nl.x, nl.y, nl.z, nl.w}, dxv[4] = {dxs.x, dxs.y, dxs.z, dxs.w};
          float pv[4], ds[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = q4 * 4 + u;
            float t = fmaf(__uint_as_float(vs_[cc * 32 + e]), sc.scale2, nlv[u] + badd_all);     // s * c - lse (+ bias)
            if (bias_mixed && row0_g + e >= sc.bias_row0) t += sc.bias2;
            float pe = exp2f(t);
            if (causal && key_g > row0_g + e) pe = 0.f;
            pv[u] = pe;
            ds[u] = pe * fmaf(__uint_as_float(vd_[cc * 32 + e]), scale, -dxv[u]);                  // P (dP - D) / sqrt(hd)
          }
          pp[q4 * 2] = pack_h16x2(pv[0], pv[1]); pp[q4 * 2 + 1] = pack_h16x2(pv[2], pv[3]);
          dd[q4 * 2] = pack_h16x2(ds[0], ds[1]); dd[q4 * 2 + 1] = pack_h16x2(ds[2], ds[3]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) { pp[e] = 0u; dd[e] = 0u; }
      }
      // rows [32 ch, 32 ch + 32) -> packed columns [16 ch, 16 ch + 16)
      tmem_st_32x16(tlane + T_ST + static_cast<uint32_t>(ch * 16), pp);
      tmem_st_32x16(tlane + T_DPT + static_cast<uint32_t>(ch * 16), dd);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t accum = it > 0 ? 1u : 0u;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)                      // dV[key][d] += sum_rows P^T[key][row] dO[row][d]
        umma_h16_ts(tmem + T_DV, tmem + T_ST + static_cast<uint32_t>(ks * 8), mndesc(dooff, ks), id_t, (accum || ks > 0) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)                      // dK[key][d] += sum_rows dS^T[key][row] Q[row][d]
        umma_h16_ts(tmem + T_DK, tmem + T_DPT + static_cast<uint32_t>(ks * 8), mndesc(qoff, ks), id_t, (accum || ks > 0) ? 1u : 0u);
      umma_commit(bar_m2);
    }
    __syncwarp();
    ph_q[buf] ^= 1u; ph_m1 ^= 1u; ph_m2 ^= 1u;
  }
  mbar_wait(bar_m2, ph_m2 ^ 1u);
  tc_fence_after();
  // all shared memory is free now: RoPE table of the key positions -> buffer 1 (buffer 0 stages dK / dV)
  uint8_t* s_rope = sgen + K_BUF0 + K_BUFSZ;
  stage_rope_table(s_rope, p.cosT, p.sinT, key0, S, tid, LB_THREADS);
  __syncthreads();
#pragma unroll 1
  for (int which = 1; which < 3; ++which) {              // 1: dK (inverse RoPE at the key positions), 2: dV
    const uint32_t tcol = which == 1 ? T_DK : T_DV;
    const int sdst = K_BUF0 + (which == 1 ? 0 : 32768);
#pragma unroll
    for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(tlane + tcol + static_cast<uint32_t>(ch * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
        if (which == 1) inv_rope8(f, s_rope, r, ch * 4 + q);
        *reinterpret_cast<uint4*>(sgen + sdst + (ch >> 1) * 16384 + sw128_off(r, (ch & 1) * 4 + q)) = pack8(f);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tm_dqkv, sbase + sdst, which * D + c, key0, n);
      tma_store_3d(&tm_dqkv, sbase + sdst + 16384, which * D + c + 64, key0, n);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_read();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
int attn_tcl_init() {
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_tcl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LF_SMEM);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(attn_fwd_tcl): %s", cudaGetErrorString(e));
  e = cudaFuncSetAttribute(attn_bwd_tcl_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(attn_bwd_tcl_dq): %s", cudaGetErrorString(e));
  e = cudaFuncSetAttribute(attn_bwd_tcl_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K_SMEM);
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(attn_bwd_tcl_dkv): %s", cudaGetErrorString(e));
  return FVQA_OK;
}

int attn_fwd_tcl(const AttnParams& p, cudaStream_t stream) {
  const int D = p.H * 128;
  CUtensorMap tq, ta, to;
  int rc = get_tmap_seq(p.qkv, p.n_seq, p.S, 3 * D, 3 * D, 128, &tq);
  if (rc) return rc;
  rc = get_tmap(p.akv, p.A, 2 * D, p.akv_ld, 16, &ta);
  if (rc) return rc;
  rc = get_tmap_seq(p.out, p.n_seq, p.S, D, D, 128, &to);
  if (rc) return rc;
  attn_fwd_tcl_kernel<<<dim3(p.qblocks, p.H, p.n_seq), TC_THREADS, LF_SMEM, stream>>>(tq, ta, to, p);
  return check_launch("attn_fwd_tcl");
}

int attn_bwd_tcl(const AttnParams& p, cudaStream_t stream) {
  const int D = p.H * 128;
  CUtensorMap tq, ta, td, to, tg;
  int rc = get_tmap_seq(p.qkv, p.n_seq, p.S, 3 * D, 3 * D, 128, &tq);
  if (rc) return rc;
  rc = get_tmap(p.akv, p.A, 2 * D, p.akv_ld, 16, &ta);
  if (rc) return rc;
  rc = get_tmap_seq(p.dout, p.n_seq, p.S, D, D, 128, &td);
  if (rc) return rc;
  rc = get_tmap_seq(p.out, p.n_seq, p.S, D, D, 128, &to);
  if (rc) return rc;
  rc = get_tmap_seq(p.dqkv, p.n_seq, p.S, 3 * D, 3 * D, 128, &tg);
  if (rc) return rc;
  const dim3 grid(p.qblocks, p.H, p.n_seq);
  attn_bwd_tcl_dq_kernel<<<grid, LQ_THREADS, A_SMEM, stream>>>(tq, ta, td, to, tg, p);
  rc = check_launch("attn_bwd_tcl_dq");
  if (rc) return rc;
  attn_bwd_tcl_dkv_kernel<<<grid, LB_THREADS, K_SMEM, stream>>>(tq, td, tg, p);
  return check_launch("attn_bwd_tcl_dkv");
}

}  // namespace fvqa
