// Fused causal attention forward / backward for the LLaMA-VQA step, with
//   * RoPE (interleaved pairs, llama/model.py:61-67): q/k arrive already rotated (the rotation is
//     folded into the QKV GEMM epilogue); the INVERSE rotation is applied here to dQ/dK in registers
//     before they are written, so the dX GEMM consumes gradients of the un-rotated projections;
//   * the adapter-prompt branch: a SEPARATE softmax over the A adapter keys scaled by tanh(gate1)
//     (model.py:99-115) whose keys/values are shared by every sequence;
//   * the gate2 bias on text scores of rows >= vs+F, columns [vs, vs+F) (model.py:116-119), per
//     sequence (vstart < 0 = no bias, the QAV stream).
// Flash-style: scores/probabilities never leave the SM. Backward is two deterministic passes
// (no atomics): pass A owns query rows (dQ, D, gate partials), pass B owns keys (dK, dV and the
// adapter dK_a/dV_a partials); a tiny kernel reduces the per-CTA partials in a fixed order.
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 h16 tensor-core tiles fed from padded
// shared memory through ldmatrix; 8-warp CTAs own 128 query rows (or 128 keys) and stream the other
// operand in 64-row tiles through a cp.async double buffer (attention is ~0.5 % of the step's FLOPs;
// the tcgen05 budget went to the GEMM first).
#include "attention.h"
#include "common.cuh"

namespace fvqa {

constexpr int AT_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// ---------------------------------------------------------------------------------------------
// mma.sync / ldmatrix helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32." FVQA_MMA_TYPE "." FVQA_MMA_TYPE ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// acc[NT][4] += A[16 x K] * B^T.  A: smem row-major rows a_row0..+16; B: smem stored [n][k], rows b_row0..+NT*8.
template <int NT, int K, int LD>
__device__ __forceinline__ void warp_mma_nt(float (&acc)[NT][4], uint32_t sA, int a_row0, uint32_t sB, int b_row0) {
  const int lane = threadIdx.x & 31;
  const uint32_t a_addr = sA + static_cast<uint32_t>(((a_row0 + (lane & 15)) * LD + (lane >> 4) * 8) * 2);
  const uint32_t b_addr = sB + static_cast<uint32_t>(((b_row0 + (lane & 7) + ((lane >> 4) << 3)) * LD + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    uint32_t a[4];
    ldsm_x4(a, a_addr + kb * 32);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4(b, b_addr + static_cast<uint32_t>(np * 16 * LD * 2 + kb * 32));
      mma16816(acc[2 * np], a, b[0], b[1]);
      mma16816(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}

// acc[NT][4] += Afrag * B.  Afrag: KB register k-blocks; B: smem stored [k][n], rows b_k0..+KB*16, cols b_n0..+NT*8.
template <int NT, int KB, int LD>
__device__ __forceinline__ void warp_mma_ra_t(float (&acc)[NT][4], const uint32_t (&a)[KB][4], uint32_t sB, int b_k0, int b_n0) {
  const int lane = threadIdx.x & 31;
  const uint32_t b_addr = sB + static_cast<uint32_t>(((b_k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + b_n0 + (lane >> 4) * 8) * 2);
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4_t(b, b_addr + static_cast<uint32_t>((kb * 16 * LD + np * 16) * 2));
      mma16816(acc[2 * np], a[kb], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[kb], b[2], b[3]);
    }
  }
}

// C fragments (two adjacent n-tiles) -> A fragment of one k-block.
__device__ __forceinline__ void c_to_a(uint32_t (&a)[4], const float (&c0)[4], const float (&c1)[4]) {
  a[0] = pack_h16x2(c0[0], c0[1]);
  a[1] = pack_h16x2(c0[2], c0[3]);
  a[2] = pack_h16x2(c1[0], c1[1]);
  a[3] = pack_h16x2(c1[2], c1[3]);
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// reductions over the 8 lanes that share (lane & 3): i.e. over the M index of a C fragment column
__device__ __forceinline__ float col_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
}
__device__ __forceinline__ float col_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v + __shfl_xor_sync(0xffffffffu, v, 16);
}

// ---------------------------------------------------------------------------------------------
// tile movers
// ---------------------------------------------------------------------------------------------
// Load ROWS x HD h16 (row stride `stride` elements) starting at sequence position row0 into padded
// smem; rows with position >= limit are zero-filled; optional RoPE at position = row index.
template <int HD, int ROWS, bool ROPE>
__device__ __forceinline__ void load_tile(h16* s, const h16* __restrict__ g, long stride, int row0, int limit,
                                          const float* __restrict__ cosT, const float* __restrict__ sinT) {
  constexpr int LD = HD + 8;
  constexpr int VPR = HD / 8;
  for (int idx = threadIdx.x; idx < ROWS * VPR; idx += AT_THREADS) {
    const int r = idx / VPR, v = idx - r * VPR;
    const int pos = row0 + r;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (pos < limit) {
      val = __ldg(reinterpret_cast<const uint4*>(g + static_cast<long>(pos) * stride) + v);
      if (ROPE) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(cosT + static_cast<long>(pos) * (HD / 2)) + v);
        const float4 sn = __ldg(reinterpret_cast<const float4*>(sinT + static_cast<long>(pos) * (HD / 2)) + v);
        float f[8];
        unpack8(val, f);
        const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = f[2 * i], b = f[2 * i + 1];
          f[2 * i] = a * cc[i] - b * ss[i];
          f[2 * i + 1] = a * ss[i] + b * cc[i];
        }
        val = pack8(f);
      }
    }
    *reinterpret_cast<uint4*>(s + r * LD + v * 8) = val;
  }
}

// Warp writes its 16 x HD fp32 C-fragment tile (optionally inverse-rotated, scaled) as h16 through a
// private smem staging area (16 rows, padded) to global rows row0+0..15 (< limit) with 16-byte stores.
template <int HD, bool INV_ROPE>
__device__ __forceinline__ void store_tile_warp(h16* stage, float (&acc)[HD / 8][4], float scale, h16* __restrict__ g, long stride,
                                                int row0, int limit, const float* __restrict__ cosT, const float* __restrict__ sinT) {
  constexpr int LD = HD + 8;
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float a = acc[nt][2 * hh] * scale, b = acc[nt][2 * hh + 1] * scale;
      const int r = gq + 8 * hh;
      if (INV_ROPE) {
        const int pos = row0 + r;
        if (pos < limit) {
          const float c = __ldg(cosT + static_cast<long>(pos) * (HD / 2) + nt * 4 + t);
          const float sn = __ldg(sinT + static_cast<long>(pos) * (HD / 2) + nt * 4 + t);
          const float ra = a * c + b * sn, rb = -a * sn + b * c;
          a = ra; b = rb;
        }
      }
      *reinterpret_cast<uint32_t*>(stage + r * LD + nt * 8 + 2 * t) = pack_h16x2(a, b);
    }
  }
  __syncwarp();
  constexpr int VPR = HD / 8;
  for (int idx = lane; idx < 16 * VPR; idx += 32) {
    const int r = idx / VPR, v = idx - r * VPR;
    if (row0 + r < limit)
      *(reinterpret_cast<uint4*>(g + static_cast<long>(row0 + r) * stride) + v) = *reinterpret_cast<const uint4*>(stage + r * LD + v * 8);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// asynchronous tile loads (cp.async, 16 B per request; rows past `limit` are zero-filled)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int HD, int ROWS, int NT>
__device__ __forceinline__ void load_tile_async(h16* s, const h16* __restrict__ g, long stride, int row0, int limit) {
  constexpr int LD = HD + 8;
  constexpr int VPR = HD / 8;
  const uint32_t sbase = smem_u32(s);
#pragma unroll
  for (int idx = threadIdx.x; idx < ROWS * VPR; idx += NT) {
    const int r = idx / VPR, v = idx - r * VPR;
    const int pos = row0 + r;
    const bool ok = pos < limit;
    cp_async16(sbase + static_cast<uint32_t>((r * LD + v * 8) * 2), g + static_cast<long>(ok ? pos : 0) * stride + v * 8, ok);
  }
}

constexpr int AT_NT = 256;        // threads per CTA (8 warps)
constexpr int AT_QB = 128;        // query rows (pass A / fwd) or keys (pass B) owned by a CTA
constexpr int AT_T = 64;          // streamed tile (keys in fwd / pass A, query rows in pass B)

// ---------------------------------------------------------------------------------------------
// forward: CTA = (128 query rows, head, sequence); K/V tiles of 64 keys double-buffered
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(AT_NT) attn_fwd_kernel(const AttnParams p) {
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem[];
  h16* sQ = reinterpret_cast<h16*>(smem);
  h16* sK = sQ + AT_QB * LD;            // [2][64][LD]
  h16* sV = sK + 2 * AT_T * LD;         // [2][64][LD]
  h16* sKa = sV + 2 * AT_T * LD;
  h16* sVa = sKa + AT_AP * LD;
  const int qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int S = p.S, D = p.H * HD;
  const long qkv_stride = 3L * D;
  const h16* qbase = p.qkv + static_cast<long>(n) * S * qkv_stride + h * HD;
  const h16* kbase = qbase + D;
  const h16* vbase = qbase + 2 * D;
  const int r0 = qb * AT_QB;
  const float scale2 = rsqrtf(static_cast<float>(HD)) * LOG2E;
  const int vs = p.vstart[n];
  const float bias2 = (vs >= 0) ? p.gate2[h] * LOG2E : 0.f;
  const int bias_row0 = (vs >= 0) ? vs + p.F : 0x7fffffff;   // rows >= this get the bias
  const int bias_c0 = vs, bias_c1 = vs + p.F;                  // columns [c0, c1)
  const int last_key = min(S, r0 + AT_QB) - 1;                 // causal: keys beyond the block's last row are never needed
  const int n_tiles = last_key / AT_T + 1;

  load_tile_async<HD, AT_QB, AT_NT>(sQ, qbase, qkv_stride, r0, S);
  load_tile_async<HD, AT_AP, AT_NT>(sKa, p.akv + h * HD, p.akv_ld, 0, p.A);
  load_tile_async<HD, AT_AP, AT_NT>(sVa, p.akv + D + h * HD, p.akv_ld, 0, p.A);
  load_tile_async<HD, AT_T, AT_NT>(sK, kbase, qkv_stride, 0, S);
  load_tile_async<HD, AT_T, AT_NT>(sV, vbase, qkv_stride, 0, S);
  cp_async_commit();
  if (n_tiles > 1) {
    load_tile_async<HD, AT_T, AT_NT>(sK + AT_T * LD, kbase, qkv_stride, AT_T, S);
    load_tile_async<HD, AT_T, AT_NT>(sV + AT_T * LD, vbase, qkv_stride, AT_T, S);
  }
  cp_async_commit();

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int row_a = r0 + warp * 16 + gq;   // this thread's rows: row_a and row_a + 8
  const int warp_last_row = r0 + warp * 16 + 15;
  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV), sKa_u = smem_u32(sKa), sVa_u = smem_u32(sVa);

  for (int j = 0; j < n_tiles; ++j) {
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t kb_u = sK_u + static_cast<uint32_t>((j & 1) * AT_T * LD * 2);
    const uint32_t vb_u = sV_u + static_cast<uint32_t>((j & 1) * AT_T * LD * 2);
    if (j * AT_T <= warp_last_row) {       // warp-uniform causal skip
      float s[AT_T / 8][4];
#pragma unroll
      for (int i = 0; i < AT_T / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
      warp_mma_nt<AT_T / 8, HD, LD>(s, sQ_u, warp * 16, kb_u, 0);
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < AT_T / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = row_a + 8 * (e >> 1);
          const int col = j * AT_T + nt * 8 + 2 * t + (e & 1);
          float v = s[nt][e] * scale2;
          if (row >= bias_row0 && col >= bias_c0 && col < bias_c1) v += bias2;
          if (col > row) v = -INFINITY;
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
      float alpha[2], m_new[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        m_new[hh] = fmaxf(m_run[hh], quad_max(mx[hh]));
        alpha[hh] = (m_run[hh] == -INFINITY) ? 0.f : exp2f(m_run[hh] - m_new[hh]);
        m_run[hh] = m_new[hh];
        l_run[hh] *= alpha[hh];
      }
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { o[i][0] *= alpha[0]; o[i][1] *= alpha[0]; o[i][2] *= alpha[1]; o[i][3] *= alpha[1]; }
      uint32_t pf[AT_T / 16][4];
#pragma unroll
      for (int nt = 0; nt < AT_T / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = (m_new[e >> 1] == -INFINITY) ? 0.f : exp2f(s[nt][e] - m_new[e >> 1]);
          s[nt][e] = pv;
          l_run[e >> 1] += pv;
        }
      }
#pragma unroll
      for (int kb = 0; kb < AT_T / 16; ++kb) c_to_a(pf[kb], s[2 * kb], s[2 * kb + 1]);
      warp_mma_ra_t<HD / 8, AT_T / 16, LD>(o, pf, vb_u, 0, 0);
    }
    __syncthreads();                         // everyone done with buffer j&1 before it is refilled
    if (j + 2 < n_tiles) {
      load_tile_async<HD, AT_T, AT_NT>(sK + (j & 1) * AT_T * LD, kbase, qkv_stride, (j + 2) * AT_T, S);
      load_tile_async<HD, AT_T, AT_NT>(sV + (j & 1) * AT_T * LD, vbase, qkv_stride, (j + 2) * AT_T, S);
    }
    cp_async_commit();
  }
  float lse2[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const float l = quad_sum(l_run[hh]);
    const float inv = (l > 0.f) ? 1.f / l : 0.f;
    lse2[hh] = m_run[hh] + log2f(l);
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { o[i][2 * hh] *= inv; o[i][2 * hh + 1] *= inv; }
  }
  if (t == 0) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = row_a + 8 * hh;
      if (row < S) p.lse[(static_cast<long>(n) * p.H + h) * S + row] = lse2[hh] * LN2;
    }
  }
  // ---- adapter branch: separate softmax over the A adapter keys, scaled by tanh(gate1) ----
  {
    float sa[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    warp_mma_nt<2, HD, LD>(sa, sQ_u, warp * 16, sKa_u, 0);
    const float tg = tanhf(p.gate1[h]);
    float mx[2] = {-INFINITY, -INFINITY}, sm[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + 2 * t + (e & 1);
        sa[nt][e] = (key < p.A) ? sa[nt][e] * scale2 : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sa[nt][e]);
      }
    mx[0] = quad_max(mx[0]); mx[1] = quad_max(mx[1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sa[nt][e] = exp2f(sa[nt][e] - mx[e >> 1]);
        sm[e >> 1] += sa[nt][e];
      }
    sm[0] = quad_sum(sm[0]); sm[1] = quad_sum(sm[1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sa[nt][e] = sa[nt][e] / sm[e >> 1] * tg;               // softmax * tanh(gate1)
    uint32_t pa[1][4];
    c_to_a(pa[0], sa[0], sa[1]);
    warp_mma_ra_t<HD / 8, 1, LD>(o, pa, sVa_u, 0, 0);
  }
  // each warp re-uses its own 16 rows of sQ as the staging area (only this warp ever read them)
  __syncwarp();
  store_tile_warp<HD, false>(sQ + warp * 16 * LD, o, 1.f, p.out + static_cast<long>(n) * S * D + h * HD, D, r0 + warp * 16, S, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------------
// backward pass A: CTA = (128 query rows, head, sequence) -> dQ, D_x, gate1/gate2 partial sums
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(AT_NT) attn_bwd_dq_kernel(const AttnParams p) {
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem[];
  h16* sQ = reinterpret_cast<h16*>(smem);
  h16* sdO = sQ + AT_QB * LD;
  h16* sO = sdO + AT_QB * LD;
  h16* sK = sO + AT_QB * LD;             // [2][64][LD]
  h16* sV = sK + 2 * AT_T * LD;          // [2][64][LD]
  h16* sKa = sV + 2 * AT_T * LD;
  h16* sVa = sKa + AT_AP * LD;
  float* sRed = reinterpret_cast<float*>(sVa + AT_AP * LD);   // [16] cross-warp gate partials
  const int qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int S = p.S, D = p.H * HD;
  const long qkv_stride = 3L * D;
  const h16* qbase = p.qkv + static_cast<long>(n) * S * qkv_stride + h * HD;
  const h16* kbase = qbase + D;
  const h16* vbase = qbase + 2 * D;
  const h16* obase = p.out + static_cast<long>(n) * S * D + h * HD;
  const h16* dobase = p.dout + static_cast<long>(n) * S * D + h * HD;
  const int r0 = qb * AT_QB;
  const float scale = rsqrtf(static_cast<float>(HD));
  const float scale2 = scale * LOG2E;
  const int vs = p.vstart[n];
  const float bias2 = (vs >= 0) ? p.gate2[h] * LOG2E : 0.f;
  const int bias_row0 = (vs >= 0) ? vs + p.F : 0x7fffffff;
  const int bias_c0 = vs, bias_c1 = vs + p.F;
  const float tg = tanhf(p.gate1[h]);
  const int last_key = min(S, r0 + AT_QB) - 1;
  const int n_tiles = last_key / AT_T + 1;

  load_tile_async<HD, AT_QB, AT_NT>(sQ, qbase, qkv_stride, r0, S);
  load_tile_async<HD, AT_QB, AT_NT>(sdO, dobase, D, r0, S);
  load_tile_async<HD, AT_QB, AT_NT>(sO, obase, D, r0, S);
  load_tile_async<HD, AT_AP, AT_NT>(sKa, p.akv + h * HD, p.akv_ld, 0, p.A);
  load_tile_async<HD, AT_AP, AT_NT>(sVa, p.akv + D + h * HD, p.akv_ld, 0, p.A);
  load_tile_async<HD, AT_T, AT_NT>(sK, kbase, qkv_stride, 0, S);
  load_tile_async<HD, AT_T, AT_NT>(sV, vbase, qkv_stride, 0, S);
  cp_async_commit();
  if (n_tiles > 1) {
    load_tile_async<HD, AT_T, AT_NT>(sK + AT_T * LD, kbase, qkv_stride, AT_T, S);
    load_tile_async<HD, AT_T, AT_NT>(sV + AT_T * LD, vbase, qkv_stride, AT_T, S);
  }
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  const uint32_t sQ_u = smem_u32(sQ), sdO_u = smem_u32(sdO), sK_u = smem_u32(sK), sV_u = smem_u32(sV),
                 sKa_u = smem_u32(sKa), sVa_u = smem_u32(sVa);
  const int row_a = r0 + warp * 16 + gq;
  const int warp_last_row = r0 + warp * 16 + 15;

  // D_total[row] = <dO[row], O[row]> for this thread's two rows (each lane of the quad takes a quarter of hd)
  float dtot[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int lr = warp * 16 + gq + 8 * hh;
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < HD / 32; ++v) {
      const int vec = t * (HD / 32) + v;
      float a[8], b[8];
      unpack8(*reinterpret_cast<const uint4*>(sO + lr * LD + vec * 8), a);
      unpack8(*reinterpret_cast<const uint4*>(sdO + lr * LD + vec * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc += a[e] * b[e];
    }
    dtot[hh] = quad_sum(acc);
  }

  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
  float g1_part = 0.f, g2_part = 0.f;
  float dx[2];
  // ---- adapter branch ----
  {
    float sa[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float dpa[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    warp_mma_nt<2, HD, LD>(sa, sQ_u, warp * 16, sKa_u, 0);
    warp_mma_nt<2, HD, LD>(dpa, sdO_u, warp * 16, sVa_u, 0);     // dP_a' = dO V_a^T (without tanh(gate1))
    float mx[2] = {-INFINITY, -INFINITY}, sm[2] = {0.f, 0.f}, da[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + 2 * t + (e & 1);
        sa[nt][e] = (key < p.A) ? sa[nt][e] * scale2 : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sa[nt][e]);
      }
    mx[0] = quad_max(mx[0]); mx[1] = quad_max(mx[1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sa[nt][e] = exp2f(sa[nt][e] - mx[e >> 1]);
        sm[e >> 1] += sa[nt][e];
      }
    sm[0] = quad_sum(sm[0]); sm[1] = quad_sum(sm[1]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sa[nt][e] /= sm[e >> 1];
        da[e >> 1] += sa[nt][e] * dpa[nt][e];
      }
    da[0] = quad_sum(da[0]); da[1] = quad_sum(da[1]);            // D_a' = <dO, P_a V_a>
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const bool ok = (row_a + 8 * hh) < S;
      if (ok && t == 0) g1_part += da[hh];
      dx[hh] = dtot[hh] - tg * da[hh];                           // D of the text softmax
      if (ok && t == 0) p.ws_dx[(static_cast<long>(n) * p.H + h) * (p.qblocks * AT_QB) + row_a + 8 * hh] = dx[hh];
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sa[nt][e] = tg * sa[nt][e] * (dpa[nt][e] - da[e >> 1]);   // dS_a
    uint32_t dsa[1][4];
    c_to_a(dsa[0], sa[0], sa[1]);
    warp_mma_ra_t<HD / 8, 1, LD>(dq, dsa, sKa_u, 0, 0);
  }
  float lse2[2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int row = row_a + 8 * hh;
    lse2[hh] = (row < S) ? p.lse[(static_cast<long>(n) * p.H + h) * S + row] * LOG2E : 0.f;
  }
  // ---- text keys ----
  for (int j = 0; j < n_tiles; ++j) {
    if (j > 0) {
      cp_async_wait<1>();
      __syncthreads();
    }
    const uint32_t kb_u = sK_u + static_cast<uint32_t>((j & 1) * AT_T * LD * 2);
    const uint32_t vb_u = sV_u + static_cast<uint32_t>((j & 1) * AT_T * LD * 2);
    if (j * AT_T <= warp_last_row) {
      float s[AT_T / 8][4], dp[AT_T / 8][4];
#pragma unroll
      for (int i = 0; i < AT_T / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
      warp_mma_nt<AT_T / 8, HD, LD>(s, sQ_u, warp * 16, kb_u, 0);
      warp_mma_nt<AT_T / 8, HD, LD>(dp, sdO_u, warp * 16, vb_u, 0);
#pragma unroll
      for (int nt = 0; nt < AT_T / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = row_a + 8 * (e >> 1);
          const int col = j * AT_T + nt * 8 + 2 * t + (e & 1);
          float v = s[nt][e] * scale2;
          const bool biased = (row >= bias_row0 && col >= bias_c0 && col < bias_c1);
          if (biased) v += bias2;
          const float pv = (col > row || row >= S) ? 0.f : exp2f(v - lse2[e >> 1]);
          const float ds = pv * (dp[nt][e] - dx[e >> 1]);
          if (biased) g2_part += ds;
          s[nt][e] = ds;
        }
      }
      uint32_t dsf[AT_T / 16][4];
#pragma unroll
      for (int kb = 0; kb < AT_T / 16; ++kb) c_to_a(dsf[kb], s[2 * kb], s[2 * kb + 1]);
      warp_mma_ra_t<HD / 8, AT_T / 16, LD>(dq, dsf, kb_u, 0, 0);
    }
    __syncthreads();
    if (j + 2 < n_tiles) {
      load_tile_async<HD, AT_T, AT_NT>(sK + (j & 1) * AT_T * LD, kbase, qkv_stride, (j + 2) * AT_T, S);
      load_tile_async<HD, AT_T, AT_NT>(sV + (j & 1) * AT_T * LD, vbase, qkv_stride, (j + 2) * AT_T, S);
    }
    cp_async_commit();
  }
  // gate partial sums of this CTA (fixed reduction order)
  g1_part = warp_sum(g1_part);
  g2_part = warp_sum(g2_part);
  if (lane == 0) { sRed[warp] = g1_part; sRed[8 + warp] = g2_part; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* wsg = p.ws_gate + ((static_cast<long>(n) * p.H + h) * p.qblocks + qb) * 2;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += sRed[w]; b += sRed[8 + w]; }
    wsg[0] = a;
    wsg[1] = b;
  }
  // dQ = scale * (dS K), inverse-rotated; staged through this warp's rows of sQ
  store_tile_warp<HD, true>(sQ + warp * 16 * LD, dq, scale, p.dqkv + static_cast<long>(n) * S * qkv_stride + h * HD, qkv_stride,
                            r0 + warp * 16, S, p.cosT, p.sinT);
}

// ---------------------------------------------------------------------------------------------
// backward pass B: CTA = (128 keys, head, sequence) -> dK, dV ; blockIdx.x == qblocks -> adapter keys
// (transposed formulation: S^T = K Q^T so that P^T / dS^T come out as A-operand fragments);
// Q / dO tiles of 64 rows double-buffered.
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(AT_NT) attn_bwd_dkv_kernel(const AttnParams p) {
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem[];
  h16* sK = reinterpret_cast<h16*>(smem);   // [128][LD]
  h16* sV = sK + AT_QB * LD;                 // [128][LD]
  h16* sQ = sV + AT_QB * LD;                 // [2][64][LD]
  h16* sdO = sQ + 2 * AT_T * LD;             // [2][64][LD]
  float* sLse = reinterpret_cast<float*>(sdO + 2 * AT_T * LD);   // [2][64]
  float* sDx = sLse + 2 * AT_T;                                  // [2][64]
  const int kb_idx = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  const int S = p.S, D = p.H * HD;
  const long qkv_stride = 3L * D;
  const h16* qbase = p.qkv + static_cast<long>(n) * S * qkv_stride + h * HD;
  const h16* kbase = qbase + D;
  const h16* vbase = qbase + 2 * D;
  const h16* dobase = p.dout + static_cast<long>(n) * S * D + h * HD;
  const float scale = rsqrtf(static_cast<float>(HD));
  const float scale2 = scale * LOG2E;
  const uint32_t sQ_u = smem_u32(sQ), sdO_u = smem_u32(sdO), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  const long lse_base = (static_cast<long>(n) * p.H + h) * S;
  const long dx_base = (static_cast<long>(n) * p.H + h) * (p.qblocks * AT_QB);

  if (kb_idx == p.qblocks) {
    // ======================= adapter keys =======================
    // smem use here: sK region = Q block (128 rows), sV region = dO block, sQ[0..16) = K_a, sdO[0..16) = V_a
    const float tg = tanhf(p.gate1[h]);
    load_tile_async<HD, AT_AP, AT_NT>(sQ, p.akv + h * HD, p.akv_ld, 0, p.A);
    load_tile_async<HD, AT_AP, AT_NT>(sdO, p.akv + D + h * HD, p.akv_ld, 0, p.A);
    float dka[HD / 8][4], dva[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) { dka[i][0] = dka[i][1] = dka[i][2] = dka[i][3] = 0.f; dva[i][0] = dva[i][1] = dva[i][2] = dva[i][3] = 0.f; }
    for (int i = 0; i < p.qblocks; ++i) {
      __syncthreads();
      load_tile_async<HD, AT_QB, AT_NT>(sK, qbase, qkv_stride, i * AT_QB, S);
      load_tile_async<HD, AT_QB, AT_NT>(sV, dobase, D, i * AT_QB, S);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      // each warp takes 16 of the 128 query rows (they are the contraction dimension of dK_a / dV_a)
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      warp_mma_nt<2, HD, LD>(st, sQ_u, 0, sK_u, warp * 16);      // S_a^T [keys x rows]
      warp_mma_nt<2, HD, LD>(dpt, sdO_u, 0, sV_u, warp * 16);    // dP_a'^T
      float mx[2][2], sm[2][2], da[2][2];                          // [n-tile][column parity]
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          // keys of this thread: gq (e = c) and gq + 8 (e = 2 + c)
          float v0 = (gq < p.A) ? st[nt][c] * scale2 : -INFINITY;
          float v1 = (gq + 8 < p.A) ? st[nt][2 + c] * scale2 : -INFINITY;
          st[nt][c] = v0; st[nt][2 + c] = v1;
          mx[nt][c] = col_max(fmaxf(v0, v1));
        }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          st[nt][c] = exp2f(st[nt][c] - mx[nt][c]);
          st[nt][2 + c] = exp2f(st[nt][2 + c] - mx[nt][c]);
          sm[nt][c] = col_sum(st[nt][c] + st[nt][2 + c]);
        }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int row = i * AT_QB + warp * 16 + nt * 8 + 2 * t + c;
          const float inv = (row < S) ? 1.f / sm[nt][c] : 0.f;    // rows past the sequence contribute nothing
          st[nt][c] *= inv; st[nt][2 + c] *= inv;
          da[nt][c] = col_sum(st[nt][c] * dpt[nt][c] + st[nt][2 + c] * dpt[nt][2 + c]);
        }
      float pt[2][4], dst[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pt[nt][e] = tg * st[nt][e];                                         // (tanh(g1) P_a)^T
          dst[nt][e] = tg * st[nt][e] * (dpt[nt][e] - da[nt][e & 1]);       // dS_a^T
        }
      uint32_t pf[1][4], df[1][4];
      c_to_a(pf[0], pt[0], pt[1]);
      c_to_a(df[0], dst[0], dst[1]);
      warp_mma_ra_t<HD / 8, 1, LD>(dva, pf, sV_u, warp * 16, 0);
      warp_mma_ra_t<HD / 8, 1, LD>(dka, df, sK_u, warp * 16, 0);
    }
    // cross-warp reduction in a fixed order through shared memory (fp32 [8][16][HD] fits in sK+sV)
    __syncthreads();
    float* red = reinterpret_cast<float*>(sK);
    float* wsa = p.ws_akv + (static_cast<long>(n) * p.H + h) * 2 * AT_AP * HD;
    auto reduce_out = [&](float (&acc)[HD / 8][4], float mul, int which) {
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          red[(warp * AT_AP + gq + 8 * (e >> 1)) * HD + nt * 8 + 2 * t + (e & 1)] = acc[nt][e] * mul;
      __syncthreads();
      for (int idx = threadIdx.x; idx < AT_AP * HD; idx += AT_NT) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += red[w * AT_AP * HD + idx];
        wsa[which * AT_AP * HD + idx] = a;
      }
      __syncthreads();
    };
    reduce_out(dka, scale, 0);
    reduce_out(dva, 1.f, 1);
    return;
  }

  // ======================= text keys =======================
  const int k0 = kb_idx * AT_QB;
  const int n_qt = (S + AT_T - 1) / AT_T;          // 64-row query tiles in the sequence
  const int i0 = k0 / AT_T;                        // first query tile that can see these keys
  auto load_q_tile = [&](int i, int buf) {
    load_tile_async<HD, AT_T, AT_NT>(sQ + buf * AT_T * LD, qbase, qkv_stride, i * AT_T, S);
    load_tile_async<HD, AT_T, AT_NT>(sdO + buf * AT_T * LD, dobase, D, i * AT_T, S);
    if (threadIdx.x < AT_T) {
      const int row = i * AT_T + threadIdx.x;
      sLse[buf * AT_T + threadIdx.x] = (row < S) ? p.lse[lse_base + row] * LOG2E : 0.f;
      sDx[buf * AT_T + threadIdx.x] = (row < S) ? p.ws_dx[dx_base + row] : 0.f;
    }
  };
  load_tile_async<HD, AT_QB, AT_NT>(sK, kbase, qkv_stride, k0, S);
  load_tile_async<HD, AT_QB, AT_NT>(sV, vbase, qkv_stride, k0, S);
  load_q_tile(i0, 0);
  cp_async_commit();
  if (i0 + 1 < n_qt) load_q_tile(i0 + 1, 1);
  cp_async_commit();
  const int vs = p.vstart[n];
  const float bias2 = (vs >= 0) ? p.gate2[h] * LOG2E : 0.f;
  const int bias_row0 = (vs >= 0) ? vs + p.F : 0x7fffffff;
  const int bias_c0 = vs, bias_c1 = vs + p.F;
  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const int key_a = k0 + warp * 16 + gq;   // this thread's keys: key_a, key_a + 8
  const int warp_first_key = k0 + warp * 16;
  for (int i = i0; i < n_qt; ++i) {
    const int buf = (i - i0) & 1;
    cp_async_wait<1>();
    __syncthreads();
    if (i * AT_T + AT_T - 1 >= warp_first_key) {    // warp-uniform causal skip
      const uint32_t qb_u = sQ_u + static_cast<uint32_t>(buf * AT_T * LD * 2);
      const uint32_t dob_u = sdO_u + static_cast<uint32_t>(buf * AT_T * LD * 2);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float st[4][4], dpt[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { st[q][0] = st[q][1] = st[q][2] = st[q][3] = 0.f; dpt[q][0] = dpt[q][1] = dpt[q][2] = dpt[q][3] = 0.f; }
        warp_mma_nt<4, HD, LD>(st, sK_u, warp * 16, qb_u, half * 32);     // S^T [16 keys x 32 rows]
        warp_mma_nt<4, HD, LD>(dpt, sV_u, warp * 16, dob_u, half * 32);   // dP^T
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = key_a + 8 * (e >> 1);
            const int lr = half * 32 + nt * 8 + 2 * t + (e & 1);
            const int row = i * AT_T + lr;
            float v = st[nt][e] * scale2;
            if (row >= bias_row0 && key >= bias_c0 && key < bias_c1) v += bias2;
            const float pv = (key > row || row >= S) ? 0.f : exp2f(v - sLse[buf * AT_T + lr]);
            st[nt][e] = pv;
            dpt[nt][e] = pv * (dpt[nt][e] - sDx[buf * AT_T + lr]);
          }
        }
        uint32_t pf[2][4], df[2][4];
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) { c_to_a(pf[kb], st[2 * kb], st[2 * kb + 1]); c_to_a(df[kb], dpt[2 * kb], dpt[2 * kb + 1]); }
        warp_mma_ra_t<HD / 8, 2, LD>(dv, pf, dob_u, half * 32, 0);
        warp_mma_ra_t<HD / 8, 2, LD>(dk, df, qb_u, half * 32, 0);
      }
    }
    __syncthreads();
    if (i + 2 < n_qt) load_q_tile(i + 2, buf);
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncthreads();   // everyone done with sQ/sdO before they become staging areas
  // 8 warps x 16 rows = 128 staging rows: sQ (2 x 64 rows) for dK, sdO for dV
  h16* dkbase = p.dqkv + static_cast<long>(n) * S * qkv_stride + D + h * HD;
  store_tile_warp<HD, true>(sQ + warp * 16 * LD, dk, scale, dkbase, qkv_stride, k0 + warp * 16, S, p.cosT, p.sinT);
  store_tile_warp<HD, false>(sdO + warp * 16 * LD, dv, 1.f, dkbase + D, qkv_stride, k0 + warp * 16, S, nullptr, nullptr);
}

// grid (H, A + 1): blocks (h, a < A) reduce adapter row a of head h over the sequences (independent
// loads, fixed summation order); block (h, A) reduces the gate partials of head h.
__global__ void __launch_bounds__(256) attn_bwd_reduce_kernel(const float* __restrict__ ws_akv, const float* __restrict__ ws_gate,
                                                              const float* __restrict__ gate1, float* __restrict__ dakv,
                                                              float* __restrict__ dgate1, float* __restrict__ dgate2, int n_seq,
                                                              int H, int hd, int A, int qtiles, int n_akv) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, a = blockIdx.y;
  const int D = H * hd;
  if (a < A) {
    for (int idx = threadIdx.x; idx < 2 * hd; idx += blockDim.x) {
      const int which = idx / hd, c = idx - which * hd;
      const float* src = ws_akv + (static_cast<long>(h) * 2 + which) * AT_AP * hd + a * hd + c;
      const long stride = static_cast<long>(H) * 2 * AT_AP * hd;
      float acc = 0.f;
      int n = 0;
      for (; n + 8 <= n_akv; n += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (n + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      for (; n < n_akv; ++n) acc += __ldg(src + n * stride);
      dakv[static_cast<long>(a) * 2 * D + which * D + h * hd + c] = acc;
    }
    return;
  }
  float g1 = 0.f, g2 = 0.f;
  const int total = n_seq * qtiles;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int n = i / qtiles, q = i - n * qtiles;
    const float* w = ws_gate + ((static_cast<long>(n) * H + h) * qtiles + q) * 2;
    g1 += w[0];
    g2 += w[1];
  }
  g1 = block_sum(g1, red);       // fixed tree order -> deterministic
  g2 = block_sum(g2, red);
  if (threadIdx.x == 0) {
    const float tg = tanhf(gate1[h]);
    dgate1[h] = (1.f - tg * tg) * g1;
    dgate2[h] = g2;
  }
}

template <int HD> constexpr int fwd_smem() { return (AT_QB + 4 * AT_T + 2 * AT_AP) * (HD + 8) * 2; }
template <int HD> constexpr int dq_smem() { return (3 * AT_QB + 4 * AT_T + 2 * AT_AP) * (HD + 8) * 2 + 128; }
template <int HD> constexpr int dkv_smem() { return (2 * AT_QB + 4 * AT_T) * (HD + 8) * 2 + 4 * AT_T * 4; }

int attn_init() {
  cudaError_t e;
#define FVQA_ATTR(fn, bytes)                                                                   \
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);            \
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(e));
  FVQA_ATTR(attn_fwd_kernel<64>, fwd_smem<64>())
  FVQA_ATTR(attn_fwd_kernel<128>, fwd_smem<128>())
  FVQA_ATTR(attn_bwd_dq_kernel<64>, dq_smem<64>())
  FVQA_ATTR(attn_bwd_dq_kernel<128>, dq_smem<128>())
  FVQA_ATTR(attn_bwd_dkv_kernel<64>, dkv_smem<64>())
  FVQA_ATTR(attn_bwd_dkv_kernel<128>, dkv_smem<128>())
#undef FVQA_ATTR
  int rc = attn_tc_init();
  if (rc) return rc;
  return attn_tcl_init();
}

static int check_attn_args(int n_seq, int S, int H, int hd, int A, int akv_ld) {
  FVQA_REQUIRE(hd == 64 || hd == 128, FVQA_ERR_UNSUPPORTED, "attention: head_dim %d not in {64,128}", hd);
  FVQA_REQUIRE(A >= 1 && A <= AT_AP, FVQA_ERR_UNSUPPORTED, "attention: adapter_len %d not in [1,%d]", A, AT_AP);
  FVQA_REQUIRE(n_seq > 0 && S > 0 && H > 0 && H <= 65535 && n_seq <= 65535, FVQA_ERR_INVALID_ARG, "attention: bad sizes n_seq=%d S=%d H=%d", n_seq, S, H);
  FVQA_REQUIRE(akv_ld % 8 == 0 && akv_ld >= 2 * H * hd, FVQA_ERR_INVALID_ARG, "attention: akv_ld %d", akv_ld);
  return FVQA_OK;
}

}  // namespace fvqa

using namespace fvqa;

extern "C" int fvqa_attn_fwd(const fvqa_h16* qkv, const fvqa_h16* akv, int akv_ld, const float* rope_cos, const float* rope_sin,
                             const float* gate1, const float* gate2, const int32_t* vstart, fvqa_h16* out, float* lse, int n_seq,
                             int S, int H, int hd, int A, int max_feats, void* stream) {
  int rc = check_attn_args(n_seq, S, H, hd, A, akv_ld);
  if (rc) return rc;
  AttnParams p{};
  p.qkv = reinterpret_cast<const h16*>(qkv); p.akv = reinterpret_cast<const h16*>(akv); p.akv_ld = akv_ld;
  p.cosT = rope_cos; p.sinT = rope_sin; p.gate1 = gate1; p.gate2 = gate2; p.vstart = vstart;
  p.out = reinterpret_cast<h16*>(out); p.lse = lse;
  p.n_seq = n_seq; p.S = S; p.H = H; p.A = A; p.F = max_feats; p.qblocks = (S + AT_QB - 1) / AT_QB;
  dim3 grid(p.qblocks, H, n_seq);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (attn_tc_supported(S, hd, A)) return attn_fwd_tc(p, s);
  if (attn_tcl_supported(S, hd, A)) return attn_fwd_tcl(p, s);
  if (hd == 64) attn_fwd_kernel<64><<<grid, AT_NT, fwd_smem<64>(), s>>>(p);
  else attn_fwd_kernel<128><<<grid, AT_NT, fwd_smem<128>(), s>>>(p);
  return check_launch("attn_fwd");
}

extern "C" int64_t fvqa_attn_bwd_ws_bytes(int n_seq, int S, int H, int hd, int A) {
  (void)A;
  const int64_t qblocks = (S + AT_QB - 1) / AT_QB;
  const int64_t nh = static_cast<int64_t>(n_seq) * H;
  return 4 * (nh * qblocks * AT_QB + nh * qblocks * 2 + nh * qblocks * 2 * AT_AP * hd);   // D | gate partials | adapter partials
}

extern "C" int fvqa_attn_bwd(const fvqa_h16* qkv, const fvqa_h16* akv, int akv_ld, const float* rope_cos, const float* rope_sin,
                             const float* gate1, const float* gate2, const int32_t* vstart, const fvqa_h16* out, const float* lse,
                             const fvqa_h16* dout, fvqa_h16* dqkv, float* dakv, float* dgate1, float* dgate2, void* ws, int n_seq,
                             int S, int H, int hd, int A, int max_feats, void* stream) {
  int rc = check_attn_args(n_seq, S, H, hd, A, akv_ld);
  if (rc) return rc;
  FVQA_REQUIRE(ws != nullptr, FVQA_ERR_INVALID_ARG, "attn_bwd: workspace is null");
  AttnParams p{};
  p.qkv = reinterpret_cast<const h16*>(qkv); p.akv = reinterpret_cast<const h16*>(akv); p.akv_ld = akv_ld;
  p.cosT = rope_cos; p.sinT = rope_sin; p.gate1 = gate1; p.gate2 = gate2; p.vstart = vstart;
  p.out = const_cast<h16*>(reinterpret_cast<const h16*>(out)); p.lse = const_cast<float*>(lse);
  p.dout = reinterpret_cast<const h16*>(dout); p.dqkv = reinterpret_cast<h16*>(dqkv);
  p.n_seq = n_seq; p.S = S; p.H = H; p.A = A; p.F = max_feats; p.qblocks = (S + AT_QB - 1) / AT_QB;
  const int64_t nh = static_cast<int64_t>(n_seq) * H;
  p.ws_dx = reinterpret_cast<float*>(ws);
  p.ws_gate = p.ws_dx + nh * p.qblocks * AT_QB;
  p.ws_akv = p.ws_gate + nh * p.qblocks * 2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid_a(p.qblocks, H, n_seq), grid_b(p.qblocks + 1, H, n_seq);
  int n_akv = n_seq;                     // groups of adapter partials the reduce kernel sums
  if (attn_tc_supported(S, hd, A)) {
    rc = attn_bwd_tc(p, s);
    if (rc) return rc;
  } else if (attn_tcl_supported(S, hd, A)) {
    rc = attn_bwd_tcl(p, s);
    if (rc) return rc;
    n_akv = n_seq * p.qblocks;           // one group per (sequence, query tile)
  } else if (hd == 64) {
    attn_bwd_dq_kernel<64><<<grid_a, AT_NT, dq_smem<64>(), s>>>(p);
    rc = check_launch("attn_bwd_dq");
    if (rc) return rc;
    attn_bwd_dkv_kernel<64><<<grid_b, AT_NT, dkv_smem<64>(), s>>>(p);
  } else {
    attn_bwd_dq_kernel<128><<<grid_a, AT_NT, dq_smem<128>(), s>>>(p);
    rc = check_launch("attn_bwd_dq");
    if (rc) return rc;
    attn_bwd_dkv_kernel<128><<<grid_b, AT_NT, dkv_smem<128>(), s>>>(p);
  }
  rc = check_launch("attn_bwd_dkv");
  if (rc) return rc;
  launch_k(attn_bwd_reduce_kernel, dim3(H, A + 1), dim3(256), 0, s, static_cast<const float*>(p.ws_akv), static_cast<const float*>(p.ws_gate), gate1, dakv,
           dgate1, dgate2, n_seq, H, hd, A, p.qblocks, n_akv);
  return check_launch("attn_bwd_reduce");
}
