// Device helpers shared by the tcgen05 attention kernels (attention_tc.cu: S <= 128; attention_tc_long.cu: any S).
#pragma once
#include <cuda_fp16.h>

#include "attention.h"
#include "common.cuh"

namespace fvqa {
namespace {

constexpr int TC_THREADS = 128;
constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr float TC_LN2 = 0.6931471805599453f;

// MN-major, 128-byte swizzle: 64 MN-elements per 128-byte row, 8-row (K) groups 1024 B apart (SBO),
// 64-element MN blocks `lbo_bytes` apart.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// No swizzle ("interleave"): 8 x 16-byte core matrices (128 B contiguous).
//   K-major : core matrices along K are lbo apart, 8-row groups along M/N are sbo apart.
//   MN-major: 8-element chunks along M/N are sbo apart, 8-row groups along K are lbo apart.
__device__ __forceinline__ uint64_t desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_h16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (FVQA_UMMA_FMT << 7) | (FVQA_UMMA_FMT << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// byte offset of the 16-byte chunk `c16` (0..7) of row r inside a K-major SW128 [rows][64] block
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}


// RoPE table rows [pos0, pos0 + 128) -> smem as fp16 (cos, sin) pairs: [128 rows][16 chunks ^ (row & 7)][4 x half2]
// (32 KB, conflict-free for thread-per-row reads). Rows with pos >= S are skipped.
__device__ __forceinline__ void stage_rope_table(uint8_t* dst, const float* __restrict__ cosT, const float* __restrict__ sinT, int pos0, int S,
                                                 int tid, int nthreads) {
  const int rows = min(128, S - pos0);
  const int n4 = rows * 16;                              // float4 groups (4 pairs each)
  const float4* c4 = reinterpret_cast<const float4*>(cosT + static_cast<long>(pos0) * 64);
  const float4* s4 = reinterpret_cast<const float4*>(sinT + static_cast<long>(pos0) * 64);
  for (int idx = tid; idx < n4; idx += nthreads) {
    const float4 cc = __ldg(c4 + idx);
    const float4 ss = __ldg(s4 + idx);
    const int rr = idx >> 4, ch = idx & 15;
    __half2 h0 = __floats2half2_rn(cc.x, ss.x), h1 = __floats2half2_rn(cc.y, ss.y);
    __half2 h2 = __floats2half2_rn(cc.z, ss.z), h3 = __floats2half2_rn(cc.w, ss.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + rr * 256 + ((ch ^ (rr & 7)) << 4)) = u;
  }
}
// inverse rotation of the 8 values f (4 pairs) of row r, 16-byte chunk index c16 (0..15) of the head dimension
__device__ __forceinline__ void inv_rope8(float (&f)[8], const uint8_t* table, int r, int c16) {
  const uint4 t = *reinterpret_cast<const uint4*>(table + r * 256 + ((c16 ^ (r & 7)) << 4));
  const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 cs = __half22float2(*reinterpret_cast<const __half2*>(&tw[e]));
    const float a = f[2 * e], b = f[2 * e + 1];
    f[2 * e] = a * cs.x + b * cs.y;
    f[2 * e + 1] = -a * cs.y + b * cs.x;
  }
}

}  // namespace
}  // namespace fvqa
