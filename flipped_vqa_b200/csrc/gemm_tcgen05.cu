// h16 GEMM  C[M,N] = A[M,K] * B[N,K]^T (+ R)  on the 5th-gen tensor cores (tcgen05.mma, accumulators
// in TMEM), operands staged by TMA (128B swizzle) through a multi-stage mbarrier ring; persistent
// CTAs, one per SM, warp-specialised:
//   warp 0 : TMA producer (one elected thread)
//   warp 1 : MMA issuer   (one elected thread)
//   warp 2 : TMEM allocator / deallocator
//   warps 4-7 : epilogue (TMEM -> registers -> h16/fp32 -> global, optional residual add)
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//
// This kernel replaces every frozen nn.Linear of the reference (llama/model.py:89,99-100,128,142,
// 348,354) and — with the load-time transposed weight copies — its dX-only backward.
#include <cuda.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/fvqa_debug.h"
#include "common.cuh"
#include "tmap.h"

namespace fvqa {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;           // 64 h16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int PAIR_THREADS = 384;   // CTA-pair kernel: warps 0-3 TMA / MMA / TMEM alloc / idle, warps 4-11 epilogue (two per TMEM lane quadrant)
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
  // SwiGLU-fused epilogues of the CTA-pair kernel (EPI_SWIGLU_FWD / EPI_SWIGLU_BWD)
  void* aux;          // FWD: c = silu(a) * b output [M, hid];  BWD: g = [a | b] input [M, 2*hid]
  int ld_aux, hid;
  const int32_t* pos_ids;  // ROPE only: position of each row (ragged / compacted token layouts); nullptr -> row % S
};

enum { EPI_PLAIN = 0, EPI_ROPE = 1, EPI_SWIGLU_FWD = 2, EPI_SWIGLU_BWD = 3 };

// One epilogue chunk: NC (32 or 16) consecutive fp32 accumulator columns of one output row held in
// registers -> optional RoPE rotation / residual add -> global store (h16 or fp32).
template <int NC, bool OUT_F32, bool ROPE>
__device__ __forceinline__ void epilogue_store(uint32_t (&v)[32], void* __restrict__ Cout, const GemmEpi& epi, int row, int col0,
                                               int N, int ldc) {
  if constexpr (OUT_F32) {
    float* crow = reinterpret_cast<float*>(Cout) + static_cast<long>(row) * ldc + col0;
    const float* rrow = epi.R ? reinterpret_cast<const float*>(epi.R) + static_cast<long>(row) * epi.ldr + col0 : nullptr;
#pragma unroll
    for (int j = 0; j < NC / 4; ++j) {
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
    h16* crow = reinterpret_cast<h16*>(Cout) + static_cast<long>(row) * ldc + col0;
    const h16* rrow = epi.R ? reinterpret_cast<const h16*>(epi.R) + static_cast<long>(row) * epi.ldr + col0 : nullptr;
    if constexpr (ROPE) {
      const int pos = epi.pos_ids != nullptr ? __ldg(epi.pos_ids + row) : row % epi.S;
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) {
        const int col = col0 + 8 * q;                 // 8-column groups never straddle a head (hd % 8 == 0)
        if (col < epi.rope_cols) {
          const long ti = static_cast<long>(pos) * (epi.hd >> 1) + ((col % epi.hd) >> 1);
          const float4 cc = __ldg(reinterpret_cast<const float4*>(epi.cosT + ti));
          const float4 ss = __ldg(reinterpret_cast<const float4*>(epi.sinT + ti));
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
    for (int j = 0; j < NC / 8; ++j) {
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

template <int BN, bool OUT_F32, bool ROPE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
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
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  pdl_wait();

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
      constexpr uint32_t idesc = umma_idesc_h16(GEMM_BM, BN);
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
            umma_h16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
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
        if (row_ok && col0 < N) epilogue_store<32, OUT_F32, ROPE>(v, Cout, epi, row, col0, N, ldc);
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
// CTA-pair kernel (cta_group::2): a cluster of two CTAs on the two SMs of one TPC owns a 256 x BN
// output tile. Each CTA stages ITS 128 rows of A and ITS BN/2 rows of B per k-block (32 KB at
// BN = 256 instead of the 48 KB of the single-CTA tile -> deeper TMA ring, 1/3 less L2->smem
// traffic per FLOP); the leader CTA's elected thread issues tcgen05.mma.cta_group::2 (UMMA
// 256 x BN x 16) which reads both CTAs' shared memory and writes 128 accumulator rows into each
// CTA's TMEM. BN is a RUNTIME multiple of 16 in [64, 256], chosen per problem so that the 74 pairs
// finish their last wave together (N = 4096 outputs: 256 -> 2.6 waves, 176 -> 3.9 waves).
// Barriers: full[s] lives in the leader (both CTAs' TMA bytes are credited to it), empty[s] and
// tmem_full[a] exist in both CTAs and are signalled by multicast tcgen05.commit, tmem_empty[a]
// lives in the leader and collects the 8 epilogue warps of the pair.
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_MAX_STAGES = 8;
constexpr int PAIR_BAR_BYTES = 256;
constexpr int PAIR_SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA

__host__ __device__ constexpr int pair_stage_bytes(int bn) { return GEMM_BM * GEMM_BK * 2 + (bn / 2) * GEMM_BK * 2; }

// QUAD: a cluster of TWO CTA pairs (4 CTAs) owns a 256 x 2 BN output block: both pairs work on the same 256 rows and on
// adjacent column tiles, so CTA r and CTA r^2 need the same 128 x 64 A slice per k-block: each loads HALF of it (64 rows)
// and TMA-multicasts it to both (-25 % L2 -> SM operand bytes; the pair kernel sits on the L2 throughput cap, and this is
// what cuBLAS's 2x2-cluster kernels do on the N = 4096 shapes). A stage is free once BOTH pairs' MMAs have consumed it
// (the partner pair writes into it), hence two arrivals on empty[s], committed to all four CTAs. Only 33 four-CTA
// clusters fit the chip (132 of 148 SMs, profiles/r1_cluster_occupancy.txt).
// B_MN: B is given as a row-major [K, N] matrix (N contiguous) and read as an MN-major UMMA operand: C = A . B, i.e. the dX-only
// backward dX = dY . W straight from the [out, in] weight the forward uses - no transposed weight copy (13.5 GB at 7B). Per k-block
// each CTA TMA-loads (BN / 2) / 64 boxes of [64 k rows][64 columns] (same bytes as the K-major tile); BN must be a multiple of 128.
template <bool OUT_F32, int EPI, bool QUAD = false, bool B_MN = false>
__global__ void __cluster_dims__(QUAD ? 4 : 2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
gemm_nt_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         void* __restrict__ Cout, const GemmEpi epi, int M, int N, int K, int ldc, int BN, int stages, int l2_hints) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = pair_stage_bytes(BN);
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PAIR_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * PAIR_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * PAIR_MAX_STAGES + 2 + a); };
  const uint32_t holder = bar_base + 8u * (2 * PAIR_MAX_STAGES + 4);
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (holder - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();           // rank in the cluster: 0..1, or 0..3 (QUAD)
  const uint32_t rank = crank & 1u;                   // CTA within its pair; 0 = the pair's MMA leader
  const uint32_t leader = crank & ~1u;                // cluster rank of this pair's leader
  const int cq = QUAD ? static_cast<int>(crank >> 1) : 0;   // pair within the cluster
  // scheduling unit: a pair and its tile, or (QUAD) a cluster and its two column-adjacent tiles
  const int pair = QUAD ? blockIdx.x >> 2 : blockIdx.x >> 1, n_pairs = QUAD ? gridDim.x >> 2 : gridDim.x >> 1;
  constexpr bool ROPE = (EPI == EPI_ROPE);
  const int tiles_m = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  // EPI_SWIGLU_FWD: B = [W1; W3] (2*hid rows); tile tn pairs W1 rows [128 tn, +128) (CTA 0's half of B)
  // with W3 rows [128 tn, +128) (CTA 1's half) so that a row's a- and b-values meet in one accumulator.
  const int tiles_n = (EPI == EPI_SWIGLU_FWD) ? epi.hid / 128 : (N + BN - 1) / BN;
  const int num_tiles = QUAD ? tiles_m * (tiles_n >> 1) : tiles_m * tiles_n;     // scheduling units (QUAD: tiles_n is even)
  auto tile_of = [&](int unit) { return QUAD ? (unit % tiles_m) + tiles_m * (2 * (unit / tiles_m) + cq) : unit; };
  const int num_kb = K / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), QUAD ? 2 : 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * ((blockDim.x >> 5) - 4));  // epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(holder, 512);
    tmem_relinquish_pair();
  }
  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();               // barrier inits + TMEM allocation visible to both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  pdl_wait();                       // everything above overlapped the previous kernel's tail; from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < num_tiles; unit += n_pairs) {
        const int tile = tile_of(unit);
        const int m0 = (tile % tiles_m) * (2 * GEMM_BM) + static_cast<int>(rank) * GEMM_BM;
        const int n0 = (EPI == EPI_SWIGLU_FWD) ? (tile / tiles_m) * 128 + static_cast<int>(rank) * epi.hid
                                               : (tile / tiles_m) * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + GEMM_BM * GEMM_BK * 2;
          const uint32_t lfull = mapa_shared(full_bar(stage), leader);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * stage_bytes);
          if constexpr (QUAD) {
            // my half (64 rows) of the A slice I share with CTA crank ^ 2, multicast to both of us; B is this pair's own
            constexpr uint32_t kHalf = (GEMM_BM / 2) * GEMM_BK * 2;
            tma_load_2d_pair_mc(sa + cq * kHalf, &tmap_a, lfull, kb * GEMM_BK, m0 + cq * (GEMM_BM / 2),
                                static_cast<uint16_t>((1u << crank) | (1u << (crank ^ 2u))));
            if constexpr (B_MN) {
              for (int j = 0; j < (BN >> 7); ++j) tma_load_2d_pair(sb + j * 8192, &tmap_b, lfull, n0 + j * 64, kb * GEMM_BK);
            } else {
              tma_load_2d_pair(sb, &tmap_b, lfull, kb * GEMM_BK, n0);
            }
          } else if (l2_hints & 1) {      // weights stream through once per wave; the activation operand is re-read by every tile column
            tma_load_2d_pair_hint(sa, &tmap_a, lfull, kb * GEMM_BK, m0, L2_EVICT_LAST);
            tma_load_2d_pair_hint(sb, &tmap_b, lfull, kb * GEMM_BK, n0, L2_EVICT_FIRST);
          } else {
            tma_load_2d_pair(sa, &tmap_a, lfull, kb * GEMM_BK, m0);
            if constexpr (B_MN) {
              for (int j = 0; j < (BN >> 7); ++j) tma_load_2d_pair(sb + j * 8192, &tmap_b, lfull, n0 + j * 64, kb * GEMM_BK);
            } else {
              tma_load_2d_pair(sb, &tmap_b, lfull, kb * GEMM_BK, n0);
            }
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      // l2_hints bit 1 (probe, fvqa_debug_gemm_mixed_a): A is read in the OTHER 16-bit format than B (mixed fp16 x bf16 MMA)
      const uint32_t idesc = (umma_idesc_h16(2 * GEMM_BM, BN) ^ ((l2_hints & 2) ? (1u << 7) : 0u)) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = pair; unit < num_tiles; unit += n_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + GEMM_BM * GEMM_BK * 2;
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = B_MN ? umma_desc_mn_sw128(sb, 8192) : umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
            // K-major: +16 elements = 32 B inside the swizzle row; MN-major: +16 k-rows of 128 B = 2048 B (16-byte units)
            umma_h16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(B_MN ? 128 * k : 2 * k), idesc,
                              (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair_mask(empty_bar(stage), QUAD ? 0xF : 0x3);                       // QUAD: the partner pair writes into this stage too
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair_mask(tfull_bar(acc), static_cast<uint16_t>(0x3u << leader));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: own 128 rows, all BN columns) =====================
    // The epilogue of a short-K tile (K = 4096: ~21 us of MMA) is a chain of dependent global loads (residual / cos,sin /
    // g) and stores per chunk, which used to be LONGER than the main loop it should hide behind (Wo + residual: 114 us).
    // The loads of chunk c+1 are therefore issued before chunk c is processed, and the first chunk's before the
    // accumulator is even complete (88 us). A second warp per TMEM lane quadrant (384 threads, each warp half of the
    // chunks; test hook) measured no better and 5 % worse for the SwiGLU-forward epilogue, so four warps stay the default.
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const bool split = blockDim.x > 256;                  // 384 threads: two epilogue warps per quadrant (test hook: 256 = one)
    int acc = 0;
    uint32_t acc_phase = 0;
    const int n32 = BN >> 5;
    const int nch = n32 + ((BN & 16) ? 1 : 0);            // 32-column chunks incl. a 16-column tail
    const int c_begin = (split && half) ? (nch + 1) >> 1 : 0;
    const int c_end = (split && !half) ? (nch + 1) >> 1 : nch;
    for (int unit = pair; unit < num_tiles; unit += n_pairs) {
      const int tile = tile_of(unit);
      const int m0 = (tile % tiles_m) * (2 * GEMM_BM) + static_cast<int>(rank) * GEMM_BM;
      const int n0 = (tile / tiles_m) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * 256);
      if constexpr (EPI == EPI_SWIGLU_FWD) {
        // columns [0,128) = a = W1 x, [128,256) = b = W3 x of hidden units [128 tn, +128): write g = [a | b] (h16,
        // saved for backward) and c = silu(a) * b computed from the ROUNDED a, b (bit-identical to swiglu_fwd_kernel)
        const int hcol = (tile / tiles_m) * 128;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = split ? 2 * half : 0; c < (split ? 2 * half + 2 : 4); ++c) {
          uint32_t va[32], vb[32];
          tmem_ld_32x32(tbase + static_cast<uint32_t>(c * 32), va);
          tmem_ld_32x32(tbase + static_cast<uint32_t>(128 + c * 32), vb);
          tmem_ld_wait();
          if (row_ok) {
            h16* grow = reinterpret_cast<h16*>(Cout) + static_cast<long>(row) * ldc + hcol + c * 32;
            h16* crow = reinterpret_cast<h16*>(epi.aux) + static_cast<long>(row) * epi.ld_aux + hcol + c * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float fa[8], fb[8], o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { fa[e] = __uint_as_float(va[q * 8 + e]); fb[e] = __uint_as_float(vb[q * 8 + e]); }
              const uint4 pa = pack8(fa), pb = pack8(fb);
              unpack8(pa, fa);
              unpack8(pb, fb);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = fa[e] / (1.f + __expf(-fa[e])) * fb[e];
              *reinterpret_cast<uint4*>(grow + q * 8) = pa;
              *reinterpret_cast<uint4*>(grow + epi.hid + q * 8) = pb;
              *reinterpret_cast<uint4*>(crow + q * 8) = pack8(o);
            }
          }
        }
      } else if constexpr (EPI == EPI_SWIGLU_BWD) {
        // accumulator = dc = d(silu(a) * b) for hidden units [n0, n0 + BN): read g = [a | b], write
        // dg = [dc b s (1 + a (1 - s)) | dc a s] (bit-identical to swiglu_bwd_kernel on the h16-rounded dc)
        uint4 ga_n[4], gb_n[4];
        auto load_g = [&](int c) {
          const int col0 = n0 + c * 32;
          if (row_ok && col0 < N) {
            const uint4* gr = reinterpret_cast<const uint4*>(reinterpret_cast<const h16*>(epi.aux) + static_cast<long>(row) * epi.ld_aux + col0);
            const uint4* gr2 = reinterpret_cast<const uint4*>(reinterpret_cast<const h16*>(epi.aux) + static_cast<long>(row) * epi.ld_aux + epi.hid + col0);
#pragma unroll
            for (int q = 0; q < 4; ++q) { ga_n[q] = __ldg(gr + q); gb_n[q] = __ldg(gr2 + q); }
          }
        };
        // (BN is a multiple of 32 here: hid % 32 == 0 and the launcher keeps 32 | BN for this epilogue)
        if (c_begin < c_end) load_g(c_begin);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          uint4 ga[4], gb[4];
          const bool ok = row_ok && col0 < N;
#pragma unroll
          for (int q = 0; q < 4; ++q) { ga[q] = ga_n[q]; gb[q] = gb_n[q]; }
          if (c + 1 < c_end) load_g(c + 1);
          uint32_t v[32];
          tmem_ld_32x32(tbase + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait();
          if (ok) {
            h16* drow = reinterpret_cast<h16*>(Cout) + static_cast<long>(row) * ldc + col0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a[8], b[8], d[8], da[8], db[8];
              unpack8(ga[q], a);
              unpack8(gb[q], b);
#pragma unroll
              for (int e = 0; e < 8; ++e) d[e] = __uint_as_float(v[q * 8 + e]);
              unpack8(pack8(d), d);                                   // dc as the unfused path sees it (h16)
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float sg = 1.f / (1.f + __expf(-a[e]));
                da[e] = d[e] * b[e] * sg * (1.f + a[e] * (1.f - sg));
                db[e] = d[e] * a[e] * sg;
              }
              *reinterpret_cast<uint4*>(drow + q * 8) = pack8(da);
              *reinterpret_cast<uint4*>(drow + epi.hid + q * 8) = pack8(db);
            }
          }
        }
      } else {
        // aux = what the chunk needs from global memory besides the accumulator: the fp32 residual (8 x float4) or the
        // RoPE cos | sin of the row's position (4 + 4 x float4); loaded one chunk ahead
        constexpr bool PIPE = OUT_F32 || ROPE;
        float4 aux_n[8];
        const int pos = ROPE ? (epi.pos_ids != nullptr ? __ldg(epi.pos_ids + (row_ok ? row : 0)) : row % epi.S) : 0;
        auto load_aux = [&](int c) {
          const int col0 = n0 + c * 32;
          if (!(row_ok && col0 < N)) return;
          if constexpr (OUT_F32) {
            if (epi.R != nullptr) {
              const float4* rr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.R) + static_cast<long>(row) * epi.ldr + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col0 + 4 * j < N) aux_n[j] = rr[j];
            }
          } else if constexpr (ROPE) {
            // 8-column groups never straddle a head or the q|k / v boundary (hd % 8 == 0, col % 8 == 0)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = col0 + 8 * q;
              if (col < epi.rope_cols) {
                const long ti = static_cast<long>(pos) * (epi.hd >> 1) + ((col % epi.hd) >> 1);
                aux_n[q] = __ldg(reinterpret_cast<const float4*>(epi.cosT + ti));
                aux_n[4 + q] = __ldg(reinterpret_cast<const float4*>(epi.sinT + ti));
              }
            }
          }
        };
        if (PIPE && c_begin < c_end) load_aux(c_begin);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = n0 + c * 32;
          const bool tail = c >= n32;                       // the 16-column tail chunk (BN & 16)
          float4 aux[8];
          if constexpr (PIPE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) aux[j] = aux_n[j];
            if (c + 1 < c_end) load_aux(c + 1);
          }
          uint32_t v[32];
          if (tail) tmem_ld_32x16(tbase + static_cast<uint32_t>(c * 32), v);
          else tmem_ld_32x32(tbase + static_cast<uint32_t>(c * 32), v);
          tmem_ld_wait();
          const int ncol = tail ? 16 : 32;
          if (!(row_ok && col0 < N)) {
            // nothing to store for this lane / chunk
          } else if constexpr (OUT_F32) {
            float* crow = reinterpret_cast<float*>(Cout) + static_cast<long>(row) * ldc + col0;
            const bool has_r = epi.R != nullptr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (4 * j < ncol && col0 + 4 * j < N) {
                float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]));
                if (has_r) { o.x += aux[j].x; o.y += aux[j].y; o.z += aux[j].z; o.w += aux[j].w; }
                *reinterpret_cast<float4*>(crow + 4 * j) = o;
              }
            }
          } else {
            h16* crow = reinterpret_cast<h16*>(Cout) + static_cast<long>(row) * ldc + col0;
            const h16* rrow = epi.R ? reinterpret_cast<const h16*>(epi.R) + static_cast<long>(row) * epi.ldr + col0 : nullptr;
            if constexpr (ROPE) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (col0 + 8 * q < epi.rope_cols) {
                  const float cv[4] = {aux[q].x, aux[q].y, aux[q].z, aux[q].w};
                  const float sv[4] = {aux[4 + q].x, aux[4 + q].y, aux[4 + q].z, aux[4 + q].w};
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
              if (8 * j < ncol && col0 + 8 * j < N) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
                if (rrow != nullptr) {
                  float r[8];
                  unpack8(*reinterpret_cast<const uint4*>(rrow + 8 * j), r);
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] += r[e];
                }
                *reinterpret_cast<uint4*>(crow + 8 * j) = pack8(f);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), leader));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();               // the peer's shared memory / TMEM stay alive until every MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
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
static int g_quad_clusters = 0;   // co-resident 4-CTA clusters of the QUAD GEMM (33 on a B200: 132 of 148 SMs); 0 = unavailable
static std::atomic<int> g_quad_mode{1};       // fvqa_gemm_debug_quad: 0 = never, 1 = heuristic (default), 2 = every eligible plain GEMM
static std::mutex g_mu;

struct MapKey {
  const void* ptr;
  int rows, cols, ld, box_rows;
  int seq_len = 0;             // > 0: 3-D map (cols, seq_len, rows / seq_len)
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && seq_len == o.seq_len;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.cols) * 0xC2B2AE3D27D4EB4Full) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.ld) * 0x165667B19E3779F9ull) + (h << 6) + (h >> 2);
    h ^= static_cast<size_t>(k.box_rows) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.seq_len) * 0x27D4EB2F165667C5ull) + (h << 6) + (h >> 2);
    return h;
  }
};
// per host thread: launches take no lock (the library is re-entrant per device; each rank drives its GPU from one thread)
static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

int num_sms() { return g_num_sms; }

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
  e = cudaFuncSetAttribute(gemm_nt_kernel<BN, F32, ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                           GemmCfg<BN>::kSmemBytes);                                                          \
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
  FVQA_SET_SMEM(256, false, false)
  FVQA_SET_SMEM(256, true, false)
  FVQA_SET_SMEM(128, false, false)
  FVQA_SET_SMEM(128, true, false)
  FVQA_SET_SMEM(256, false, true)
  FVQA_SET_SMEM(128, false, true)
#undef FVQA_SET_SMEM
#define FVQA_SET_PAIR(F32, EPI)                                                                                   \
  e = cudaFuncSetAttribute(gemm_nt_pair_kernel<F32, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           PAIR_SMEM_LIMIT);                                                                       \
  FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(gemm pair): %s", cudaGetErrorString(e));
  FVQA_SET_PAIR(false, EPI_PLAIN)
  FVQA_SET_PAIR(true, EPI_PLAIN)
  FVQA_SET_PAIR(false, EPI_ROPE)
  FVQA_SET_PAIR(false, EPI_SWIGLU_FWD)
  FVQA_SET_PAIR(false, EPI_SWIGLU_BWD)
#undef FVQA_SET_PAIR
  {
    auto k1 = gemm_nt_pair_kernel<false, EPI_PLAIN, false, true>;
    auto k2 = gemm_nt_pair_kernel<false, EPI_SWIGLU_BWD, false, true>;
    auto k3 = gemm_nt_pair_kernel<false, EPI_PLAIN, true, true>;
    e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_LIMIT);
    FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(gemm nn): %s", cudaGetErrorString(e));
  }
  // 2x2-cluster (QUAD) variants of the plain-epilogue kernel: opt-in shared memory, and how many 4-CTA clusters fit the chip
  {
    auto kq16 = gemm_nt_pair_kernel<false, EPI_PLAIN, true>;
    auto kq32 = gemm_nt_pair_kernel<true, EPI_PLAIN, true>;
    e = cudaFuncSetAttribute(kq16, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_LIMIT);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kq32, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_LIMIT);
    FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(gemm quad): %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * (g_num_sms / 4));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = PAIR_SMEM_LIMIT;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 4; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kq16, &cfg) == cudaSuccess && n > 0) g_quad_clusters = n;
    (void)cudaGetLastError();
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return FVQA_OK;
}

int get_tmap(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, rows, cols, ld, box_rows};
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
  CUresult r = g_encode(&m, FVQA_TMAP_DTYPE, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FVQA_REQUIRE(r == CUDA_SUCCESS, FVQA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%d cols=%d ld=%d box_rows=%d",
               static_cast<int>(r), ptr, rows, cols, ld, box_rows);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return FVQA_OK;
}

// 3-D view of a token-major matrix: (column, position in sequence, sequence). Boxes are 64 columns x
// box_rows positions of ONE sequence: positions >= seq_len are out of bounds (zero-filled on load, dropped
// on store), so a 128-row box never touches the next sequence when seq_len < 128.
int get_tmap_seq(const void* ptr, int n_seq, int seq_len, int cols, int ld, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, n_seq * seq_len, cols, ld, box_rows, seq_len};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return FVQA_OK;
  }
  CUtensorMap m;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(seq_len), static_cast<cuuint64_t>(n_seq)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * static_cast<cuuint64_t>(seq_len)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(&m, FVQA_TMAP_DTYPE, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FVQA_REQUIRE(r == CUDA_SUCCESS, FVQA_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p n_seq=%d S=%d cols=%d ld=%d",
               static_cast<int>(r), ptr, n_seq, seq_len, cols, ld);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return FVQA_OK;
}

template <int BN, bool OUT_F32, bool ROPE>
static int launch_gemm(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, const GemmEpi& epi, int M,
                       int N, int K, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = get_tmap(A, M, K, lda, GEMM_BM, &ta);
  if (rc) return rc;
  rc = get_tmap(B, N, K, ldb, BN, &tb);
  if (rc) return rc;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_nt_kernel<BN, OUT_F32, ROPE><<<grid, GEMM_THREADS, GemmCfg<BN>::kSmemBytes, stream>>>(ta, tb, C, epi, M, N, K, ldc);
  return check_launch("gemm_nt");
}

// M <= 16 (adapter-prompt projections): gemm_skinny.cu
bool gemm_skinny_supported(int M, int K, const void* R);
int gemm_skinny(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, int M, int N, int K, int out_fp32, int num_sms,
                cudaStream_t stream);
int gemm_skinny_grouped(const h16* A, long strideA, int lda, const h16* B, const h16* const* Bptrs, int ldb, void* C, long strideC,
                        int ldc, int M, int N, int K, int groups, int out_fp32, int num_sms, cudaStream_t stream);
bool gemm_skinny_shape_ok(int M, int N, int K);
int gemm_skinny_epi(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, int M, int N, int K, int out_fp32, const float* R,
                    int ldr, const float* cosT, const float* sinT, int rope_cols, int hd, int S, const int32_t* pos_ids, int num_sms,
                    cudaStream_t stream);
extern std::atomic<int> g_skinny_force_nt;

// ---- CTA-pair path ---------------------------------------------------------------------------
static std::atomic<int> g_mixed_a{0};    // probe (fvqa_gemm_debug_mixed_a)
static std::atomic<int> g_l2_hints{0};         // test hook (fvqa_gemm_debug_l2_hints)
static std::atomic<int> g_force_bn{0};        // test hook (fvqa_gemm_debug_force_bn): 0 = heuristic, -1 = single-CTA kernel only
static std::atomic<int> g_pair_threads{GEMM_THREADS};   // test hook (fvqa_gemm_debug_epilogue_warps): 256 = 4 epilogue warps (default), 384 = 8

// Output-tile width for the pair kernel: maximise wave efficiency x per-tile efficiency. The per-tile
// factors are MEASURED (B200, 3072 x 22016 x 4096, tools/gemm_diag.py): the UMMA 256 x N x 16 issue time
// barely shrinks with N, so narrow tiles only pay when they remove most of a wave.
static int choose_pair_bn(int M, int N) {
  static const int kBn[] = {256, 240, 224, 208, 192, 176, 160, 144, 128};
  static const double kTileEff[] = {1.0, 0.964, 0.872, 0.82, 0.778, 0.675, 0.62, 0.56, 0.50};
  const int pairs = g_num_sms / 2;
  const long tm = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  int best = 256;
  double best_score = -1.0;
  for (int i = 0; i < 9; ++i) {
    const int bn = kBn[i];
    const long tn = (N + bn - 1) / bn;
    const long tiles = tm * tn;
    const long waves = (tiles + pairs - 1) / pairs;
    const double eff = (static_cast<double>(M) * N) / (static_cast<double>(waves) * pairs * 2 * GEMM_BM * bn);
    const double score = eff * kTileEff[i];
    // The training step runs power-capped (sustained): padded columns of a narrower tile cost energy even when
    // they fill a wave (N = 4096: bn 240 is 1-2 % SLOWER than 256 sustained, profiles/r1_gemm_sustained_power.txt),
    // so a narrower tile must win clearly.
    if (score > best_score + (i == 0 ? 0.0 : 0.08)) { best_score = score; best = bn; }
  }
  return best;
}

template <bool OUT_F32, int EPI>
static int launch_gemm_pair(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, const GemmEpi& epi, int M,
                            int N, int K, int bn, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = get_tmap(A, M, K, lda, GEMM_BM, &ta);
  if (rc) return rc;
  rc = get_tmap(B, N, K, ldb, bn / 2, &tb);
  if (rc) return rc;
  const int stage_bytes = pair_stage_bytes(bn);
  int stages = (PAIR_SMEM_LIMIT - 1024 - PAIR_BAR_BYTES) / stage_bytes;
  if (stages > PAIR_MAX_STAGES) stages = PAIR_MAX_STAGES;
  const int smem = stages * stage_bytes + PAIR_BAR_BYTES + 1024;
  const int tiles_n = (EPI == EPI_SWIGLU_FWD) ? epi.hid / 128 : (N + bn - 1) / bn;
  const int tiles = ((M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * tiles_n;
  const int pairs = tiles < g_num_sms / 2 ? tiles : g_num_sms / 2;
  launch_k(gemm_nt_pair_kernel<OUT_F32, EPI, false>, dim3(2 * pairs), dim3(g_pair_threads.load()), smem, stream, ta, tb, C, epi, M, N, K, ldc, bn, stages,
           g_l2_hints | (g_mixed_a << 1));
  return check_launch("gemm_nt_pair");
}

// C = A . B with B row-major [K, N] (see B_MN): h16 out, plain or SwiGLU-backward epilogue, BN = 256; the 2x2-cluster variant where
// the pair schedule would end in a partial wave (same rule as the NT launcher).
static bool use_quad(int M, int N);
template <int EPI>
static int launch_gemm_nn(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, const GemmEpi& epi, int M, int N, int K,
                          cudaStream_t stream) {
  constexpr int bn = 256;
  const bool quad = (EPI == EPI_PLAIN) && use_quad(M, N);
  CUtensorMap ta, tb;
  int rc = get_tmap(A, M, K, lda, quad ? GEMM_BM / 2 : GEMM_BM, &ta);
  if (rc) return rc;
  rc = get_tmap(B, K, N, ldb, GEMM_BK, &tb);                     // boxes of [64 k rows][64 columns]
  if (rc) return rc;
  const int stage_bytes = pair_stage_bytes(bn);
  int stages = (PAIR_SMEM_LIMIT - 1024 - PAIR_BAR_BYTES) / stage_bytes;
  if (stages > PAIR_MAX_STAGES) stages = PAIR_MAX_STAGES;
  const int smem = stages * stage_bytes + PAIR_BAR_BYTES + 1024;
  const int tiles_m = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM), tiles_n = (N + bn - 1) / bn;
  if (quad) {
    if constexpr (EPI == EPI_PLAIN) {
      const int units = tiles_m * (N / (2 * bn));
      const int clusters = units < g_quad_clusters ? units : g_quad_clusters;
      launch_k(gemm_nt_pair_kernel<false, EPI_PLAIN, true, true>, dim3(4 * clusters), dim3(PAIR_THREADS), smem, stream, ta, tb, C, epi, M, N, K, ldc,
               bn, stages, 0);
      return check_launch("gemm_nn_quad");
    }
  }
  const int tiles = tiles_m * tiles_n;
  const int pairs = tiles < g_num_sms / 2 ? tiles : g_num_sms / 2;
  launch_k(gemm_nt_pair_kernel<false, EPI, false, true>, dim3(2 * pairs), dim3(g_pair_threads.load()), smem, stream, ta, tb, C, epi, M, N, K, ldc, bn,
           stages, 0);
  return check_launch("gemm_nn_pair");
}

// 2x2-cluster multicast variant (plain epilogue, BN = 256, an even number of column tiles)
template <bool OUT_F32>
static int launch_gemm_quad(const h16* A, int lda, const h16* B, int ldb, void* C, int ldc, const GemmEpi& epi, int M, int N, int K,
                            cudaStream_t stream) {
  constexpr int bn = 256;
  CUtensorMap ta, tb;
  int rc = get_tmap(A, M, K, lda, GEMM_BM / 2, &ta);          // 64-row boxes: each CTA loads half of the slice it shares
  if (rc) return rc;
  rc = get_tmap(B, N, K, ldb, bn / 2, &tb);
  if (rc) return rc;
  const int stage_bytes = pair_stage_bytes(bn);
  int stages = (PAIR_SMEM_LIMIT - 1024 - PAIR_BAR_BYTES) / stage_bytes;
  if (stages > PAIR_MAX_STAGES) stages = PAIR_MAX_STAGES;
  const int smem = stages * stage_bytes + PAIR_BAR_BYTES + 1024;
  const int units = ((M + 2 * GEMM_BM - 1) / (2 * GEMM_BM)) * (N / (2 * bn));
  const int clusters = units < g_quad_clusters ? units : g_quad_clusters;
  // 384 threads = two epilogue warps per TMEM lane quadrant: neutral for the pair kernel, but with the multicast operand stream the
  // faster accumulator drain pays (K = 11008 burst 191.7 vs 199.0 us with four warps; in-step A/B -0.5 % vs 0.0 %)
  launch_k(gemm_nt_pair_kernel<OUT_F32, EPI_PLAIN, true>, dim3(4 * clusters), dim3(PAIR_THREADS), smem, stream, ta, tb, C, epi, M, N, K, ldc, bn, stages, 0);
  return check_launch("gemm_nt_quad");
}
// Used when the pair schedule would end in a clearly partial wave and the cluster schedule does not (N = 4096 outputs of a
// 3072-row step: 192 tiles on 74 pairs = 86 % vs 96 units on 33 clusters = 97 %): measured sustained +0.6 % (K = 11008) to
// +2.5 % (K = 22016) over the pair kernel, -0.5 % step time; on shapes with full pair waves (N >= 11008) the 16 idle SMs of the
// cluster schedule cost 4-5 % instead (profiles/r1_gemm_quad_cluster.txt).
static bool use_quad(int M, int N) {
  if (g_quad_mode == 0 || g_quad_clusters <= 0 || g_force_bn != 0 || M <= GEMM_BM || N % 512 != 0) return false;
  if (g_quad_mode == 2) return true;
  const long tm = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const long tiles = tm * (N / 256), units = tm * (N / 512), pairs = g_num_sms / 2;
  const double eff_pair = static_cast<double>(tiles) / (static_cast<double>((tiles + pairs - 1) / pairs) * pairs);
  const double eff_quad = static_cast<double>(units) / (static_cast<double>((units + g_quad_clusters - 1) / g_quad_clusters) * g_quad_clusters);
  return eff_pair < 0.9 && eff_quad > eff_pair + 0.05;
}

// M > 128 rows: CTA-pair kernel; tiny-M problems (adapter prompts, a handful of labelled rows) stay on
// the single-CTA kernel.
static bool use_pair(int M, int N) { return g_force_bn >= 0 && M > GEMM_BM && N >= 64; }
static int pair_bn(int M, int N) {
  if (g_force_bn > 0) return g_force_bn.load();
  if (N < 128) return ((N + 15) / 16) * 16 < 64 ? 64 : ((N + 15) / 16) * 16;
  return choose_pair_bn(M, N);
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

extern "C" int fvqa_gemm_nt(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, void* C, int ldc,
                                 const void* R, int ldr, int M, int N, int K, int out_fp32, void* stream) {
  int rc = check_gemm_args(A, lda, B, ldb, C, ldc, R, ldr, M, N, K);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const h16* a = reinterpret_cast<const h16*>(A);
  const h16* b = reinterpret_cast<const h16*>(B);
  GemmEpi epi{R, ldr, nullptr, nullptr, 0, 0, 1, nullptr, 0, 0};
  if (g_force_bn == 0 && gemm_skinny_supported(M, K, R)) return gemm_skinny(a, lda, b, ldb, C, ldc, M, N, K, out_fp32, g_num_sms, s);
  // M <= 16 with the fp32 residual stream (decode steps of the generation evaluator): HBM-bound weight streaming on every SM
  // instead of N / 256 single-CTA tiles
  if (g_force_bn == 0 && out_fp32 && R != nullptr && gemm_skinny_shape_ok(M, N, K))
    return gemm_skinny_epi(a, lda, b, ldb, C, ldc, M, N, K, 1, reinterpret_cast<const float*>(R), ldr, nullptr, nullptr, 0, 0, 1, nullptr,
                           g_num_sms, s);
  if (use_quad(M, N)) {
    return out_fp32 ? launch_gemm_quad<true>(a, lda, b, ldb, C, ldc, epi, M, N, K, s) : launch_gemm_quad<false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
  }
  if (use_pair(M, N)) {
    const int bn = pair_bn(M, N);
    return out_fp32 ? launch_gemm_pair<true, EPI_PLAIN>(a, lda, b, ldb, C, ldc, epi, M, N, K, bn, s)
                    : launch_gemm_pair<false, EPI_PLAIN>(a, lda, b, ldb, C, ldc, epi, M, N, K, bn, s);
  }
  if (prefer_bn128(M, N)) {
    return out_fp32 ? launch_gemm<128, true, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s)
                    : launch_gemm<128, false, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
  }
  return out_fp32 ? launch_gemm<256, true, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s)
                  : launch_gemm<256, false, false>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
}

/* Grouped skinny GEMM (M <= 16): C_g[M,N] = A_g[M,K] * B_g[N,K]^T for g < groups in ONE launch; see include/fvqa.h. */
extern "C" int fvqa_gemm_skinny_grouped(const fvqa_h16* A, int64_t strideA, int lda, const void* const* B_ptrs_dev, int ldb, void* C,
                                        int64_t strideC, int ldc, int M, int N, int K, int groups, int out_fp32, void* stream) {
  FVQA_REQUIRE(g_encode != nullptr, FVQA_ERR_INVALID_ARG, "fvqa_init() has not been called");
  FVQA_REQUIRE(M > 0 && M <= 16 && N > 0 && K > 0 && groups > 0 && groups <= 65535, FVQA_ERR_INVALID_ARG,
               "gemm_skinny_grouped: M=%d (<= 16) N=%d K=%d groups=%d", M, N, K, groups);
  FVQA_REQUIRE(gemm_skinny_supported(M, K, nullptr) && N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && strideA % 8 == 0, FVQA_ERR_UNSUPPORTED,
               "gemm_skinny_grouped: K=%d must be a multiple of 256, N / lda / ldb / strideA multiples of 8", K);
  FVQA_REQUIRE(A != nullptr && B_ptrs_dev != nullptr && C != nullptr && lda >= K && ldb >= K && ldc >= N, FVQA_ERR_INVALID_ARG,
               "gemm_skinny_grouped: null pointer or leading dimension too small");
  return gemm_skinny_grouped(reinterpret_cast<const h16*>(A), static_cast<long>(strideA), lda, nullptr,
                             reinterpret_cast<const h16* const*>(B_ptrs_dev), ldb, C, static_cast<long>(strideC), ldc, M, N, K, groups,
                             out_fp32, g_num_sms, static_cast<cudaStream_t>(stream));
}

static int gemm_rope_impl(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc, int M, int N, int K,
                          const float* rope_cos, const float* rope_sin, int rope_cols, int hd, int S, const int32_t* pos_ids,
                          void* stream) {
  int rc = check_gemm_args(A, lda, B, ldb, C, ldc, nullptr, 0, M, N, K);
  if (rc) return rc;
  FVQA_REQUIRE((hd == 64 || hd == 128) && rope_cols % hd == 0 && rope_cols <= N && S > 0 && rope_cos && rope_sin,
               FVQA_ERR_INVALID_ARG, "gemm_rope: bad rope arguments (hd=%d rope_cols=%d S=%d)", hd, rope_cols, S);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const h16* a = reinterpret_cast<const h16*>(A);
  const h16* b = reinterpret_cast<const h16*>(B);
  GemmEpi epi{nullptr, 0, rope_cos, rope_sin, rope_cols, hd, S, nullptr, 0, 0, pos_ids};
  if (g_force_bn == 0 && gemm_skinny_shape_ok(M, N, K))
    return gemm_skinny_epi(a, lda, b, ldb, C, ldc, M, N, K, 0, nullptr, 0, rope_cos, rope_sin, rope_cols, hd, S, pos_ids, g_num_sms, s);
  if (use_pair(M, N)) return launch_gemm_pair<false, EPI_ROPE>(a, lda, b, ldb, C, ldc, epi, M, N, K, pair_bn(M, N), s);
  if (prefer_bn128(M, N)) return launch_gemm<128, false, true>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
  return launch_gemm<256, false, true>(a, lda, b, ldb, C, ldc, epi, M, N, K, s);
}

extern "C" int fvqa_gemm_nt_rope(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc, int M, int N,
                                      int K, const float* rope_cos, const float* rope_sin, int rope_cols, int hd, int S,
                                      void* stream) {
  return gemm_rope_impl(A, lda, B, ldb, C, ldc, M, N, K, rope_cos, rope_sin, rope_cols, hd, S, nullptr, stream);
}

/* Ragged / compacted token layouts (shared-prefix option scoring): row r is rotated by the angle of position pos_ids[r]. */
extern "C" int fvqa_gemm_nt_rope_pos(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc, int M,
                                          int N, int K, const float* rope_cos, const float* rope_sin, int rope_cols, int hd,
                                          const int32_t* pos_ids, void* stream) {
  FVQA_REQUIRE(pos_ids != nullptr, FVQA_ERR_INVALID_ARG, "gemm_rope_pos: pos_ids is NULL");
  return gemm_rope_impl(A, lda, B, ldb, C, ldc, M, N, K, rope_cos, rope_sin, rope_cols, hd, 1, pos_ids, stream);
}

/* Test / tuning hook: epilogue warps per CTA of the CTA-pair kernel, 4 (default) or 8 (two per TMEM lane quadrant).
 * Returns the previous value. */
extern "C" int fvqa_gemm_debug_epilogue_warps(int n) {
  const int prev = (g_pair_threads >> 5) - 4;
  if (n == 4 || n == 8) g_pair_threads = (4 + n) * 32;
  return prev;
}

/* Test / tuning hook for the 2x2-cluster multicast kernel (plain epilogue, M > 128, N % 512 == 0): 0 = never, 1 = when the pair
 * schedule would end in a partial wave (default), 2 = always when eligible. Returns the previous mode;
 * fvqa_gemm_quad_clusters() = how many 4-CTA clusters fit the device (0 = variant unavailable). */
extern "C" int fvqa_gemm_debug_quad(int mode) {
  const int prev = g_quad_mode;
  if (mode >= 0 && mode <= 2) g_quad_mode = mode;
  return prev;
}
extern "C" int fvqa_gemm_quad_clusters(void) { return g_quad_clusters; }

/* Tuning hook: the skinny (M <= 16) kernel's CTA covers 8 * nt output columns, nt in {1, 2, 4}; 0 restores the heuristic. */
extern "C" int fvqa_gemm_debug_skinny_nt(int nt) {
  const int prev = g_skinny_force_nt;
  g_skinny_force_nt = nt;
  return prev;
}

/* Test / tuning hook: force the CTA-pair tile width (multiple of 16 in [64,256]); 0 restores the
 * heuristic, -1 forces the single-CTA kernel. Returns the previous setting. */
extern "C" int fvqa_gemm_debug_force_bn(int bn) {
  const int prev = g_force_bn;
  if (bn == 0 || bn == -1 || (bn >= 64 && bn <= 256 && bn % 16 == 0)) g_force_bn = bn;
  return prev;
}

/* W1|W3 projection with SwiGLU in the epilogue (llama/model.py:142): g[M, 2*hid] = x W13^T (h16, saved for
 * backward) and c[M, hid] = silu(g[:, :hid]) * g[:, hid:] in one pass. W13 = [W1; W3] is [2*hid, K]. */
extern "C" int fvqa_gemm_swiglu_fwd(const fvqa_h16* X, int ldx, const fvqa_h16* W13, int ldw, fvqa_h16* G, int ldg, fvqa_h16* Cc,
                                    int ldcc, int M, int hid, int K, void* stream) {
  int rc = check_gemm_args(X, ldx, W13, ldw, G, ldg, nullptr, 0, M, 2 * hid, K);
  if (rc) return rc;
  FVQA_REQUIRE(hid % 128 == 0 && ldcc % 8 == 0 && ldcc >= hid && Cc != nullptr && (reinterpret_cast<uintptr_t>(Cc) & 15) == 0,
               FVQA_ERR_UNSUPPORTED, "gemm_swiglu_fwd: hid=%d must be a multiple of 128 (ldc=%d)", hid, ldcc);
  GemmEpi epi{nullptr, 0, nullptr, nullptr, 0, 0, 1, Cc, ldcc, hid};
  return launch_gemm_pair<false, EPI_SWIGLU_FWD>(reinterpret_cast<const h16*>(X), ldx, reinterpret_cast<const h16*>(W13), ldw, G, ldg, epi,
                                                 M, 2 * hid, K, 256, static_cast<cudaStream_t>(stream));
}

/* Backward of the above through W2 and the SwiGLU: dg[M, 2*hid] = swiglu'(g) . (dY W2t^T) where the
 * intermediate dc = dY W2t^T [M, hid] never leaves the SM. W2t is [hid, K] (the transposed w2). */
extern "C" int fvqa_gemm_swiglu_bwd(const fvqa_h16* dY, int ldy, const fvqa_h16* W2t, int ldw, const fvqa_h16* G, int ldg,
                                    fvqa_h16* dG, int lddg, int M, int hid, int K, void* stream) {
  int rc = check_gemm_args(dY, ldy, W2t, ldw, dG, lddg, nullptr, 0, M, hid, K);
  if (rc) return rc;
  FVQA_REQUIRE(hid % 32 == 0 && ldg % 8 == 0 && ldg >= 2 * hid && lddg >= 2 * hid && G != nullptr && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
               FVQA_ERR_UNSUPPORTED, "gemm_swiglu_bwd: hid=%d must be a multiple of 32 (ldg=%d lddg=%d)", hid, ldg, lddg);
  GemmEpi epi{nullptr, 0, nullptr, nullptr, 0, 0, 1, const_cast<fvqa_h16*>(G), ldg, hid};
  return launch_gemm_pair<false, EPI_SWIGLU_BWD>(reinterpret_cast<const h16*>(dY), ldy, reinterpret_cast<const h16*>(W2t), ldw, dG, lddg, epi,
                                                 M, hid, K, 256, static_cast<cudaStream_t>(stream));
}

/* C[M,N] = A[M,K] . B[K,N] with B ROW-MAJOR [K, N] (leading dimension ldb >= N): the dX-only backward dX = dY . W read straight from
 * the [out, in] weight of the forward pass (no transposed copy). h16 out. K % 64 == 0, N % 8 == 0. */
extern "C" int fvqa_gemm_nn(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc, int M, int N, int K, void* stream) {
  FVQA_REQUIRE(g_encode != nullptr, FVQA_ERR_INVALID_ARG, "fvqa_init() has not been called");
  FVQA_REQUIRE(M > 0 && N > 0 && K > 0, FVQA_ERR_INVALID_ARG, "gemm_nn: empty problem M=%d N=%d K=%d", M, N, K);
  FVQA_REQUIRE(K % GEMM_BK == 0 && N % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && lda >= K && ldb >= N && ldc >= N,
               FVQA_ERR_UNSUPPORTED, "gemm_nn: K=%d must be a multiple of 64, N / lda / ldb / ldc multiples of 8 (N=%d lda=%d ldb=%d ldc=%d)", K, N, lda,
               ldb, ldc);
  FVQA_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0, FVQA_ERR_INVALID_ARG,
               "gemm_nn: pointers must be 16-byte aligned");
  GemmEpi epi{nullptr, 0, nullptr, nullptr, 0, 0, 1, nullptr, 0, 0};
  return launch_gemm_nn<EPI_PLAIN>(reinterpret_cast<const h16*>(A), lda, reinterpret_cast<const h16*>(B), ldb, C, ldc, epi, M, N, K,
                                   static_cast<cudaStream_t>(stream));
}

/* fvqa_gemm_swiglu_bwd with W2 given as the forward's [K = d, hid] row-major weight instead of its transposed copy:
 * dG[M, 2*hid] = swiglu'(G) applied to dc = dY[M,K] . W2[K, hid]. hid % 32 == 0. */
extern "C" int fvqa_gemm_swiglu_bwd_nn(const fvqa_h16* dY, int ldy, const fvqa_h16* W2, int ldw, const fvqa_h16* G, int ldg, fvqa_h16* dG,
                                       int lddg, int M, int hid, int K, void* stream) {
  FVQA_REQUIRE(g_encode != nullptr, FVQA_ERR_INVALID_ARG, "fvqa_init() has not been called");
  FVQA_REQUIRE(M > 0 && hid > 0 && K > 0 && K % GEMM_BK == 0 && hid % 32 == 0 && ldy % 8 == 0 && ldw % 8 == 0 && ldg % 8 == 0 && lddg % 8 == 0 &&
                   ldy >= K && ldw >= hid && ldg >= 2 * hid && lddg >= 2 * hid && G != nullptr,
               FVQA_ERR_UNSUPPORTED, "gemm_swiglu_bwd_nn: M=%d hid=%d K=%d ldy=%d ldw=%d ldg=%d lddg=%d", M, hid, K, ldy, ldw, ldg, lddg);
  FVQA_REQUIRE(((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(dG)) & 15) == 0,
               FVQA_ERR_INVALID_ARG, "gemm_swiglu_bwd_nn: pointers must be 16-byte aligned");
  GemmEpi epi{nullptr, 0, nullptr, nullptr, 0, 0, 1, const_cast<fvqa_h16*>(G), ldg, hid};
  return launch_gemm_nn<EPI_SWIGLU_BWD>(reinterpret_cast<const h16*>(dY), ldy, reinterpret_cast<const h16*>(W2), ldw, dG, lddg, epi, M, hid, K,
                                        static_cast<cudaStream_t>(stream));
}

/* Tuning hook: 1 = TMA loads of the CTA-pair kernel carry L2 eviction hints (A evict_last, B evict_first). */
extern "C" int fvqa_gemm_debug_l2_hints(int on) {
  const int prev = g_l2_hints;
  g_l2_hints = on;
  return prev;
}

/* Probe: 1 = the A operand of the CTA-pair kernel is read in the other 16-bit format than B (see fvqa_debug.h). */
extern "C" int fvqa_gemm_debug_mixed_a(int on) {
  const int prev = g_mixed_a;
  g_mixed_a = on ? 1 : 0;
  return prev;
}
