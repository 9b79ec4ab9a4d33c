// Shared declarations of the attention kernels (attention.cu: mma.sync path for any S / hd in {64,128};
// attention_tc.cu: tcgen05/TMEM path for S <= 128, hd = 128).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace fvqa {

constexpr int AT_AP = 16;   // adapter keys padded to one MMA k-block

struct AttnParams {
  const h16* qkv; const h16* akv; int akv_ld;
  const float* cosT; const float* sinT; const float* gate1; const float* gate2; const int32_t* vstart;
  h16* out; float* lse;                       // fwd outputs / bwd inputs
  const h16* dout; h16* dqkv;       // bwd
  float* ws_dx; float* ws_gate; float* ws_akv;
  int n_seq, S, H, A, F, qblocks;                       // qblocks = ceil(S / 128)
};

// tcgen05 path (attention_tc.cu). `*_supported` says whether the shape is handled; the launchers
// return FVQA_OK or an error code.
bool attn_tc_supported(int S, int hd, int A);
int attn_tc_init();
int attn_fwd_tc(const AttnParams& p, cudaStream_t stream);
int attn_bwd_tc(const AttnParams& p, cudaStream_t stream);
// tiled tcgen05 path for S > 128, hd = 128 (attention_tc_long.cu)
bool attn_tcl_supported(int S, int hd, int A);
int attn_tcl_init();
int attn_fwd_tcl(const AttnParams& p, cudaStream_t stream);
int attn_bwd_tcl(const AttnParams& p, cudaStream_t stream);   // dq + dkv kernels; caller runs the reduce kernel

}  // namespace fvqa
