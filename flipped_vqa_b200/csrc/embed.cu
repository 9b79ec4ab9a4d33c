// Input side of the step: visual projection (fp32, trainable), token-embedding gather with video
// injection for the three objective streams, and the matching gradient gathers.
// Reference: llama/model.py:286-336 (forward); backward derived in SURVEY.md §8(a) addendum.
#include "common.cuh"

namespace fvqa {

constexpr int VP_COLS = 8;      // output columns per CTA
constexpr int VP_THREADS = 256;

// vf32[r, c] = sum_k video[r, k] * wv[c, k] (+ bias[c]) (+ add[r, c])     fp32 Linear: model.py:322 (visual_proj, no bias);
// also the frozen audio projections / cross-attention q,k,v of the audio-fusion variants (model.py:307-320, :148-163)
__global__ void __launch_bounds__(VP_THREADS) visual_proj_fwd_kernel(const float* __restrict__ video,
                                                                     const float* __restrict__ wv,
                                                                     const float* __restrict__ bias, const float* __restrict__ add,
                                                                     float* __restrict__ vf, int rows, int dim, int vdim) {
  extern __shared__ float ws[];  // [VP_COLS][vdim]
  const int c0 = blockIdx.x * VP_COLS;
  for (int i = threadIdx.x; i < VP_COLS * vdim; i += VP_THREADS) {
    const int j = i / vdim, k = i - j * vdim;
    ws[i] = (c0 + j < dim) ? wv[static_cast<long>(c0 + j) * vdim + k] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < rows; r += VP_THREADS / 32) {
    float acc[VP_COLS];
#pragma unroll
    for (int j = 0; j < VP_COLS; ++j) acc[j] = 0.f;
    const float* vrow = video + static_cast<long>(r) * vdim;
    for (int k = lane; k < vdim; k += 32) {
      const float v = __ldg(vrow + k);
#pragma unroll
      for (int j = 0; j < VP_COLS; ++j) acc[j] += v * ws[j * vdim + k];
    }
#pragma unroll
    for (int j = 0; j < VP_COLS; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < VP_COLS; ++j)
        if (c0 + j < dim) {
          float o = acc[j];
          if (bias != nullptr) o += bias[c0 + j];
          if (add != nullptr) o += add[static_cast<long>(r) * dim + c0 + j];
          vf[static_cast<long>(r) * dim + c0 + j] = o;
        }
    }
  }
}

// Cross-attention of the 'attention' audio-fusion variant (CrossAttentionModule.forward, model.py:153-169):
// out[b, f, :] = softmax_j(<q[b, f], k[b, j]> / sqrt(D)) . v[b, j, :]   over the Fa audio tokens of sample b. One CTA per (b, f).
constexpr int XA_MAX_TOKENS = 64;
__global__ void __launch_bounds__(256) cross_attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                             const float* __restrict__ v, float* __restrict__ out, int F, int Fa, int D,
                                                             float scale) {
  __shared__ float red[32];
  __shared__ float p[XA_MAX_TOKENS];
  const int b = blockIdx.x / F;
  const float* qrow = q + static_cast<long>(blockIdx.x) * D;
  for (int j = 0; j < Fa; ++j) {
    const float* krow = k + (static_cast<long>(b) * Fa + j) * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s += qrow[i] * krow[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) p[j] = s * scale;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float m = p[0];
    for (int j = 1; j < Fa; ++j) m = fmaxf(m, p[j]);
    float z = 0.f;
    for (int j = 0; j < Fa; ++j) { p[j] = expf(p[j] - m); z += p[j]; }
    for (int j = 0; j < Fa; ++j) p[j] /= z;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float o = 0.f;
    for (int j = 0; j < Fa; ++j) o += p[j] * v[(static_cast<long>(b) * Fa + j) * D + i];
    out[static_cast<long>(blockIdx.x) * D + i] = o;
  }
}

// dwv[c, k] = sum_r dvf[r, c] * video[r, k]
__global__ void __launch_bounds__(VP_THREADS) visual_proj_bwd_kernel(const float* __restrict__ dvf,
                                                                     const float* __restrict__ video,
                                                                     float* __restrict__ dwv, int rows, int dim, int vdim) {
  const int c0 = blockIdx.x * VP_COLS;
  for (int k = threadIdx.x; k < vdim; k += VP_THREADS) {
    float acc[VP_COLS];
#pragma unroll
    for (int j = 0; j < VP_COLS; ++j) acc[j] = 0.f;
    for (int r = 0; r < rows; ++r) {
      const float v = __ldg(video + static_cast<long>(r) * vdim + k);
#pragma unroll
      for (int j = 0; j < VP_COLS; ++j)
        if (c0 + j < dim) acc[j] += __ldg(dvf + static_cast<long>(r) * dim + c0 + j) * v;
    }
#pragma unroll
    for (int j = 0; j < VP_COLS; ++j)
      if (c0 + j < dim) dwv[static_cast<long>(c0 + j) * vdim + k] = acc[j];
  }
}

// One CTA per token row; output is the fp32 residual stream (values are h16-representable).
__global__ void __launch_bounds__(256) build_h0_fwd_kernel(
    const h16* __restrict__ tok_emb, const int32_t* __restrict__ ids, const int32_t* __restrict__ labels,
    const int32_t* __restrict__ vstart, const int32_t* __restrict__ seq_video, const int32_t* __restrict__ qav_index,
    const float* __restrict__ vf32, const float* __restrict__ temporal, float* __restrict__ h0, int S, int dim, int F) {
  const int row = blockIdx.x;
  const int n = row / S, p = row - n * S;
  const int vs = vstart[n];
  const int b = seq_video[n];
  const int nvec = dim >> 3;
  const uint4* erow = reinterpret_cast<const uint4*>(tok_emb + static_cast<long>(ids[row]) * dim);
  float4* orow = reinterpret_cast<float4*>(h0 + static_cast<long>(row) * dim);
  auto put = [&](int v, const float (&o)[8]) {
    orow[2 * v] = make_float4(o[0], o[1], o[2], o[3]);
    orow[2 * v + 1] = make_float4(o[4], o[5], o[6], o[7]);
  };
  if (vs >= 0) {
    if (p >= vs && p < vs + F) {  // h[:, vs:vs+F] = (vf + temporal).half()   (model.py:324-332)
      const int f = p - vs;
      const float* vrow = vf32 + (static_cast<long>(b) * F + f) * dim;
      const float* trow = temporal + static_cast<long>(f) * dim;
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = h16_round(vrow[v * 8 + j] + trow[v * 8 + j]);
        put(v, o);
      }
    } else {
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float o[8];
        unpack8(__ldg(erow + v), o);
        put(v, o);
      }
    }
  } else {
    // QAV: h = emb * ~(label >= 0); h.scatter_add_(1, index, video_feature)   (model.py:335-336)
    const bool masked = labels[row] >= 0;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
      float o[8];
      if (masked) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      } else {
        unpack8(__ldg(erow + v), o);
      }
      for (int f = 0; f < F; ++f) {
        if (qav_index[b * F + f] == p) {
          const float* vrow = vf32 + (static_cast<long>(b) * F + f) * dim;
          const float* trow = temporal + static_cast<long>(f) * dim;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = h16_round(o[j] + h16_round(vrow[v * 8 + j] + trow[v * 8 + j]));
        }
      }
      put(v, o);
    }
  }
}

// One CTA per (video sample b, frame f): dvf[b,f,:] = sum over sequences of dh0 at that frame's slot.
__global__ void __launch_bounds__(256) build_h0_bwd_kernel(const float* __restrict__ dh0, const int32_t* __restrict__ vstart,
                                                            const int32_t* __restrict__ seq_video,
                                                            const int32_t* __restrict__ qav_index, float* __restrict__ dvf,
                                                            int n_seq, int S, int dim, int F) {
  const int b = blockIdx.x / F, f = blockIdx.x - b * F;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float acc = 0.f;
    for (int n = 0; n < n_seq; ++n) {
      if (seq_video[n] != b) continue;
      const int vs = vstart[n];
      const int pos = vs >= 0 ? vs + f : qav_index[b * F + f];
      if (pos < 0 || pos >= S) continue;
      acc += __ldg(dh0 + (static_cast<long>(n) * S + pos) * dim + c);
    }
    dvf[(static_cast<long>(b) * F + f) * dim + c] = acc;
  }
}

// One CTA per frame f: dtemporal[f] = sum_b dvf[b,f]; then dvf[b,f] += dvf_qav[b,f].
__global__ void __launch_bounds__(256) video_grad_finish_kernel(float* __restrict__ dvf, const float* __restrict__ dvf_qav,
                                                                 float* __restrict__ dtemporal, int n_video, int dim, int F) {
  const int f = blockIdx.x;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < n_video; ++b) {
      const long o = (static_cast<long>(b) * F + f) * dim + c;
      const float g = dvf[o];
      acc += g;
      if (dvf_qav) dvf[o] = g + dvf_qav[o];
    }
    dtemporal[static_cast<long>(f) * dim + c] = acc;
  }
}

}  // namespace fvqa

using namespace fvqa;

extern "C" int fvqa_linear_f32(const float* x, const float* w, const float* bias, const float* add, float* y, int rows, int dim,
                               int in_dim, void* stream) {
  if (rows <= 0) return FVQA_OK;
  const size_t smem = static_cast<size_t>(VP_COLS) * in_dim * sizeof(float);
  FVQA_REQUIRE(smem <= 96 * 1024, FVQA_ERR_UNSUPPORTED, "linear_f32: in_dim %d too large", in_dim);
  static bool big_smem_enabled = false;       // 768 + 1024 concatenated features need 56 KB (> the 48 KB default)
  if (smem > 48 * 1024 && !big_smem_enabled) {
    const cudaError_t e = cudaFuncSetAttribute(visual_proj_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    FVQA_REQUIRE(e == cudaSuccess, FVQA_ERR_CUDA, "cudaFuncSetAttribute(visual_proj_fwd): %s", cudaGetErrorString(e));
    big_smem_enabled = true;
  }
  visual_proj_fwd_kernel<<<(dim + VP_COLS - 1) / VP_COLS, VP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(x, w, bias, add, y, rows,
                                                                                                                dim, in_dim);
  return check_launch("linear_f32");
}

extern "C" int fvqa_visual_proj_fwd(const float* video, const float* wv, float* vf32, int rows, int dim, int vdim, void* stream) {
  return fvqa_linear_f32(video, wv, nullptr, nullptr, vf32, rows, dim, vdim, stream);
}

extern "C" int fvqa_cross_attn_fwd(const float* q, const float* k, const float* v, float* out, int n_samples, int frames, int tokens,
                                   int dim, void* stream) {
  FVQA_REQUIRE(tokens >= 1 && tokens <= XA_MAX_TOKENS, FVQA_ERR_UNSUPPORTED, "cross_attn: %d audio tokens (max %d)", tokens, XA_MAX_TOKENS);
  if (n_samples * frames <= 0) return FVQA_OK;
  cross_attn_fwd_kernel<<<n_samples * frames, 256, 0, static_cast<cudaStream_t>(stream)>>>(q, k, v, out, frames, tokens, dim,
                                                                                            rsqrtf(static_cast<float>(dim)));
  return check_launch("cross_attn_fwd");
}

extern "C" int fvqa_visual_proj_bwd(const float* dvf, const float* video, float* dwv, int rows, int dim, int vdim, void* stream) {
  visual_proj_bwd_kernel<<<(dim + VP_COLS - 1) / VP_COLS, VP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(dvf, video, dwv, rows, dim, vdim);
  return check_launch("visual_proj_bwd");
}

extern "C" int fvqa_build_h0_fwd(const fvqa_h16* tok_emb, const int32_t* ids, const int32_t* labels, const int32_t* vstart,
                                 const int32_t* seq_video, const int32_t* qav_index, const float* vf32, const float* temporal,
                                 float* h0, int n_seq, int S, int dim, int max_feats, void* stream) {
  FVQA_REQUIRE(dim % 8 == 0, FVQA_ERR_UNSUPPORTED, "build_h0: dim %d must be a multiple of 8", dim);
  if (n_seq * S <= 0) return FVQA_OK;
  build_h0_fwd_kernel<<<n_seq * S, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const h16*>(tok_emb), ids, labels, vstart, seq_video, qav_index, vf32, temporal, h0, S, dim, max_feats);
  return check_launch("build_h0_fwd");
}

extern "C" int fvqa_build_h0_bwd(const float* dh0, const int32_t* vstart, const int32_t* seq_video, const int32_t* qav_index,
                                 float* dvf, int n_seq, int n_video, int S, int dim, int max_feats, void* stream) {
  FVQA_REQUIRE(dim % 8 == 0, FVQA_ERR_UNSUPPORTED, "build_h0_bwd: dim %d must be a multiple of 8", dim);
  if (n_video * max_feats <= 0) return FVQA_OK;
  build_h0_bwd_kernel<<<n_video * max_feats, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dh0, vstart, seq_video, qav_index, dvf, n_seq, S, dim, max_feats);
  return check_launch("build_h0_bwd");
}

extern "C" int fvqa_video_grad_finish(float* dvf, const float* dvf_qav, float* dtemporal, int n_video, int dim, int max_feats,
                                      void* stream) {
  if (max_feats <= 0) return FVQA_OK;
  video_grad_finish_kernel<<<max_feats, 256, 0, static_cast<cudaStream_t>(stream)>>>(dvf, dvf_qav, dtemporal, n_video, dim, max_feats);
  return check_launch("video_grad_finish");
}
