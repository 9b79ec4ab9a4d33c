// Input side of the step: visual projection (fp32, trainable), token-embedding gather with video
// injection for the three objective streams, and the matching gradient gathers.
// Reference: llama/model.py:286-336 (forward); backward derived in SURVEY.md §8(a) addendum.
#include "common.cuh"

namespace fvqa {

constexpr int VP_THREADS = 256;
constexpr int LIN_COLS = 32;     // output columns per CTA (one per lane)
constexpr int LIN_KC = 64;       // k-chunk staged in shared memory
constexpr int LIN_ROWS = 128;    // rows per pass: warp w owns rows w, w + 8, ... (16 per warp)

// y[r, c] = sum_k x[r, k] * w[c, k] (+ bias[c]) (+ add[r, c])     fp32 Linear: model.py:322 (visual_proj, no bias); also the frozen
// audio projections / cross-attention q,k,v of the audio-fusion variants (model.py:307-320, :148-163).
// CTA = 32 output columns x up to 128 rows; the k dimension goes through shared memory in chunks of 64: w tile transposed to
// [k][c] (lane = column: conflict-free), x tile [row][k] read as float4 broadcasts; each lane keeps 16 row accumulators.
// grid = (dim / 32, row passes): 128 CTAs at d = 4096 (was: 512 CTAs whose warps each reduced 8 columns per row with shuffles).
template <bool VEC>
__global__ void __launch_bounds__(VP_THREADS) linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, const float* __restrict__ add,
                                                                float* __restrict__ y, int rows, int dim, int in_dim) {
  __shared__ float wt[LIN_KC][LIN_COLS + 1];
  __shared__ __align__(16) float xs[LIN_ROWS][LIN_KC];
  const int c0 = blockIdx.x * LIN_COLS, r0 = blockIdx.y * LIN_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nrows = min(LIN_ROWS, rows - r0);
  float acc[LIN_ROWS / 8];
#pragma unroll
  for (int i = 0; i < LIN_ROWS / 8; ++i) acc[i] = 0.f;
  // VEC (in_dim % 4 == 0, 16-byte aligned rows): the next chunk's global loads (2 + 8 float4 per thread) are in flight in
  // registers while the current chunk is multiplied - with one CTA per SM nothing else hides the DRAM latency
  constexpr int WV = LIN_COLS * LIN_KC / 4 / VP_THREADS, XV = LIN_ROWS * LIN_KC / 4 / VP_THREADS;
  float4 wreg[WV], xreg[XV];
  auto gload = [&](int k0) {
#pragma unroll
    for (int j = 0; j < WV; ++j) {
      const int idx = threadIdx.x + j * VP_THREADS, c = idx / (LIN_KC / 4), k4 = idx % (LIN_KC / 4);
      wreg[j] = (c0 + c < dim && k0 + 4 * k4 < in_dim) ? __ldg(reinterpret_cast<const float4*>(w + static_cast<long>(c0 + c) * in_dim + k0) + k4)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < XV; ++j) {
      const int idx = threadIdx.x + j * VP_THREADS, r = idx / (LIN_KC / 4), k4 = idx % (LIN_KC / 4);
      xreg[j] = (r < nrows && k0 + 4 * k4 < in_dim) ? __ldg(reinterpret_cast<const float4*>(x + static_cast<long>(r0 + r) * in_dim + k0) + k4)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int j = 0; j < WV; ++j) {
      const int idx = threadIdx.x + j * VP_THREADS, c = idx / (LIN_KC / 4), k4 = idx % (LIN_KC / 4);
      wt[4 * k4][c] = wreg[j].x; wt[4 * k4 + 1][c] = wreg[j].y; wt[4 * k4 + 2][c] = wreg[j].z; wt[4 * k4 + 3][c] = wreg[j].w;
    }
#pragma unroll
    for (int j = 0; j < XV; ++j) {
      const int idx = threadIdx.x + j * VP_THREADS, r = idx / (LIN_KC / 4), k4 = idx % (LIN_KC / 4);
      *reinterpret_cast<float4*>(&xs[r][4 * k4]) = xreg[j];
    }
  };
  if constexpr (VEC) gload(0);
  for (int k0 = 0; k0 < in_dim; k0 += LIN_KC) {
    __syncthreads();
    if constexpr (VEC) {
      sstore();
    } else {
      for (int i = threadIdx.x; i < LIN_COLS * LIN_KC; i += VP_THREADS) {        // coalesced along k, transposed store (stride 33)
        const int c = i / LIN_KC, k = i - c * LIN_KC;
        wt[k][c] = (c0 + c < dim && k0 + k < in_dim) ? __ldg(w + static_cast<long>(c0 + c) * in_dim + k0 + k) : 0.f;
      }
      for (int i = threadIdx.x; i < nrows * LIN_KC; i += VP_THREADS) {
        const int r = i / LIN_KC, k = i - r * LIN_KC;
        xs[r][k] = (k0 + k < in_dim) ? __ldg(x + static_cast<long>(r0 + r) * in_dim + k0 + k) : 0.f;
      }
    }
    __syncthreads();
    if constexpr (VEC)
      if (k0 + LIN_KC < in_dim) gload(k0 + LIN_KC);
#pragma unroll 4
    for (int k = 0; k < LIN_KC; k += 4) {
      const float w0 = wt[k][lane], w1 = wt[k + 1][lane], w2 = wt[k + 2][lane], w3 = wt[k + 3][lane];
#pragma unroll
      for (int i = 0; i < LIN_ROWS / 8; ++i) {
        const int r = warp + 8 * i;
        if (r < nrows) {                                                         // warp-uniform
          const float4 xv = *reinterpret_cast<const float4*>(&xs[r][k]);
          acc[i] += xv.x * w0 + xv.y * w1 + xv.z * w2 + xv.w * w3;
        }
      }
    }
  }
  const int c = c0 + lane;
  if (c < dim) {
    const float bv = bias != nullptr ? bias[c] : 0.f;
#pragma unroll
    for (int i = 0; i < LIN_ROWS / 8; ++i) {
      const int r = warp + 8 * i;
      if (r < nrows) {
        float o = acc[i] + bv;
        if (add != nullptr) o += add[static_cast<long>(r0 + r) * dim + c];
        y[static_cast<long>(r0 + r) * dim + c] = o;
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

// dwv[c, k] = sum_r dvf[r, c] * video[r, k]   (fp32 outer-product accumulation, fixed row order -> deterministic)
// CTA = 32 c x 64 k output tile; rows go through shared memory in chunks of 64; thread = one k and 8 consecutive c
// (x tile: conflict-free, d tile: float4 broadcasts). grid = (dim / 32, vdim / 64): 1536 CTAs at 4096 x 768 (was: 512 CTAs whose
// threads walked all rows with 9 global loads per 8 FMAs, 122 us).
constexpr int VPB_C = 32, VPB_K = 64, VPB_R = 64;
__global__ void __launch_bounds__(VP_THREADS) visual_proj_bwd_kernel(const float* __restrict__ dvf,
                                                                     const float* __restrict__ video,
                                                                     float* __restrict__ dwv, int rows, int dim, int vdim) {
  __shared__ __align__(16) float ds[VPB_R][VPB_C];
  __shared__ float vs[VPB_R][VPB_K];
  const int c0 = blockIdx.x * VPB_C, k0 = blockIdx.y * VPB_K;
  const int kk = threadIdx.x & (VPB_K - 1), cg = (threadIdx.x >> 6) * 8;        // 4 groups of 8 columns
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int r0 = 0; r0 < rows; r0 += VPB_R) {
    const int n = min(VPB_R, rows - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < n * VPB_C; i += VP_THREADS) {
      const int r = i / VPB_C, c = i - r * VPB_C;
      ds[r][c] = (c0 + c < dim) ? __ldg(dvf + static_cast<long>(r0 + r) * dim + c0 + c) : 0.f;
    }
    for (int i = threadIdx.x; i < n * VPB_K; i += VP_THREADS) {
      const int r = i / VPB_K, k = i - r * VPB_K;
      vs[r][k] = (k0 + k < vdim) ? __ldg(video + static_cast<long>(r0 + r) * vdim + k0 + k) : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < n; ++r) {
      const float v = vs[r][kk];
      const float4 d0 = *reinterpret_cast<const float4*>(&ds[r][cg]), d1 = *reinterpret_cast<const float4*>(&ds[r][cg + 4]);
      acc[0] += d0.x * v; acc[1] += d0.y * v; acc[2] += d0.z * v; acc[3] += d0.w * v;
      acc[4] += d1.x * v; acc[5] += d1.y * v; acc[6] += d1.z * v; acc[7] += d1.w * v;
    }
  }
  if (k0 + kk < vdim) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + cg + j < dim) dwv[static_cast<long>(c0 + cg + j) * vdim + k0 + kk] = acc[j];
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

// build_h0_bwd + video_grad_finish in ONE full-grid launch: for frame f and a 256-column slab, walk the video samples b:
//   g = sum over the sequences of sample b of dh0 at frame f's slot;  dtemporal[f] += g;  dvf[b, f] = g + dvf_qav[b, f].
// grid (F, dim / 256) = 160 CTAs at d = 4096 (the two separate kernels ran on 80 and 10 CTAs: 56 + 61 us).
__global__ void __launch_bounds__(256) video_grad_kernel(const float* __restrict__ dh0, const int32_t* __restrict__ vstart,
                                                         const int32_t* __restrict__ seq_video, const int32_t* __restrict__ qav_index,
                                                         const float* __restrict__ dvf_qav, float* __restrict__ dvf,
                                                         float* __restrict__ dtemporal, int n_seq, int n_video, int S, int dim, int F) {
  const int f = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x;
  if (c >= dim) return;
  float tsum = 0.f;
  for (int b = 0; b < n_video; ++b) {
    float g = 0.f;
    for (int n = 0; n < n_seq; ++n) {
      if (seq_video[n] != b) continue;
      const int vs = vstart[n];
      const int pos = vs >= 0 ? vs + f : qav_index[b * F + f];
      if (pos < 0 || pos >= S) continue;
      g += __ldg(dh0 + (static_cast<long>(n) * S + pos) * dim + c);
    }
    tsum += g;
    const long o = (static_cast<long>(b) * F + f) * dim + c;
    dvf[o] = dvf_qav != nullptr ? g + dvf_qav[o] : g;
  }
  dtemporal[static_cast<long>(f) * dim + c] = tsum;
}

}  // namespace fvqa

using namespace fvqa;

extern "C" int fvqa_video_grad(const float* dh0, const int32_t* vstart, const int32_t* seq_video, const int32_t* qav_index,
                               const float* dvf_qav, float* dvf, float* dtemporal, int n_seq, int n_video, int S, int dim, int max_feats,
                               void* stream) {
  if (max_feats <= 0 || dim <= 0) return FVQA_OK;
  video_grad_kernel<<<dim3(max_feats, (dim + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dh0, vstart, seq_video, qav_index, dvf_qav, dvf, dtemporal, n_seq, n_video, S, dim, max_feats);
  return check_launch("video_grad");
}

extern "C" int fvqa_linear_f32(const float* x, const float* w, const float* bias, const float* add, float* y, int rows, int dim,
                               int in_dim, void* stream) {
  if (rows <= 0) return FVQA_OK;
  FVQA_REQUIRE(dim > 0 && in_dim > 0, FVQA_ERR_INVALID_ARG, "linear_f32: dim %d in_dim %d", dim, in_dim);
  const dim3 grid((dim + LIN_COLS - 1) / LIN_COLS, (rows + LIN_ROWS - 1) / LIN_ROWS);
  const bool vec = in_dim % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
  if (vec) linear_f32_kernel<true><<<grid, VP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, add, y, rows, dim, in_dim);
  else linear_f32_kernel<false><<<grid, VP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, add, y, rows, dim, in_dim);
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
  if (dim <= 0 || vdim <= 0) return FVQA_OK;
  const dim3 grid((dim + VPB_C - 1) / VPB_C, (vdim + VPB_K - 1) / VPB_K);
  visual_proj_bwd_kernel<<<grid, VP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(dvf, video, dwv, rows, dim, vdim);
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
