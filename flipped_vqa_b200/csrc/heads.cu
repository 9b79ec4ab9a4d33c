// Loss heads: vocabulary cross-entropy over labelled rows (fwd/bwd), QAV video-feature
// reconstruction loss (fwd/bwd), per-token loss scatter and multiple-choice option scoring.
// Reference: llama/model.py:347-361, llama/model_my_original_mod.py:375-377, engine.py:88-93.
#include "common.cuh"

namespace fvqa {

constexpr int CE_THREADS = 512;
constexpr int QAV_MAXF = 16;

// row_loss = logsumexp(logits[row]) - logits[row, target]; one CTA per row, online max/sum.
__global__ void __launch_bounds__(CE_THREADS) ce_fwd_kernel(const float* __restrict__ logits, int ld,
                                                            const int32_t* __restrict__ target, float* __restrict__ row_loss,
                                                            float* __restrict__ row_lse, int V) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int t = target[row];
  if (t < 0) {
    if (threadIdx.x == 0) { row_loss[row] = 0.f; row_lse[row] = 0.f; }
    return;
  }
  const float* l = logits + static_cast<long>(row) * ld;
  float m = -INFINITY, s = 0.f;
  const int nv = V >> 2;
  const float4* l4 = reinterpret_cast<const float4*>(l);
  for (int i = threadIdx.x; i < nv; i += CE_THREADS) {
    const float4 v = __ldg(l4 + i);
    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (mx > m) { s *= __expf(m - mx); m = mx; }
    s += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  }
  for (int i = (nv << 2) + threadIdx.x; i < V; i += CE_THREADS) {
    const float v = l[i];
    if (v > m) { s *= __expf(m - v); m = v; }
    s += __expf(v - m);
  }
  const float gm = block_max(m, red);
  s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
  const float gs = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float lse = gm + logf(gs);
    row_lse[row] = lse;
    row_loss[row] = lse - l[t];
  }
}

// dlogits = (softmax - onehot) * gscale * inv_count  (h16); padding rows are zero-filled.
__global__ void __launch_bounds__(CE_THREADS) ce_bwd_kernel(const float* __restrict__ logits, int ld,
                                                            const int32_t* __restrict__ target, const float* __restrict__ row_lse,
                                                            const float* __restrict__ gscale, float inv_count,
                                                            h16* __restrict__ dlogits, int ldd, int V) {
  const int row = blockIdx.x;
  const int t = target[row];
  h16* d = dlogits + static_cast<long>(row) * ldd;
  const int nv = V >> 3;
  if (t < 0) {
    for (int i = threadIdx.x; i < nv; i += CE_THREADS) reinterpret_cast<uint4*>(d)[i] = make_uint4(0, 0, 0, 0);
    for (int i = (nv << 3) + threadIdx.x; i < V; i += CE_THREADS) d[i] = f2h(0.f);
    return;
  }
  const float* l = logits + static_cast<long>(row) * ld;
  const float lse = row_lse[row];
  const float sc = gscale[0] * inv_count;
  for (int i = threadIdx.x; i < nv; i += CE_THREADS) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(l) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(l) + 2 * i + 1);
    float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (__expf(o[j] - lse) - ((i * 8 + j) == t ? 1.f : 0.f)) * sc;
    reinterpret_cast<uint4*>(d)[i] = pack8(o);
  }
  for (int i = (nv << 3) + threadIdx.x; i < V; i += CE_THREADS)
    d[i] = f2h((__expf(l[i] - lse) - (i == t ? 1.f : 0.f)) * sc);
}

// Deterministic single-CTA reduction: out[0] = scale * sum(v[0..rows)).
__global__ void __launch_bounds__(256) sum_scale_kernel(const float* __restrict__ v, int rows, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < rows; i += 256) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// QAV forward: one CTA per gathered row, warp j computes logit j = <hn[row], vf32[b, j, :]> / tau (F warps), then one thread does
// the F-way softmax. (Was: every thread strided over the row for all F frames and F block reductions in sequence - 96 us for 80 rows.)
__global__ void __launch_bounds__(32 * QAV_MAXF) qav_fwd_kernel(const h16* __restrict__ hn, const float* __restrict__ vf32,
                                                                 const int32_t* __restrict__ row_video, const int32_t* __restrict__ target,
                                                                 float inv_tau, float* __restrict__ row_loss, float* __restrict__ prob,
                                                                 int dim, int F) {
  __shared__ float logit[QAV_MAXF];
  const int row = blockIdx.x;
  const int b = row_video[row];
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (b < 0) {
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    if (threadIdx.x < F) prob[row * F + threadIdx.x] = 0.f;
    return;
  }
  const uint4* h = reinterpret_cast<const uint4*>(hn + static_cast<long>(row) * dim);
  const float* vb = vf32 + (static_cast<long>(b) * F + j) * dim;
  float acc = 0.f;
  for (int v = lane; v < (dim >> 3); v += 32) {
    float x[8];
    unpack8(__ldg(h + v), x);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(vb + v * 8)), a1 = __ldg(reinterpret_cast<const float4*>(vb + v * 8) + 1);
    acc += x[0] * a0.x + x[1] * a0.y + x[2] * a0.z + x[3] * a0.w + x[4] * a1.x + x[5] * a1.y + x[6] * a1.z + x[7] * a1.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) logit[j] = acc * inv_tau;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int q = 0; q < F; ++q) m = fmaxf(m, logit[q]);
    float z = 0.f;
    for (int q = 0; q < F; ++q) z += expf(logit[q] - m);
    const float lse = m + logf(z);
    for (int q = 0; q < F; ++q) prob[row * F + q] = expf(logit[q] - lse);
    row_loss[row] = lse - logit[target[row]];
  }
}

// dhn[row, c] = sum_j dlogit[row, j] * vf32[b, j, c]: grid (row, 2048-column slab), 8 columns per thread
__global__ void __launch_bounds__(256) qav_bwd_dh_kernel(const float* __restrict__ vf32, const int32_t* __restrict__ row_video,
                                                         const int32_t* __restrict__ target, const float* __restrict__ prob,
                                                         const float* __restrict__ gscale, float coef, h16* __restrict__ dhn,
                                                         int dim, int F) {
  const int row = blockIdx.x;
  const int v = blockIdx.y * 256 + threadIdx.x;               // 8-column vector index
  if (v >= (dim >> 3)) return;
  const int b = row_video[row];
  uint4* o = reinterpret_cast<uint4*>(dhn + static_cast<long>(row) * dim) + v;
  if (b < 0) {
    *o = make_uint4(0, 0, 0, 0);
    return;
  }
  const float sc = gscale[0] * coef;
  const int tg = target[row];
  float a[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = 0.f;
  const float* vb = vf32 + static_cast<long>(b) * F * dim + v * 8;
  for (int j = 0; j < F; ++j) {
    const float dl = (prob[row * F + j] - (j == tg ? 1.f : 0.f)) * sc;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(vb + static_cast<long>(j) * dim)), x1 = __ldg(reinterpret_cast<const float4*>(vb + static_cast<long>(j) * dim) + 1);
    a[0] += dl * x0.x; a[1] += dl * x0.y; a[2] += dl * x0.z; a[3] += dl * x0.w;
    a[4] += dl * x1.x; a[5] += dl * x1.y; a[6] += dl * x1.z; a[7] += dl * x1.w;
  }
  *o = pack8(a);
}

// dvf_qav[b, j, c] = sum_{rows r of sample b} dlogit[r, j] * hn[r, c] (fixed row order -> deterministic): grid (b * F + j,
// 256-column slab), one column per thread; the rows of the sample and their dlogit are found once per CTA (shared memory) instead
// of by every thread for every column (was 168 us for 0.7 MB on 80 CTAs).
constexpr int QAV_ROW_CHUNK = 256;
__global__ void __launch_bounds__(256) qav_bwd_dv_kernel(const h16* __restrict__ hn, const int32_t* __restrict__ row_video,
                                                         const int32_t* __restrict__ target, const float* __restrict__ prob,
                                                         const float* __restrict__ gscale, float coef, float* __restrict__ dvf,
                                                         int rows, int dim, int F) {
  __shared__ float dl_s[QAV_ROW_CHUNK];
  __shared__ int hit_s[QAV_ROW_CHUNK];
  const int b = blockIdx.x / F, j = blockIdx.x - b * F;
  const int c = blockIdx.y * 256 + threadIdx.x;
  const float sc = gscale[0] * coef;
  float a = 0.f;
  for (int r0 = 0; r0 < rows; r0 += QAV_ROW_CHUNK) {
    const int r = r0 + threadIdx.x;
    __syncthreads();
    if (r < rows) {
      const bool hit = row_video[r] == b;
      hit_s[threadIdx.x] = hit;
      dl_s[threadIdx.x] = hit ? (prob[r * F + j] - (j == target[r] ? 1.f : 0.f)) * sc : 0.f;
    }
    __syncthreads();
    const int n = min(QAV_ROW_CHUNK, rows - r0);
    if (c < dim)
      for (int q = 0; q < n; ++q)
        if (hit_s[q]) a += dl_s[q] * h2f(hn[static_cast<long>(r0 + q) * dim + c]);
  }
  if (c < dim) dvf[(static_cast<long>(b) * F + j) * dim + c] = a;
}

__global__ void scatter_rows_kernel(const float* __restrict__ v, const int32_t* __restrict__ idx, float* __restrict__ dst, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows && idx[i] >= 0) dst[idx[i]] = v[i];
}

// engine.py:88-93: count = (loss != 0).sum(-1); prediction = (loss.sum(-1) / count).argmin(-1)
__global__ void option_score_kernel(const float* __restrict__ tok, int32_t* __restrict__ pred, float* __restrict__ mean_loss,
                                    int n_items, int n_opt, int len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  int best = 0;
  float bestv = 0.f;
  bool have = false, nan_hit = false;
  for (int o = 0; o < n_opt; ++o) {
    const float* t = tok + (static_cast<long>(i) * n_opt + o) * len;
    float s = 0.f;
    int cnt = 0;
    for (int k = 0; k < len; ++k) {
      s += t[k];
      cnt += (t[k] != 0.f);
    }
    const float m = s / static_cast<float>(cnt);
    if (mean_loss) mean_loss[i * n_opt + o] = m;
    if (nan_hit) continue;
    if (m != m) { best = o; nan_hit = true; continue; }  // torch.argmin propagates NaN
    if (!have || m < bestv) { best = o; bestv = m; have = true; }
  }
  pred[i] = best;
}

}  // namespace fvqa

using namespace fvqa;

extern "C" int fvqa_ce_fwd(const float* logits, int ld, const int32_t* target, float* row_loss, float* row_lse, int rows, int V,
                           void* stream) {
  FVQA_REQUIRE(ld % 4 == 0, FVQA_ERR_UNSUPPORTED, "ce_fwd: ld %d must be a multiple of 4", ld);
  if (rows <= 0) return FVQA_OK;
  ce_fwd_kernel<<<rows, CE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, target, row_loss, row_lse, V);
  return check_launch("ce_fwd");
}

extern "C" int fvqa_ce_bwd(const float* logits, int ld, const int32_t* target, const float* row_lse, const float* gscale_dev,
                           float inv_count, fvqa_h16* dlogits, int ldd, int rows, int V, void* stream) {
  FVQA_REQUIRE(ld % 4 == 0 && ldd % 8 == 0, FVQA_ERR_UNSUPPORTED, "ce_bwd: ld %d / ldd %d alignment", ld, ldd);
  if (rows <= 0) return FVQA_OK;
  ce_bwd_kernel<<<rows, CE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, target, row_lse, gscale_dev, inv_count,
                                                                             reinterpret_cast<h16*>(dlogits), ldd, V);
  return check_launch("ce_bwd");
}

extern "C" int fvqa_sum_scale(const float* v, int rows, float scale, float* out, void* stream) {
  sum_scale_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, rows, scale, out);
  return check_launch("sum_scale");
}

extern "C" int fvqa_qav_loss_fwd(const fvqa_h16* hn, const float* vf32, const int32_t* row_video, const int32_t* target, float tau,
                                 float* row_loss, float* prob, int rows, int dim, int max_feats, void* stream) {
  FVQA_REQUIRE(max_feats <= QAV_MAXF, FVQA_ERR_UNSUPPORTED, "qav_loss: max_feats %d > %d", max_feats, QAV_MAXF);
  if (rows <= 0) return FVQA_OK;
  FVQA_REQUIRE(dim % 8 == 0 && max_feats >= 1, FVQA_ERR_UNSUPPORTED, "qav_loss: dim %d must be a multiple of 8", dim);
  qav_fwd_kernel<<<rows, 32 * max_feats, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const h16*>(hn), vf32, row_video, target,
                                                                                  1.f / tau, row_loss, prob, dim, max_feats);
  return check_launch("qav_loss_fwd");
}

extern "C" int fvqa_qav_loss_bwd(const fvqa_h16* hn, const float* vf32, const int32_t* row_video, const int32_t* target,
                                 const float* prob, const float* gscale_dev, float inv_count, float tau, fvqa_h16* dhn,
                                 float* dvf_qav, int rows, int n_video, int dim, int max_feats, void* stream) {
  FVQA_REQUIRE(max_feats <= QAV_MAXF, FVQA_ERR_UNSUPPORTED, "qav_loss: max_feats %d > %d", max_feats, QAV_MAXF);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float coef = inv_count / tau;
  if (rows > 0) {
    FVQA_REQUIRE(dim % 8 == 0, FVQA_ERR_UNSUPPORTED, "qav_loss: dim %d must be a multiple of 8", dim);
    qav_bwd_dh_kernel<<<dim3(rows, (dim / 8 + 255) / 256), 256, 0, s>>>(vf32, row_video, target, prob, gscale_dev, coef,
                                                                         reinterpret_cast<h16*>(dhn), dim, max_feats);
    int rc = check_launch("qav_loss_bwd(dh)");
    if (rc) return rc;
  }
  if (n_video * max_feats > 0) {
    qav_bwd_dv_kernel<<<dim3(n_video * max_feats, (dim + 255) / 256), 256, 0, s>>>(reinterpret_cast<const h16*>(hn), row_video, target, prob,
                                                                                    gscale_dev, coef, dvf_qav, rows, dim, max_feats);
    return check_launch("qav_loss_bwd(dv)");
  }
  return FVQA_OK;
}

namespace fvqa {
// Greedy decoding step of the generation evaluator (llama/model.py:429-467): for row b, tok = argmax_v logits[b, v] (lowest index
// on ties, like torch.max), written to ids[b, pos[b] + 1] and out_tokens[b, step]; x_next[b, :] = tok_emb[tok, :] (fp32 residual
// stream of the next decode step); margin[b] (optional) = best - second best logit. One CTA per row.
__global__ void __launch_bounds__(256) greedy_next_kernel(const float* __restrict__ logits, int ld, int V, const h16* __restrict__ tok_emb,
                                                          int dim, int32_t* __restrict__ ids, int S, const int32_t* __restrict__ pos,
                                                          int32_t* __restrict__ out_tokens, int out_ld, int step, float* __restrict__ x_next,
                                                          float* __restrict__ margin) {
  __shared__ float bv[8], sv[8];
  __shared__ int bi[8];
  __shared__ int tok_s;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* l = logits + static_cast<long>(b) * ld;
  float best = -INFINITY, second = -INFINITY;
  int arg = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float x = l[v];
    if (x > best || (x == best && v < arg)) { second = best; best = x; arg = v; }
    else if (x > second) second = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { second = fmaxf(best, os); best = ob; arg = oa; }
    else second = fmaxf(second, ob);
  }
  if (lane == 0) { bv[warp] = best; sv[warp] = second; bi[warp] = arg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    best = bv[0]; second = sv[0]; arg = bi[0];
    for (int w = 1; w < 8; ++w) {
      if (bv[w] > best || (bv[w] == best && bi[w] < arg)) { second = fmaxf(best, sv[w]); best = bv[w]; arg = bi[w]; }
      else second = fmaxf(second, bv[w]);
    }
    const int p = pos[b];
    if (p + 1 < S) ids[static_cast<long>(b) * S + p + 1] = arg;
    out_tokens[static_cast<long>(b) * out_ld + step] = arg;
    if (margin != nullptr) margin[static_cast<long>(b) * out_ld + step] = best - second;
    tok_s = arg;
  }
  __syncthreads();
  const uint4* e = reinterpret_cast<const uint4*>(tok_emb + static_cast<long>(tok_s) * dim);
  float4* o = reinterpret_cast<float4*>(x_next + static_cast<long>(b) * dim);
  for (int v = threadIdx.x; v < (dim >> 3); v += 256) {
    float f[8];
    unpack8(__ldg(e + v), f);
    o[2 * v] = make_float4(f[0], f[1], f[2], f[3]);
    o[2 * v + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// One thread: k = 2^round(log2(target / max|g|)), gs_out = g * k, inv_k = 1 / k (see include/fvqa.h).
__global__ void grad_scale_prepare_kernel(const float* __restrict__ g, float target, float* __restrict__ gs, float* __restrict__ inv_k) {
  const float m = fmaxf(fabsf(g[0]), fmaxf(fabsf(g[1]), fabsf(g[2])));
  float k = 1.f;
  if (target > 0.f && m > 0.f && isfinite(m)) {
    int e = static_cast<int>(rintf(log2f(target / m)));
    e = max(-60, min(60, e));
    k = exp2f(static_cast<float>(e));
  }
  gs[0] = g[0] * k; gs[1] = g[1] * k; gs[2] = g[2] * k;
  inv_k[0] = 1.f / k;
}
__global__ void scale_f32_kernel(float* __restrict__ x, const float* __restrict__ factor, long n) {
  const float f = __ldg(factor);
  const long n4 = n >> 2;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 v = x4[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    x4[i] = v;
  }
  for (long i = (n4 << 2) + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) x[i] *= f;
}
}  // namespace fvqa

extern "C" int fvqa_greedy_next(const float* logits, int ld, int V, const fvqa_h16* tok_emb, int dim, int32_t* ids, int S, const int32_t* pos,
                                int32_t* out_tokens, int out_ld, int step, float* x_next, float* margin, int rows, void* stream) {
  FVQA_REQUIRE(logits && tok_emb && ids && pos && out_tokens && x_next && dim % 8 == 0 && V > 0 && step >= 0 && step < out_ld,
               FVQA_ERR_INVALID_ARG, "greedy_next: bad arguments (dim %d V %d step %d of %d)", dim, V, step, out_ld);
  if (rows <= 0) return FVQA_OK;
  fvqa::greedy_next_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, V, reinterpret_cast<const fvqa::h16*>(tok_emb), dim,
                                                                                 ids, S, pos, out_tokens, out_ld, step, x_next, margin);
  return fvqa::check_launch("greedy_next");
}

extern "C" int fvqa_grad_scale_prepare(const float* gscale, float target, float* gs_out, float* inv_k, void* stream) {
  FVQA_REQUIRE(gscale && gs_out && inv_k, FVQA_ERR_INVALID_ARG, "grad_scale_prepare: null pointer");
  fvqa::grad_scale_prepare_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(gscale, target, gs_out, inv_k);
  return fvqa::check_launch("grad_scale_prepare");
}

extern "C" int fvqa_scale_f32(float* x, const float* factor_dev, int64_t n, void* stream) {
  FVQA_REQUIRE(x && factor_dev && n >= 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, FVQA_ERR_INVALID_ARG, "scale_f32: bad arguments");
  if (n == 0) return FVQA_OK;
  long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  fvqa::scale_f32_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, factor_dev, static_cast<long>(n));
  return fvqa::check_launch("scale_f32");
}

extern "C" int fvqa_scatter_rows(const float* row_val, const int32_t* dst_index, float* dst, int rows, void* stream) {
  if (rows <= 0) return FVQA_OK;
  scatter_rows_kernel<<<(rows + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(row_val, dst_index, dst, rows);
  return check_launch("scatter_rows");
}

extern "C" int fvqa_option_score(const float* token_loss, int32_t* prediction, float* mean_loss, int n_items, int n_opt, int len,
                                 void* stream) {
  if (n_items <= 0) return FVQA_OK;
  option_score_kernel<<<(n_items + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(token_loss, prediction, mean_loss, n_items,
                                                                                            n_opt, len);
  return check_launch("option_score");
}
