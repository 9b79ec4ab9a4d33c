/* flipped-vqa-b200 — C ABI of the LLaMA-VQA training-step kernels (sm_100a).
 *
 * The reference (inesriahi/Flipped-VQA) has no FFI layer: its boundary is the Python class API of
 * llama/model.py. This header is the boundary OUR Python host binds (flipped_vqa_b200/_lib.py,
 * ctypes): one entry point per fused region of the reference's hot path, forward and backward.
 * Each entry point cites the reference lines it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory
 *    (no hidden cudaMalloc, no host synchronisation inside any call);
 *  - `stream` is a cudaStream_t / CUstream passed as void*;
 *  - "h16" is the 16-bit tensor-core operand format this library was BUILT for: IEEE fp16 (default; the reference's own
 *    dtype, llama_vqa.py:63) or bf16 (libfvqa_bf16.so, -DFVQA_BF16); fvqa_operand_dtype() says which. tcgen05.mma
 *    kind::f16 rejects mixed fp16 x bf16 operands, so weights, activations and gradients all use it. `fvqa_h16` tensors
 *    are `uint16_t`-sized elements, row-major, leading dimension in
 *    ELEMENTS; "tokens" are the rows of all objective streams concatenated:
 *    sequence n = stream * B + b, token row = n * S + position;
 *  - return value: 0 = ok, <0 = FVQA_ERR_*; fvqa_last_error() returns a thread-local message.
 */
#ifndef FVQA_H_
#define FVQA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVQA_ABI_VERSION 2

#define FVQA_OK 0
#define FVQA_ERR_INVALID_ARG (-1)
#define FVQA_ERR_CUDA (-2)
#define FVQA_ERR_UNSUPPORTED (-3)

typedef uint16_t fvqa_h16;
#define FVQA_DTYPE_FP16 0
#define FVQA_DTYPE_BF16 1

int fvqa_abi_version(void);
/* FVQA_DTYPE_FP16 or FVQA_DTYPE_BF16: the format of every `fvqa_h16` tensor of this build. */
int fvqa_operand_dtype(void);
const char* fvqa_last_error(void);
/* One-time per-process setup (kernel attributes, driver entry points). Idempotent. */
int fvqa_init(void);

/* ---- RMSNorm (llama/model.py:31-42) on the fp32 residual stream. y = h16(x * rstd * w); rstd saved. -- */
int fvqa_rmsnorm_fwd(const float* x, const fvqa_h16* w, fvqa_h16* y, float* rstd,
                     int rows, int dim, float eps, void* stream);
/* dX only (weight is frozen). dx = (dres ? dres : 0) + d rmsnorm(x)/dx . dy   (fp32);
 * dx_h16 (optional) receives the same values rounded to h16 = A operand of the next GEMM. */
int fvqa_rmsnorm_bwd(const fvqa_h16* dy, const float* x, const fvqa_h16* w, const float* rstd,
                     const float* dres, float* dx, fvqa_h16* dx_h16, int rows, int dim, void* stream);
/* Final norm applied only to the gathered rows `idx[i]` (>=0) of x; rows with idx<0 give zeros.
 * (llama/model.py:347,352,358 restricted to the positions the losses read.) */
int fvqa_rmsnorm_gather_fwd(const float* x, const int32_t* idx, const fvqa_h16* w, fvqa_h16* y,
                            float* rstd, int rows_out, int dim, float eps, void* stream);
/* dx[idx[i]] = rmsnorm backward of row i (dx / dx_h16 zero-initialised by the caller; each source
 * row may appear at most once per call). */
int fvqa_rmsnorm_scatter_bwd(const fvqa_h16* dy, const float* x, const int32_t* idx,
                             const fvqa_h16* w, const float* rstd, float* dx, fvqa_h16* dx_h16,
                             int rows_out, int dim, void* stream);

/* ---- SwiGLU (llama/model.py:142). g = [rows, 2*hid] holding a=W1x | b=W3x; c = silu(a)*b. -------- */
int fvqa_swiglu_fwd(const fvqa_h16* g, fvqa_h16* c, int rows, int hid, void* stream);
int fvqa_swiglu_bwd(const fvqa_h16* dc, const fvqa_h16* g, fvqa_h16* dg, int rows, int hid, void* stream);

/* ---- h16 GEMM on tcgen05/TMEM fed by TMA (replaces every frozen nn.Linear: llama/model.py:89,
 *      99-100,128,142,348,354 and their dX-only backward). C[M,N] = A[M,K] * B[N,K]^T (+ R[M,N]).
 *      A, B h16 K-contiguous; fp32 accumulation in TMEM. out_fp32 != 0 -> C is float, else h16.
 *      R (optional residual, SAME dtype as C, leading dimension ldr) is added in fp32 before the store:
 *      with out_fp32 this is the fp32 residual stream  h = x + attn,  out = h + ffn (model.py:185-186).
 *      Requirements: K % 64 == 0, N % 8 == 0, lda/ldb/ldc/ldr % 8 == 0, 16-byte aligned pointers. */
int fvqa_gemm_nt(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, void* C, int ldc,
                      const void* R, int ldr, int M, int N, int K, int out_fp32, void* stream);
/* Same GEMM (h16 out) with RoPE folded into the epilogue for the fused Wq|Wk|Wv projection
 * (llama/model.py:89 + :61-67): columns [0, rope_cols) are q|k heads of width hd whose interleaved
 * pairs are rotated by the angle of position (row % S); rope_cos/rope_sin are [>=S, hd/2] fp32. */
int fvqa_gemm_nt_rope(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc,
                           int M, int N, int K, const float* rope_cos, const float* rope_sin,
                           int rope_cols, int hd, int S, void* stream);
/* Same, for ragged / compacted token layouts (shared-prefix option scoring, llama/model_my_original_mod.py:332-377
 * evaluated once per option-invariant prefix): row r is rotated by the angle of position pos_ids[r] (device int32 [M],
 * each in [0, rows of the rope tables)). */
int fvqa_gemm_nt_rope_pos(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc,
                               int M, int N, int K, const float* rope_cos, const float* rope_sin,
                               int rope_cols, int hd, const int32_t* pos_ids, void* stream);

/* Grouped skinny GEMM for the adapter-prompt projections of ALL layers in one launch (llama/model.py:99-100 `wk/wv(adapter)` and
 * its backward d adapter = dK_a Wk + dV_a Wv): C_g[M,N] = A_g[M,K] * B_g[N,K]^T, g < groups, M <= 16, K % 256 == 0.
 * A_g = A + g*strideA, C_g = C + g*strideC (strides in elements of the respective type); B_g = B_ptrs_dev[g], a DEVICE array of
 * `groups` device pointers (the per-layer weight blocks are separate allocations), each with leading dimension ldb. HBM-bound:
 * 2*N*K bytes per group. */
int fvqa_gemm_skinny_grouped(const fvqa_h16* A, int64_t strideA, int lda, const void* const* B_ptrs_dev, int ldb, void* C,
                             int64_t strideC, int ldc, int M, int N, int K, int groups, int out_fp32, void* stream);

/* SwiGLU fused into the GEMM epilogues (llama/model.py:142 `w2(silu(w1 x) * w3 x)` and its backward):
 *  fwd: G[M, 2*hid] = X[M,K] * W13[2*hid, K]^T (h16, saved for backward; W13 = [W1; W3]) and
 *       C[M, hid] = silu(G[:, :hid]) * G[:, hid:], bit-identical to fvqa_gemm_nt + fvqa_swiglu_fwd. hid % 128 == 0.
 *  bwd: dG[M, 2*hid] = swiglu'(G) applied to dc = dY[M,K] * W2t[hid, K]^T; dc never reaches HBM;
 *       bit-identical to fvqa_gemm_nt + fvqa_swiglu_bwd. hid % 32 == 0. */
int fvqa_gemm_swiglu_fwd(const fvqa_h16* X, int ldx, const fvqa_h16* W13, int ldw, fvqa_h16* G, int ldg,
                         fvqa_h16* C, int ldc, int M, int hid, int K, void* stream);
int fvqa_gemm_swiglu_bwd(const fvqa_h16* dY, int ldy, const fvqa_h16* W2t, int ldw, const fvqa_h16* G, int ldg,
                         fvqa_h16* dG, int lddg, int M, int hid, int K, void* stream);


/* dX-only backward WITHOUT transposed weight copies: C[M,N] = A[M,K] . B[K,N] with B the forward's row-major [out = K, in = N] weight
 * (leading dimension ldb >= N), read as an MN-major tcgen05 operand; h16 in / out, fp32 accumulation, same tiles and k order as
 * fvqa_gemm_nt on the transposed copy (bit-identical results). K % 64 == 0, N % 8 == 0. fvqa_gemm_swiglu_bwd_nn is
 * fvqa_gemm_swiglu_bwd with W2 [K = d, hid] instead of W2t [hid, K]. */
int fvqa_gemm_nn(const fvqa_h16* A, int lda, const fvqa_h16* B, int ldb, fvqa_h16* C, int ldc, int M, int N, int K,
                 void* stream);
int fvqa_gemm_swiglu_bwd_nn(const fvqa_h16* dY, int ldy, const fvqa_h16* W2, int ldw, const fvqa_h16* G, int ldg,
                            fvqa_h16* dG, int lddg, int M, int hid, int K, void* stream);

/* ---- fused attention (llama/model.py:61-67 RoPE, :87-126 attention incl. adapter branch). --------
 * qkv  [n_seq*S, 3*H*hd] h16 (q | k | v) with q,k ALREADY rotated (fvqa_gemm_nt_rope).
 * akv  [>=A rows, ld akv_ld] h16: adapter keys (cols 0..H*hd) | adapter values (cols H*hd..2*H*hd),
 *      no RoPE (model.py:99-100), shared by every sequence.
 * rope [S, hd/2] fp32 cos and sin tables (model.py:45-50); used by backward for the inverse rotation.
 * gate1/gate2 [H] fp32 (model.py:84-85). vstart[n_seq] int32: video_start of the sequence, or -1
 * for "no bias" sequences (QAV, model.py:121-122). max_feats = F.
 * out  [n_seq*S, H*hd] h16 = tanh(gate1)*softmax(q ka^T/sqrt(hd)) va + softmax(q k^T/sqrt(hd)+causal+bias) v
 * lse  [n_seq, H, S] fp32 log-sum-exp of the text softmax (saved for backward).
 * hd in {64, 128}; A <= 16. */
int fvqa_attn_fwd(const fvqa_h16* qkv, const fvqa_h16* akv, int akv_ld, const float* rope_cos,
                  const float* rope_sin, const float* gate1, const float* gate2, const int32_t* vstart,
                  fvqa_h16* out, float* lse, int n_seq, int S, int H, int hd, int A, int max_feats,
                  void* stream);
/* Backward. dqkv [n_seq*S, 3*H*hd] h16 receives dq|dk|dv with the inverse RoPE already applied.
 * Per-CTA partials of the shared-parameter gradients go to `ws` (size from fvqa_attn_bwd_ws_bytes)
 * and are reduced in a fixed order into dakv [A, 2*H*hd] fp32 (dK_a | dV_a), dgate1[H], dgate2[H]
 * (fp32, overwritten). */
/* 1 if this shape takes the tcgen05 attention path (backward = 2 kernel launches instead of 3). */
int fvqa_attn_uses_tc(int S, int hd, int A);
int64_t fvqa_attn_bwd_ws_bytes(int n_seq, int S, int H, int hd, int A);
int fvqa_attn_bwd(const fvqa_h16* qkv, const fvqa_h16* akv, int akv_ld, const float* rope_cos,
                  const float* rope_sin, const float* gate1, const float* gate2, const int32_t* vstart,
                  const fvqa_h16* out, const float* lse, const fvqa_h16* dout, fvqa_h16* dqkv,
                  float* dakv, float* dgate1, float* dgate2, void* ws, int n_seq, int S, int H, int hd,
                  int A, int max_feats, void* stream);

/* ---- input embedding + video injection (llama/model.py:286-336). ---------------------------------
 * vproj: vf32[B*F, d] = video[B*F, vdim] * Wv[d, vdim]^T in fp32 (model.py:322). */
int fvqa_visual_proj_fwd(const float* video, const float* wv, float* vf32, int rows, int dim, int vdim, void* stream);
/* Generic fp32 Linear y[rows, dim] = x[rows, in_dim] * w[dim, in_dim]^T (+ bias[dim]) (+ add[rows, dim]); bias / add may be
 * NULL. Used for the input-fusion variants of llama/model.py:306-322: the concatenated [video | audio] projection (:310-311),
 * the frozen audio projection (:307, :314 with add = nothing / sum of both projections) and the q / k / v Linears (with
 * bias) of CrossAttentionModule (:148-163). */
int fvqa_linear_f32(const float* x, const float* w, const float* bias, const float* add, float* y, int rows, int dim, int in_dim,
                    void* stream);
/* CrossAttentionModule.forward (llama/model.py:153-169) after its three Linears: out[b, f, :] =
 * softmax_j(<q[b, f], k[b, j]> / sqrt(dim)) . v[b, j, :] over the `tokens` audio tokens of sample b (tokens <= 64).
 * q, out [n_samples*frames, dim]; k, v [n_samples*tokens, dim]; fp32. Forward only: everything upstream of visual_proj is frozen. */
int fvqa_cross_attn_fwd(const float* q, const float* k, const float* v, float* out, int n_samples, int frames, int tokens,
                        int dim, void* stream);
/* dWv[d, vdim] = dvf[rows, d]^T * video[rows, vdim] (fp32, overwritten). */
int fvqa_visual_proj_bwd(const float* dvf, const float* video, float* dwv, int rows, int dim, int vdim, void* stream);
/* h0[n*S+p, :] for every sequence n:
 *   mode vstart[n] >= 0 (VQA/VAQ): tok_emb[ids] except positions [vs, vs+F) <- h16(vf32[vid[n],f] + temporal[f])
 *   mode vstart[n] <  0 (QAV):     tok_emb[ids] * (labels[n,p] < 0), then += h16(vf32+temporal) at qav_index[vid[n], f]
 * ids/labels [n_seq, S] int32; seq_video[n] = video sample of sequence n; qav_index [B, F] int32.
 * h0 is the fp32 residual stream (values are h16-representable: embeddings / h16(vf+temporal)). */
int fvqa_build_h0_fwd(const fvqa_h16* tok_emb, const int32_t* ids, const int32_t* labels,
                      const int32_t* vstart, const int32_t* seq_video, const int32_t* qav_index,
                      const float* vf32, const float* temporal, float* h0,
                      int n_seq, int S, int dim, int max_feats, void* stream);
/* dvf[B, F, d] (fp32) = sum over sequences of dh0 at the video slots. Overwrites dvf. */
int fvqa_build_h0_bwd(const float* dh0, const int32_t* vstart, const int32_t* seq_video,
                      const int32_t* qav_index, float* dvf, int n_seq, int n_video, int S, int dim,
                      int max_feats, void* stream);
/* dtemporal[F, d] = sum_b dvf[b] (overwritten); then dvf[b] += dvf_qav[b] in place (dvf_qav may be NULL). */
int fvqa_video_grad_finish(float* dvf, const float* dvf_qav, float* dtemporal, int n_video, int dim,
                           int max_feats, void* stream);

/* The two calls above in one full-grid launch: dvf[b, f] = (sum over the sequences of sample b of dh0 at frame f's slot) + dvf_qav[b, f]
 * (dvf_qav may be NULL), dtemporal[f] = sum_b of the first term. Both outputs overwritten. */
int fvqa_video_grad(const float* dh0, const int32_t* vstart, const int32_t* seq_video, const int32_t* qav_index,
                    const float* dvf_qav, float* dvf, float* dtemporal, int n_seq, int n_video, int S, int dim,
                    int max_feats, void* stream);

/* ---- vocabulary cross-entropy over labelled rows (llama/model.py:348-356, ignore_index=0). -------
 * logits [rows, V] fp32 (row stride ld), target[rows] int32 (<0 = padding row).
 * row_loss[rows] = lse - logit[target] (0 for padding rows); row_lse saved for backward. */
int fvqa_ce_fwd(const float* logits, int ld, const int32_t* target, float* row_loss, float* row_lse,
                int rows, int V, void* stream);
/* dlogits[rows, V] h16 = (softmax - onehot) * (*gscale_dev) * inv_count; padding rows -> 0. */
int fvqa_ce_bwd(const float* logits, int ld, const int32_t* target, const float* row_lse,
                const float* gscale_dev, float inv_count, fvqa_h16* dlogits, int ldd, int rows, int V,
                void* stream);
/* out[0] = sum(row_loss[0..rows)) * scale  (deterministic single-block reduction). */
int fvqa_sum_scale(const float* v, int rows, float scale, float* out, void* stream);

/* ---- QAV video-feature reconstruction loss (llama/model.py:358-361, ignore_index=-1). -------------
 * hn [rows, d] h16 = final-normed hidden rows (gathered), row_video[rows] = video sample (<0 pad),
 * target[rows] in [0,F). logits[j] = <hn[i], vf32[row_video[i], j]> / tau. prob [rows, F] saved. */
int fvqa_qav_loss_fwd(const fvqa_h16* hn, const float* vf32, const int32_t* row_video,
                      const int32_t* target, float tau, float* row_loss, float* prob, int rows, int dim,
                      int max_feats, void* stream);
/* dhn [rows, d] h16 and dvf_qav [n_video, F, d] fp32 (overwritten; deterministic per-sample order). */
int fvqa_qav_loss_bwd(const fvqa_h16* hn, const float* vf32, const int32_t* row_video,
                      const int32_t* target, const float* prob, const float* gscale_dev, float inv_count,
                      float tau, fvqa_h16* dhn, float* dvf_qav, int rows, int n_video, int dim,
                      int max_feats, void* stream);

/* ---- multiple-choice option scoring (llama/model_my_original_mod.py:375-377 + engine.py:88-93). ---
 * token_loss [n_items, n_opt, S-1] fp32 -> prediction[i] = argmin_o( sum / count(loss != 0) ). */
int fvqa_scatter_rows(const float* row_val, const int32_t* dst_index, float* dst, int rows, void* stream);
int fvqa_option_score(const float* token_loss, int32_t* prediction, float* mean_loss, int n_items,
                      int n_opt, int len, void* stream);

/* ---- greedy decoding step of the generation evaluator (llama/model.py:429-467: `pred = output[:, start_idx].max(1)[1]`, token
 *      written to position start_idx + 1, its embedding fed back). Row b: tok = argmax_v logits[b, v] (lowest index on ties);
 *      ids[b, pos[b] + 1] = tok (if in range); out_tokens[b, step] = tok; x_next[b, :] = tok_emb[tok, :] as fp32;
 *      margin[b, step] (optional) = best - second-best logit. ids [rows, S] int32, pos [rows] int32, out_tokens / margin [rows, out_ld]. */
int fvqa_greedy_next(const float* logits, int ld, int V, const fvqa_h16* tok_emb, int dim, int32_t* ids, int S,
                     const int32_t* pos, int32_t* out_tokens, int out_ld, int step, float* x_next, float* margin, int rows,
                     void* stream);

/* ---- live-row pruning of the LAST layer: only the rows the losses read (labelled positions) go through
 *      its wo / FFN GEMMs, like the vocabulary projection (llama/model.py:347-356 never needs the other rows' outputs).
 *      dst[i, :] = src[idx[i], :]   and   dst[idx[i], :] = src[i, :]   for rows of row_bytes (multiple of 16); idx < 0 skipped. */
int fvqa_gather_rows(const void* src, const int32_t* idx, void* dst, int rows, int row_bytes, void* stream);
int fvqa_scatter_row_vectors(const void* src, const int32_t* idx, void* dst, int rows, int row_bytes, void* stream);
/* ---- padding-free row set: the row-wise ops (norms, frozen GEMMs, SwiGLU, residuals) of llama/model.py:172-187 only run on
 *      the rows [0, last loss-relevant position] of each sequence (rows after it cannot influence any loss under the causal
 *      mask of model.py:298-299); attention still sees the full [n_seq, S] layout. dst[r, :] = idx[r] >= 0 ? src[idx[r], :] : 0
 *      rebuilds that layout (zero rows where nothing was computed) from the compact rows. */
int fvqa_expand_rows(const void* src, const int32_t* idx, void* dst, int rows, int row_bytes, void* stream);

/* ---- gradient range management of the fp16 build (the role torch's GradScaler plays for the reference's fp16 autograd,
 *      util/misc.py:253-273, done inside the step): every loss-gradient kernel takes an upstream scale `gscale_dev`; before
 *      backward the three upstream gradients are multiplied by ONE power of two k = 2^round(log2(target / max|gscale|)) so
 *      that the 16-bit gradient operands sit mid-range whatever loss scale / accum_iter the caller uses, and the trainable
 *      gradients are multiplied by 1/k (exact) when they are final. gs_out[3] = gscale[3] * k, inv_k[0] = 1/k (k = 1 when
 *      target <= 0 or every gscale is 0). */
int fvqa_grad_scale_prepare(const float* gscale, float target, float* gs_out, float* inv_k, void* stream);
/* x[i] *= *factor_dev for i < n (fp32, in place). */
int fvqa_scale_f32(float* x, const float* factor_dev, int64_t n, void* stream);

/* ---- small utilities --------------------------------------------------------------------------- */
int fvqa_f32_to_h16(const float* src, fvqa_h16* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVQA_H_ */
