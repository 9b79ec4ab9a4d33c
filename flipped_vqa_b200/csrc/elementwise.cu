// HBM-bound row kernels: fused RMSNorm fwd/bwd (+gather/scatter variants), fused SwiGLU fwd/bwd.
// Reference semantics: llama/model.py:31-42 (RMSNorm), :142 (SwiGLU). The residual stream (and its
// gradient) is fp32, GEMM operands are h16; 16-byte vector accesses; math in fp32. One CTA per row for the norms (row cached in registers), flat
// grid-stride for SwiGLU.
#include "common.cuh"

namespace fvqa {

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAXV = 4;  // vectors (8 elements) per thread kept in registers -> dim <= 8192 (kernels are instantiated for 2, 3, 4)

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8f(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// ---------------------------------------------------------------------------------------------
// RMSNorm forward on the fp32 residual stream:  y = h16(x * rstd * w)
// idx == nullptr: row r reads x[r]; otherwise row r reads x[idx[r]] (idx<0 -> zero row).
// ---------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_fwd_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ idx, const h16* __restrict__ w,
    h16* __restrict__ y, float* __restrict__ rstd_out, int dim, float eps) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int nvec = dim >> 3;
  long src = row;
  if (idx != nullptr) src = idx[row];
  uint4* yrow = reinterpret_cast<uint4*>(y + static_cast<long>(row) * dim);
  if (src < 0) {  // padding row
    for (int v = threadIdx.x; v < nvec; v += NORM_THREADS) yrow[v] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0 && rstd_out) rstd_out[row] = 0.f;
    return;
  }
  const float* xrow = x + src * dim;
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  float xr[MAXV][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nvec) {
      load8f(xrow + v * 8, xr[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += xr[i][j] * xr[i][j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / static_cast<float>(dim) + eps);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nvec) {
      float g[8], o[8];
      unpack8(__ldg(wv + v), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = xr[i][j] * rstd * g[j];
      yrow[v] = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm backward (dX only). With n = x*rstd, dn = dy*w:
//   dx = rstd * (dn - n * mean(dn * n)) (+ dres)      fp32 out (+ optional h16 copy: next GEMM operand)
// scatter variant: output row = idx[r] (rows with idx<0 skipped), no residual.
// ---------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_bwd_kernel(
    const h16* __restrict__ dy, const float* __restrict__ x, const int32_t* __restrict__ idx,
    const h16* __restrict__ w, const float* __restrict__ rstd_in, const float* __restrict__ dres,
    float* __restrict__ dx, h16* __restrict__ dx_h16, int dim) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int nvec = dim >> 3;
  long src = row;
  if (idx != nullptr) {
    src = idx[row];
    if (src < 0) return;
  }
  const uint4* dyrow = reinterpret_cast<const uint4*>(dy + static_cast<long>(row) * dim);
  const float* xrow = x + src * dim;
  const uint4* wv = reinterpret_cast<const uint4*>(w);
  const float rstd = rstd_in[row];
  float xr[MAXV][8], dn[MAXV][8], fr[MAXV][8];
  float dot = 0.f;
  const float* rrow = dres ? dres + src * dim : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nvec) {
      load8f(xrow + v * 8, xr[i]);
      // the residual gradient is only needed after the row reduction: loading it here keeps ONE round trip to memory per row
      if (rrow) load8f(rrow + v * 8, fr[i]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) fr[i][j] = 0.f;
      }
      float fd[8], fw[8];
      unpack8(__ldg(dyrow + v), fd);
      unpack8(__ldg(wv + v), fw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dn[i][j] = fd[j] * fw[j];
        dot += dn[i][j] * xr[i][j] * rstd;
      }
    }
  }
  dot = block_sum(dot, red) / static_cast<float>(dim);
  float* dxrow = dx + src * dim;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * NORM_THREADS;
    if (v < nvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fr[i][j] + rstd * (dn[i][j] - xr[i][j] * rstd * dot);
      store8f(dxrow + v * 8, o);
      if (dx_h16) reinterpret_cast<uint4*>(dx_h16 + src * dim)[v] = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SwiGLU. g = [rows, 2*hid] (a | b). fwd: c = h16( silu(a) * b ).
// bwd: da = dc * b * s * (1 + a * (1 - s)), db = dc * silu(a), s = sigmoid(a).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const h16* __restrict__ g, h16* __restrict__ c,
                                                          long rows, int hid) {
  const int hv = hid >> 3;
  const long total = rows * hv;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / hv;
    const int v = static_cast<int>(i - r * hv);
    const uint4* grow = reinterpret_cast<const uint4*>(g + r * 2 * hid);
    float a[8], b[8], o[8];
    unpack8(__ldg(grow + v), a);
    unpack8(__ldg(grow + hv + v), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = a[j] / (1.f + __expf(-a[j]));
      o[j] = s * b[j];
    }
    reinterpret_cast<uint4*>(c + r * hid)[v] = pack8(o);
  }
}

__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const h16* __restrict__ dc, const h16* __restrict__ g,
                                                          h16* __restrict__ dg, long rows, int hid) {
  const int hv = hid >> 3;
  const long total = rows * hv;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / hv;
    const int v = static_cast<int>(i - r * hv);
    const uint4* grow = reinterpret_cast<const uint4*>(g + r * 2 * hid);
    float a[8], b[8], d[8], da[8], db[8];
    unpack8(__ldg(grow + v), a);
    unpack8(__ldg(grow + hv + v), b);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dc + r * hid) + v), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = 1.f / (1.f + __expf(-a[j]));
      da[j] = d[j] * b[j] * s * (1.f + a[j] * (1.f - s));
      db[j] = d[j] * a[j] * s;
    }
    uint4* orow = reinterpret_cast<uint4*>(dg + r * 2 * hid);
    orow[v] = pack8(da);
    orow[hv + v] = pack8(db);
  }
}

__global__ void f32_to_h16_kernel(const float* __restrict__ src, h16* __restrict__ dst, long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x)
    dst[i] = f2h(src[i]);
}

// dst[i, :] = src[idx[i], :] (GATHER) or dst[idx[i], :] = src[i, :] (!GATHER); rows of `vecs` 16-byte vectors
template <bool GATHER>
__global__ void __launch_bounds__(256) move_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx, uint4* __restrict__ dst,
                                                        long rows, int vecs) {
  const long total = rows * vecs;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / vecs;
    const int v = static_cast<int>(i - r * vecs);
    const long o = idx[r];
    if (o < 0) continue;
    if (GATHER) dst[r * vecs + v] = __ldg(src + o * vecs + v);
    else dst[o * vecs + v] = __ldg(src + r * vecs + v);
  }
}

// dst[r] = idx[r] >= 0 ? src[idx[r]] : 0 — expands a compacted row set back to the full [n_seq * S] token layout
__global__ void __launch_bounds__(256) expand_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx, uint4* __restrict__ dst,
                                                          long rows, int vecs) {
  const long total = rows * vecs;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / vecs;
    const int v = static_cast<int>(i - r * vecs);
    const long o = idx[r];
    dst[i] = o >= 0 ? __ldg(src + o * vecs + v) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// registers follow the row length: 2 vectors per thread up to dim 4096 (7B), 3 up to 6144 (13B), else 4
#define FVQA_NORM_PICK(kernel, dim) ((dim) <= 8 * NORM_THREADS * 2 ? kernel<2> : (dim) <= 8 * NORM_THREADS * 3 ? kernel<3> : kernel<4>)

static int elementwise_grid(long work_items, int threads) {
  long blocks = (work_items + threads - 1) / threads;
  const long cap = 148L * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace fvqa

using namespace fvqa;

extern "C" int fvqa_rmsnorm_fwd(const float* x, const fvqa_h16* w, fvqa_h16* y, float* rstd, int rows,
                                int dim, float eps, void* stream) {
  FVQA_REQUIRE(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAXV, FVQA_ERR_UNSUPPORTED,
               "rmsnorm: dim %d must be a multiple of 8 and <= %d", dim, 8 * NORM_THREADS * NORM_MAXV);
  if (rows <= 0) return FVQA_OK;
  auto kfn = FVQA_NORM_PICK(rmsnorm_fwd_kernel, dim);
  launch_k(kfn, dim3(rows), dim3(NORM_THREADS), 0, static_cast<cudaStream_t>(stream), x, static_cast<const int32_t*>(nullptr),
           reinterpret_cast<const h16*>(w), reinterpret_cast<h16*>(y), rstd, dim, eps);
  return check_launch("rmsnorm_fwd");
}

extern "C" int fvqa_rmsnorm_gather_fwd(const float* x, const int32_t* idx, const fvqa_h16* w, fvqa_h16* y,
                                       float* rstd, int rows_out, int dim, float eps, void* stream) {
  FVQA_REQUIRE(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAXV, FVQA_ERR_UNSUPPORTED, "rmsnorm_gather: bad dim %d", dim);
  FVQA_REQUIRE(idx != nullptr, FVQA_ERR_INVALID_ARG, "rmsnorm_gather: idx is null");
  if (rows_out <= 0) return FVQA_OK;
  auto kfn = FVQA_NORM_PICK(rmsnorm_fwd_kernel, dim);
  launch_k(kfn, dim3(rows_out), dim3(NORM_THREADS), 0, static_cast<cudaStream_t>(stream), x, idx, reinterpret_cast<const h16*>(w),
           reinterpret_cast<h16*>(y), rstd, dim, eps);
  return check_launch("rmsnorm_gather_fwd");
}

extern "C" int fvqa_rmsnorm_bwd(const fvqa_h16* dy, const float* x, const fvqa_h16* w, const float* rstd,
                                const float* dres, float* dx, fvqa_h16* dx_h16, int rows, int dim, void* stream) {
  FVQA_REQUIRE(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAXV, FVQA_ERR_UNSUPPORTED, "rmsnorm_bwd: bad dim %d", dim);
  if (rows <= 0) return FVQA_OK;
  auto kfn = FVQA_NORM_PICK(rmsnorm_bwd_kernel, dim);
  launch_k(kfn, dim3(rows), dim3(NORM_THREADS), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const h16*>(dy), x,
           static_cast<const int32_t*>(nullptr), reinterpret_cast<const h16*>(w), rstd, dres, dx, reinterpret_cast<h16*>(dx_h16), dim);
  return check_launch("rmsnorm_bwd");
}

extern "C" int fvqa_rmsnorm_scatter_bwd(const fvqa_h16* dy, const float* x, const int32_t* idx, const fvqa_h16* w,
                                        const float* rstd, float* dx, fvqa_h16* dx_h16, int rows_out, int dim, void* stream) {
  FVQA_REQUIRE(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAXV, FVQA_ERR_UNSUPPORTED, "rmsnorm_scatter_bwd: bad dim %d", dim);
  FVQA_REQUIRE(idx != nullptr, FVQA_ERR_INVALID_ARG, "rmsnorm_scatter_bwd: idx is null");
  if (rows_out <= 0) return FVQA_OK;
  auto kfn = FVQA_NORM_PICK(rmsnorm_bwd_kernel, dim);
  launch_k(kfn, dim3(rows_out), dim3(NORM_THREADS), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const h16*>(dy), x, idx,
           reinterpret_cast<const h16*>(w), rstd, static_cast<const float*>(nullptr), dx, reinterpret_cast<h16*>(dx_h16), dim);
  return check_launch("rmsnorm_scatter_bwd");
}

extern "C" int fvqa_swiglu_fwd(const fvqa_h16* g, fvqa_h16* c, int rows, int hid, void* stream) {
  FVQA_REQUIRE(hid % 8 == 0, FVQA_ERR_UNSUPPORTED, "swiglu: hid %d must be a multiple of 8", hid);
  if (rows <= 0) return FVQA_OK;
  const long items = static_cast<long>(rows) * (hid >> 3);
  swiglu_fwd_kernel<<<elementwise_grid(items, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const h16*>(g), reinterpret_cast<h16*>(c), rows, hid);
  return check_launch("swiglu_fwd");
}

extern "C" int fvqa_swiglu_bwd(const fvqa_h16* dc, const fvqa_h16* g, fvqa_h16* dg, int rows, int hid, void* stream) {
  FVQA_REQUIRE(hid % 8 == 0, FVQA_ERR_UNSUPPORTED, "swiglu: hid %d must be a multiple of 8", hid);
  if (rows <= 0) return FVQA_OK;
  const long items = static_cast<long>(rows) * (hid >> 3);
  swiglu_bwd_kernel<<<elementwise_grid(items, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const h16*>(dc), reinterpret_cast<const h16*>(g), reinterpret_cast<h16*>(dg), rows, hid);
  return check_launch("swiglu_bwd");
}

extern "C" int fvqa_f32_to_h16(const float* src, fvqa_h16* dst, int64_t n, void* stream) {
  if (n <= 0) return FVQA_OK;
  launch_k(f32_to_h16_kernel, dim3(elementwise_grid(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, reinterpret_cast<h16*>(dst),
           static_cast<long>(n));
  return check_launch("f32_to_h16");
}

extern "C" int fvqa_gather_rows(const void* src, const int32_t* idx, void* dst, int rows, int row_bytes, void* stream) {
  FVQA_REQUIRE(row_bytes % 16 == 0, FVQA_ERR_UNSUPPORTED, "gather_rows: row_bytes %d must be a multiple of 16", row_bytes);
  if (rows <= 0) return FVQA_OK;
  const int vecs = row_bytes / 16;
  move_rows_kernel<true><<<elementwise_grid(static_cast<long>(rows) * vecs, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src), idx, reinterpret_cast<uint4*>(dst), rows, vecs);
  return check_launch("gather_rows");
}

extern "C" int fvqa_expand_rows(const void* src, const int32_t* idx, void* dst, int rows, int row_bytes, void* stream) {
  FVQA_REQUIRE(row_bytes % 16 == 0, FVQA_ERR_UNSUPPORTED, "expand_rows: row_bytes %d must be a multiple of 16", row_bytes);
  if (rows <= 0) return FVQA_OK;
  const int vecs = row_bytes / 16;
  expand_rows_kernel<<<elementwise_grid(static_cast<long>(rows) * vecs, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src), idx, reinterpret_cast<uint4*>(dst), rows, vecs);
  return check_launch("expand_rows");
}

extern "C" int fvqa_scatter_row_vectors(const void* src, const int32_t* idx, void* dst, int rows, int row_bytes, void* stream) {
  FVQA_REQUIRE(row_bytes % 16 == 0, FVQA_ERR_UNSUPPORTED, "scatter_row_vectors: row_bytes %d must be a multiple of 16", row_bytes);
  if (rows <= 0) return FVQA_OK;
  const int vecs = row_bytes / 16;
  move_rows_kernel<false><<<elementwise_grid(static_cast<long>(rows) * vecs, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(src), idx, reinterpret_cast<uint4*>(dst), rows, vecs);
  return check_launch("scatter_row_vectors");
}
