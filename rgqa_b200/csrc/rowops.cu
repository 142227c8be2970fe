// rgqa_b200 — row-wise HBM-bound kernels: embedding gather + LayerNorm, LayerNorm, the visual
// feature/box combine, dtype casts.  One warp owns one row; a row of H = 128*NV floats sits in
// registers as NV float4 per lane (fully coalesced 512 B per warp per load), statistics are
// two-pass fp32 (mean, then variance of the centred values) with warp-shuffle reductions.
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int kWarpsPerBlock = 4;

template <int NV>
__device__ __forceinline__ void row_stats(const float4 (&x)[NV], int H, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
  mean = warp_sum(s) / static_cast<float>(H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) / static_cast<float>(H);
  rstd = 1.0f / sqrtf(var + eps);
}

template <int NV>
__device__ __forceinline__ void normalize(float4 (&x)[NV], float mean, float rstd, const float* gamma,
                                          const float* beta, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    x[i].x = (x[i].x - mean) * rstd * g.x + b.x;
    x[i].y = (x[i].y - mean) * rstd * g.y + b.y;
    x[i].z = (x[i].z - mean) * rstd * g.z + b.z;
    x[i].w = (x[i].w - mean) * rstd * g.w + b.w;
  }
}

// `lo_row` / `stats_row` (bf16 outputs only): the row additionally leaves as an identity deferred-LayerNorm state —
// lo = bf16(y - bf16(y)) and statistics (sum 0, sumsq H) so that mean = 0, rstd = 1 (see identity_state)
template <int NV, typename OutT>
__device__ __forceinline__ void store_row(const float4 (&x)[NV], OutT* out_row, int lane, float* copy_row = nullptr,
                                          __nv_bfloat16* lo_row = nullptr, float* stats_row = nullptr) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    // optional fp32 copy: the residual stream stays fp32 while the GEMM operand copy is bf16
    if (copy_row != nullptr) *reinterpret_cast<float4*>(copy_row + c) = x[i];
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_row) + c) = x[i];
    } else if constexpr (std::is_same<OutT, __half>::value) {  // the fp16 single-word deferred-LayerNorm state: no lo row
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out_row) + c) = make_uint2(pack_f16x2_sat(x[i].x, x[i].y), pack_f16x2_sat(x[i].z, x[i].w));
    } else {
      uint2 u;
      u.x = pack_bf16x2(x[i].x, x[i].y);
      u.y = pack_bf16x2(x[i].z, x[i].w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_row) + c) = u;
      if (lo_row != nullptr) {
        const float2 h0 = unpack_bf16x2(u.x), h1 = unpack_bf16x2(u.y);
        *reinterpret_cast<uint2*>(lo_row + c) = make_uint2(pack_bf16x2(x[i].x - h0.x, x[i].y - h0.y), pack_bf16x2(x[i].z - h1.x, x[i].w - h1.y));
      }
    }
  }
  if (stats_row != nullptr && lane < 3)
    reinterpret_cast<float4*>(stats_row)[lane] = make_float4(0.f, lane == 0 ? static_cast<float>(NV * 128) : 0.f, 0.f, 0.f);
}

// ---- embeddings + LayerNorm (src/lxrt/modeling.py:278-292) --------------------------------------
template <int NV, typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ seg, const float* __restrict__ word,
                const float* __restrict__ pos, const float* __restrict__ type, const float* __restrict__ gamma,
                const float* __restrict__ beta, OutT* __restrict__ out, float* __restrict__ out32, __nv_bfloat16* __restrict__ lo,
                float* __restrict__ stats, int rows, int L, int H, float eps, int vocab,
                int n_types) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // ids are validated by the host shim; the clamp only keeps a corrupt id from faulting the GPU
  const int id = min(max(ids[row], 0), vocab - 1);
  const int sg = seg != nullptr ? min(max(seg[row], 0), n_types - 1) : 0;
  const int ps = row % L;
  const float* w = word + static_cast<long long>(id) * H;
  const float* p = pos + static_cast<long long>(ps) * H;
  const float* t = type + static_cast<long long>(sg) * H;
  float4 x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + c));
    const float4 d = __ldg(reinterpret_cast<const float4*>(t + c));
    // reference order: (words + position) + token_type  (modeling.py:289)
    x[i] = make_float4((a.x + b.x) + d.x, (a.y + b.y) + d.y, (a.z + b.z) + d.z, (a.w + b.w) + d.w);
  }
  float mean, rstd;
  row_stats<NV>(x, H, eps, mean, rstd);
  normalize<NV>(x, mean, rstd, gamma, beta, lane);
  store_row<NV, OutT>(x, out + static_cast<long long>(row) * H, lane,
                      out32 != nullptr ? out32 + static_cast<long long>(row) * H : nullptr,
                      lo != nullptr ? lo + static_cast<long long>(row) * H : nullptr,
                      stats != nullptr ? stats + static_cast<long long>(row) * 12 : nullptr);
}

// ---- plain LayerNorm over fp32 rows ------------------------------------------------------------
template <int NV, typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const float* __restrict__ xin, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, OutT* __restrict__ out, float* __restrict__ out32, int rows, int H, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = xin + static_cast<long long>(row) * ldx;
  float4 x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) x[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  float mean, rstd;
  row_stats<NV>(x, H, eps, mean, rstd);
  normalize<NV>(x, mean, rstd, gamma, beta, lane);
  store_row<NV, OutT>(x, out + static_cast<long long>(row) * H, lane,
                      out32 != nullptr ? out32 + static_cast<long long>(row) * H : nullptr);
}

// ---- small-K projection of a per-row vector (boxes, K = 4 or 7) ---------------------------------
// The [H, K] weight is staged once per CTA into shared memory TRANSPOSED ([K][H] followed by the bias [H]) so a
// lane's four output columns are one conflict-free 16-byte read per k; CTAs then walk many rows.
__device__ __forceinline__ void stage_small_weight(float* wsm, const float* __restrict__ W, const float* __restrict__ bias,
                                                   int H, int K) {
  for (int idx = threadIdx.x; idx < H * K; idx += blockDim.x) {
    const int c = idx / K, k = idx - c * K;
    wsm[k * H + c] = W[idx];
  }
  for (int c = threadIdx.x; c < H; c += blockDim.x) wsm[K * H + c] = bias[c];
  __syncthreads();
}
template <int NV>
__device__ __forceinline__ void small_proj(float4 (&y)[NV], const float* vec, int K, const float* wsm, int H, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    // torch addmm order: sum_k x[k] * W[c,k], then + bias; K is tiny so the order is fixed here too
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(wsm + k * H + c);
      a.x = fmaf(vec[k], w.x, a.x);
      a.y = fmaf(vec[k], w.y, a.y);
      a.z = fmaf(vec[k], w.z, a.z);
      a.w = fmaf(vec[k], w.w, a.w);
    }
    const float4 b = *reinterpret_cast<const float4*>(wsm + K * H + c);
    y[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// ---- VisualFeatEncoder tail (src/lxrt/modeling.py:507-517) --------------------------------------
template <int NV, typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
visn_combine_kernel(const float* __restrict__ x_pre, const float* __restrict__ boxes, int box_dim,
                    const float* __restrict__ Wb, const float* __restrict__ bb, const float* __restrict__ g1,
                    const float* __restrict__ b1, const float* __restrict__ g2, const float* __restrict__ b2,
                    OutT* __restrict__ out, float* __restrict__ out32, __nv_bfloat16* __restrict__ lo, float* __restrict__ stats,
                    int rows, int H, float eps) {
  extern __shared__ __align__(16) float wsm[];
  pdl_launch_dependents();
  pdl_wait();
  stage_small_weight(wsm, Wb, bb, H, box_dim);
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += gridDim.x * kWarpsPerBlock) {
  float4 x[NV], y[NV];
  const float* xr = x_pre + static_cast<long long>(row) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) x[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  float bx[8];
  for (int k = 0; k < box_dim; ++k) bx[k] = boxes[static_cast<long long>(row) * box_dim + k];
  small_proj<NV>(y, bx, box_dim, wsm, H, lane);
  float mean, rstd;
  row_stats<NV>(x, H, eps, mean, rstd);
  normalize<NV>(x, mean, rstd, g1, b1, lane);
  row_stats<NV>(y, H, eps, mean, rstd);
  normalize<NV>(y, mean, rstd, g2, b2, lane);
#pragma unroll
  for (int i = 0; i < NV; ++i)
    x[i] = make_float4((x[i].x + y[i].x) / 2.f, (x[i].y + y[i].y) / 2.f, (x[i].z + y[i].z) / 2.f,
                       (x[i].w + y[i].w) / 2.f);
  store_row<NV, OutT>(x, out + static_cast<long long>(row) * H, lane,
                      out32 != nullptr ? out32 + static_cast<long long>(row) * H : nullptr,
                      lo != nullptr ? lo + static_cast<long long>(row) * H : nullptr,
                      stats != nullptr ? stats + static_cast<long long>(row) * 12 : nullptr);
  }
}

// ---- UniterImageEmbeddings tail (src/uniter/modeling.py:606-612) --------------------------------
template <int NV, typename OutT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
uniter_img_kernel(const float* __restrict__ x_pre, const float* __restrict__ posf, int pos_dim,
                  const float* __restrict__ Wp, const float* __restrict__ bp, const float* __restrict__ g1,
                  const float* __restrict__ b1, const float* __restrict__ g2, const float* __restrict__ b2,
                  const float* __restrict__ type_row, const float* __restrict__ g3, const float* __restrict__ b3,
                  OutT* __restrict__ out, float* __restrict__ out32, int rows, int H, float eps) {
  extern __shared__ __align__(16) float wsm[];
  pdl_launch_dependents();
  pdl_wait();
  stage_small_weight(wsm, Wp, bp, H, pos_dim);
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); row < rows; row += gridDim.x * kWarpsPerBlock) {
  float4 x[NV], y[NV];
  const float* xr = x_pre + static_cast<long long>(row) * H;
#pragma unroll
  for (int i = 0; i < NV; ++i) x[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
  float pv[8];
  for (int k = 0; k < pos_dim; ++k) pv[k] = posf[static_cast<long long>(row) * pos_dim + k];
  small_proj<NV>(y, pv, pos_dim, wsm, H, lane);
  float mean, rstd;
  row_stats<NV>(x, H, eps, mean, rstd);
  normalize<NV>(x, mean, rstd, g1, b1, lane);
  row_stats<NV>(y, H, eps, mean, rstd);
  normalize<NV>(y, mean, rstd, g2, b2, lane);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(type_row + i * 128 + lane * 4));
    // reference order: (transformed_im + transformed_pos) + type_embeddings  (uniter/modeling.py:610)
    x[i] = make_float4((x[i].x + y[i].x) + t.x, (x[i].y + y[i].y) + t.y, (x[i].z + y[i].z) + t.z,
                       (x[i].w + y[i].w) + t.w);
  }
  row_stats<NV>(x, H, eps, mean, rstd);
  normalize<NV>(x, mean, rstd, g3, b3, lane);
  store_row<NV, OutT>(x, out + static_cast<long long>(row) * H, lane,
                      out32 != nullptr ? out32 + static_cast<long long>(row) * H : nullptr);
  }
}

// ---- casts -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in,
                                                            __nv_bfloat16* __restrict__ out, long long n8) {
  // 8 elements per thread per step: 2 x 16 B loads, 1 x 16 B store; grid-stride
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y);
    u.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(out)[i] = u;
  }
}
__global__ void cast_f32_bf16_tail(const float* in, __nv_bfloat16* out, long long start, long long n) {
  const long long i = start + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, long long src_ld, T* __restrict__ dst,
                                   long long dst_ld, int rows, int H) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < static_cast<long long>(rows) * H) {
    const int r = static_cast<int>(i / H), c = static_cast<int>(i % H);
    dst[r * dst_ld + c] = src[r * src_ld + c];
  }
}

// LayerNorm folding into the consumer's weight (deferred LayerNorm): one warp per output row n.
// The consumer epilogue computes rstd * acc + b' and DROPS the exact form's - rstd * mean * c[n], c[n] = sum_k W'[n,k]: the
// folded row is centred in fp32, so c is only the sum of its 768 bf16 rounding errors (~2e-3 for |w| ~ 0.03) — harmless for
// rows with mean / std ~ 0.1, but a trained checkpoint's hidden rows can sit at mean / std ~ 5-10 (outlier channels), where
// the dropped term would reach the bf16 noise floor.  So the rounding is done with ERROR FEEDBACK along k (each element
// absorbs the rounding error accumulated so far, staying within one bf16 ulp of its target) and the element of SMALLEST
// magnitude goes last: it takes whatever the row sum still owes (<= half an ulp of the largest weight, ~1e-4) and leaves
// only ITS rounding error, half an ulp of ~1e-4: |c[n]| ~ 1e-7 |w|max instead of ~2e-3.
__global__ void __launch_bounds__(128) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      int N, int K, __nv_bfloat16* __restrict__ Wo,
                                                      float* __restrict__ c_out, float* __restrict__ b_out, int f16) {
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  // centre the folded row: (x - mean(x)) · w == x · (w - mean(w)), so the consumer never needs the row mean
  float ms = 0.f;
  for (int k = lane; k < K; k += 32) ms = fmaf(W[static_cast<long long>(n) * K + k], gamma[k], ms);
  const float wmean = warp_sum(ms) / static_cast<float>(K);
  float bs = 0.f;
  for (int k = lane; k < K; k += 32) bs = fmaf(W[static_cast<long long>(n) * K + k], beta[k], bs);
  bs = warp_sum(bs);
  if (lane == 0) {
    // sequential error-feedback rounding (load time only: ~2 x 768 steps per row, rows in parallel)
    int klast = 0;
    float amin = INFINITY;
    for (int k = 0; k < K; ++k) {
      const float a = fabsf(fmaf(W[static_cast<long long>(n) * K + k], gamma[k], -wmean));
      if (a < amin) {
        amin = a;
        klast = k;
      }
    }
    float err = 0.f;
    double cs = 0.0;
    for (int i = 0; i <= K; ++i) {
      if (i == klast) continue;          // the smallest element is rounded last (i == K)
      const int k = i < K ? i : klast;
      const float target = fmaf(W[static_cast<long long>(n) * K + k], gamma[k], -wmean);
      float got;
      if (f16) {  // fp16 copy (consumers of the single-word fp16 state): same error feedback, 11 significand bits
        const __half wh = __float2half_rn(target + err);
        reinterpret_cast<__half*>(Wo)[static_cast<long long>(n) * K + k] = wh;
        got = __half2float(wh);
      } else {
        const __nv_bfloat16 wb = __float2bfloat16_rn(target + err);
        Wo[static_cast<long long>(n) * K + k] = wb;
        got = __bfloat162float(wb);
      }
      err += target - got;   // what the row sum still owes
      cs += got;
    }
    c_out[n] = static_cast<float>(cs);
    b_out[n] = bias[n] + bs;
  }
}

// identity deferred-LayerNorm state from a normalised row pair (bf16 hi + fp32 copy)
__global__ void __launch_bounds__(256) identity_state_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, float* __restrict__ stats,
                                                             long long n4, long long nstat, float width) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    if (lo == nullptr) {  // fp16 single-word state: `hi` is REWRITTEN as fp16(y)
      reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_f16x2_sat(v.x, v.y), pack_f16x2_sat(v.z, v.w));
      continue;
    }
    const uint2 h = reinterpret_cast<const uint2*>(hi)[i];
    const float2 h0 = unpack_bf16x2(h.x), h1 = unpack_bf16x2(h.y);
    reinterpret_cast<uint2*>(lo)[i] = make_uint2(pack_bf16x2(v.x - h0.x, v.y - h0.y), pack_bf16x2(v.z - h1.x, v.w - h1.y));
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nstat; i += stride)
    stats[i] = (i % 12 == 1) ? width : 0.f;
}

// fp32 [M, K] (row stride ld) -> bf16 [M, 3K]: hi = bf16(x), lo = bf16(x - hi) laid out hi | hi | lo (activation side) or
// hi | lo | hi (weight side), so that one bf16 GEMM over K' = 3K computes hi·hi + hi·lo + lo·hi (fp32 accumulate)
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ in, long long ld, __nv_bfloat16* __restrict__ out,
                                                           long long M, int K, int weight_side) {
  const int k4 = K / 4;
  const long long n4 = M * k4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / k4;
    const int c = static_cast<int>(i - r * k4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + r * ld + c);
    const uint2 h = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    const float2 h0 = unpack_bf16x2(h.x), h1 = unpack_bf16x2(h.y);
    const uint2 l = make_uint2(pack_bf16x2(v.x - h0.x, v.y - h0.y), pack_bf16x2(v.z - h1.x, v.w - h1.y));
    __nv_bfloat16* o = out + r * 3 * K + c;
    *reinterpret_cast<uint2*>(o) = h;
    *reinterpret_cast<uint2*>(o + K) = weight_side ? l : h;
    *reinterpret_cast<uint2*>(o + 2 * K) = weight_side ? h : l;
  }
}

// fp32 [R, C] (row stride ld) -> bf16 [C, R]: the operand re-layout of a weight-gradient product (G = dy^T x runs over the
// ROWS of dy and x, the tensor-core GEMM wants that dimension contiguous) in one pass — 6 bytes per element instead of the
// 14 of a transposing copy followed by a cast.  64 x 32 tiles through padded shared memory: 128-byte reads along C, 128-byte
// bf16x2 writes along R.
__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ in, long long ld, __nv_bfloat16* __restrict__ out,
                                                             int R, int C, float* __restrict__ colsum) {
  __shared__ float tile[32][65];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[tx][ty + 8 * i] = (r < R && c < C) ? in[static_cast<long long>(r) * ld + c] : 0.f;
  }
  __syncthreads();
  // optional column sums of the fp32 input (the bias gradient of the same nn.Linear: its dy is read here anyway)
  if (colsum != nullptr && ty == 0 && c0 + tx < C) {
    float s = 0.f;
#pragma unroll 16
    for (int r = 0; r < 64; ++r) s += tile[tx][r];
    atomicAdd(colsum + c0 + tx, s);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + 2 * tx;
    if (c < C && r < R) {
      __nv_bfloat16* o = out + static_cast<long long>(c) * R + r;
      if (r + 1 < R && (R & 1) == 0)
        *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(tile[ty + 8 * i][2 * tx], tile[ty + 8 * i][2 * tx + 1]);
      else {
        o[0] = __float2bfloat16_rn(tile[ty + 8 * i][2 * tx]);
        if (r + 1 < R) o[1] = __float2bfloat16_rn(tile[ty + 8 * i][2 * tx + 1]);
      }
    }
  }
}

// fp32 [R, C] (row stride ld) -> bf16 [R, C] (row stride ldo), with the column sums of the fp32 input added to colsum (optional): the
// training step's ONE pass over a linear layer's dy — its bf16 copy feeds both the dgrad and the (transposed-operand) wgrad GEMM,
// the column sums are the bias gradient.  Block = 64 rows x 128 columns; a thread owns 4 columns of 8 rows.
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ in, long long ld, __nv_bfloat16* __restrict__ out,
                                                          long long ldo, int R, int C, float* __restrict__ colsum) {
  __shared__ float part[8][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 31) * 4, ry = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 64;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + ry + 8 * i;
      if (r < R) {
        const float4 v = *reinterpret_cast<const float4*>(in + static_cast<long long>(r) * ld + c);
        *reinterpret_cast<uint2*>(out + static_cast<long long>(r) * ldo + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
  }
  if (colsum == nullptr) return;
  *reinterpret_cast<float4*>(&part[ry][(threadIdx.x & 31) * 4]) = s;
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.x * 128 + threadIdx.x < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(colsum + blockIdx.x * 128 + threadIdx.x, t);
  }
}

inline int row_grid(int rows) { return (rows + kWarpsPerBlock - 1) / kWarpsPerBlock; }

}  // namespace

#define RGQA_DISPATCH_ROW_SMEM(KERNEL, OUT_T, SMEM, ...)                                                   \
  do {                                                                                                     \
    const int nv = H / 128;                                                                                \
    int grid_ = row_grid(M);                                                                               \
    if (grid_ > 148 * 8) grid_ = 148 * 8; /* a few resident waves; warps stride over the rows */           \
    if (H % 128 == 0 && nv == 6)                                                                           \
      RGQA_CUDA_OK(launch_pdl(KERNEL<6, OUT_T>, dim3(grid_), dim3(kWarpsPerBlock * 32), (SMEM), stream, __VA_ARGS__)); \
    else {                                                                                                 \
      set_error("row kernel: unsupported hidden size %d (need 768)", H);                                   \
      return 1;                                                                                            \
    }                                                                                                      \
  } while (0)

#define RGQA_DISPATCH_ROW(KERNEL, OUT_T, ...)                                                              \
  do {                                                                                                     \
    const int nv = H / 128;                                                                                \
    if (H % 128 == 0 && nv == 6)                                                                           \
      RGQA_CUDA_OK(launch_pdl(KERNEL<6, OUT_T>, dim3(row_grid(M)), dim3(kWarpsPerBlock * 32), 0, stream, __VA_ARGS__)); \
    else if (H % 128 == 0 && nv == 12)                                                                     \
      RGQA_CUDA_OK(launch_pdl(KERNEL<12, OUT_T>, dim3(row_grid(M)), dim3(kWarpsPerBlock * 32), 0, stream, __VA_ARGS__)); \
    else {                                                                                                 \
      set_error("row kernel: unsupported hidden size %d (need 768 or 1536)", H);                           \
      return 1;                                                                                            \
    }                                                                                                      \
  } while (0)

int embed_ln(cudaStream_t stream, int out_dtype, const int32_t* ids, const int32_t* seg, const float* word,
             const float* pos, const float* type, const float* gamma, const float* beta, void* out, float* out32, int B,
             int L, int H, int vocab, int max_pos, int n_types, float eps, void* lo_out, float* stats_out) {
  RGQA_CHECK(B > 0 && L > 0 && L <= max_pos, "embed_ln: bad B=%d L=%d (max_pos=%d)", B, L, max_pos);
  const int M = B * L;
  if (out_dtype == DT_F16) {  // fp16 identity state (single word, statistics (0, H)); lo_out is not used
    RGQA_CHECK(stats_out != nullptr, "embed_ln: the fp16 state needs its statistics buffer");
    RGQA_DISPATCH_ROW(embed_ln_kernel, __half, ids, seg, word, pos, type, gamma, beta, static_cast<__half*>(out), out32, static_cast<__nv_bfloat16*>(nullptr), stats_out, M, L, H, eps, vocab, n_types);
  } else if (out_dtype == DT_BF16) {
    RGQA_DISPATCH_ROW(embed_ln_kernel, __nv_bfloat16, ids, seg, word, pos, type, gamma, beta, static_cast<__nv_bfloat16*>(out), out32, static_cast<__nv_bfloat16*>(lo_out), stats_out, M, L, H, eps, vocab, n_types);
  } else {
    RGQA_DISPATCH_ROW(embed_ln_kernel, float, ids, seg, word, pos, type, gamma, beta, static_cast<float*>(out), out32, static_cast<__nv_bfloat16*>(nullptr), static_cast<float*>(nullptr), M, L, H, eps, vocab, n_types);
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm(cudaStream_t stream, int out_dtype, const float* x, long long ldx, const float* gamma,
              const float* beta, void* y, float* y32, int M, int H, float eps) {
  RGQA_CHECK(M > 0, "layernorm: empty input");
  RGQA_CHECK(ldx % 4 == 0, "layernorm: row stride must be a multiple of 4 floats");
  if (out_dtype == DT_BF16) {
    RGQA_DISPATCH_ROW(layernorm_kernel, __nv_bfloat16, x, ldx, gamma, beta, static_cast<__nv_bfloat16*>(y), y32, M, H, eps);
  } else {
    RGQA_DISPATCH_ROW(layernorm_kernel, float, x, ldx, gamma, beta, static_cast<float*>(y), y32, M, H, eps);
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int visn_combine(cudaStream_t stream, int out_dtype, const float* x_pre, const float* boxes, int box_dim,
                 const float* Wb, const float* bb, const float* g1, const float* b1, const float* g2,
                 const float* b2, void* out, float* out32, int M, int H, float eps, void* lo_out, float* stats_out) {
  RGQA_CHECK(M > 0 && box_dim >= 1 && box_dim <= 8, "visn_combine: bad M=%d box_dim=%d", M, box_dim);
  if (out_dtype == DT_F16) {
    RGQA_CHECK(stats_out != nullptr, "visn_combine: the fp16 state needs its statistics buffer");
    RGQA_DISPATCH_ROW_SMEM(visn_combine_kernel, __half, (box_dim + 1) * H * sizeof(float), x_pre, boxes, box_dim, Wb, bb, g1, b1, g2, b2, static_cast<__half*>(out), out32, static_cast<__nv_bfloat16*>(nullptr), stats_out, M, H, eps);
  } else if (out_dtype == DT_BF16) {
    RGQA_DISPATCH_ROW_SMEM(visn_combine_kernel, __nv_bfloat16, (box_dim + 1) * H * sizeof(float), x_pre, boxes, box_dim, Wb, bb, g1, b1, g2, b2, static_cast<__nv_bfloat16*>(out), out32, static_cast<__nv_bfloat16*>(lo_out), stats_out, M, H, eps);
  } else {
    RGQA_DISPATCH_ROW_SMEM(visn_combine_kernel, float, (box_dim + 1) * H * sizeof(float), x_pre, boxes, box_dim, Wb, bb, g1, b1, g2, b2, static_cast<float*>(out), out32, static_cast<__nv_bfloat16*>(nullptr), static_cast<float*>(nullptr), M, H, eps);
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int uniter_img_combine(cudaStream_t stream, int out_dtype, const float* x_pre, const float* pos, int pos_dim,
                       const float* Wp, const float* bp, const float* g1, const float* b1, const float* g2,
                       const float* b2, const float* type_row, const float* g3, const float* b3, void* out,
                       float* out32, int M, int H, float eps) {
  RGQA_CHECK(M > 0 && pos_dim >= 1 && pos_dim <= 8, "uniter_img_combine: bad M=%d pos_dim=%d", M, pos_dim);
  if (out_dtype == DT_BF16) {
    RGQA_DISPATCH_ROW_SMEM(uniter_img_kernel, __nv_bfloat16, (pos_dim + 1) * H * sizeof(float), x_pre, pos, pos_dim, Wp, bp, g1, b1, g2, b2, type_row, g3, b3, static_cast<__nv_bfloat16*>(out), out32, M, H, eps);
  } else {
    RGQA_DISPATCH_ROW_SMEM(uniter_img_kernel, float, (pos_dim + 1) * H * sizeof(float), x_pre, pos, pos_dim, Wp, bp, g1, b1, g2, b2, type_row, g3, b3, static_cast<float*>(out), out32, M, H, eps);
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int cast_f32_to_f16(cudaStream_t stream, const float* in, void* out, long long n) {
  RGQA_CHECK(n > 0 && in != nullptr && out != nullptr, "cast_f32_to_f16: bad arguments");
  long long grid = (n + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  cast_f32_f16_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(in, static_cast<__half*>(out), n);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int identity_state(cudaStream_t stream, const float* y_f32, void* hi, void* lo, float* stats, long long M, int H) {
  RGQA_CHECK(M > 0 && H % 4 == 0, "identity_state: bad shape");
  const long long n4 = M * H / 4;
  long long grid = (n4 + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  identity_state_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(y_f32, static_cast<__nv_bfloat16*>(hi),
                                                                       static_cast<__nv_bfloat16*>(lo), stats, n4, M * 12,
                                                                       static_cast<float>(H));
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int fold_ln_weight(cudaStream_t stream, const float* W, const float* bias, const float* gamma, const float* beta, int N,
                   int K, void* W_out_bf16, float* c_out, float* bias_out, int f16) {
  RGQA_CHECK(N > 0 && K > 0, "fold_ln_weight: empty weight");
  fold_ln_kernel<<<(N + 3) / 4, 128, 0, stream>>>(W, bias, gamma, beta, N, K, static_cast<__nv_bfloat16*>(W_out_bf16), c_out,
                                                  bias_out, f16);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int split_bf16x3(cudaStream_t stream, const float* in, long long ld, void* out, long long M, int K, int weight_side) {
  RGQA_CHECK(M > 0 && K > 0 && K % 4 == 0 && ld % 4 == 0, "split_bf16x3: bad shape M=%lld K=%d ld=%lld", M, K, ld);
  RGQA_CHECK(reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0, "split_bf16x3: unaligned operand");
  long long blocks = (M * (K / 4) + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  split_bf16x3_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, ld, static_cast<__nv_bfloat16*>(out), M, K, weight_side);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int transpose_cast_bf16(cudaStream_t stream, const float* in, long long ld, void* out, int R, int C, float* colsum) {
  RGQA_CHECK(R > 0 && C > 0 && ld >= C, "transpose_cast_bf16: bad shape R=%d C=%d ld=%lld", R, C, ld);
  RGQA_CHECK(reinterpret_cast<uintptr_t>(out) % 4 == 0, "transpose_cast_bf16: unaligned output");
  transpose_cast_kernel<<<dim3((C + 31) / 32, (R + 63) / 64), 256, 0, stream>>>(in, ld, static_cast<__nv_bfloat16*>(out), R, C, colsum);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int cast_bf16_colsum(cudaStream_t stream, const float* in, long long ld, void* out, long long ldo, int R, int C, float* colsum) {
  RGQA_CHECK(in && out && R > 0 && C > 0 && C % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && ldo >= C &&
                 reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0,
             "cast_bf16_colsum: needs C and the row strides multiples of 4 and a 16-byte aligned input");
  cast_colsum_kernel<<<dim3((C + 127) / 128, (R + 63) / 64), 256, 0, stream>>>(in, ld, static_cast<__nv_bfloat16*>(out), ldo, R, C, colsum);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int cast_f32_to_bf16(cudaStream_t stream, const float* in, void* out, long long n) {
  if (n <= 0) return 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const long long n8 = aligned ? n / 8 : 0;
  if (n8 > 0) {
    long long blocks = (n8 + 255) / 256;
    const long long cap = 148LL * 16;  // a few resident waves; grid-stride covers the rest
    if (blocks > cap) blocks = cap;
    cast_f32_bf16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), n8);
  }
  const long long start = n8 * 8;
  if (start < n) {
    const long long rem = n - start;
    cast_f32_bf16_tail<<<static_cast<int>((rem + 255) / 256), 256, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out),
                                                                              start, n);
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int cast_bf16_to_f32(cudaStream_t stream, const void* in, float* out, long long n) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  cast_bf16_f32_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, n);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int gather_rows(cudaStream_t stream, int dtype, const void* src, long long src_ld, void* dst, long long dst_ld,
                int rows, int H) {
  if (rows <= 0) return 0;
  const long long n = static_cast<long long>(rows) * H;
  const int blocks = static_cast<int>((n + 255) / 256);
  if (dtype == DT_F32)
    gather_rows_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(src), src_ld,
                                                         static_cast<float*>(dst), dst_ld, rows, H);
  else
    gather_rows_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), src_ld,
                                                                 static_cast<__nv_bfloat16*>(dst), dst_ld, rows, H);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
