// rgqa_b200 — input-gradient (dgrad) kernels of the fp32 path: what ODIN (src/tasks/gqa_odin.py:97-131) and the
// Mahalanobis input pre-processing (src/tasks/gqa_maha.py:238-268) need from `loss.backward()` is d loss / d (feats, boxes)
// only — no weight gradients.  The linear layers' dgrad is the forward fp32 GEMM against transposed weight copies
// (gemm_f32 with EPI_RESID folds the residual branch's gradient in); this file holds the rest:
//
//   layernorm_bwd   BertLayerNorm (modeling.py:258-270):  xh = (x-mu)*rstd, g = dy*gamma,
//                   dx = rstd * (g - mean(g) - xh * mean(g*xh))  [* scale, + add]
//   gelu_fwd/bwd    gelu (modeling.py:112-118): y = x Phi(x);  dx = dy (Phi(x) + x phi(x))
//   tanh_bwd        BertPooler (modeling.py:575-581): dx = dy (1 - y^2)
//   attention_bwd   BertAttention core (modeling.py:330-346), probabilities recomputed:
//                   dV = P^T dO;  dP = dO V^T;  dS = P (dP - rowsum(dP P));  dQ = dS K / 8;  dK = dS^T Q / 8
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int HD = 64;

__global__ void __launch_bounds__(128)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ add, float* __restrict__ dx, int M, int H, float eps, float scale) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<long long>(row) * H;
  const float* dr = dy + static_cast<long long>(row) * H;
  float s = 0.f;
  for (int c = lane; c < H; c += 32) s += xr[c];
  const float mu = warp_sum(s) / static_cast<float>(H);
  float q = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float d = xr[c] - mu;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(H) + eps);
  float m1 = 0.f, m2 = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float g = dr[c] * gamma[c] * scale;
    m1 += g;
    m2 = fmaf(g, (xr[c] - mu) * rstd, m2);
  }
  m1 = warp_sum(m1) / static_cast<float>(H);
  m2 = warp_sum(m2) / static_cast<float>(H);
  for (int c = lane; c < H; c += 32) {
    const float g = dr[c] * gamma[c] * scale;
    float v = rstd * (g - m1 - (xr[c] - mu) * rstd * m2);
    if (add != nullptr) v += add[static_cast<long long>(row) * H + c];
    dx[static_cast<long long>(row) * H + c] = v;
  }
}

// layernorm_bwd + ln_param_grad in ONE pass over (dy, x): the training step calls them back to back on the same rows (61 pairs per
// step).  A warp walks rows, keeps its per-column dgamma / dbeta partials in registers, the CTA merges them through shared-memory
// atomics and issues ONE global atomic per column (the stand-alone ln_param_grad issues one per warp).
template <int NV>  // NV = H / 32 columns per lane
__global__ void __launch_bounds__(256)
layernorm_bwd_params_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                            const float* __restrict__ add, float* __restrict__ dx, float* __restrict__ dgamma,
                            float* __restrict__ dbeta, int M, float eps, float scale) {
  constexpr int H = NV * 32;
  __shared__ float sdg[H], sdb[H];
  for (int c = threadIdx.x; c < H; c += 256) sdg[c] = sdb[c] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gm[NV], dg[NV], db[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    gm[j] = gamma[lane + 32 * j];
    dg[j] = db[j] = 0.f;
  }
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const float* xr = x + static_cast<long long>(row) * H;
    const float* dr = dy + static_cast<long long>(row) * H;
    float xv[NV], gv[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xv[j] = xr[lane + 32 * j];
      gv[j] = dr[lane + 32 * j] * scale;
      s += xv[j];
    }
    const float mu = warp_sum(s) / static_cast<float>(H);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xv[j] -= mu;
      q = fmaf(xv[j], xv[j], q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(H) + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xv[j] *= rstd;                        // xh
      dg[j] = fmaf(gv[j], xv[j], dg[j]);    // parameter gradients see scale * dy
      db[j] += gv[j];
      gv[j] *= gm[j];                       // g = scale * dy * gamma
      m1 += gv[j];
      m2 = fmaf(gv[j], xv[j], m2);
    }
    m1 = warp_sum(m1) / static_cast<float>(H);
    m2 = warp_sum(m2) / static_cast<float>(H);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float v = rstd * (gv[j] - m1 - xv[j] * m2);
      if (add != nullptr) v += add[static_cast<long long>(row) * H + lane + 32 * j];
      dx[static_cast<long long>(row) * H + lane + 32 * j] = v;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    atomicAdd(&sdg[lane + 32 * j], dg[j]);
    atomicAdd(&sdb[lane + 32 * j], db[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += 256) {
    atomicAdd(dgamma + c, sdg[c]);
    atomicAdd(dbeta + c, sdb[c]);
  }
}

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = gelu_erf(x[i]);
}
// the same activation with a bf16 result: in the mixed-precision step GELU(x) is only ever a GEMM operand (FFN2, and its weight gradient)
__global__ void __launch_bounds__(256) gelu_fwd_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n2) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const float2 v = reinterpret_cast<const float2*>(x)[i];
    reinterpret_cast<uint32_t*>(y)[i] = pack_bf16x2(gelu_erf(v.x), gelu_erf(v.y));
  }
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    dx[i] = dy[i] * fmaf(v, pdf, cdf);
  }
}
__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dx[i] = dy[i] * (1.0f - y[i] * y[i]);
}

// one CTA per (sample, head); everything of the head lives in shared memory (Lq, Lk <= 64).  Rows are staged at a stride of
// 68 floats (16-byte aligned, four banks of skew per row: eight consecutive rows tile the 32 banks for float4 reads) and every
// product is register-tiled 4 x 4 (scores / dP: 4 query rows x 4 key rows per thread, 32 FMAs per 16 LDS.128; dQ / dK / dV:
// 4 rows x 4 columns, 16 - 32 FMAs per 2 - 5 loads) — the first version read two shared-memory words per FMA and ran at an
// eighth of the FMA rate (119 us per launch at B = 128, L = 36).
constexpr int BLD = 68;
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void axpy4(float4& acc, float s, const float4 x) {
  acc.x = fmaf(s, x.x, acc.x);
  acc.y = fmaf(s, x.y, acc.y);
  acc.z = fmaf(s, x.z, acc.z);
  acc.w = fmaf(s, x.w, acc.w);
}
__global__ void __launch_bounds__(256)
attention_bwd_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                     const float* __restrict__ v, long long ldv, const int32_t* __restrict__ key_mask,
                     const float* __restrict__ dctx, long long ldo, float* __restrict__ dq, long long lddq,
                     float* __restrict__ dk, long long lddk, float* __restrict__ dv, long long lddv, int Lq, int Lk, int heads,
                     float drop_p, unsigned long long drop_seed, int vec) {
  extern __shared__ __align__(16) float sm_bwd[];
  // attention-probability dropout of the forward (same counter-based mask): O = (P o M) V with M = keep / (1 - p)
  const float keep_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  auto mask_at = [&](int i, int j) -> float {
    if (drop_p <= 0.f) return 1.0f;
    return dropout_keep(drop_seed, (static_cast<unsigned long long>(blockIdx.x) * Lq + i) * Lk + j, drop_p) ? keep_scale : 0.f;
  };
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  float* Qs = sm_bwd;               // [Lq][BLD]
  float* Ks = Qs + Lq * BLD;        // [Lk][BLD]
  float* Vs = Ks + Lk * BLD;        // [Lk][BLD]
  float* Os = Vs + Lk * BLD;        // dO [Lq][BLD]
  float* P = Os + Lq * BLD;         // [Lq][BLD] scores, then probabilities
  float* dS = P + Lq * BLD;         // [Lq][BLD] dP o M, then dS
  const int tid = threadIdx.x;
  const float* qg = q + static_cast<long long>(b) * Lq * ldq + h * HD;
  const float* og = dctx + static_cast<long long>(b) * Lq * ldo + h * HD;
  const float* kg = k + static_cast<long long>(b) * Lk * ldk + h * HD;
  const float* vg = v + static_cast<long long>(b) * Lk * ldv + h * HD;
  if (vec) {
    for (int e = tid; e < Lq * 16; e += 256) {
      const int r = e >> 4, c = (e & 15) * 4;
      *reinterpret_cast<float4*>(Qs + r * BLD + c) = *reinterpret_cast<const float4*>(qg + r * ldq + c);
      *reinterpret_cast<float4*>(Os + r * BLD + c) = *reinterpret_cast<const float4*>(og + r * ldo + c);
    }
    for (int e = tid; e < Lk * 16; e += 256) {
      const int r = e >> 4, c = (e & 15) * 4;
      *reinterpret_cast<float4*>(Ks + r * BLD + c) = *reinterpret_cast<const float4*>(kg + r * ldk + c);
      *reinterpret_cast<float4*>(Vs + r * BLD + c) = *reinterpret_cast<const float4*>(vg + r * ldv + c);
    }
  } else {
    for (int e = tid; e < Lq * HD; e += 256) {
      const int r = e >> 6, c = e & 63;
      Qs[r * BLD + c] = qg[r * ldq + c];
      Os[r * BLD + c] = og[r * ldo + c];
    }
    for (int e = tid; e < Lk * HD; e += 256) {
      const int r = e >> 6, c = e & 63;
      Ks[r * BLD + c] = kg[r * ldk + c];
      Vs[r * BLD + c] = vg[r * ldv + c];
    }
  }
  __syncthreads();
  const int nI = (Lq + 3) >> 2, nJ = (Lk + 3) >> 2;
  // scores + additive mask (modeling.py:333-336) and dP = dO V^T.  Thread tile: rows i = ti + nI * ii, keys j = tj + nJ * jj
  // (interleaved, so that the lanes of a warp read CONSECUTIVE K / V rows: conflict-free float4 reads)
  if (tid < nI * nJ) {
    const int ti = tid / nJ, tj = tid - ti * nJ;
    int ri[4], rj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      ri[a] = min(ti + nI * a, Lq - 1);   // rows past the end recompute the last one; nothing is stored for them
      rj[a] = min(tj + nJ * a, Lk - 1);
    }
    float s[4][4], dp[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[a][c] = dp[a][c] = 0.f;
#pragma unroll 2
    for (int c4 = 0; c4 < 16; ++c4) {
      float4 q4[4], o4[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        q4[a] = *reinterpret_cast<const float4*>(Qs + ri[a] * BLD + c4 * 4);
        o4[a] = *reinterpret_cast<const float4*>(Os + ri[a] * BLD + c4 * 4);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 k4 = *reinterpret_cast<const float4*>(Ks + rj[c] * BLD + c4 * 4);
        const float4 v4 = *reinterpret_cast<const float4*>(Vs + rj[c] * BLD + c4 * 4);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          s[a][c] = dot4(q4[a], k4, s[a][c]);
          dp[a][c] = dot4(o4[a], v4, dp[a][c]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ti + nI * a;
      if (i >= Lq) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = tj + nJ * c;
        if (j >= Lk) continue;
        float sc = s[a][c] * 0.125f;
        if (key_mask != nullptr && key_mask[b * Lk + j] == 0) sc += -10000.0f;
        P[i * BLD + j] = sc;
        dS[i * BLD + j] = dp[a][c] * mask_at(i, j);   // d loss / d P (the gradient reaches P through the dropout mask)
      }
    }
  }
  __syncthreads();
  // row softmax, then dS = P (dP - sum_j dP P): one warp per row
  for (int i = tid >> 5; i < Lq; i += 8) {
    const int lane = tid & 31;
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, P[i * BLD + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float p = expf(P[i * BLD + j] - mx);
      P[i * BLD + j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float dot = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float p = P[i * BLD + j] * inv;
      P[i * BLD + j] = p;
      dot = fmaf(dS[i * BLD + j], p, dot);
    }
    dot = warp_sum(dot);
    for (int j = lane; j < Lk; j += 32) {
      dS[i * BLD + j] = P[i * BLD + j] * (dS[i * BLD + j] - dot);
      if (drop_p > 0.f) P[i * BLD + j] *= mask_at(i, j);   // dV sees the dropped probabilities
    }
    // the float4 reads of the dK / dV pass run up to three keys past Lk: zero them (a stale NaN times zero would still be NaN)
    for (int j = Lk + lane; j < ((Lk + 3) & ~3); j += 32) P[i * BLD + j] = dS[i * BLD + j] = 0.f;
  }
  __syncthreads();
  // dQ = dS K / 8: 4 query rows x 4 columns per thread
  if (tid < nI * 16) {
    const int ti = tid >> 4, c0 = (tid & 15) * 4;
    int ri[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ri[a] = min(ti + nI * a, Lq - 1);
    float4 acc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < Lk; ++j) {
      const float4 k4 = *reinterpret_cast<const float4*>(Ks + j * BLD + c0);
#pragma unroll
      for (int a = 0; a < 4; ++a) axpy4(acc[a], dS[ri[a] * BLD + j], k4);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ti + nI * a;
      if (i >= Lq) continue;
      float* dst = dq + (static_cast<long long>(b) * Lq + i) * lddq + h * HD + c0;
      const float4 o = make_float4(acc[a].x * 0.125f, acc[a].y * 0.125f, acc[a].z * 0.125f, acc[a].w * 0.125f);
      if (vec) {
        *reinterpret_cast<float4*>(dst) = o;
      } else {
        dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w;
      }
    }
  }
  // dK = dS^T Q / 8, dV = (P o M)^T dO: 4 consecutive keys x 4 columns per thread
  if (tid < nJ * 16) {
    const int j0 = (tid >> 4) * 4, c0 = (tid & 15) * 4;
    float4 ak[4], av[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) ak[a] = av[a] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int i = 0; i < Lq; ++i) {
      const float4 ds4 = *reinterpret_cast<const float4*>(dS + i * BLD + j0);
      const float4 p4 = *reinterpret_cast<const float4*>(P + i * BLD + j0);
      const float4 q4 = *reinterpret_cast<const float4*>(Qs + i * BLD + c0);
      const float4 o4 = *reinterpret_cast<const float4*>(Os + i * BLD + c0);
      axpy4(ak[0], ds4.x, q4); axpy4(ak[1], ds4.y, q4); axpy4(ak[2], ds4.z, q4); axpy4(ak[3], ds4.w, q4);
      axpy4(av[0], p4.x, o4); axpy4(av[1], p4.y, o4); axpy4(av[2], p4.z, o4); axpy4(av[3], p4.w, o4);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = j0 + a;
      if (j >= Lk) continue;
      float* dkp = dk + (static_cast<long long>(b) * Lk + j) * lddk + h * HD + c0;
      float* dvp = dv + (static_cast<long long>(b) * Lk + j) * lddv + h * HD + c0;
      const float4 okk = make_float4(ak[a].x * 0.125f, ak[a].y * 0.125f, ak[a].z * 0.125f, ak[a].w * 0.125f);
      if (vec) {
        *reinterpret_cast<float4*>(dkp) = okk;
        *reinterpret_cast<float4*>(dvp) = av[a];
      } else {
        dkp[0] = okk.x; dkp[1] = okk.y; dkp[2] = okk.z; dkp[3] = okk.w;
        dvp[0] = av[a].x; dvp[1] = av[a].y; dvp[2] = av[a].z; dvp[3] = av[a].w;
      }
    }
  }
}

// y = dropout(x) (+ resid): BertEmbeddings / BertAttOutput / BertOutput / VisualFeatEncoder dropout sites (modeling.py:291, :359,
// :413, :516); the backward of the dropout itself is the same call on the gradient (same seed -> same mask)
__global__ void __launch_bounds__(256)
dropout_add_kernel(const float* __restrict__ x, const float* __restrict__ resid, float* __restrict__ y, long long n, float p,
                   unsigned long long seed) {
  const float scale = 1.0f / (1.0f - p);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = dropout_keep(seed, static_cast<unsigned long long>(i), p) ? x[i] * scale : 0.f;
    if (resid != nullptr) v += resid[i];
    y[i] = v;
  }
}

inline int ew_grid(long long n) {
  long long g = (n + 255) / 256;
  return static_cast<int>(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

}  // namespace

int layernorm_bwd(cudaStream_t stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx, int M,
                  int H, float eps, float scale) {
  RGQA_CHECK(dy && x && gamma && dx && M > 0 && H > 0, "layernorm_bwd: bad arguments");
  layernorm_bwd_kernel<<<(M + 3) / 4, 128, 0, stream>>>(dy, x, gamma, add, dx, M, H, eps, scale);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int layernorm_bwd_params(cudaStream_t stream, const float* dy, const float* x, const float* gamma, const float* add, float* dx,
                         float* dgamma, float* dbeta, int M, int H, float eps, float scale) {
  RGQA_CHECK(dy && x && gamma && dx && dgamma && dbeta && M > 0, "layernorm_bwd_params: bad arguments");
  int grid = (M + 7) / 8;
  if (grid > 148 * 2) grid = 148 * 2;
  if (H == 768)
    layernorm_bwd_params_kernel<24><<<grid, 256, 0, stream>>>(dy, x, gamma, add, dx, dgamma, dbeta, M, eps, scale);
  else if (H == 1536)
    layernorm_bwd_params_kernel<48><<<grid, 256, 0, stream>>>(dy, x, gamma, add, dx, dgamma, dbeta, M, eps, scale);
  else {
    set_error("layernorm_bwd_params: unsupported width %d (768 or 1536)", H);
    return 1;
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int gelu_fwd(cudaStream_t stream, const float* x, float* y, long long n) {
  RGQA_CHECK(x && y && n > 0, "gelu_fwd: bad arguments");
  gelu_fwd_kernel<<<ew_grid(n), 256, 0, stream>>>(x, y, n);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int gelu_fwd_bf16(cudaStream_t stream, const float* x, void* y, long long n) {
  RGQA_CHECK(x && y && n > 0 && n % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 8 == 0 && reinterpret_cast<uintptr_t>(y) % 4 == 0,
             "gelu_fwd_bf16: bad arguments (n must be even, pointers aligned)");
  gelu_fwd_bf16_kernel<<<ew_grid(n / 2), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y), n / 2);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int gelu_bwd(cudaStream_t stream, const float* dy, const float* x, float* dx, long long n) {
  RGQA_CHECK(dy && x && dx && n > 0, "gelu_bwd: bad arguments");
  gelu_bwd_kernel<<<ew_grid(n), 256, 0, stream>>>(dy, x, dx, n);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int tanh_bwd(cudaStream_t stream, const float* dy, const float* y, float* dx, long long n) {
  RGQA_CHECK(dy && y && dx && n > 0, "tanh_bwd: bad arguments");
  tanh_bwd_kernel<<<ew_grid(n), 256, 0, stream>>>(dy, y, dx, n);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int dropout_add(cudaStream_t stream, const float* x, const float* resid, float* y, long long n, float p, unsigned long long seed) {
  RGQA_CHECK(x && y && n > 0 && p >= 0.f && p < 1.f, "dropout_add: bad arguments");
  dropout_add_kernel<<<ew_grid(n), 256, 0, stream>>>(x, resid, y, n, p, seed);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int attention_bwd(cudaStream_t stream, const float* q, long long ldq, const float* k, long long ldk, const float* v,
                  long long ldv, const int32_t* key_mask, const float* dctx, long long ldo, float* dq, long long lddq, float* dk,
                  long long lddk, float* dv, long long lddv, int B, int Lq, int Lk, int heads, float drop_p,
                  unsigned long long drop_seed) {
  RGQA_CHECK(q && k && v && dctx && dq && dk && dv, "attention_bwd: null argument");
  RGQA_CHECK(drop_p >= 0.f && drop_p < 1.f, "attention_bwd: dropout probability out of range");
  RGQA_CHECK(B > 0 && heads > 0 && Lq >= 1 && Lq <= 64 && Lk >= 1 && Lk <= 64, "attention_bwd: Lq=%d Lk=%d out of range [1,64]", Lq, Lk);
  const size_t smem = static_cast<size_t>(4 * Lq + 2 * Lk) * BLD * sizeof(float);
  auto al16 = [](const void* p, long long ld) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % 4 == 0; };
  const int vec = (al16(q, ldq) && al16(k, ldk) && al16(v, ldv) && al16(dctx, ldo) && al16(dq, lddq) && al16(dk, lddk) && al16(dv, lddv)) ? 1 : 0;
  static bool cfg = false;
  if (!cfg) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    cfg = true;
  }
  attention_bwd_kernel<<<B * heads, 256, smem, stream>>>(q, ldq, k, ldk, v, ldv, key_mask, dctx, ldo, dq, lddq, dk, lddk, dv, lddv,
                                                         Lq, Lk, heads, drop_p, drop_seed, vec);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa

// ------------------------------------------------------------------------------------------------------------------
// Weight-gradient (wgrad) kernels of the fp32 path — the parameter half of `loss.backward()` in GQA.train
// (src/tasks/gqa_conf.py:196).  All of them ACCUMULATE into their outputs (shared modules — the cross-attention block
// serves both directions — and tied uses simply call twice); the caller zeroes the gradient buffers once per step.
//   gemm_f32_tn        dW[N, K] += dY[M, N]^T · X[M, K]                (nn.Linear weight)
//   colsum             db[N]    += sum_m dY[m, n]                      (nn.Linear bias, LayerNorm beta is in ln_param_grad)
//   ln_param_grad      dgamma[c] += sum_m dy[m,c]·xh[m,c],  dbeta[c] += sum_m dy[m,c]     (BertLayerNorm, xh recomputed)
//   embedding_scatter  dE[idx[r], :] += g[r, :] unless idx[r] == padding_idx  (nn.Embedding; the reference builds all three
//                      BertEmbeddings tables with padding_idx = 0, modeling.py:272-274, so word 0, POSITION 0 and token
//                      type 0 never receive a gradient)
// ------------------------------------------------------------------------------------------------------------------
namespace rgqa {
namespace {

__global__ void __launch_bounds__(256)
gemm_f32_tn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bm, long long ldb, float* __restrict__ out,
                   long long ldo, int M, int N, int K, int m_chunk) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int n0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;   // loader: 16 rows (m) x 64 columns
  const int m_lo = blockIdx.z * m_chunk, m_hi = min(M, m_lo + m_chunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int m0 = m_lo; m0 < m_hi; m0 += 16) {
    const int m = m0 + lr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lr][lc + i] = (m < m_hi && n0 + lc + i < N) ? A[static_cast<long long>(m) * lda + n0 + lc + i] : 0.f;
      Bs[lr][lc + i] = (m < m_hi && k0 + lc + i < K) ? Bm[static_cast<long long>(m) * ldb + k0 + lc + i] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const float4 av = *reinterpret_cast<const float4*>(&As[s][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[s][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty * 4 + i, k = k0 + tx * 4 + j;
      if (n < N && k < K) atomicAdd(out + static_cast<long long>(n) * ldo + k, acc[i][j]);
    }
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ out, int M, int N, int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += X[static_cast<long long>(r) * ldx + c];
  atomicAdd(out + c, s);
}

template <int NV>  // NV = H / 32 columns per lane
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int M, float eps, float scale) {
  constexpr int H = NV * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dg[NV], db[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) dg[j] = db[j] = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < M; row += gridDim.x * 8) {
    const float* xr = x + static_cast<long long>(row) * H;
    const float* dr = dy + static_cast<long long>(row) * H;
    float xv[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xv[j] = xr[lane + 32 * j];
      s += xv[j];
    }
    const float mu = warp_sum(s) / static_cast<float>(H);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float d = xv[j] - mu;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(H) + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float g = dr[lane + 32 * j] * scale;
      dg[j] = fmaf(g, (xv[j] - mu) * rstd, dg[j]);
      db[j] += g;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    atomicAdd(dgamma + lane + 32 * j, dg[j]);
    atomicAdd(dbeta + lane + 32 * j, db[j]);
  }
}

__global__ void __launch_bounds__(256)
embedding_scatter_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx, float* __restrict__ table, int rows, int H,
                         int table_rows, int padding_idx) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int t = idx[r];
  if (t < 0 || t >= table_rows || t == padding_idx) return;  // nn.Embedding(padding_idx=...) rows get no gradient
  for (int c = threadIdx.x; c < H; c += 256) atomicAdd(table + static_cast<long long>(t) * H + c, g[static_cast<long long>(r) * H + c]);
}

}  // namespace

int gemm_f32_tn(cudaStream_t stream, const float* A, long long lda, const float* B, long long ldb, float* out, long long ldo, int M,
                int N, int K) {
  RGQA_CHECK(A && B && out && M > 0 && N > 0 && K > 0, "gemm_f32_tn: bad arguments");
  // split the reduction over M so that small weight matrices still fill the GPU
  const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
  int splits = tiles >= 296 ? 1 : (296 + tiles - 1) / tiles;
  int m_chunk = ((M + splits - 1) / splits + 15) / 16 * 16;
  if (m_chunk < 16) m_chunk = 16;
  splits = (M + m_chunk - 1) / m_chunk;
  gemm_f32_tn_kernel<<<dim3((K + 63) / 64, (N + 63) / 64, splits), 256, 0, stream>>>(A, lda, B, ldb, out, ldo, M, N, K, m_chunk);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int colsum(cudaStream_t stream, const float* X, long long ldx, float* out, int M, int N) {
  RGQA_CHECK(X && out && M > 0 && N > 0, "colsum: bad arguments");
  const int rows_per_block = 64;
  colsum_kernel<<<dim3((N + 255) / 256, (M + rows_per_block - 1) / rows_per_block), 256, 0, stream>>>(X, ldx, out, M, N, rows_per_block);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int ln_param_grad(cudaStream_t stream, const float* dy, const float* x, float* dgamma, float* dbeta, int M, int H, float eps,
                  float scale) {
  RGQA_CHECK(dy && x && dgamma && dbeta && M > 0, "ln_param_grad: bad arguments");
  int grid = (M + 7) / 8;
  if (grid > 148 * 2) grid = 148 * 2;
  if (H == 768)
    ln_param_grad_kernel<24><<<grid, 256, 0, stream>>>(dy, x, dgamma, dbeta, M, eps, scale);
  else if (H == 1536)
    ln_param_grad_kernel<48><<<grid, 256, 0, stream>>>(dy, x, dgamma, dbeta, M, eps, scale);
  else {
    set_error("ln_param_grad: unsupported width %d (768 or 1536)", H);
    return 1;
  }
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}
int embedding_scatter(cudaStream_t stream, const float* g, const int32_t* idx, float* table, int rows, int H, int table_rows,
                      int padding_idx) {
  RGQA_CHECK(g && idx && table && rows > 0 && H > 0, "embedding_scatter: bad arguments");
  embedding_scatter_kernel<<<rows, 256, 0, stream>>>(g, idx, table, rows, H, table_rows, padding_idx);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

// x - noise * sign(grad), sign = +1 where grad >= 0 else -1: the input pre-processing step of ODIN / Mahalanobis with --noise
// (src/tasks/gqa_odin.py:118-120, gqa_maha.py:246-252), one pass over the [B, 36, 2048] features
namespace {
__global__ void __launch_bounds__(256) sign_step_kernel(const float* __restrict__ x, const float* __restrict__ grad, float* __restrict__ y,
                                                        long long n4, long long n, float noise) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(x)[i], g = reinterpret_cast<const float4*>(grad)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(g.x >= 0.f ? a.x - noise : a.x + noise, g.y >= 0.f ? a.y - noise : a.y + noise,
                                                  g.z >= 0.f ? a.z - noise : a.z + noise, g.w >= 0.f ? a.w - noise : a.w + noise);
  }
  for (long long i = 4 * n4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = grad[i] >= 0.f ? x[i] - noise : x[i] + noise;
}
}  // namespace
int sign_step(cudaStream_t stream, const float* x, const float* grad, float* y, long long n, float noise) {
  if (n <= 0) return 0;
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
  const long long n4 = al ? n / 4 : 0;
  long long blocks = ((n4 > 0 ? n4 : n) + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  sign_step_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(x, grad, y, n4, n, noise);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
