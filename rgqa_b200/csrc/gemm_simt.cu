// rgqa_b200 — fp32 GEMM on the CUDA cores (FFMA), used by the precise path (fp32 weights and
// activations end to end, tolerance 1e-4 on logits) and by the Mahalanobis scorer, whose quadratic
// forms need full fp32 operands.  Same contract as gemm_bf16:
//   out[M,N] = epilogue(A[M,K] · W[N,K]ᵀ + bias[N] (+ resid[M,N]))
// 64x64 tile per CTA, 16-deep K slices through shared memory, 4x4 micro-tile per thread.
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <int EPI>
__device__ __forceinline__ float act_f32(float v) {
  if constexpr (EPI == EPI_GELU) return gelu_erf(v);
  if constexpr (EPI == EPI_TANH) return tanhf(v);
  return v;
}

template <int EPI>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                const float* __restrict__ bias, const float* __restrict__ resid, long long ldr,
                float* __restrict__ out, long long ldo, int M, int N, int K) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  const int tx = tid & 15;  // 16 column groups of 4
  const int ty = tid >> 4;  // 16 row groups of 4
  // loader mapping: 64 rows x 16 k = 256 float4 along K
  const int lr = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  const bool k_vec = ((lda & 3) == 0) && ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(W) & 15) == 0);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, w[4] = {0.f, 0.f, 0.f, 0.f};
    const int ar = m0 + lr, wr = n0 + lr;
    if (ar < M) {
      const float* src = A + static_cast<long long>(ar) * lda + k0 + lk;
      if (k_vec && k0 + lk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + lk + i < K) a[i] = src[i];
      }
    }
    if (wr < N) {
      const float* src = W + static_cast<long long>(wr) * ldw + k0 + lk;
      if (k_vec && k0 + lk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + lk + i < K) w[i] = src[i];
      }
    }
    __syncthreads();  // previous slice fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = a[i];
      Ws[lk + i][lr] = w[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float ar4[4] = {av.x, av.y, av.z, av.w};
      const float wr4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar4[i], wr4[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[c];
      if constexpr (EPI == EPI_RESID) v += resid[static_cast<long long>(r) * ldr + c];
      out[static_cast<long long>(r) * ldo + c] = act_f32<EPI>(v);
    }
  }
}

// Large problems: 128 x 128 tile, 8 x 8 outputs per thread (two 4-wide halves per dimension so that the shared-memory reads stay
// 16-byte vectors without bank conflicts), K slices of 16 double-buffered in shared memory (one __syncthreads per slice).
// 64 FMAs per four LDS.128 instead of 16 per two: ~2x the 64 x 64 kernel on the fp32 path's encoder / backward GEMMs.
template <int EPI>
__global__ void __launch_bounds__(256)
gemm_f32_big_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                    const float* __restrict__ bias, const float* __restrict__ resid, long long ldr,
                    float* __restrict__ out, long long ldo, int M, int N, int K) {
  constexpr int BT = 128, BKK = 16;
  __shared__ float As[2][BKK][BT + 4];
  __shared__ float Ws[2][BKK][BT + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BT, n0 = blockIdx.x * BT;
  const int tx = tid & 15, ty = tid >> 4;
  // loader: 128 rows x 16 k = 512 float4 per operand; thread -> rows lr and lr + 64, k offset lk
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const bool k_vec = ((lda & 3) == 0) && ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto fetch = [&](const float* __restrict__ base, long long ld, int row, int rows, int k0, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (row < rows) {
      const float* src = base + static_cast<long long>(row) * ld + k0 + lk;
      if (k_vec && k0 + lk + 3 < K) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + lk + i < K) v[i] = src[i];
      }
    }
  };
  float a0[4], a1[4], w0[4], w1[4];
  fetch(A, lda, m0 + lr, M, 0, a0);
  fetch(A, lda, m0 + lr + 64, M, 0, a1);
  fetch(W, ldw, n0 + lr, N, 0, w0);
  fetch(W, ldw, n0 + lr + 64, N, 0, w1);
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BKK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][lk + i][lr] = a0[i];
      As[buf][lk + i][lr + 64] = a1[i];
      Ws[buf][lk + i][lr] = w0[i];
      Ws[buf][lk + i][lr + 64] = w1[i];
    }
    __syncthreads();  // slice k0 visible; the other buffer was fully consumed before the previous barrier
    if (k0 + BKK < K) {  // next slice's global loads in flight during this slice's math
      fetch(A, lda, m0 + lr, M, k0 + BKK, a0);
      fetch(A, lda, m0 + lr + 64, M, k0 + BKK, a1);
      fetch(W, ldw, n0 + lr, N, k0 + BKK, w0);
      fetch(W, ldw, n0 + lr + 64, N, k0 + BKK, w1);
    }
#pragma unroll
    for (int k = 0; k < BKK; ++k) {
      const float4 av0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 av1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 wv0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 wv1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float ar[8] = {av0.x, av0.y, av0.z, av0.w, av1.x, av1.y, av1.z, av1.w};
      const float wr[8] = {wv0.x, wv0.y, wv0.z, wv0.w, wv1.x, wv1.y, wv1.z, wv1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (c >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr) v += bias[c];
      if constexpr (EPI == EPI_RESID) v += resid[static_cast<long long>(r) * ldr + c];
      out[static_cast<long long>(r) * ldo + c] = act_f32<EPI>(v);
    }
  }
}

}  // namespace

int gemm_f32(cudaStream_t stream, const float* A, long long lda, const float* W, long long ldw, const float* bias,
             const float* resid, long long ldr, float* out, long long ldo, int M, int N, int K, int epi) {
  RGQA_CHECK(M > 0 && N > 0 && K > 0, "gemm_f32: empty problem M=%d N=%d K=%d", M, N, K);
  RGQA_CHECK(epi != EPI_RESID || resid != nullptr, "gemm_f32: EPI_RESID needs a residual pointer");
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  void* tok = prof_begin(stream);
  // enough 128 x 128 tiles to cover the GPU: the big-tile kernel; otherwise 64 x 64 tiles (more CTAs)
  const long long big_tiles = static_cast<long long>((N + 127) / 128) * ((M + 127) / 128);
  if (big_tiles >= 96) {
    dim3 g2((N + 127) / 128, (M + 127) / 128);
    switch (epi) {
      case EPI_BIAS: gemm_f32_big_kernel<EPI_BIAS><<<g2, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K); break;
      case EPI_GELU: gemm_f32_big_kernel<EPI_GELU><<<g2, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K); break;
      case EPI_TANH: gemm_f32_big_kernel<EPI_TANH><<<g2, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K); break;
      case EPI_RESID: gemm_f32_big_kernel<EPI_RESID><<<g2, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K); break;
      default: set_error("gemm_f32: unknown epilogue %d", epi); return 1;
    }
    prof_end(tok, stream, 2.0 * M * static_cast<double>(N) * K, PROF_GEMM_F32);
    note_launch();
    RGQA_CUDA_OK(cudaGetLastError());
    return 0;
  }
  switch (epi) {
    case EPI_BIAS:
      gemm_f32_kernel<EPI_BIAS><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K);
      break;
    case EPI_GELU:
      gemm_f32_kernel<EPI_GELU><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K);
      break;
    case EPI_TANH:
      gemm_f32_kernel<EPI_TANH><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K);
      break;
    case EPI_RESID:
      gemm_f32_kernel<EPI_RESID><<<grid, 256, 0, stream>>>(A, lda, W, ldw, bias, resid, ldr, out, ldo, M, N, K);
      break;
    default:
      set_error("gemm_f32: unknown epilogue %d", epi);
      return 1;
  }
  prof_end(tok, stream, 2.0 * M * static_cast<double>(N) * K, PROF_GEMM_F32);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
