// rgqa_b200 — multi-head attention for the tiny LXMERT / UNITER sequences (Lq, Lk <= 64; d = 64).
//
//   out[b, i, h*64:(h+1)*64] = softmax_j( q_i·k_j / 8 + (mask[b,j] ? 0 : -10000) ) · v_j
//
// Restates BertAttention.forward (src/lxrt/modeling.py:320-347): additive -10000 key mask built in
// LXRTModel.forward (:857-865); no mask at all for visual keys (src/lxrt/entry.py:109).
// One CTA per (sample, head); K and V of that head live in shared memory as fp32, every thread owns
// one query row (q and the output accumulator in registers), scores go through a thread-private
// shared-memory column so the softmax is the exact two-pass max/sum form in fp32.
// Q/K/V are read straight out of the fused [tokens, 3*768] projection buffer (strided rows), so no
// transpose / permute / contiguous copies exist (the reference makes four per call, :315-318,:344).
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int HD = 64;       // head dim
constexpr int MAXL = 64;     // max sequence length handled
constexpr int kThreads = 64;

template <typename T>
__device__ __forceinline__ void load_row8(const T* p, float (&dst)[8]);
template <>
__device__ __forceinline__ void load_row8<float>(const float* p, float (&dst)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w;
  dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
}
template <>
__device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16* p, float (&dst)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  dst[0] = a.x; dst[1] = a.y; dst[2] = b.x; dst[3] = b.y;
  dst[4] = c.x; dst[5] = c.y; dst[6] = d.x; dst[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void store_row8(T* p, const float (&src)[8]);
template <>
__device__ __forceinline__ void store_row8<float>(float* p, const float (&src)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(src[0], src[1], src[2], src[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(src[4], src[5], src[6], src[7]);
}
template <>
__device__ __forceinline__ void store_row8<__nv_bfloat16>(__nv_bfloat16* p, const float (&src)[8]) {
  uint4 u;
  u.x = pack_bf16x2(src[0], src[1]);
  u.y = pack_bf16x2(src[2], src[3]);
  u.z = pack_bf16x2(src[4], src[5]);
  u.w = pack_bf16x2(src[6], src[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
attention_kernel(const T* __restrict__ q, long long ldq, const T* __restrict__ k, long long ldk,
                 const T* __restrict__ v, long long ldv, const int32_t* __restrict__ key_mask, T* __restrict__ out,
                 long long ldo, int Lq, int Lk, int heads) {
  extern __shared__ float sm[];
  float* Ks = sm;                   // [Lk][64]
  float* Vs = Ks + Lk * HD;         // [Lk][64]
  float* Ps = Vs + Lk * HD;         // [Lk][kThreads]  (column = thread)
  float* Ms = Ps + Lk * kThreads;   // [Lk] additive mask

  const int b = blockIdx.x / heads;
  const int h = blockIdx.x % heads;
  const int tid = threadIdx.x;

  // cooperative K/V load: Lk rows x 8 chunks of 8 elements
  for (int idx = tid; idx < Lk * 8; idx += kThreads) {
    const int j = idx >> 3, c = (idx & 7) * 8;
    float t[8];
    load_row8<T>(k + (static_cast<long long>(b) * Lk + j) * ldk + h * HD + c, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) Ks[j * HD + c + i] = t[i];
    load_row8<T>(v + (static_cast<long long>(b) * Lk + j) * ldv + h * HD + c, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) Vs[j * HD + c + i] = t[i];
  }
  for (int j = tid; j < Lk; j += kThreads)
    Ms[j] = (key_mask == nullptr || key_mask[b * Lk + j] != 0) ? 0.f : -10000.f;
  __syncthreads();

  if (tid < Lq) {
    float qr[HD];
    const T* qp = q + (static_cast<long long>(b) * Lq + tid) * ldq + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      float t[8];
      load_row8<T>(qp + c, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) qr[c + i] = t[i];
    }
    float mx = -INFINITY;
    for (int j = 0; j < Lk; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(Ks + j * HD);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 kk = kr[c];  // same address across the warp: broadcast
        s = fmaf(qr[4 * c + 0], kk.x, s);
        s = fmaf(qr[4 * c + 1], kk.y, s);
        s = fmaf(qr[4 * c + 2], kk.z, s);
        s = fmaf(qr[4 * c + 3], kk.w, s);
      }
      s = s * 0.125f + Ms[j];  // / sqrt(64), then the additive mask (modeling.py:333-336)
      Ps[j * kThreads + tid] = s;
      mx = fmaxf(mx, s);
    }
    float acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    float sum = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float d = Ps[j * kThreads + tid] - mx;
      const float pj = (sizeof(T) == 4) ? expf(d) : __expf(d);
      sum += pj;
      const float4* vr = reinterpret_cast<const float4*>(Vs + j * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        const float4 vv = vr[c];
        acc[4 * c + 0] = fmaf(pj, vv.x, acc[4 * c + 0]);
        acc[4 * c + 1] = fmaf(pj, vv.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(pj, vv.z, acc[4 * c + 2]);
        acc[4 * c + 3] = fmaf(pj, vv.w, acc[4 * c + 3]);
      }
    }
    const float inv = 1.0f / sum;
    T* op = out + (static_cast<long long>(b) * Lq + tid) * ldo + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = acc[c + i] * inv;
      store_row8<T>(op + c, t);
    }
  }
}

}  // namespace

int attention(cudaStream_t stream, int dtype, const void* q, long long ldq, const void* k, long long ldk,
              const void* v, long long ldv, const int32_t* key_mask, void* out, long long ldo, int B, int Lq,
              int Lk, int heads) {
  RGQA_CHECK(B > 0 && heads > 0, "attention: empty batch");
  RGQA_CHECK(Lq >= 1 && Lq <= MAXL && Lk >= 1 && Lk <= MAXL, "attention: Lq=%d Lk=%d out of range [1,%d]", Lq, Lk,
             MAXL);
  const int esz = dtype == DT_F32 ? 4 : 2;
  RGQA_CHECK((ldq * esz) % 16 == 0 && (ldk * esz) % 16 == 0 && (ldv * esz) % 16 == 0 && (ldo * esz) % 16 == 0,
             "attention: row strides must be multiples of 16 bytes");
  const size_t smem = static_cast<size_t>(Lk) * (2 * HD + kThreads + 1) * sizeof(float);
  const int grid = B * heads;
  void* tok = prof_begin(stream);
  if (dtype == DT_F32) {
    static bool cfg = false;
    if (!cfg) {
      RGQA_CUDA_OK(cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      cfg = true;
    }
    attention_kernel<float><<<grid, kThreads, smem, stream>>>(
        static_cast<const float*>(q), ldq, static_cast<const float*>(k), ldk, static_cast<const float*>(v), ldv,
        key_mask, static_cast<float*>(out), ldo, Lq, Lk, heads);
  } else {
    static bool cfg = false;
    if (!cfg) {
      RGQA_CUDA_OK(cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        64 * 1024));
      cfg = true;
    }
    attention_kernel<__nv_bfloat16><<<grid, kThreads, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k), ldk,
        static_cast<const __nv_bfloat16*>(v), ldv, key_mask, static_cast<__nv_bfloat16*>(out), ldo, Lq, Lk, heads);
  }
  prof_end(tok, stream, 4.0 * B * static_cast<double>(Lq) * Lk * heads * HD, PROF_ATTENTION);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
