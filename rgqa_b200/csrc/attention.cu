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
#include <cuda.h>

#include <cstdlib>

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
                 long long ldo, int Lq, int Lk, int heads, float drop_p, unsigned long long drop_seed) {
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
      float pj = (sizeof(T) == 4) ? expf(d) : __expf(d);
      sum += pj;
      // attention-probability dropout (training, modeling.py:341): applied to the normalised row, i.e. after the sum
      if (drop_p > 0.f)
        pj = dropout_keep(drop_seed, (static_cast<unsigned long long>(blockIdx.x) * Lq + tid) * Lk + j, drop_p) ? pj / (1.0f - drop_p) : 0.f;
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

// ------------------------------------------------------------------------------------------------
// bf16 path: the same attention on the warp-level tensor-core path (mma.sync m16n8k16, bf16 -> fp32).
// The tiles are far too small for tcgen05 (its minimum M is 64 and one instruction is a whole-CTA
// operation); one WARP owns one (sample, head): S = QK^T as 16-row query tiles against the K rows in
// shared memory, the softmax on the accumulator fragments (quad shuffles), P re-used in registers as
// the A operand of P·V, V^T fragments fetched with ldmatrix.trans.  Keys are padded to 16*KS with
// zero rows (score -inf, probability 0); masked keys get the reference's additive -10000.
// ------------------------------------------------------------------------------------------------
// 2^x on the MUFU alone (exp2f adds a range split for denormal results: 3 more instructions per probability, and a softmax
// term below 2^-126 contributes nothing); qkv_attn_tcgen05.cu uses the same expression so the two paths stay bit-identical
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.44269504088896340736f;
constexpr float kScaleLog2e = 0.125f * kLog2e;
constexpr int kLds = 72;  // smem row stride in bf16: 144 B -> the 8 rows of a fragment load hit 32 distinct banks

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

// 4 CTAs (16 warps) per SM: the kernel is bound by loads in flight (46 % of its stall samples are long-scoreboard
// waits at 12 warps / SM), so the register budget is capped at 128 and the shared-memory carve-out is maximised.
template <int KS>
__global__ void __launch_bounds__(128, 4)
attention_mma_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                     long long ldk, const __nv_bfloat16* __restrict__ v, long long ldv,
                     const int32_t* __restrict__ key_mask, __nv_bfloat16* __restrict__ out, long long ldo, int Lq,
                     int Lk, int heads, int items) {
  constexpr int NT = 2 * KS;     // 8-key score tiles
  constexpr int ROWS = 16 * KS;  // padded key count
  extern __shared__ __align__(16) unsigned char smem_att[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int item = blockIdx.x * 4 + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (item >= items) return;  // warp-uniform; no block-wide barrier below
  const int b = item / heads, h = item % heads;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_att) + static_cast<size_t>(warp) * (2 * ROWS * kLds);
  __nv_bfloat16* Vs = Ks + ROWS * kLds;

  // stage this head's K and V rows with cp.async (16 B per lane per copy, no register staging, all 8*KS copies
  // in flight at once; rows >= Lk are zero-filled through src-size 0)
#pragma unroll
  for (int it = 0; it < ROWS * 8 / 32; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx >> 3, c = (idx & 7) * 8;
    const bool ok = r < Lk;
    const long long grow = static_cast<long long>(b) * Lk + (ok ? r : 0);
    cp_async_16(Ks + r * kLds + c, k + grow * ldk + h * HD + c, ok ? 16u : 0u);
    cp_async_16(Vs + r * kLds + c, v + grow * ldv + h * HD + c, ok ? 16u : 0u);
  }
  cp_async_commit();

  float madd[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = 8 * j + 2 * t + e;
      madd[j][e] = key >= Lk ? -INFINITY : ((key_mask != nullptr && key_mask[b * Lk + key] == 0) ? -10000.f * kLog2e : 0.f);
    }

  // Q fragments come straight from global memory (4 B per lane); the next tile's are requested before the
  // current tile is computed, and the first tile's before the K/V copies are waited for
  auto load_q = [&](int m0, uint32_t (&qf)[4][4]) {
    const int r0 = m0 + g, r1 = r0 + 8;
    const __nv_bfloat16* q0 = q + (static_cast<long long>(b) * Lq + r0) * ldq + h * HD;
    const __nv_bfloat16* q1 = q + (static_cast<long long>(b) * Lq + r1) * ldq + h * HD;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + 2 * t;
      qf[ks][0] = r0 < Lq ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c)) : 0u;
      qf[ks][1] = r1 < Lq ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c)) : 0u;
      qf[ks][2] = r0 < Lq ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c + 8)) : 0u;
      qf[ks][3] = r1 < Lq ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c + 8)) : 0u;
    }
  };
  uint32_t qa[4][4], qn[4][4];
  load_q(0, qa);
  cp_async_wait_all();
  __syncwarp();

  for (int m0 = 0; m0 < Lq; m0 += 16) {
    const int r0 = m0 + g, r1 = r0 + 8;
    if (m0 + 16 < Lq) load_q(m0 + 16, qn);
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      // one ldmatrix.x4 = the (b0, b1) fragments of two 16-wide k-steps for these 8 keys
      const __nv_bfloat16* kr = Ks + (8 * j + (lane & 7)) * kLds + 8 * (lane >> 3);
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kb[4];
        ldmatrix_x4(kb, kr + kp * 32);
        mma_bf16_16816(s[j], qa[2 * kp], kb[0], kb[1]);
        mma_bf16_16816(s[j], qa[2 * kp + 1], kb[2], kb[3]);
      }
    }
    // scores / sqrt(64) + additive mask (modeling.py:333-336); row max over this row's 8*NT keys
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      // (q.k / 8 + mask) * log2(e): the softmax below is then a bare ex2
      s[j][0] = fmaf(s[j][0], kScaleLog2e, madd[j][0]);
      s[j][1] = fmaf(s[j][1], kScaleLog2e, madd[j][1]);
      s[j][2] = fmaf(s[j][2], kScaleLog2e, madd[j][0]);
      s[j][3] = fmaf(s[j][3], kScaleLog2e, madd[j][1]);
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[KS][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float p0 = ex2_fast(s[j][0] - mx0), p1 = ex2_fast(s[j][1] - mx0);
      const float p2 = ex2_fast(s[j][2] - mx1), p3 = ex2_fast(s[j][3] - mx1);
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      // accumulator tiles (2ks, 2ks+1) of S are exactly the A fragment of k-step ks of P·V
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vb[4];
        const __nv_bfloat16* vrow = Vs + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kLds + dp * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(vb, vrow);
        mma_bf16_16816(o[2 * dp], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(o[2 * dp + 1], pa[ks], vb[2], vb[3]);
      }
    }
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    __nv_bfloat16* o0 = out + (static_cast<long long>(b) * Lq + r0) * ldo + h * HD + 2 * t;
    __nv_bfloat16* o1 = out + (static_cast<long long>(b) * Lq + r1) * ldo + h * HD + 2 * t;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      if (r0 < Lq) *reinterpret_cast<uint32_t*>(o0 + 8 * d) = pack_bf16x2(o[d][0] * inv0, o[d][1] * inv0);
      if (r1 < Lq) *reinterpret_cast<uint32_t*>(o1 + 8 * d) = pack_bf16x2(o[d][2] * inv1, o[d][3] * inv1);
    }
    if (m0 + 16 < Lq) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int i = 0; i < 4; ++i) qa[ks][i] = qn[ks][i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-fed persistent variant (large batches).  The per-warp kernel above spends 46 % of its stall samples waiting for
// its own loads (12-16 warps / SM, each load -> compute -> store in sequence).  Here one CTA per SM streams through a
// contiguous range of (sample, head) units.  Each of the kAttnWarps compute warps owns DEPTH slots of shared memory
// (Q, K, V tiles of one unit: [L rows x 128 B], 128-byte swizzle -> conflict-free ldmatrix); a producer warp TMA-loads
// unit i into slot (i / W) % DEPTH of warp i % W, so every warp always has its next unit(s) in flight while it computes,
// and every mbarrier has exactly one producer and one consumer walking it in order.  The output tile is written over the
// consumed Q rows (same swizzle) and leaves through a TMA store; the slot goes back to the producer when that store has
// read it.  Tiles are exactly L rows: fragment rows beyond a tile (query / key padding up to a multiple of 16) alias the
// next tile -- always finite bf16 data or the zeros of the one-time clear -- under a score of -inf / probability 0 /
// a row guard, so they never reach the output.  The swizzle is computed from absolute shared-memory address bits (what
// the TMA unit does), which is why tiles need only 128-byte alignment.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 8;
// zeroed bytes behind the last slot: fragment rows up to 15 lines past a tile (line stride up to 384 B) land here
constexpr int kAttnPad = 6144;

struct AttnTmaParams {
  const int32_t* key_mask;
  int Lq, Lk, heads, items;
  int slot_bytes;       // (Lq + 2 Lk) * 128
  int q_stride;         // bytes between consecutive Q (= output) lines of a slot: 384 merged, 128 otherwise
  int kv_stride;        // bytes between consecutive K (or V) lines: 384 merged, 256 otherwise
  int k_off, v_off;     // first K / V line, from the slot base
  int merged;           // self-attention: ONE 3-D box (64 cols x {q,k,v} x L rows) per unit; else Q box + {k,v} box
  int planes;           // operands come from the head-major plane tensor of gemm_bf16_qkv_planes: boxes are (64, L, n planes)
  int q_row0, kv_row0;  // planes: first query / key row of this call inside the plane tensor
  int depth;            // slots per compute warp
  int debug;            // measurement switches (RGQA_ATTN_DEBUG): 1 no stores, 2 no P.V, 4 no math at all
};

// shared-memory address of 16-byte chunk `chunk` (0..7) of the 128-byte line at `line_addr` under the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_addr(uint32_t line_addr, int chunk) {
  return line_addr + ((static_cast<uint32_t>(chunk) ^ ((line_addr >> 7) & 7u)) << 4);
}

// tmA: merged -> the (64, 3, L) box over [rows][q|k|v][768]; else the 2-D Q box.  tmB: the (64, 2, Lk) {k, v} box (unused
// when merged).  tmO: merged -> 3-D map whose middle extent is 1, so only the q-lines of the (64, 3, Lq) box are stored
// (TMA drops out-of-range box elements); else the 2-D output box.
template <int KS>
__global__ void __launch_bounds__((1 + kAttnWarps) * 32, 1)
attention_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const AttnTmaParams p) {
  constexpr int NT = 2 * KS;
  extern __shared__ uint8_t smem_attn_raw[];
  const uint32_t raw = smem_u32(smem_attn_raw);
  uint8_t* smem = smem_attn_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int slot_bytes = p.slot_bytes;
  const int nslot = kAttnWarps * p.depth;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(nslot) * slot_bytes + kAttnPad);
  uint64_t* empty_bar = full_bar + nslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // clear the ring once (fragment rows that alias a never-written neighbour must not hold NaN patterns), then hand it
  // to the async proxy
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (nslot * slot_bytes + kAttnPad) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslot; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  fence_proxy_async_smem();
  __syncthreads();
  // launched with the programmatic-stream-serialization attribute (launch_pdl): everything above overlapped the producing
  // GEMM's tail; Q / K / V and the key mask are read only below (the producer warp could otherwise TMA-load a
  // projection the Q|K|V GEMM was still writing when option "pdl" is on)
  pdl_launch_dependents();
  pdl_wait();

  // contiguous range of units for this CTA: consecutive heads of consecutive samples (adjacent 128-byte columns of the
  // same token rows) are requested back to back
  const int per = (p.items + gridDim.x - 1) / gridDim.x;
  const int u0 = blockIdx.x * per;
  const int n = min(p.items, u0 + per) - u0;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t bytes = static_cast<uint32_t>(slot_bytes);
      for (int i = 0; i < n; ++i) {
        const int w = i % kAttnWarps, k = i / kAttnWarps;
        const int slot = w * p.depth + k % p.depth;
        const int u = u0 + i;
        const int b = u / p.heads, h = u % p.heads;
        uint8_t* dst = smem + static_cast<size_t>(slot) * slot_bytes;
        if (k >= p.depth) mbar_wait(&empty_bar[slot], ((k / p.depth) - 1) & 1u);
        mbar_expect_tx(&full_bar[slot], bytes);
        if (p.planes) {  // plane index = head * 3 + {q, k, v}; Q, K, V tiles land back to back
          if (p.merged) {
            tma_load_3d(dst, &tmA, &full_bar[slot], 0, p.q_row0 + b * p.Lq, 3 * h);
          } else {
            tma_load_3d(dst, &tmA, &full_bar[slot], 0, p.q_row0 + b * p.Lq, 3 * h);
            tma_load_3d(dst + p.k_off, &tmB, &full_bar[slot], 0, p.kv_row0 + b * p.Lk, 3 * h + 1);
          }
        } else if (p.merged) {
          tma_load_3d(dst, &tmA, &full_bar[slot], h * HD, 0, b * p.Lq);
        } else {
          tma_load_2d(dst, &tmA, &full_bar[slot], h * HD, b * p.Lq);
          tma_load_3d(dst + p.k_off, &tmB, &full_bar[slot], h * HD, 0, b * p.Lk);
        }
      }
    }
    return;
  }

  const int cw = warp - 1;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t qs = static_cast<uint32_t>(p.q_stride), kvs = static_cast<uint32_t>(p.kv_stride);
  int pending_slot = -1;  // slot whose output store may still be reading shared memory
  int k = 0;
  for (int i = cw; i < n; i += kAttnWarps, ++k) {
    const int slot = cw * p.depth + k % p.depth;
    const int u = u0 + i;
    const int b = u / p.heads, h = u % p.heads;
    const uint32_t Qs = smem_u32(smem + static_cast<size_t>(slot) * slot_bytes);
    const uint32_t Ks = Qs + p.k_off;
    const uint32_t Vs = Qs + p.v_off;

    float madd[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = 8 * j + 2 * t + e;
        madd[j][e] = key >= p.Lk ? -INFINITY
                                 : ((p.key_mask != nullptr && __ldg(p.key_mask + b * p.Lk + key) == 0) ? -10000.f * kLog2e : 0.f);
      }
    if (pending_slot >= 0) {  // the previous unit's store has had this long to read its lines: give the slot back
      if (lane == 0) {
        tma_store_wait_read<0>();
        mbar_arrive(&empty_bar[pending_slot]);
      }
      pending_slot = -1;
    }
    mbar_wait(&full_bar[slot], (k / p.depth) & 1u);
    if (p.debug & 4) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
      continue;
    }

    for (int m0 = 0; m0 < p.Lq; m0 += 16) {
      const int r0 = m0 + g, r1 = r0 + 8;
      // A fragments of the 16 query rows: one ldmatrix.x4 per 16-wide k-step
      uint32_t qa[4][4];
      {
        const uint32_t line = Qs + static_cast<uint32_t>(m0 + (lane & 15)) * qs;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(qa[ks][0]), "=r"(qa[ks][1]), "=r"(qa[ks][2]), "=r"(qa[ks][3])
                       : "r"(sw128_addr(line, 2 * ks + (lane >> 4))));
        }
      }
      float s[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const uint32_t line = Ks + static_cast<uint32_t>(8 * j + (lane & 7)) * kvs;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t kb[4];
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(kb[0]), "=r"(kb[1]), "=r"(kb[2]), "=r"(kb[3])
                       : "r"(sw128_addr(line, 4 * kp + (lane >> 3))));
          mma_bf16_16816(s[j], qa[2 * kp], kb[0], kb[1]);
          mma_bf16_16816(s[j], qa[2 * kp + 1], kb[2], kb[3]);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        s[j][0] = fmaf(s[j][0], kScaleLog2e, madd[j][0]);
        s[j][1] = fmaf(s[j][1], kScaleLog2e, madd[j][1]);
        s[j][2] = fmaf(s[j][2], kScaleLog2e, madd[j][0]);
        s[j][3] = fmaf(s[j][3], kScaleLog2e, madd[j][1]);
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pa[KS][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float p0 = ex2_fast(s[j][0] - mx0), p1 = ex2_fast(s[j][1] - mx0);
        const float p2 = ex2_fast(s[j][2] - mx1), p3 = ex2_fast(s[j][3] - mx1);
        sum0 += p0 + p1;
        sum1 += p2 + p3;
        pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

      float o[8][4];
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if (p.debug & 2) break;
        const uint32_t line = Vs + static_cast<uint32_t>(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kvs;
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vb[4];
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(vb[0]), "=r"(vb[1]), "=r"(vb[2]), "=r"(vb[3])
                       : "r"(sw128_addr(line, 2 * dp + (lane >> 4))));
          mma_bf16_16816(o[2 * dp], pa[ks], vb[0], vb[1]);
          mma_bf16_16816(o[2 * dp + 1], pa[ks], vb[2], vb[3]);
        }
      }
      // the 16 query rows of this tile are in registers since the top of the iteration: their lines now take the output
      // (rows >= Lq alias other lines of the ring and are skipped)
      const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
      __syncwarp();  // every lane's ldmatrix of these Q lines has completed
      const uint32_t l0 = Qs + static_cast<uint32_t>(r0) * qs, l1 = Qs + static_cast<uint32_t>(r1) * qs;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        if (r0 < p.Lq)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw128_addr(l0, d) + 4u * t), "r"(pack_bf16x2(o[d][0] * inv0, o[d][1] * inv0)) : "memory");
        if (r1 < p.Lq)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sw128_addr(l1, d) + 4u * t), "r"(pack_bf16x2(o[d][2] * inv1, o[d][3] * inv1)) : "memory");
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (!(p.debug & 1)) {
        const uint8_t* src = smem + static_cast<size_t>(slot) * slot_bytes;
        if (p.merged && !p.planes)
          tma_store_3d(&tmO, src, h * HD, 0, b * p.Lq);
        else
          tma_store_2d(&tmO, src, h * HD, b * p.Lq);
      }
      tma_store_commit();
    }
    pending_slot = slot;
  }
  if (lane == 0) tma_store_wait<0>();  // shared memory must outlive the last store's read
}

template <int KS>
int launch_attention_tma(cudaStream_t stream, const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, long long ldk,
                         const __nv_bfloat16* v, long long ldv, const int32_t* key_mask, __nv_bfloat16* out, long long ldo,
                         int B, int Lq, int Lk, int heads, int depth) {
  CUtensorMap tmA, tmB, tmO;
  const long long cols = static_cast<long long>(heads) * HD;
  const long long kv_gap = (v - k) * 2;  // bytes between a row's K and V blocks
  const bool merged = (k - q) * 2 == kv_gap && ldq == ldk && Lq == Lk;
  AttnTmaParams p;
  if (merged) {  // self-attention over one fused [rows][q|k|v] buffer
    if (make_tmap_3d(&tmA, q, cols, 3, static_cast<long long>(B) * Lq, kv_gap, ldq * 2, HD, 3, Lq) ||
        make_tmap_3d(&tmO, out, cols, 1, static_cast<long long>(B) * Lq, ldo * 2 /*unused: extent 1*/, ldo * 2, HD, 3, Lq))
      return 1;
    tmB = tmA;
    p.q_stride = p.kv_stride = 384;
    p.k_off = 128;
    p.v_off = 256;
  } else {
    if (make_tmap_ex(&tmA, q, static_cast<long long>(B) * Lq, cols, ldq, Lq, HD, 2, 128) ||
        make_tmap_3d(&tmB, k, cols, 2, static_cast<long long>(B) * Lk, kv_gap, ldk * 2, HD, 2, Lk) ||
        make_tmap_ex(&tmO, out, static_cast<long long>(B) * Lq, cols, ldo, Lq, HD, 2, 128))
      return 1;
    p.q_stride = 128;
    p.kv_stride = 256;
    p.k_off = Lq * 128;
    p.v_off = Lq * 128 + 128;
  }
  p.merged = merged ? 1 : 0;
  p.planes = 0;
  p.q_row0 = p.kv_row0 = 0;
  p.key_mask = key_mask;
  p.Lq = Lq;
  p.Lk = Lk;
  p.heads = heads;
  p.items = B * heads;
  p.slot_bytes = (Lq + 2 * Lk) * 128;
  p.depth = depth;
  static const int debug_env = [] {  // measurement switches, read once
    const char* d = getenv("RGQA_ATTN_DEBUG");
    return d ? atoi(d) : 0;
  }();
  p.debug = debug_env;
  const int smem = kAttnWarps * depth * p.slot_bytes + 1024 + kAttnPad + 1024;
  static bool cfg = false;
  if (!cfg) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(attention_tma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg = true;
  }
  int grid = num_sms();
  if (grid > p.items) grid = p.items;
  RGQA_CUDA_OK(launch_pdl(attention_tma_kernel<KS>, dim3(grid), dim3((1 + kAttnWarps) * 32), smem, stream, tmA, tmB, tmO, p));
  return 0;
}

template <int KS>
int launch_attention_planes(cudaStream_t stream, const __nv_bfloat16* planes, long long plane_rows, long long q_row0,
                            long long kv_row0, const int32_t* key_mask, __nv_bfloat16* out, long long ldo, int B, int Lq, int Lk,
                            int heads, int depth) {
  CUtensorMap tmA, tmB, tmO;
  const bool self = (q_row0 == kv_row0) && (Lq == Lk);
  const long long np = 3LL * heads;
  if (make_tmap_3d(&tmA, planes, HD, plane_rows, np, 128, plane_rows * 128, HD, Lq, self ? 3 : 1) ||
      make_tmap_3d(&tmB, planes, HD, plane_rows, np, 128, plane_rows * 128, HD, Lk, 2) ||
      make_tmap_ex(&tmO, out, static_cast<long long>(B) * Lq, static_cast<long long>(heads) * HD, ldo, Lq, HD, 2, 128))
    return 1;
  AttnTmaParams p;
  p.key_mask = key_mask;
  p.Lq = Lq;
  p.Lk = Lk;
  p.heads = heads;
  p.items = B * heads;
  p.slot_bytes = (Lq + 2 * Lk) * 128;
  p.q_stride = p.kv_stride = 128;
  p.k_off = Lq * 128;
  p.v_off = (Lq + Lk) * 128;
  p.merged = self ? 1 : 0;
  p.planes = 1;
  p.q_row0 = static_cast<int>(q_row0);
  p.kv_row0 = static_cast<int>(kv_row0);
  p.depth = depth;
  p.debug = 0;
  const int smem = kAttnWarps * depth * p.slot_bytes + 1024 + kAttnPad + 1024;
  static bool cfg = false;
  if (!cfg) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(attention_tma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cfg = true;
  }
  int grid = num_sms();
  if (grid > p.items) grid = p.items;
  RGQA_CUDA_OK(launch_pdl(attention_tma_kernel<KS>, dim3(grid), dim3((1 + kAttnWarps) * 32), smem, stream, tmA, tmB, tmO, p));
  return 0;
}

// slots per compute warp that fit next to the barriers and the overrun pad; < 2 means no prefetch: use the per-warp kernel
inline int attention_tma_depth(int Lq, int Lk) {
  const int slot = (Lq + 2 * Lk) * 128;
  const int budget = 227 * 1024 - 1024 /*alignment*/ - kAttnPad /*overrun pad*/ - 1024 /*barriers*/;
  int depth = budget / (kAttnWarps * slot);
  return depth > 4 ? 4 : depth;
}

template <int KS>
int launch_attention_mma(cudaStream_t stream, const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, long long ldk,
                         const __nv_bfloat16* v, long long ldv, const int32_t* key_mask, __nv_bfloat16* out, long long ldo,
                         int B, int Lq, int Lk, int heads) {
  const int items = B * heads;
  const size_t smem = static_cast<size_t>(4) * 2 * (16 * KS) * kLds * sizeof(__nv_bfloat16);
  static bool cfg = false;
  if (!cfg) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    RGQA_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel<KS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cfg = true;
  }
  RGQA_CUDA_OK(launch_pdl(attention_mma_kernel<KS>, dim3((items + 3) / 4), dim3(128), smem, stream, q, ldq, k, ldk, v, ldv, key_mask, out, ldo, Lq, Lk, heads, items));
  return 0;
}

}  // namespace

int attention_planes_ok(int B, int Lq, int Lk, int heads) {
  return option_get(OPT_ATTN_TMA) != 0 && attention_tma_depth(Lq, Lk) >= 2 && static_cast<long long>(B) * heads >= 16LL * num_sms();
}

int attention_planes(cudaStream_t stream, const void* planes, long long plane_rows, long long q_row0, long long kv_row0,
                     const int32_t* key_mask, void* out, long long ldo, int B, int Lq, int Lk, int heads) {
  RGQA_CHECK(planes && out && B > 0 && heads > 0, "attention_planes: bad arguments");
  RGQA_CHECK(attention_planes_ok(B, Lq, Lk, heads), "attention_planes: shape B=%d Lq=%d Lk=%d is not served by the TMA kernel", B, Lq, Lk);
  RGQA_CHECK(q_row0 + static_cast<long long>(B) * Lq <= plane_rows && kv_row0 + static_cast<long long>(B) * Lk <= plane_rows,
             "attention_planes: rows outside the plane tensor");
  const int depth = attention_tma_depth(Lq, Lk);
  const __nv_bfloat16* pb = static_cast<const __nv_bfloat16*>(planes);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out);
  void* tok = prof_begin(stream);
  int rc = 0;
  switch ((Lk + 15) / 16) {
    case 1: rc = launch_attention_planes<1>(stream, pb, plane_rows, q_row0, kv_row0, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
    case 2: rc = launch_attention_planes<2>(stream, pb, plane_rows, q_row0, kv_row0, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
    case 3: rc = launch_attention_planes<3>(stream, pb, plane_rows, q_row0, kv_row0, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
    default: rc = launch_attention_planes<4>(stream, pb, plane_rows, q_row0, kv_row0, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
  }
  if (rc) return 1;
  prof_end(tok, stream, 4.0 * B * static_cast<double>(Lq) * Lk * heads * HD, PROF_ATTENTION);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

int attention(cudaStream_t stream, int dtype, const void* q, long long ldq, const void* k, long long ldk,
              const void* v, long long ldv, const int32_t* key_mask, void* out, long long ldo, int B, int Lq,
              int Lk, int heads, float drop_p, unsigned long long drop_seed) {
  RGQA_CHECK(drop_p == 0.f || (dtype == DT_F32 && drop_p > 0.f && drop_p < 1.f), "attention: dropout is an fp32 training feature");
  RGQA_CHECK(B > 0 && heads > 0, "attention: empty batch");
  RGQA_CHECK(Lq >= 1 && Lq <= MAXL && Lk >= 1 && Lk <= MAXL, "attention: Lq=%d Lk=%d out of range [1,%d]", Lq, Lk,
             MAXL);
  const int esz = dtype == DT_F32 ? 4 : 2;
  RGQA_CHECK((ldq * esz) % 16 == 0 && (ldk * esz) % 16 == 0 && (ldv * esz) % 16 == 0 && (ldo * esz) % 16 == 0,
             "attention: row strides must be multiples of 16 bytes");
  const size_t smem = static_cast<size_t>(Lk) * (2 * HD + kThreads + 1) * sizeof(float);
  const int grid = B * heads;
  (void)smem;
  void* tok = prof_begin(stream);
  if (dtype == DT_F32) {
    static bool cfg = false;
    if (!cfg) {
      RGQA_CUDA_OK(cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      cfg = true;
    }
    attention_kernel<float><<<grid, kThreads, smem, stream>>>(
        static_cast<const float*>(q), ldq, static_cast<const float*>(k), ldk, static_cast<const float*>(v), ldv,
        key_mask, static_cast<float*>(out), ldo, Lq, Lk, heads, drop_p, drop_seed);
  } else {
    const __nv_bfloat16* qb = static_cast<const __nv_bfloat16*>(q);
    const __nv_bfloat16* kb = static_cast<const __nv_bfloat16*>(k);
    const __nv_bfloat16* vb = static_cast<const __nv_bfloat16*>(v);
    __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out);
    int rc = 0;
    // large batches: TMA-fed persistent kernel (needs >= 2 units per compute warp, room for two slots per warp, and
    // 16-byte aligned operands)
    const int depth = attention_tma_depth(Lq, Lk);
    const bool tma_ok = option_get(OPT_ATTN_TMA) != 0 && depth >= 2 &&
                        static_cast<long long>(B) * heads >= 2LL * num_sms() * kAttnWarps &&
                        (reinterpret_cast<uintptr_t>(q) % 16 == 0) && (reinterpret_cast<uintptr_t>(k) % 16 == 0) &&
                        (reinterpret_cast<uintptr_t>(v) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                        ldk == ldv && vb > kb && ((vb - kb) * 2) % 16 == 0 && (vb - kb) <= ldk - heads * HD;
    if (tma_ok) {
      switch ((Lk + 15) / 16) {
        case 1: rc = launch_attention_tma<1>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
        case 2: rc = launch_attention_tma<2>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
        case 3: rc = launch_attention_tma<3>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
        default: rc = launch_attention_tma<4>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads, depth); break;
      }
    } else
    switch ((Lk + 15) / 16) {
      case 1: rc = launch_attention_mma<1>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads); break;
      case 2: rc = launch_attention_mma<2>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads); break;
      case 3: rc = launch_attention_mma<3>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads); break;
      default: rc = launch_attention_mma<4>(stream, qb, ldq, kb, ldk, vb, ldv, key_mask, ob, ldo, B, Lq, Lk, heads); break;
    }
    if (rc) return 1;
  }
  prof_end(tok, stream, 4.0 * B * static_cast<double>(Lq) * Lk * heads * HD, PROF_ATTENTION);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
