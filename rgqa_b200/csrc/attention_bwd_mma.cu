// rgqa_b200 — BertAttention core backward (src/lxrt/modeling.py:330-346) on the tensor cores, for the mixed-precision training
// step (bf16 matrix products, fp32 softmax and accumulation):
//
//   S = Q K^T / 8 + mask,  P = softmax(S),  dP = dO V^T,  dS = P o (dP - rowsum(dP o P)),
//   dQ = dS K / 8,  dK = dS^T Q / 8,  dV = P^T dO
//
// One CTA of four warps per (sample, head).  Q, K, V and dO (fp32 in global memory, the layout of the fp32 path's kernel) are cast
// to bf16 into padded shared-memory rows (72 elements: conflict-free ldmatrix); warp w owns the 16-query tile w for the
// row-wise half (S, dP as m16n8k16 MMAs, softmax / dS in fp32 registers, dQ with dS as the A operand straight from the
// accumulators) and leaves P and dS as bf16 in shared memory; after one barrier warp w owns the 16-key tile w for the
// column-wise half (dV = P^T dO, dK = dS^T Q: A operands read TRANSPOSED from the stored P / dS with ldmatrix.trans).
// The fp32 path (`attention_bwd_kernel`, backward.cu) stays the reference implementation and the one the input-gradient scorers
// pin against autograd; this one is checked against it and against autograd at bf16 tolerances.
#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int HD = 64;
constexpr int LDS = 72;   // bf16 elements per shared-memory row (144 B: eight consecutive rows hit eight different 16-byte bank groups)
constexpr float kLog2e = 1.44269504088896340736f;

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// A fragment (m16 x k16) of a matrix stored row-major [m][k]: rows m0.., columns k0..
__device__ __forceinline__ void load_a(uint32_t (&a)[4], uint32_t base, int m0, int k0, int lane) {
  ldsm4(a, base + static_cast<uint32_t>(((m0 + (lane & 15)) * LDS + k0 + (lane >> 4) * 8) * 2));
}
// A fragment (m16 x k16) of the TRANSPOSE of a matrix stored row-major [k][m]: A[m][k] = T[k][m]
__device__ __forceinline__ void load_a_t(uint32_t (&a)[4], uint32_t base, int m0, int k0, int lane) {
  const int i = lane >> 3;   // 8x8 block: (m block, k block) = (i & 1, i >> 1)
  ldsm4t(a, base + static_cast<uint32_t>(((k0 + (i >> 1) * 8 + (lane & 7)) * LDS + m0 + (i & 1) * 8) * 2));
}
// B fragments (k16 x n8, "col") for TWO n blocks from a matrix stored [n][k] (K in Q K^T): {b0, b1} = n block 0, {b2, b3} = n block 1
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], uint32_t base, int n0, int k0, int lane) {
  const int i = lane >> 3;   // (n block, k block) = (i >> 1, i & 1)
  ldsm4(b, base + static_cast<uint32_t>(((n0 + (i >> 1) * 8 + (lane & 7)) * LDS + k0 + (i & 1) * 8) * 2));
}
// B fragments for TWO n blocks from a matrix stored [k][n] (V in P V): {b0, b1} = n block 0 (k 0-7, 8-15), {b2, b3} = n block 1
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], uint32_t base, int k0, int n0, int lane) {
  const int i = lane >> 3;   // (n block, k block) = (i >> 1, i & 1)
  ldsm4t(b, base + static_cast<uint32_t>(((k0 + (i & 1) * 8 + (lane & 7)) * LDS + n0 + (i >> 1) * 8) * 2));
}

// fp32 rows [L][64] (row stride ld) -> bf16 smem rows [Lpad][LDS], rows >= L zero
__device__ __forceinline__ void stage(const float* __restrict__ g, long long ld, int L, int Lpad, __nv_bfloat16* s, int tid, int nthreads, bool vec) {
  for (int e = tid; e < Lpad * 16; e += nthreads) {
    const int r = e >> 4, c = (e & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < L) {
      if (vec) v = *reinterpret_cast<const float4*>(g + r * ld + c);
      else v = make_float4(g[r * ld + c], g[r * ld + c + 1], g[r * ld + c + 2], g[r * ld + c + 3]);
    }
    *reinterpret_cast<uint2*>(s + r * LDS + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// NT = ceil(Lk / 8) score tiles, MT = ceil(Lq / 16) query tiles, KT = ceil(Lk / 16) key tiles (all <= 8 / 4 / 4)
template <int NT>
__global__ void __launch_bounds__(128)
attention_bwd_mma_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ k, long long ldk,
                         const float* __restrict__ v, long long ldv, const int32_t* __restrict__ key_mask,
                         const float* __restrict__ dctx, long long ldo, float* __restrict__ dq, long long lddq,
                         float* __restrict__ dk, long long lddk, float* __restrict__ dv, long long lddv, int Lq, int Lk, int heads, int vec) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  constexpr int KT = (NT + 1) / 2;
  constexpr int LKP = KT * 16;                      // key rows staged (zero padded)
  const int MT = (Lq + 15) >> 4;
  const int LQP = MT * 16;
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);     // [LQP][LDS]
  __nv_bfloat16* Os = Qs + LQP * LDS;                                // dO [LQP][LDS]
  __nv_bfloat16* Ks = Os + LQP * LDS;                                // [LKP][LDS]
  __nv_bfloat16* Vs = Ks + LKP * LDS;                                // [LKP][LDS]
  __nv_bfloat16* Ps = Vs + LKP * LDS;                                // P  [LQP][LDS] (columns = keys)
  __nv_bfloat16* Ds = Ps + LQP * LDS;                                // dS [LQP][LDS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  stage(q + static_cast<long long>(b) * Lq * ldq + h * HD, ldq, Lq, LQP, Qs, tid, 128, vec != 0);
  stage(dctx + static_cast<long long>(b) * Lq * ldo + h * HD, ldo, Lq, LQP, Os, tid, 128, vec != 0);
  stage(k + static_cast<long long>(b) * Lk * ldk + h * HD, ldk, Lk, LKP, Ks, tid, 128, vec != 0);
  stage(v + static_cast<long long>(b) * Lk * ldv + h * HD, ldv, Lk, LKP, Vs, tid, 128, vec != 0);
  __syncthreads();
  const uint32_t Qb = smem_u32(Qs), Ob = smem_u32(Os), Kb = smem_u32(Ks), Vb = smem_u32(Vs), Pb = smem_u32(Ps), Db = smem_u32(Ds);
  const int g = lane >> 2, t = lane & 3;

  // ---- row-wise half: query tile `warp` ----
  for (int mt = warp; mt < MT; mt += 4) {
    const int q0 = mt * 16;
    uint32_t qa[4][4], oa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      load_a(qa[ks], Qb, q0, ks * 16, lane);
      load_a(oa[ks], Ob, q0, ks * 16, lane);
    }
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
#pragma unroll
    for (int jp = 0; jp < KT; ++jp) {      // two score tiles (16 keys) per pass
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kb[4], vb[4];
        load_b_nk(kb, Kb, jp * 16, ks * 16, lane);
        load_b_nk(vb, Vb, jp * 16, ks * 16, lane);   // dP = dO V^T: V stored [key][d] is the [n][k] operand, like K
        mma_16816(s[2 * jp], qa[ks], kb[0], kb[1]);
        mma_16816(dp[2 * jp], oa[ks], vb[0], vb[1]);
        if (2 * jp + 1 < NT) {
          mma_16816(s[2 * jp + 1], qa[ks], kb[2], kb[3]);
          mma_16816(dp[2 * jp + 1], oa[ks], vb[2], vb[3]);
        }
      }
    }
    // (q k / 8 + mask) log2(e); masked keys take the reference's additive -10000 (modeling.py:865), keys past Lk -inf
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int ee = 0; ee < 2; ++ee) {
        const int key = 8 * j + 2 * t + ee;
        float madd = -INFINITY;
        if (key < Lk) madd = (key_mask != nullptr && key_mask[b * Lk + key] == 0) ? -10000.0f * kLog2e : 0.f;
        s[j][ee] = fmaf(s[j][ee], 0.125f * kLog2e, madd);
        s[j][2 + ee] = fmaf(s[j][2 + ee], 0.125f * kLog2e, madd);
      }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = ex2f(s[j][0] - mx0);
      s[j][1] = ex2f(s[j][1] - mx0);
      s[j][2] = ex2f(s[j][2] - mx1);
      s[j][3] = ex2f(s[j][3] - mx1);
      sum0 += s[j][0] + s[j][1];
      sum1 += s[j][2] + s[j][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;          // P
      dot0 = fmaf(dp[j][0], s[j][0], fmaf(dp[j][1], s[j][1], dot0));
      dot1 = fmaf(dp[j][2], s[j][2], fmaf(dp[j][3], s[j][3], dot1));
    }
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1);
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
    // rows of this tile past Lq were staged as zeros (uniform P, zero dO): they must not reach dK / dV
    const bool ok0 = q0 + g < Lq, ok1 = q0 + g + 8 < Lq;
    uint32_t da[KT][4];   // dS as the A operand of dQ = dS K (accumulator layout == A fragment layout, two score tiles per k16 step)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float d0 = ok0 ? s[j][0] * (dp[j][0] - dot0) : 0.f, d1 = ok0 ? s[j][1] * (dp[j][1] - dot0) : 0.f;
      const float d2 = ok1 ? s[j][2] * (dp[j][2] - dot1) : 0.f, d3 = ok1 ? s[j][3] * (dp[j][3] - dot1) : 0.f;
      const uint32_t p01 = pack_bf16x2(ok0 ? s[j][0] : 0.f, ok0 ? s[j][1] : 0.f), p23 = pack_bf16x2(ok1 ? s[j][2] : 0.f, ok1 ? s[j][3] : 0.f);
      const uint32_t d01 = pack_bf16x2(d0, d1), d23 = pack_bf16x2(d2, d3);
      *reinterpret_cast<uint32_t*>(Ps + (q0 + g) * LDS + 8 * j + 2 * t) = p01;
      *reinterpret_cast<uint32_t*>(Ps + (q0 + g + 8) * LDS + 8 * j + 2 * t) = p23;
      *reinterpret_cast<uint32_t*>(Ds + (q0 + g) * LDS + 8 * j + 2 * t) = d01;
      *reinterpret_cast<uint32_t*>(Ds + (q0 + g + 8) * LDS + 8 * j + 2 * t) = d23;
      da[j >> 1][(j & 1) * 2 + 0] = d01;
      da[j >> 1][(j & 1) * 2 + 1] = d23;
    }
    if (NT & 1) {   // the odd last score tile has no partner: its half of the k16 step is zero, in registers and in shared memory
      da[KT - 1][2] = da[KT - 1][3] = 0u;
      *reinterpret_cast<uint32_t*>(Ps + (q0 + g) * LDS + 8 * NT + 2 * t) = 0u;
      *reinterpret_cast<uint32_t*>(Ps + (q0 + g + 8) * LDS + 8 * NT + 2 * t) = 0u;
      *reinterpret_cast<uint32_t*>(Ds + (q0 + g) * LDS + 8 * NT + 2 * t) = 0u;
      *reinterpret_cast<uint32_t*>(Ds + (q0 + g + 8) * LDS + 8 * NT + 2 * t) = 0u;
    }
    // dQ = dS K / 8: K stored [key][d] is the [k][n] operand
    float o[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KT; ++ks) {
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t kb[4];
        load_b_kn(kb, Kb, ks * 16, dpair * 16, lane);
        mma_16816(o[2 * dpair], da[ks], kb[0], kb[1]);
        mma_16816(o[2 * dpair + 1], da[ks], kb[2], kb[3]);
      }
    }
    float* dq0 = dq + (static_cast<long long>(b) * Lq + q0 + g) * lddq + h * HD + 2 * t;
    float* dq1 = dq0 + 8 * lddq;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      if (ok0) *reinterpret_cast<float2*>(dq0 + 8 * d) = make_float2(o[d][0] * 0.125f, o[d][1] * 0.125f);
      if (ok1) *reinterpret_cast<float2*>(dq1 + 8 * d) = make_float2(o[d][2] * 0.125f, o[d][3] * 0.125f);
    }
  }
  __syncthreads();
  // ---- column-wise half: key tile `warp`: dV = P^T dO, dK = dS^T Q / 8 (sum over the query rows, k16 steps of 16 queries) ----
  for (int kt = warp; kt < KT; kt += 4) {
    const int k0 = kt * 16;
    float ov[8][4], ok[8][4];
#pragma unroll
    for (int d = 0; d < 8; ++d) ov[d][0] = ov[d][1] = ov[d][2] = ov[d][3] = ok[d][0] = ok[d][1] = ok[d][2] = ok[d][3] = 0.f;
    for (int qs = 0; qs < MT; ++qs) {
      uint32_t pa[4], sa[4];
      load_a_t(pa, Pb, k0, qs * 16, lane);    // A[m = key][k = query] = P[query][key]
      load_a_t(sa, Db, k0, qs * 16, lane);
#pragma unroll
      for (int dpair = 0; dpair < 4; ++dpair) {
        uint32_t ob[4], qb[4];
        load_b_kn(ob, Ob, qs * 16, dpair * 16, lane);   // B[k = query][n = d] = dO[query][d]
        load_b_kn(qb, Qb, qs * 16, dpair * 16, lane);
        mma_16816(ov[2 * dpair], pa, ob[0], ob[1]);
        mma_16816(ov[2 * dpair + 1], pa, ob[2], ob[3]);
        mma_16816(ok[2 * dpair], sa, qb[0], qb[1]);
        mma_16816(ok[2 * dpair + 1], sa, qb[2], qb[3]);
      }
    }
    const bool r0 = k0 + g < Lk, r1 = k0 + g + 8 < Lk;
    float* dv0 = dv + (static_cast<long long>(b) * Lk + k0 + g) * lddv + h * HD + 2 * t;
    float* dk0 = dk + (static_cast<long long>(b) * Lk + k0 + g) * lddk + h * HD + 2 * t;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      if (r0) {
        *reinterpret_cast<float2*>(dv0 + 8 * d) = make_float2(ov[d][0], ov[d][1]);
        *reinterpret_cast<float2*>(dk0 + 8 * d) = make_float2(ok[d][0] * 0.125f, ok[d][1] * 0.125f);
      }
      if (r1) {
        *reinterpret_cast<float2*>(dv0 + 8 * lddv + 8 * d) = make_float2(ov[d][2], ov[d][3]);
        *reinterpret_cast<float2*>(dk0 + 8 * lddk + 8 * d) = make_float2(ok[d][2] * 0.125f, ok[d][3] * 0.125f);
      }
    }
  }
}

template <int NT>
int launch_bwd(cudaStream_t stream, const float* q, long long ldq, const float* k, long long ldk, const float* v, long long ldv,
               const int32_t* key_mask, const float* dctx, long long ldo, float* dq, long long lddq, float* dk, long long lddk,
               float* dv, long long lddv, int B, int Lq, int Lk, int heads, int vec) {
  const int lqp = ((Lq + 15) / 16) * 16, lkp = ((NT + 1) / 2) * 16;
  const size_t smem = static_cast<size_t>(4 * lqp + 2 * lkp) * LDS * 2;
  auto kern = attention_bwd_mma_kernel<NT>;
  static bool cfg = false;
  if (!cfg) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    cfg = true;
  }
  kern<<<B * heads, 128, smem, stream>>>(q, ldq, k, ldk, v, ldv, key_mask, dctx, ldo, dq, lddq, dk, lddk, dv, lddv, Lq, Lk, heads, vec);
  return 0;
}

}  // namespace

int attention_bwd_bf16(cudaStream_t stream, const float* q, long long ldq, const float* k, long long ldk, const float* v,
                       long long ldv, const int32_t* key_mask, const float* dctx, long long ldo, float* dq, long long lddq, float* dk,
                       long long lddk, float* dv, long long lddv, int B, int Lq, int Lk, int heads) {
  RGQA_CHECK(q && k && v && dctx && dq && dk && dv, "attention_bwd_bf16: null argument");
  RGQA_CHECK(B > 0 && heads > 0 && Lq >= 1 && Lq <= 64 && Lk >= 1 && Lk <= 64, "attention_bwd_bf16: Lq=%d Lk=%d out of range [1,64]", Lq, Lk);
  auto al8 = [](const void* p, long long ld) { return reinterpret_cast<uintptr_t>(p) % 8 == 0 && ld % 2 == 0; };
  RGQA_CHECK(al8(dq, lddq) && al8(dk, lddk) && al8(dv, lddv), "attention_bwd_bf16: gradient rows must be 8-byte aligned");
  auto al16 = [](const void* p, long long ld) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % 4 == 0; };
  const int vec = (al16(q, ldq) && al16(k, ldk) && al16(v, ldv) && al16(dctx, ldo)) ? 1 : 0;
  int rc;
#define RGQA_BWD(NTV) rc = launch_bwd<NTV>(stream, q, ldq, k, ldk, v, ldv, key_mask, dctx, ldo, dq, lddq, dk, lddk, dv, lddv, B, Lq, Lk, heads, vec)
  switch ((Lk + 7) / 8) {
    case 1: RGQA_BWD(1); break;
    case 2: RGQA_BWD(2); break;
    case 3: RGQA_BWD(3); break;
    case 4: RGQA_BWD(4); break;
    case 5: RGQA_BWD(5); break;
    case 6: RGQA_BWD(6); break;
    case 7: RGQA_BWD(7); break;
    default: RGQA_BWD(8); break;
  }
#undef RGQA_BWD
  if (rc) return 1;
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
