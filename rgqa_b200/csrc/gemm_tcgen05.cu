// rgqa_b200 — bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), TMA-fed.
//
//   out[M,N] = epilogue( A[M,K] (bf16, row stride lda) · W[N,K]ᵀ (bf16, nn.Linear layout) + bias[N] )
//
// This file carries every nn.Linear of the LXMERT / UNITER forward in the bf16 path
// (reference call sites: src/lxrt/modeling.py:321-323 Q/K/V, :358 and :412 output dense,
// :399 intermediate dense, :578 pooler, src/tasks/gqa_model.py:22-27 answer head).
//
// Two persistent, warp-specialised kernels share one structure:
//   warp 0      TMA producer   global -> 128B-swizzled smem ring, mbarrier tx-count
//   warp 1      MMA issuer     one lane issues tcgen05.mma, accumulators in TMEM (2 x 256 fp32 columns:
//                              double buffered, so the epilogue of tile i overlaps the main loop of tile i+1)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       tcgen05.ld 32x32b -> registers (one accumulator ROW per thread), fused bias /
//                              erf-GELU / tanh / fp32 residual, then either
//                                staged : 16-byte st.shared into a 128B-swizzled slab + TMA bulk store
//                                         (full-line writes; the TMA clips the M and N tails), or
//                                direct : predicated global stores (tiny or unaligned outputs)
//   * gemm_bf16_tcgen05_kernel  one CTA per tile (128 x BN), for small M
//   * gemm_bf16_pair_kernel     a CTA PAIR (cta_group::2) per 256 x 256 tile: each CTA stages its own 128 rows of
//                               A and only half of the W tile, the leader issues one 256-row MMA for both SMs
// Tails: TMA zero-fills out-of-bounds rows of A (M tail) and W (N tail).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = (4 + kNumEpiWarps) * 32;
constexpr int kSlabBytes = 128 * 128;  // staged epilogue: 128 rows x 128 B, 128B-swizzled, one TMA store box

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  // ring + barriers (kStages*2 + 4) + tmem ptr, plus 1024 B alignment slack
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct GemmParams {
  int M, N, K;
  const float* bias;   // [N] fp32 or nullptr
  const float* resid;  // [M, ldr] fp32 residual stream (EPI_RESID) or nullptr
  long long ldr;
  void* out;
  long long ldo;
  // deferred LayerNorm of the A operand: A holds bf16(x) of the PRE-LayerNorm rows, W is scaled by
  // gamma, bias carries W·beta; the epilogue applies the per-row (x - mu) * rstd: out = rstd*(acc - mu*c[n]) + bias[n]
  // EPI_RESID_HL (producer of a deferred-LayerNorm hidden state): the residual is LN_prev(x_prev) recomputed from the
  // previous pre-LN rows (bf16 hi + lo pair), their statistics and the previous LayerNorm's gamma/beta
  const float* in_stats;   // [M][6][2]
  const float* in_gamma;   // [N]
  const float* in_beta;    // [N]
  float* out_stats;        // [M][6][2] partial (sum, sumsq) of the rows this GEMM writes
  const float* ln_stats;  // [M][6][2] partial (sum, sum of squares) of each 768-wide row, or nullptr
  const float* ln_c;      // [N] sum_k W'[n,k]
  float ln_eps;
  long long ln_stats_ld;  // floats between the statistics of consecutive GEMM rows (12 when dense)
  int out_heads;  // > 0: the output is the head-major Q|K|V plane tensor (gemm_bf16_qkv_planes); column block cb of 64 goes
                  // to plane (cb % heads) * 3 + cb / heads
  int tn;      // EPI_ACCUM, CTA-pair kernel: both operands are given TRANSPOSED ([K, M] and [K, N] row-major) and read MN-major
  int splits;  // EPI_ACCUM: every output tile is computed by `splits` CTAs (pairs), each over K / splits (0 or 1: no split)
  int a_f16;   // A and W are fp16 (the single-word deferred-LayerNorm state against the fp16 copy of its consumer's weight)
  int vec_ok;  // direct path: rows of out (and resid) are 16-byte aligned -> vector stores allowed
  int debug;   // measurement only (rgqa_set_option "gemm_debug"): bit0 skip TMA stores, bit1 skip epilogue math+smem
  RowReduce red;  // EPI_REDUCE: row reductions fused into the epilogue (answer-head scorers, Mahalanobis)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. far below one bf16 ulp of the output);
// two MUFU ops + a 5-term Horner instead of erff's branchy path: the FFN1 epilogue has ~24 issue
// slots per output element before it, not the MMA, paces the tile.
// 13 instructions per element: erf(|x|/sqrt2) = 1 - t(a1 + t(a2 + t(a3 + t(a4 + t a5)))) exp(-x^2/2),
// t = 1/(1 + p|x|/sqrt2); gelu = x/2 + |x/2| erf(|x|/sqrt2)  (odd extension folded into the last FMA).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  q *= t;
  const float e = ex2_approx((x * -0.72134752044448170368f) * x);  // exp(-x^2/2) = 2^(-x^2 * log2(e)/2)
  const float erf_abs = fmaf(-q, e, 1.0f);
  const float h = 0.5f * x;
  return fmaf(fabsf(h), erf_abs, h);
}

// Two elements per instruction on Blackwell's packed-fp32 pipe (FFMA2 / FMUL2): same formula as gelu_erf_fast, 18
// issue slots per pair instead of 26.
__device__ __forceinline__ float2 gelu_erf_fast2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 d = __ffma2_rn(make_float2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f), ax,
                              make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  // the polynomial with negated coefficients: q = -(a1 t + ... + a5 t^5)
  float2 q = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
  q = __ffma2_rn(q, t, make_float2(-1.421413741f, -1.421413741f));
  q = __ffma2_rn(q, t, make_float2(0.284496736f, 0.284496736f));
  q = __ffma2_rn(q, t, make_float2(-0.254829592f, -0.254829592f));
  q = __fmul2_rn(q, t);
  float2 a = __fmul2_rn(x, make_float2(-0.72134752044448170368f, -0.72134752044448170368f));
  a = __fmul2_rn(a, x);  // -x^2 * log2(e) / 2
  const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
  const float2 erf_abs = __ffma2_rn(q, e, make_float2(1.0f, 1.0f));
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  const float2 ah = __fmul2_rn(ax, make_float2(0.5f, 0.5f));
  return __ffma2_rn(ah, erf_abs, h);
}

template <int EPI>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (EPI == EPI_GELU) return gelu_erf_fast(v);
  if constexpr (EPI == EPI_TANH) return tanhf(v);
  return v;
}

// mean and 1/std of one 768-wide row from its six partial (sum, sum of squares) pairs
struct RowNorm {
  float mu, rstd;
};
__device__ __forceinline__ RowNorm row_norm(const float* __restrict__ stats, int row, bool row_ok, float eps,
                                            long long ld = 12) {
  RowNorm r{0.f, 1.f};
  if (stats != nullptr && row_ok) {
    const float4* s4 = reinterpret_cast<const float4*>(stats + static_cast<long long>(row) * ld);
    const float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
    const float sum = ((a.x + a.z) + (b.x + b.z)) + (c.x + c.z);
    const float sq = ((a.y + a.w) + (b.y + b.w)) + (c.y + c.w);
    r.mu = sum * (1.0f / 768.0f);
    const float var = fmaxf(sq * (1.0f / 768.0f) - r.mu * r.mu, 0.f);
    r.rstd = 1.0f / sqrtf(var + eps);
  }
  return r;
}

// acc + bias (+ fp32 residual) -> activation, for one 32-column chunk of one accumulator row
template <int EPI>
__device__ __forceinline__ void epilogue_math(const uint32_t (&v)[32], float (&f)[32], int row, bool row_ok, int col0,
                                              const GemmParams& p, bool aligned, RowNorm rn = RowNorm{0.f, 1.f}) {
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  // deferred LayerNorm: rstd * (acc - mu * c[n]) + bias'[n].  The folded weight rows are centred (zero sum up to bf16
  // rounding), so c is normally omitted (ln_c == nullptr) and the whole correction is the rstd factor of the bias FMA.
  if (p.ln_stats != nullptr && p.ln_c != nullptr) {
    if (col0 + 32 <= p.N && aligned) {
      const float4* c4 = reinterpret_cast<const float4*>(p.ln_c + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 c = __ldg(c4 + i);
        f[4 * i + 0] = fmaf(-rn.mu, c.x, f[4 * i + 0]);
        f[4 * i + 1] = fmaf(-rn.mu, c.y, f[4 * i + 1]);
        f[4 * i + 2] = fmaf(-rn.mu, c.z, f[4 * i + 2]);
        f[4 * i + 3] = fmaf(-rn.mu, c.w, f[4 * i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) f[i] = fmaf(-rn.mu, __ldg(p.ln_c + col0 + i), f[i]);
    }
  }
  if (col0 + 32 <= p.N && aligned) {
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i);
        // rstd == 1 outside the deferred-LayerNorm consumers; packed fp32 (FFMA2): two columns per instruction
        const float2 lo2 = __ffma2_rn(make_float2(rn.rstd, rn.rstd), make_float2(f[4 * i + 0], f[4 * i + 1]), make_float2(b.x, b.y));
        const float2 hi2 = __ffma2_rn(make_float2(rn.rstd, rn.rstd), make_float2(f[4 * i + 2], f[4 * i + 3]), make_float2(b.z, b.w));
        f[4 * i + 0] = lo2.x;
        f[4 * i + 1] = lo2.y;
        f[4 * i + 2] = hi2.x;
        f[4 * i + 3] = hi2.y;
      }
    }
    if constexpr (EPI == EPI_RESID) {
      if (row_ok) {
        const float4* r4 = reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row) * p.ldr + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 r = __ldg(r4 + i);
          f[4 * i + 0] += r.x;
          f[4 * i + 1] += r.y;
          f[4 * i + 2] += r.z;
          f[4 * i + 3] += r.w;
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = col0 + i;
      if (col < p.N) {
        if (p.bias != nullptr) f[i] = fmaf(rn.rstd, f[i], __ldg(p.bias + col));
        if constexpr (EPI == EPI_RESID) {
          if (row_ok) f[i] += p.resid[static_cast<long long>(row) * p.ldr + col];
        }
      }
    }
  }
  if constexpr (EPI == EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 g = gelu_erf_fast2(make_float2(f[i], f[i + 1]));
      f[i] = g.x;
      f[i + 1] = g.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = apply_act<EPI>(f[i]);
  }
}

// direct path: one 32-column chunk straight to global memory (vector stores when aligned, else scalar, guarded)
template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int row, bool row_ok, int col0,
                                               const GemmParams& p, RowNorm rn = RowNorm{0.f, 1.f}) {
  float f[32];
  epilogue_math<EPI>(v, f, row, row_ok, col0, p, p.vec_ok != 0, rn);
  if (!row_ok) return;
  if constexpr (EPI == EPI_ACCUM) {  // out += acc, one red.global.add per element (several CTAs share the tile along K)
    float* orow = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (col0 + i < p.N) atomicAdd(orow + i, f[i]);
    return;
  }
  if (col0 + 32 <= p.N && p.vec_ok) {
    if constexpr (sizeof(OutT) == 2) {
      uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o;
        o.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
        o.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
        o.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
        o.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
        o4[i] = o;
      }
    } else {
      float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = col0 + i;
      if (col < p.N) {
        if constexpr (sizeof(OutT) == 2)
          reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(row) * p.ldo + col] = __float2bfloat16_rn(f[i]);
        else
          reinterpret_cast<float*>(p.out)[static_cast<long long>(row) * p.ldo + col] = f[i];
      }
    }
  }
}

// staged path: the chunk goes into the 128B-swizzled slab (row r_local, 16-byte piece j at physical piece
// j ^ (r_local & 7)) that a TMA bulk store then writes out as full lines
template <typename OutT>
__device__ __forceinline__ void store_chunk_swizzled(uint8_t* slab, int r_local, int chunk_in_slab, const float (&f)[32]) {
  uint8_t* rowp = slab + r_local * 128;
  const int sw = r_local & 7;
  if constexpr (sizeof(OutT) == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 o;
      o.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]);
      o.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
      o.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
      o.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
      *reinterpret_cast<uint4*>(rowp + (((chunk_in_slab * 4 + i) ^ sw) << 4)) = o;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(rowp + ((i ^ sw) << 4)) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- EPI_REDUCE: per-row reductions over the tile's columns (see RowReduce in kernels.h) ----------------------------
struct RedArgMax {
  float v;
  int i;
};
// larger value wins; on equal values the lower index wins (first-occurrence semantics of torch.max)
__device__ __forceinline__ RedArgMax red_better(RedArgMax a, RedArgMax b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ float red_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }  // == scorers.cu sigmoid_ref
__device__ __forceinline__ float red_softplus(float x) { return logf(1.0f + expf(x)); }     // as written in the reference
// one thread's running record over the columns it sees
struct RedRec {
  RedArgMax best{-INFINITY, 0x7fffffff};
  float m1 = -INFINITY, m2 = -INFINITY;  // two largest values (duplicates count twice, like topk)
  float sp = 0.f;                        // sum of log(1 + exp(v))
  float lmax = -INFINITY, lsum = 0.f;    // online logsumexp
  float q = 0.f;                         // RED_MAHA: partial <t, x>
};
__device__ __forceinline__ void red_top2(float& m1, float& m2, float v) {
  if (v > m1) {
    m2 = m1;
    m1 = v;
  } else if (v > m2) {
    m2 = v;
  }
}
__device__ __forceinline__ void red_lse(float& lmax, float& lsum, float m, float s) {  // merge (m, s) into (lmax, lsum)
  if (m == -INFINITY) return;
  if (m > lmax) {
    lsum = lsum * expf(lmax - m) + s;  // lmax = -inf: lsum is 0 and exp(-inf) = 0
    lmax = m;
  } else {
    lsum += s * expf(m - lmax);
  }
}

// ------------------------------------------------------------------------------------------------
// single-CTA kernel (small M): 128 x BN tiles, direct epilogue
// ------------------------------------------------------------------------------------------------
template <int BN, int EPI, typename OutT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int splits = (EPI == EPI_ACCUM && p.splits > 1) ? p.splits : 1;   // work item t = (tile t / splits, K slice t % splits)
  const int num_tiles = tiles_m * tiles_n * splits;
  const int num_kb = (p.K / BK) / splits;   // k-blocks per work item (the host makes K / BK a multiple of splits)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int tile = t / splits, kb0 = (t - tile * splits) * num_kb;
      const int m0 = (tile / tiles_n) * BM;
      const int n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (lane == 0) {
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], (kb0 + kb) * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], (kb0 + kb) * BK, n0);
        }
        __syncwarp();
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = umma_idesc_bf16(BM, BN) & ~(p.a_f16 ? kIdescABf16 : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t adesc = umma_smem_desc_sw128(sa);
          const uint64_t bdesc = umma_smem_desc_sw128(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per 16-element K step inside the 128 B swizzle atom (address field is in 16 B units)
            umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == C::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int e = warp - 4;
    const int quarter = e & 3;  // == warp % 4: the TMEM lane quarter this warp may touch
    const int half = e >> 2;    // column half of the tile
    constexpr int kColsPerWarp = BN / 2;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int tile = t / splits;
      const int m0 = (tile / tiles_n) * BM;
      const int n0 = (tile % tiles_n) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const RowNorm rn = row_norm(p.ln_stats, row, row_ok, p.ln_eps, p.ln_stats_ld);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + half * kColsPerWarp);
      [[maybe_unused]] RedRec rec;
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += 32) {
        const int col0 = n0 + half * kColsPerWarp + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(t_row + static_cast<uint32_t>(c), v);
        tmem_ld_wait_dep(v);
        if constexpr (EPI == EPI_REDUCE) {
          if (p.red.kind == RED_LOGITS) {
            float* orow = (p.out != nullptr && row_ok) ? reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo : nullptr;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = col0 + i;
              if (col < p.N) {
                const float x = __uint_as_float(v[i]) + __ldg(p.bias + col);
                if (orow != nullptr) orow[col] = x;
                rec.best = red_better(rec.best, RedArgMax{red_sigmoid(x), col});
                red_top2(rec.m1, rec.m2, x);
                if (p.red.flags & 1) rec.sp += red_softplus(x);
                if (p.red.flags & 2) red_lse(rec.lmax, rec.lsum, x, 1.0f);
              }
            }
          } else {  // RED_MAHA
            if (col0 < p.red.split) {  // t = x·P columns (split is a multiple of 32): q += t * x
              const float* xr = p.red.x + static_cast<long long>(row_ok ? row : 0) * p.red.split + col0;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 xx = __ldg(reinterpret_cast<const float4*>(xr + i));
                rec.q = fmaf(__uint_as_float(v[i + 0]), xx.x, rec.q);
                rec.q = fmaf(__uint_as_float(v[i + 1]), xx.y, rec.q);
                rec.q = fmaf(__uint_as_float(v[i + 2]), xx.z, rec.q);
                rec.q = fmaf(__uint_as_float(v[i + 3]), xx.w, rec.q);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int col = col0 + i;
                if (col < p.N) {
                  const int cls = col - p.red.split;
                  const float s = fmaf(-0.5f, __ldg(p.red.m + cls), __uint_as_float(v[i]));
                  rec.best = red_better(rec.best, RedArgMax{s, cls});
                }
              }
            }
          }
        } else {
          epilogue_chunk<EPI, OutT>(v, row, row_ok, col0, p, rn);
        }
      }
      // all TMEM reads of this accumulator buffer are complete (wait::ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if constexpr (EPI == EPI_REDUCE) {
        // this thread's record of (row, column tile, half), then the last CTA of the row tile merges
        const int tm = t / tiles_n, tn = t - tm * tiles_n;
        const int slots = 2 * tiles_n;
        float4* rp = reinterpret_cast<float4*>(p.red.part + (static_cast<size_t>(row) * slots + (2 * tn + half)) * 8);
        rp[0] = make_float4(rec.best.v, __int_as_float(rec.best.i), rec.m1, rec.m2);
        rp[1] = make_float4(rec.sp, rec.lmax, rec.lsum, rec.q);
        volatile uint32_t* last_flag = tmem_ptr_smem + 1;
        __threadfence();
        named_bar_sync(2, kNumEpiWarps * 32);
        if (threadIdx.x == 4 * 32) *last_flag = (atomicAdd(p.red.counter + tm, 1) == tiles_n - 1) ? 1u : 0u;
        named_bar_sync(2, kNumEpiWarps * 32);
        if (*last_flag != 0u) {
          __threadfence();
          if (half == 0 && row_ok) {
            RedRec tot;
            const float4* pr = reinterpret_cast<const float4*>(p.red.part + static_cast<size_t>(row) * slots * 8);
            for (int s = 0; s < slots; ++s) {  // ascending columns: ties keep the lowest index
              const float4 a = __ldcg(pr + 2 * s), b = __ldcg(pr + 2 * s + 1);
              tot.best = red_better(tot.best, RedArgMax{a.x, __float_as_int(a.y)});
              red_top2(tot.m1, tot.m2, a.z);
              red_top2(tot.m1, tot.m2, a.w);
              tot.sp += b.x;
              red_lse(tot.lmax, tot.lsum, b.y, b.z);
              tot.q += b.w;
            }
            if (p.red.kind == RED_LOGITS) {
              p.red.scores[row] = tot.best.v;
              p.red.scores[static_cast<size_t>(p.M) + row] = red_softplus(tot.m1) + (p.N > 1 ? red_softplus(tot.m2) : 0.f);
              p.red.scores[2 * static_cast<size_t>(p.M) + row] = tot.sp;
              p.red.scores[3 * static_cast<size_t>(p.M) + row] = tot.lmax + logf(tot.lsum);
            } else {
              p.red.scores[row] = fmaf(-0.5f, tot.q, tot.best.v);
            }
            p.red.labels[row] = tot.best.i;
          }
          if (threadIdx.x == 4 * 32) p.red.counter[tm] = 0;  // ready for the next launch (stream order)
        }
        named_bar_sync(2, kNumEpiWarps * 32);  // last_flag is rewritten by the next tile
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): two SMs of a TPC form a cluster and issue ONE 256 x 256 x 16 MMA per
// step.  Each CTA stages its own 128 rows of A and only HALF of the W tile (128 of the 256 N rows); the
// pair's tensor cores read both halves, so a CTA pulls 32 KB per k-block instead of 48 KB.
//   leader (cluster rank 0): arms the full barriers for BOTH CTAs' bytes, issues every MMA, multicasts the
//                            tcgen05.commit arrivals (smem-slot release, accumulator ready) to both CTAs
//   both CTAs             : TMA producer (credits the leader's barrier), 8 epilogue warps on their own
//                            128 TMEM lanes; accumulator hand-back arrives on the leader's barrier.
// STAGED epilogue: the 4 warps that share a column half fill a [128 rows x 128 B] swizzled slab (64 bf16 or
// 32 fp32 columns), one of them issues the TMA bulk store; two slabs per half alternate.
// ------------------------------------------------------------------------------------------------
constexpr int kPairBN = 256;
constexpr int kPairStageBytes = BM * BK * 2 + (kPairBN / 2) * BK * 2;  // 16 KB A + 16 KB half of W
// RESID_SMEM (staged EPI_RESID with fp32 output): the residual slab is TMA-LOADED into the very buffer the
// output is staged in — each thread reads its own row of it, adds, and writes the sum back in place — so both
// the residual read and the pre-LayerNorm write are full-line bulk copies; 3 buffers per half, 2-deep prefetch.
// EW = number of epilogue warps (8 or 16).  Warps come in groups of four (one per TMEM lane quarter); a group owns
// 256 / (EW/4) columns of the tile.  16 warps (bf16 outputs): each group stages exactly one 64-column slab per
// tile, so one buffer per group suffices and the erf-GELU math gets four warps per scheduler instead of two.
// residual-in-smem epilogue: ring depth per column half (prefetch distance = ring - 1) and main-loop stages
// (ring 4 / 3 stages was measured: O-proj 68.5 vs 66.4 us, FFN2 166.9 vs 145.3 us — the K = 3072 main loop needs
// its 4 stages more than the epilogue needs a deeper residual prefetch)
constexpr int kResidRing = 3;
constexpr int kResidStages = 4;
// RING = 2 (K >= 2048): one residual slab of prefetch is enough when a tile's main loop lasts 48 k-blocks, and the
// smem it frees buys a fifth main-loop stage.
// H16 (fp16 single-word state, EPI_RESID_H16*): a ring buffer is 128 rows x 64 B = 8 KB instead of 16 KB (hi + lo), which
// buys the K = 768 producer a fifth and the K = 3072 producer a sixth main-loop stage.
template <bool STAGED, bool RESID_SMEM, int EW, int RING = kResidRing, bool H16 = false>
struct PairCfg {
  static constexpr int kGroups = EW / 4;
  static constexpr int kBufBytes = H16 ? kSlabBytes / 2 : kSlabBytes;
  // plain staged epilogues (no residual ring): RING == 2 selects the "deep" variant — one staging slab per group instead of two,
  // the 32 KB it frees is a sixth main-loop stage (A/B option pair_deep)
  static constexpr int kStages = RESID_SMEM ? (H16 ? (RING == 2 ? 6 : 5) : (RING == 2 ? kResidStages + 1 : kResidStages))
                                            : (STAGED ? (RING == 2 ? 6 : 5) : 6);
  static constexpr int kBufsPerGroup = RESID_SMEM ? RING : (STAGED ? ((EW == 16 || RING == 2) ? 1 : 2) : 0);
  static constexpr int kSlabRegion = kGroups * kBufsPerGroup * kBufBytes;
  static constexpr int kSmemBytes = kStages * kPairStageBytes + kSlabRegion + 1024 + 512;
};
constexpr bool epi_is_h16(int epi) { return epi == EPI_RESID_H16 || epi == EPI_RESID_H16_BETA; }
constexpr bool epi_is_hilo(int epi) { return epi == EPI_RESID_HL || epi == EPI_RESID_HL_BETA; }

template <int EPI, typename OutT, bool STAGED, int EW, int RING>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((4 + EW) * 32, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                      const __grid_constant__ CUtensorMap tmOutLo, const __grid_constant__ CUtensorMap tmResLo,
                      const GemmParams p) {
  // deferred-LayerNorm producer (bf16 hi/lo pair + statistics)
  constexpr bool kHiLo = STAGED && epi_is_hilo(EPI);
  constexpr bool kH16 = STAGED && epi_is_h16(EPI);  // deferred-LayerNorm producer, fp16 single-word state + statistics
  constexpr bool kHasBeta = EPI == EPI_RESID_HL_BETA || EPI == EPI_RESID_H16_BETA;  // else the bias already carries beta (folded at load time)
  constexpr bool kResidSmem = (STAGED && EPI == EPI_RESID && sizeof(OutT) == 4) || kHiLo || kH16;
  static_assert(EW == 8 || (EW == 16 && STAGED && !kResidSmem), "16 epilogue warps: staged, non-residual epilogues only");
  using PC = PairCfg<STAGED, kResidSmem, EW, RING, kH16>;
  constexpr int kBufBytes = PC::kBufBytes;
  constexpr int kRing = RING;
  constexpr int kStages = PC::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* slabs = smem + kStages * kPairStageBytes;  // 1024-aligned: stage size is a multiple of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slabs + PC::kSlabRegion);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* resid_bar = tmem_empty_bar + 2;  // [EW warps][kRing], used by the RESID_SMEM epilogue only
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(resid_bar + EW * kRing);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int tiles_n = (p.N + kPairBN - 1) / kPairBN;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int splits = (EPI == EPI_ACCUM && p.splits > 1) ? p.splits : 1;   // work item t = (tile t / splits, K slice t % splits)
  const int num_tiles = tiles_m * tiles_n * splits;
  const int num_kb = (p.K / BK) / splits;   // k-blocks per work item

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (STAGED) tma_prefetch_desc(&tmOut);
    if constexpr (kResidSmem) tma_prefetch_desc(&tmRes);
    if constexpr (kHiLo) {
      tma_prefetch_desc(&tmOutLo);
      tma_prefetch_desc(&tmResLo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's producer: arrive.expect_tx for both CTAs' bytes
      mbar_init(&empty_bar[s], 1);  // one multicast commit arrival per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * EW);  // epilogue warps of both CTAs
    }
    for (int i = 0; i < EW * kRing; ++i) mbar_init(&resid_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int tile = t / splits, kb0 = (t - tile * splits) * num_kb;
      const int m0 = (tile / tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = (tile % tiles_n) * kPairBN + static_cast<int>(rank) * (kPairBN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (lane == 0) {
          uint8_t* sa = smem + stage * kPairStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          if (EPI == EPI_ACCUM && p.tn) {  // transposed operands: two [64 MN x 64 k] boxes each (tensor maps over [K, M] / [K, N])
            tma_load_2d_pair(sa, &tmA, leader_full, m0, (kb0 + kb) * BK);
            tma_load_2d_pair(sa + 8192, &tmA, leader_full, m0 + 64, (kb0 + kb) * BK);
            tma_load_2d_pair(sb, &tmB, leader_full, n0, (kb0 + kb) * BK);
            tma_load_2d_pair(sb + 8192, &tmB, leader_full, n0 + 64, (kb0 + kb) * BK);
          } else {
            tma_load_2d_pair(sa, &tmA, leader_full, (kb0 + kb) * BK, m0);
            tma_load_2d_pair(sb, &tmB, leader_full, (kb0 + kb) * BK, n0);
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (rank == 0) {
      const bool tn = EPI == EPI_ACCUM && p.tn != 0;
      const uint32_t idesc = (umma_idesc_bf16(2 * BM, kPairBN) & ~(p.a_f16 ? kIdescABf16 : 0u)) | (tn ? kIdescMnMajorAB : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kPairBN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
            if (tn) {  // MN-major tiles: 64-element MN groups 8 KB apart, a k16 step = 16 lines = 2048 B = 128 address units
              const uint64_t adesc = umma_smem_desc_sw128_mn(sa, 8192u);
              const uint64_t bdesc = umma_smem_desc_sw128_mn(sa + BM * BK * 2, 8192u);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16_pair(d_tmem, adesc + 128u * k, bdesc + 128u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            } else {
            const uint64_t adesc = umma_smem_desc_sw128(sa);
            const uint64_t bdesc = umma_smem_desc_sw128(sa + BM * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    const int e = warp - 4;
    const int quarter = e & 3;
    const int half = e >> 2;  // column group of this warp: 0..EW/4-1 ("half" when EW == 8)
    constexpr int kColsPerWarp = kPairBN / (EW / 4);
    const uint32_t leader_tmem_empty0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    const uint32_t leader_tmem_empty1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    // Every epilogue warp owns its 32 rows of each slab end to end: it loads (residual), fills and TMA-stores that
    // 32-row sub-slab itself, so the warps of a CTA never wait for each other (no named barriers in the epilogue).
    [[maybe_unused]] const bool issuer = (lane == 0);
    [[maybe_unused]] constexpr int kSubSlab = kSlabBytes / 4;  // this warp's 32 rows of a slab
    [[maybe_unused]] int slab_ctr = 0;
    if constexpr (kResidSmem) {
      // flat sequence of (tile, 32-column slab) for this column half; 4 slabs per tile
      const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;
      const int total = my_tiles * 4;
      const int r_local = quarter * 32 + lane;
      // tile coordinates are divided out once per tile (not per slab): (row0, col0) of this warp's part of the current tile
      // and of the next one, which the residual prefetch (kRing - 1 slabs ahead) and the statistics prefetch reach into
      auto tile_coords = [&](int it, int& row0, int& col0) {
        const int t = pair + it * num_pairs;
        const int tm = t / tiles_n;
        row0 = tm * (2 * BM) + static_cast<int>(rank) * BM + quarter * 32;  // this warp's first row
        col0 = (t - tm * tiles_n) * kPairBN + half * kColsPerWarp;
      };
      int cur_row0, cur_col0, nxt_row0, nxt_col0;
      tile_coords(0, cur_row0, cur_col0);
      tile_coords(1, nxt_row0, nxt_col0);
      auto issue_load = [&](int k, int row0, int col) {
        const int b = k % kRing;
        uint64_t* bar = &resid_bar[e * kRing + b];
        if (p.debug & 4) {  // measurement: no residual traffic
          mbar_arrive(bar);
          return;
        }
        mbar_expect_tx(bar, kBufBytes / 4);
        uint8_t* dst = slabs + (half * kRing + b) * kBufBytes;
        if constexpr (kH16) {  // 128 x 64 B of fp16 rows, 64-byte swizzle
          tma_load_2d(dst + quarter * (kBufBytes / 4), &tmRes, bar, col, row0);
        } else if constexpr (kHiLo) {  // hi rows then lo rows, 128 x 64 B each, 64-byte swizzle
          tma_load_2d(dst + quarter * (kSubSlab / 2), &tmRes, bar, col, row0);
          tma_load_2d(dst + kSlabBytes / 2 + quarter * (kSubSlab / 2), &tmResLo, bar, col, row0);
        } else {
          tma_load_2d(dst + quarter * kSubSlab, &tmRes, bar, col, row0);
        }
      };
      static_assert(kRing - 1 <= 4, "the initial residual loads stay inside the first tile");
      if (issuer) {
        for (int j = 0; j < kRing - 1; ++j)
          if (total > j) issue_load(j, cur_row0, cur_col0 + j * 32);
      }
      [[maybe_unused]] RowNorm rin{0.f, 1.f};
      [[maybe_unused]] float2 st_sum2 = make_float2(0.f, 0.f), st_sq2 = make_float2(0.f, 0.f);
      for (int k = 0; k < total; ++k) {
        const int itile = k >> 2, sl = k & 3;
        const int acc = itile & 1;
        if (sl == 0 && k > 0) {
          cur_row0 = nxt_row0;
          cur_col0 = nxt_col0;
          tile_coords(itile + 1, nxt_row0, nxt_col0);
        }
        const int row0 = cur_row0, col = cur_col0 + sl * 32;
        const int row = row0 + lane;
        if (sl == 0) {
          if constexpr (kHiLo || kH16) {
            rin = row_norm(p.in_stats, row, row < p.M, p.ln_eps);
            st_sum2 = st_sq2 = make_float2(0.f, 0.f);
            // this row's statistics of the NEXT tile: pull the line into L1 while this tile is processed
            if (k + 4 < total && nxt_row0 + lane < p.M)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p.in_stats + static_cast<long long>(nxt_row0 + lane) * 12));
          }
          mbar_wait(&tmem_full_bar[acc], (itile >> 1) & 1u);
          tc_fence_after();
        }
        const int b = k % kRing;
        uint8_t* buf = slabs + (half * kRing + b) * kBufBytes;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                      static_cast<uint32_t>(acc * kPairBN + half * kColsPerWarp + sl * 32), v);
        mbar_wait(&resid_bar[e * kRing + b], (k / kRing) & 1u);  // this warp's residual rows have landed
        tmem_ld_wait_dep(v);
        float f[32];
        if (p.debug & 2) {
          // measurement: no epilogue math / shared-memory traffic
        } else if constexpr (kH16) {
          // residual = LN_prev(x_prev), x_prev = the fp16 state row; v = acc + bias + residual; v goes back as fp16 (saturating)
          // and the statistics are those of the ROUNDED values — what every later reader of the state sees
          uint8_t* hrow = buf + r_local * 64;
          const int sw = (r_local >> 1) & 3;
          const float ra = rin.rstd, rb = -rin.mu * rin.rstd;
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
          const float4* g4 = reinterpret_cast<const float4*>(p.in_gamma + col);
          [[maybe_unused]] const float4* e4 = reinterpret_cast<const float4*>((kHasBeta ? p.in_beta : p.bias) + col);
          uint32_t hw[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 h4 = *reinterpret_cast<const uint4*>(hrow + ((i ^ sw) << 4));
            hw[4 * i + 0] = h4.x; hw[4 * i + 1] = h4.y; hw[4 * i + 2] = h4.z; hw[4 * i + 3] = h4.w;
          }
          const float2 ra2 = make_float2(ra, ra), rb2 = make_float2(rb, rb);
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // columns 4i .. 4i+3 = fp16 words 2i, 2i+1
            float4 bb = __ldg(b4 + i);
            const float4 gg = __ldg(g4 + i);
            if constexpr (kHasBeta) {
              const float4 ee = __ldg(e4 + i);
              bb.x += ee.x;
              bb.y += ee.y;
              bb.z += ee.z;
              bb.w += ee.w;
            }
            const float2 h0 = unpack_f16x2(hw[2 * i]), h1 = unpack_f16x2(hw[2 * i + 1]);
            const float2 ra0 = __ffma2_rn(__ffma2_rn(h0, ra2, rb2), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
            const float2 ra1 = __ffma2_rn(__ffma2_rn(h1, ra2, rb2), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
            const float2 o0 = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])), ra0);
            const float2 o1 = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), ra1);
            hw[2 * i] = pack_f16x2_sat(o0.x, o0.y);
            hw[2 * i + 1] = pack_f16x2_sat(o1.x, o1.y);
            const float2 q0 = unpack_f16x2(hw[2 * i]), q1 = unpack_f16x2(hw[2 * i + 1]);
            st_sum2 = __fadd2_rn(st_sum2, __fadd2_rn(q0, q1));
            st_sq2 = __ffma2_rn(q1, q1, __ffma2_rn(q0, q0, st_sq2));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(hrow + ((i ^ sw) << 4)) = make_uint4(hw[4 * i + 0], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
          if (sl == 3 && row < p.M) {  // this thread's 128 columns of the row are done: publish their partial statistics
            float2* so = reinterpret_cast<float2*>(p.out_stats + static_cast<long long>(row) * 12) + (cur_col0 >> 7);  // 128-column slice
            *so = make_float2(st_sum2.x + st_sum2.y, st_sq2.x + st_sq2.y);
          }
        } else if constexpr (kHiLo) {
          // residual = LN_prev(x_prev) with x_prev = hi + lo; v = acc + bias + residual; then split v again
          // 64-byte swizzle: 16-byte piece i of row r sits at piece i ^ ((r >> 1) & 3) -> conflict-free row-per-lane access
          uint8_t* hrow = buf + r_local * 64;
          uint8_t* lrow = buf + kSlabBytes / 2 + r_local * 64;
          const int sw = (r_local >> 1) & 3;
          const float ra = rin.rstd, rb = -rin.mu * rin.rstd;
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
          const float4* g4 = reinterpret_cast<const float4*>(p.in_gamma + col);
          [[maybe_unused]] const float4* e4 = reinterpret_cast<const float4*>((kHasBeta ? p.in_beta : p.bias) + col);
          uint32_t hw[16], lw[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 h4 = *reinterpret_cast<const uint4*>(hrow + ((i ^ sw) << 4));
            const uint4 l4 = *reinterpret_cast<const uint4*>(lrow + ((i ^ sw) << 4));
            hw[4 * i + 0] = h4.x; hw[4 * i + 1] = h4.y; hw[4 * i + 2] = h4.z; hw[4 * i + 3] = h4.w;
            lw[4 * i + 0] = l4.x; lw[4 * i + 1] = l4.y; lw[4 * i + 2] = l4.z; lw[4 * i + 3] = l4.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // columns 4i .. 4i+3 = bf16 words 2i, 2i+1
            float4 bb = __ldg(b4 + i);
            const float4 gg = __ldg(g4 + i);
            if constexpr (kHasBeta) {
              const float4 ee = __ldg(e4 + i);
              bb.x += ee.x;
              bb.y += ee.y;
              bb.z += ee.z;
              bb.w += ee.w;
            }
            const float2 h0 = unpack_bf16x2(hw[2 * i]), l0 = unpack_bf16x2(lw[2 * i]);
            const float2 h1 = unpack_bf16x2(hw[2 * i + 1]), l1 = unpack_bf16x2(lw[2 * i + 1]);
            // packed fp32 math (FFMA2 / FADD2): two columns per instruction
            const float2 ra2 = make_float2(ra, ra), rb2 = make_float2(rb, rb);
            const float2 ra0 = __ffma2_rn(__ffma2_rn(__fadd2_rn(h0, l0), ra2, rb2), make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
            const float2 ra1 = __ffma2_rn(__ffma2_rn(__fadd2_rn(h1, l1), ra2, rb2), make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
            const float2 o0 = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])), ra0);
            const float2 o1 = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), ra1);
            f[4 * i + 0] = o0.x;
            f[4 * i + 1] = o0.y;
            f[4 * i + 2] = o1.x;
            f[4 * i + 3] = o1.y;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t ho[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 ab = make_float2(f[8 * i + 2 * j], f[8 * i + 2 * j + 1]);
              st_sum2 = __fadd2_rn(st_sum2, ab);
              st_sq2 = __ffma2_rn(ab, ab, st_sq2);
              ho[j] = pack_bf16x2(ab.x, ab.y);
              const float2 rem = __ffma2_rn(unpack_bf16x2(ho[j]), make_float2(-1.0f, -1.0f), ab);
              lo[j] = pack_bf16x2(rem.x, rem.y);
            }
            *reinterpret_cast<uint4*>(hrow + ((i ^ sw) << 4)) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
            *reinterpret_cast<uint4*>(lrow + ((i ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          if (sl == 3 && row < p.M) {  // this thread's 128 columns of the row are done: publish their partial statistics
            float2* so = reinterpret_cast<float2*>(p.out_stats + static_cast<long long>(row) * 12) + (cur_col0 >> 7);  // 128-column slice
            *so = make_float2(st_sum2.x + st_sum2.y, st_sq2.x + st_sq2.y);
          }
        } else {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
          const uint8_t* rowp = buf + r_local * 128;
          const int sw = r_local & 7;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr) bb = __ldg(b4 + i);
            const float4 rr = *reinterpret_cast<const float4*>(rowp + ((i ^ sw) << 4));
            const float2 o0 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * i + 0]), __uint_as_float(v[4 * i + 1])),
                                                    make_float2(bb.x, bb.y)), make_float2(rr.x, rr.y));
            const float2 o1 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])),
                                                    make_float2(bb.z, bb.w)), make_float2(rr.z, rr.w));
            f[4 * i + 0] = o0.x;
            f[4 * i + 1] = o0.y;
            f[4 * i + 2] = o1.x;
            f[4 * i + 3] = o1.y;
          }
          store_chunk_swizzled<float>(buf, r_local, 0, f);  // in place: this thread's own row
        }
        (void)row;
        if (sl == 3) {  // the tile's four chunks are in registers / smem: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc == 0 ? leader_tmem_empty0 : leader_tmem_empty1);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (issuer) {
          if (row0 < p.M && !(p.debug & 1)) {
            if constexpr (kH16) {
              tma_store_2d(&tmOut, buf + quarter * (kBufBytes / 4), col, row0);
            } else if constexpr (kHiLo) {
              tma_store_2d(&tmOut, buf + quarter * (kSubSlab / 2), col, row0);
              tma_store_2d(&tmOutLo, buf + kSlabBytes / 2 + quarter * (kSubSlab / 2), col, row0);
            } else {
              tma_store_2d(&tmOut, buf + quarter * kSubSlab, col, row0);
            }
          }
          tma_store_commit();
          if (k + kRing - 1 < total) {
            tma_store_wait_read<1>();  // the store of slab k-1 has drained the buffer slab k+2 will land in
            const int sn = sl + kRing - 1;  // slab k + kRing - 1: still in this tile, or slab sn - 4 of the next one
            if (sn < 4) issue_load(k + kRing - 1, cur_row0, cur_col0 + sn * 32);
            else issue_load(k + kRing - 1, nxt_row0, nxt_col0 + (sn - 4) * 32);
          }
        }
      }
      if (issuer) tma_store_wait<0>();
    } else {
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int tile = t / splits;
      const int m0 = (tile / tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = (tile % tiles_n) * kPairBN;
      const int r_local = quarter * 32 + lane;
      const int row = m0 + r_local;
      const bool row_ok = row < p.M;
      const RowNorm rn = row_norm(p.ln_stats, row, row_ok, p.ln_eps, p.ln_stats_ld);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * kPairBN + half * kColsPerWarp);
      if constexpr (STAGED) {
        constexpr int kColsPerSlab = 128 / static_cast<int>(sizeof(OutT));  // 64 bf16 / 32 fp32
        constexpr int kChunksPerSlab = kColsPerSlab / 32;
        constexpr int kSlabsPerHalf = kColsPerWarp / kColsPerSlab;
        const bool tile_rows_live = m0 + quarter * 32 < p.M;  // out-of-range rows store nothing (TMA would clip anyway)
#pragma unroll 1
        for (int sl = 0; sl < kSlabsPerHalf; ++sl) {
          const int col_slab = n0 + half * kColsPerWarp + sl * kColsPerSlab;
          if (col_slab >= p.N) break;  // uniform over the 4 warps of this half
          constexpr int kBufs = PC::kBufsPerGroup;
          uint8_t* buf = slabs + (half * kBufs + (slab_ctr % kBufs)) * kSlabBytes;
          if (issuer) tma_store_wait_read<kBufs - 1>();  // this warp's store that last read the buffer is done with it
          __syncwarp();
          // both 32-column chunks of the slab are requested from TMEM before either is consumed
          uint32_t v[kChunksPerSlab][32];
#pragma unroll
          for (int cc = 0; cc < kChunksPerSlab; ++cc)
            tmem_ld32(t_row + static_cast<uint32_t>(sl * kColsPerSlab + cc * 32), v[cc]);
#pragma unroll
          for (int cc = 0; cc < kChunksPerSlab; ++cc) tmem_ld_wait_dep(v[cc]);
#pragma unroll
          for (int cc = 0; cc < kChunksPerSlab; ++cc) {
            if (p.debug & 2) continue;
            float f[32];
            epilogue_math<EPI>(v[cc], f, row, row_ok, col_slab + cc * 32, p, true, rn);
            store_chunk_swizzled<OutT>(buf, r_local, cc, f);
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (issuer && tile_rows_live && !(p.debug & 1)) {
            if (p.out_heads > 0) {
              const int cb = col_slab >> 6;
              tma_store_3d(&tmOut, buf + quarter * kSubSlab, 0, m0 + quarter * 32, (cb % p.out_heads) * 3 + cb / p.out_heads);
            } else if constexpr (EPI == EPI_ACCUM) {
              tma_reduce_add_2d(&tmOut, buf + quarter * kSubSlab, col_slab, m0 + quarter * 32);   // out += this K slice's tile
            } else {
              tma_store_2d(&tmOut, buf + quarter * kSubSlab, col_slab, m0 + quarter * 32);
            }
            tma_store_commit();
          }
          ++slab_ctr;
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp; c += 32) {
          const int col0 = n0 + half * kColsPerWarp + c;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_row + static_cast<uint32_t>(c), v);
          tmem_ld_wait_dep(v);
          epilogue_chunk<EPI, OutT>(v, row, row_ok, col0, p, rn);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc == 0 ? leader_tmem_empty0 : leader_tmem_empty1);
    }
    if constexpr (STAGED) {
      if (issuer) tma_store_wait<0>();  // smem must outlive every bulk store that reads it
    }
    }  // !kResidSmem
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency,
// so the library loads on a CPU-only box for the symbol check).
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D tensor map over a row-major [rows, cols] matrix with row stride ld (elements); box = [box_rows, 128 B of
// columns], 128-byte swizzle, out-of-bounds elements read as zero / are not written.
int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, int elem_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  RGQA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  RGQA_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA operand must be 16-byte aligned");
  RGQA_CHECK((ld * elem_bytes) % 16 == 0, "TMA operand row stride must be a multiple of 16 bytes (ld=%lld)", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RGQA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// general form: explicit box width (elements) and swizzle span in bytes (0 = none, 64, 128)
int make_tmap_ex(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, int box_cols,
                 int elem_bytes, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  RGQA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  RGQA_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA operand must be 16-byte aligned");
  RGQA_CHECK((ld * elem_bytes) % 16 == 0, "TMA operand row stride must be a multiple of 16 bytes (ld=%lld)", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                         : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RGQA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// 3-D form (bf16): dims (d0 fastest, stride 1) x (d1, stride1 bytes) x (d2, stride2 bytes); box (b0, b1, b2); 128-byte
// swizzle.  Box extents may exceed the tensor's: out-of-range elements load as zero and are not stored.
int make_tmap_3d(CUtensorMap* map, const void* ptr, long long d0, long long d1, long long d2, long long stride1_bytes,
                 long long stride2_bytes, int b0, int b1, int b2) {
  EncodeTiledFn enc = get_encode_fn();
  RGQA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  RGQA_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA operand must be 16-byte aligned");
  RGQA_CHECK(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA strides must be multiples of 16 bytes");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(stride1_bytes), static_cast<cuuint64_t>(stride2_bytes)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(b0), static_cast<cuuint32_t>(b1), static_cast<cuuint32_t>(b2)};
  cuuint32_t estride[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RGQA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", static_cast<int>(r));
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int g_launch_sms = 0;
int launch_sms() { return (g_launch_sms > 0 && g_launch_sms < num_sms()) ? g_launch_sms : num_sms(); }
void set_launch_sms(int sms) { g_launch_sms = sms; }

namespace {

template <int BN, int EPI, typename OutT>
int launch(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p) {
  using C = Cfg<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI, OutT>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * ((EPI == EPI_ACCUM && p.splits > 1) ? p.splits : 1);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  void* tok = prof_begin(stream);
  RGQA_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kNumThreads), C::kSmemBytes, stream, tmA, tmB, p));
  prof_end(tok, stream, 2.0 * p.M * static_cast<double>(p.N) * p.K, PROF_GEMM_TC);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int BN>
int dispatch_epi(cudaStream_t s, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int epi,
                 int out_f32) {
  switch (epi) {
    case EPI_BIAS:
      return out_f32 ? launch<BN, EPI_BIAS, float>(s, a, b, p) : launch<BN, EPI_BIAS, __nv_bfloat16>(s, a, b, p);
    case EPI_GELU:
      return out_f32 ? launch<BN, EPI_GELU, float>(s, a, b, p) : launch<BN, EPI_GELU, __nv_bfloat16>(s, a, b, p);
    case EPI_TANH:
      return out_f32 ? launch<BN, EPI_TANH, float>(s, a, b, p) : launch<BN, EPI_TANH, __nv_bfloat16>(s, a, b, p);
    case EPI_RESID:
      return out_f32 ? launch<BN, EPI_RESID, float>(s, a, b, p) : launch<BN, EPI_RESID, __nv_bfloat16>(s, a, b, p);
    default:
      set_error("gemm_bf16: unknown epilogue %d", epi);
      return 1;
  }
}

template <int EPI, typename OutT, bool STAGED, int RING = kResidRing, int EW = 8>
int launch_pair(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                const CUtensorMap& tmRes, const GemmParams& p, const CUtensorMap* tmOutLo = nullptr,
                const CUtensorMap* tmResLo = nullptr) {
  // 16 epilogue warps were measured on B200 (bf16 outputs): FFN1 1043 vs 1056 TFLOP/s, QKV 1180 vs 1274 — no gain,
  // the erf-GELU epilogue is bound by its dependent FMA/MUFU chains, not by warps in flight.  8 it is.
  auto kern = gemm_bf16_pair_kernel<EPI, OutT, STAGED, EW, RING>;
  constexpr int smem =
      PairCfg<STAGED, (STAGED && ((EPI == EPI_RESID && sizeof(OutT) == 4) || epi_is_hilo(EPI) || epi_is_h16(EPI))), EW, RING, STAGED && epi_is_h16(EPI)>::kSmemBytes;
  static bool configured = false;
  if (!configured) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + kPairBN - 1) / kPairBN) * ((EPI == EPI_ACCUM && p.splits > 1) ? p.splits : 1);
  const int max_pairs = launch_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  void* tok = prof_begin(stream);
  RGQA_CUDA_OK(launch_pdl(kern, dim3(2 * pairs), dim3((4 + EW) * 32), smem, stream, tmA, tmB, tmOut, tmRes,
                          tmOutLo ? *tmOutLo : tmA, tmResLo ? *tmResLo : tmA, p));
  prof_end(tok, stream, 2.0 * p.M * static_cast<double>(p.N) * p.K, PROF_GEMM_TC);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int EPI, typename OutT>
int launch_pair_sel(cudaStream_t s, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& r,
                    const GemmParams& p, bool staged) {
  if constexpr (EPI == EPI_RESID && sizeof(OutT) == 4) {
    if (staged && p.K >= 2048) return launch_pair<EPI, OutT, true, 2>(s, a, b, o, r, p);
  }
  if constexpr (EPI == EPI_GELU && sizeof(OutT) == 2) {  // option "gelu_ew16": 16 epilogue warps (96 registers) for the erf-GELU
    if (staged && option_get(OPT_GELU_EW16) != 0) return launch_pair<EPI, OutT, true, kResidRing, 16>(s, a, b, o, r, p);
  }
  if constexpr ((EPI == EPI_GELU || EPI == EPI_BIAS) && sizeof(OutT) == 2) {  // option "pair_deep": six stages, one staging slab
    if (staged && option_get(OPT_PAIR_DEEP) != 0) return launch_pair<EPI, OutT, true, 2>(s, a, b, o, r, p);
  }
  return staged ? launch_pair<EPI, OutT, true>(s, a, b, o, r, p) : launch_pair<EPI, OutT, false>(s, a, b, o, r, p);
}

int dispatch_pair(cudaStream_t s, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& r,
                  const GemmParams& p, int epi, int out_f32, bool staged) {
  switch (epi) {
    case EPI_BIAS:
      return out_f32 ? launch_pair_sel<EPI_BIAS, float>(s, a, b, o, r, p, staged)
                     : launch_pair_sel<EPI_BIAS, __nv_bfloat16>(s, a, b, o, r, p, staged);
    case EPI_GELU:
      return out_f32 ? launch_pair_sel<EPI_GELU, float>(s, a, b, o, r, p, staged)
                     : launch_pair_sel<EPI_GELU, __nv_bfloat16>(s, a, b, o, r, p, staged);
    case EPI_TANH:
      return out_f32 ? launch_pair_sel<EPI_TANH, float>(s, a, b, o, r, p, staged)
                     : launch_pair_sel<EPI_TANH, __nv_bfloat16>(s, a, b, o, r, p, staged);
    case EPI_RESID:
      return out_f32 ? launch_pair_sel<EPI_RESID, float>(s, a, b, o, r, p, staged)
                     : launch_pair_sel<EPI_RESID, __nv_bfloat16>(s, a, b, o, r, p, staged);
    default:
      set_error("gemm_bf16: unknown epilogue %d", epi);
      return 1;
  }
}

}  // namespace

int gemm_bf16_resid_hl(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                       const void* in_hi, const void* in_lo, const float* in_stats, const float* in_gamma,
                       const float* in_beta, float eps, void* out_hi, void* out_lo, float* out_stats, int M, int N, int K) {
  RGQA_CHECK(M > 0 && K > 0 && K % BK == 0, "gemm_bf16_resid_hl: bad problem M=%d K=%d", M, K);
  RGQA_CHECK(N == 768, "gemm_bf16_resid_hl: built for N = 768 (got %d)", N);
  RGQA_CHECK(bias && in_hi && in_lo && in_stats && in_gamma && out_hi && out_lo && out_stats,
             "gemm_bf16_resid_hl: null argument");
  CUtensorMap tmA, tmB, tmOutHi, tmOutLo, tmInHi, tmInLo;
  if (make_tmap(&tmA, A, M, K, lda, BM, 2)) return 1;
  if (make_tmap(&tmB, W, N, K, ldw, kPairBN / 2, 2)) return 1;
  // 32 rows (one epilogue warp) x 32 bf16 columns (64 B), 64-byte swizzle: hi and lo halves of a 32-column slab
  if (make_tmap_ex(&tmOutHi, out_hi, M, N, N, 32, 32, 2, 64) || make_tmap_ex(&tmOutLo, out_lo, M, N, N, 32, 32, 2, 64) ||
      make_tmap_ex(&tmInHi, in_hi, M, N, N, 32, 32, 2, 64) || make_tmap_ex(&tmInLo, in_lo, M, N, N, 32, 32, 2, 64))
    return 1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.in_stats = in_stats;
  p.in_gamma = in_gamma;
  p.in_beta = in_beta;
  p.out_stats = out_stats;
  p.ln_eps = eps;
  p.ln_stats_ld = 12;
  p.out = out_hi;
  p.ldo = N;
  p.vec_ok = 1;
  p.debug = option_get(OPT_GEMM_DEBUG);
  if (in_beta != nullptr) {
    if (K >= 2048) return launch_pair<EPI_RESID_HL_BETA, __nv_bfloat16, true, 2>(stream, tmA, tmB, tmOutHi, tmInHi, p, &tmOutLo, &tmInLo);
    return launch_pair<EPI_RESID_HL_BETA, __nv_bfloat16, true>(stream, tmA, tmB, tmOutHi, tmInHi, p, &tmOutLo, &tmInLo);
  }
  if (K >= 2048) return launch_pair<EPI_RESID_HL, __nv_bfloat16, true, 2>(stream, tmA, tmB, tmOutHi, tmInHi, p, &tmOutLo, &tmInLo);
  return launch_pair<EPI_RESID_HL, __nv_bfloat16, true>(stream, tmA, tmB, tmOutHi, tmInHi, p, &tmOutLo, &tmInLo);
}

int gemm_bf16_resid_h16(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                        const void* in_h, const float* in_stats, const float* in_gamma, const float* in_beta, float eps,
                        void* out_h, float* out_stats, int M, int N, int K) {
  RGQA_CHECK(M > 0 && K > 0 && K % BK == 0, "gemm_bf16_resid_h16: bad problem M=%d K=%d", M, K);
  RGQA_CHECK(N == 768, "gemm_bf16_resid_h16: built for N = 768 (got %d)", N);
  RGQA_CHECK(bias && in_h && in_stats && in_gamma && out_h && out_stats, "gemm_bf16_resid_h16: null argument");
  CUtensorMap tmA, tmB, tmOut, tmIn;
  if (make_tmap(&tmA, A, M, K, lda, BM, 2)) return 1;
  if (make_tmap(&tmB, W, N, K, ldw, kPairBN / 2, 2)) return 1;
  // 32 rows (one epilogue warp) x 32 fp16 columns (64 B), 64-byte swizzle
  if (make_tmap_ex(&tmOut, out_h, M, N, N, 32, 32, 2, 64) || make_tmap_ex(&tmIn, in_h, M, N, N, 32, 32, 2, 64)) return 1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.in_stats = in_stats;
  p.in_gamma = in_gamma;
  p.in_beta = in_beta;
  p.out_stats = out_stats;
  p.ln_eps = eps;
  p.ln_stats_ld = 12;
  p.out = out_h;
  p.ldo = N;
  p.vec_ok = 1;
  p.debug = option_get(OPT_GEMM_DEBUG);
  if (in_beta != nullptr) {
    if (K >= 2048) return launch_pair<EPI_RESID_H16_BETA, __nv_bfloat16, true, 2>(stream, tmA, tmB, tmOut, tmIn, p);
    return launch_pair<EPI_RESID_H16_BETA, __nv_bfloat16, true>(stream, tmA, tmB, tmOut, tmIn, p);
  }
  if (K >= 2048) return launch_pair<EPI_RESID_H16, __nv_bfloat16, true, 2>(stream, tmA, tmB, tmOut, tmIn, p);
  return launch_pair<EPI_RESID_H16, __nv_bfloat16, true>(stream, tmA, tmB, tmOut, tmIn, p);
}

int gemm_bf16_accum_tn(cudaStream_t stream, const void* At, long long ldat, const void* Bt, long long ldbt, float* out, long long ldo,
                       int M, int N, int K) {
  RGQA_CHECK(At && Bt && out && M > 0 && N > 0 && K > 0 && K % BK == 0, "gemm_bf16_accum_tn: bad problem M=%d N=%d K=%d", M, N, K);
  RGQA_CHECK(option_get(OPT_GEMM_PAIR) != 0 && option_get(OPT_GEMM_STAGED) != 0, "gemm_bf16_accum_tn: needs the staged CTA-pair kernel");
  RGQA_CHECK((reinterpret_cast<uintptr_t>(out) % 16 == 0) && ((ldo * 4) % 16 == 0) && N % 32 == 0,
             "gemm_bf16_accum_tn: out must be 16-byte aligned with N a multiple of 32");
  const int num_kb = K / BK;
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + kPairBN - 1) / kPairBN);
  const int want = num_sms() / 2;
  int splits = 1;
  while (splits < 16 && tiles * splits * 2 <= want + tiles && num_kb % (splits * 2) == 0 && num_kb / (splits * 2) >= 8) splits *= 2;
  CUtensorMap tmA, tmB, tmOut;
  // [K rows, M (N) columns] row-major, box = 64 columns (128 B) x 64 k-rows, 128-byte swizzle; columns / rows past the end read as zero
  if (make_tmap_ex(&tmA, At, K, M, ldat, BK, 64, 2, 128) || make_tmap_ex(&tmB, Bt, K, N, ldbt, BK, 64, 2, 128) ||
      make_tmap(&tmOut, out, M, N, ldo, 32, 4))
    return 1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.out = out;
  p.ldo = ldo;
  p.ln_stats_ld = 12;
  p.splits = splits;
  p.tn = 1;
  p.vec_ok = 1;
  return launch_pair<EPI_ACCUM, float, true>(stream, tmA, tmB, tmOut, tmA, p);
}

int gemm_bf16_accum(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, float* out, long long ldo,
                    int M, int N, int K) {
  RGQA_CHECK(A && W && out && M > 0 && N > 0 && K > 0 && K % BK == 0, "gemm_bf16_accum: bad problem M=%d N=%d K=%d", M, N, K);
  const int num_kb = K / BK;
  const bool out_al = (reinterpret_cast<uintptr_t>(out) % 16 == 0) && ((ldo * 4) % 16 == 0);
  // the CTA-pair kernel for every weight-sized output: its epilogue adds whole 128-byte lines (TMA reduce-add), the single-CTA
  // kernel's direct epilogue one element per atomic with a row per lane (32 sectors per instruction)
  const bool use_pair = option_get(OPT_GEMM_PAIR) != 0 && option_get(OPT_GEMM_STAGED) != 0 && M >= 256 && out_al && N % 32 == 0;
  const int bn = use_pair ? kPairBN : ((N % 128 == 0 || N >= 512) ? 128 : 64);
  const int tiles = use_pair ? ((M + 2 * BM - 1) / (2 * BM)) * ((N + kPairBN - 1) / kPairBN) : ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  // split K until every SM (pair) has a work item, keeping at least 8 k-blocks per item and K / 64 a multiple of the split
  const int want = use_pair ? num_sms() / 2 : num_sms();
  int splits = 1;
  while (splits < 16 && tiles * splits * 2 <= want + tiles && num_kb % (splits * 2) == 0 && num_kb / (splits * 2) >= 8) splits *= 2;
  CUtensorMap tmA, tmB, tmOut;
  if (make_tmap(&tmA, A, M, K, lda, BM, 2)) return 1;
  if (make_tmap(&tmB, W, N, K, ldw, use_pair ? kPairBN / 2 : bn, 2)) return 1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.out = out;
  p.ldo = ldo;
  p.ln_stats_ld = 12;
  p.splits = splits;
  p.vec_ok = out_al ? 1 : 0;
  p.debug = 0;
  if (use_pair) {
    if (make_tmap(&tmOut, out, M, N, ldo, 32, 4)) return 1;
    return launch_pair<EPI_ACCUM, float, true>(stream, tmA, tmB, tmOut, tmA, p);
  }
  return bn == 128 ? launch<128, EPI_ACCUM, float>(stream, tmA, tmB, p) : launch<64, EPI_ACCUM, float>(stream, tmA, tmB, p);
}

int gemm_bf16_reduce(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                     float* out, long long ldo, int M, int N, int K, const RowReduce& red) {
  RGQA_CHECK(M > 0 && N > 0 && K > 0 && K % BK == 0, "gemm_bf16_reduce: bad problem M=%d N=%d K=%d", M, N, K);
  RGQA_CHECK(red.kind == RED_LOGITS || red.kind == RED_MAHA, "gemm_bf16_reduce: unknown reduction %d", red.kind);
  RGQA_CHECK(red.part && red.counter && red.scores && red.labels, "gemm_bf16_reduce: null workspace / output");
  RGQA_CHECK(red.kind != RED_LOGITS || bias != nullptr, "gemm_bf16_reduce: the logit reduction needs the bias");
  RGQA_CHECK(red.kind != RED_MAHA || (red.x && red.m && red.split > 0 && red.split % 128 == 0 && red.split < N && out == nullptr &&
                                      reinterpret_cast<uintptr_t>(red.x) % 16 == 0),
             "gemm_bf16_reduce: Mahalanobis reduction needs x, m, a split that is a multiple of 128 and no matrix output");
  CUtensorMap tmA, tmB;
  if (make_tmap(&tmA, A, M, K, lda, BM, 2) || make_tmap(&tmB, W, N, K, ldw, 128, 2)) return 1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.out = out;
  p.ldo = ldo;
  p.ln_stats_ld = 12;
  p.red = red;
  return launch<128, EPI_REDUCE, float>(stream, tmA, tmB, p);
}

int gemm_bf16(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
              const void* resid, long long ldr, void* out, long long ldo, int out_f32, int M, int N, int K, int epi) {
  return gemm_bf16_folded(stream, A, lda, W, ldw, bias, resid, ldr, out, ldo, out_f32, M, N, K, epi, nullptr, nullptr, 0.f);
}

// Q|K|V projection written HEAD-MAJOR: out[plane][row][64] with plane = head * 3 + {q, k, v}, every plane a contiguous
// [plane_rows, 64] bf16 array (`out` points at this call's first row inside plane 0).  A (sample, head) unit of the attention
// kernel is then three contiguous L x 128-byte blocks instead of 3 L pieces at a 4 608-byte stride, and every 32-row store of
// this epilogue one contiguous 4 KB block.  bf16, bias epilogue, CTA-pair + staged kernel only (M >= 1024, N = 3 * heads * 64).
int gemm_bf16_qkv_planes(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                         void* out, long long plane_rows, int heads, int M, int K, const float* ln_stats, float ln_eps, int a_f16) {
  const int N = 3 * heads * 64;
  RGQA_CHECK(M >= 1024 && K > 0 && K % BK == 0 && heads > 0 && plane_rows >= M, "gemm_bf16_qkv_planes: bad problem M=%d K=%d", M, K);
  RGQA_CHECK(option_get(OPT_GEMM_PAIR) != 0 && option_get(OPT_GEMM_STAGED) != 0, "gemm_bf16_qkv_planes: needs the staged CTA-pair kernel");
  RGQA_CHECK(ln_stats == nullptr || (bias != nullptr && K == 768), "gemm_bf16_qkv_planes: deferred LayerNorm needs K = 768 and the folded bias");
  CUtensorMap tmA, tmB, tmOut;
  if (make_tmap(&tmA, A, M, K, lda, BM, 2) || make_tmap(&tmB, W, N, K, ldw, kPairBN / 2, 2) ||
      make_tmap_3d(&tmOut, out, 64, M, 3LL * heads, 128, plane_rows * 128, 64, 32, 1))
    return 1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.out = out;
  p.ldo = 64;
  p.ln_stats = ln_stats;
  p.ln_eps = ln_eps;
  p.ln_stats_ld = 12;
  p.out_heads = heads;
  p.a_f16 = a_f16;
  p.vec_ok = 1;
  p.debug = option_get(OPT_GEMM_DEBUG);
  return launch_pair<EPI_BIAS, __nv_bfloat16, true>(stream, tmA, tmB, tmOut, tmA, p);
}

int gemm_bf16_folded(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                     const void* resid, long long ldr, void* out, long long ldo, int out_f32, int M, int N, int K, int epi,
                     const float* ln_stats, const float* ln_c, float ln_eps, long long ln_stats_ld, int a_f16) {
  RGQA_CHECK(ln_stats == nullptr || (bias != nullptr && K == 768 && epi != EPI_RESID),
             "gemm_bf16: deferred LayerNorm needs K = 768, the folded bias and a non-residual epilogue");
  RGQA_CHECK(M > 0 && N > 0 && K > 0, "gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  RGQA_CHECK(K % BK == 0, "gemm_bf16: K=%d must be a multiple of %d", K, BK);
  RGQA_CHECK(epi != EPI_RESID || resid != nullptr, "gemm_bf16: EPI_RESID needs a residual pointer");
  // Tile width: 256 when it does not waste more than a quarter of the last N tile, else 128; small problems
  // (B = 32: 5-14 row tiles) shrink the tile further so that more than a handful of SMs get a CTA.
  int bn = (N % 256 == 0 || N >= 1024) ? 256 : 128;
  {
    const int tm = (M + BM - 1) / BM;
    if (M < 1024) {
      if (tm * ((N + bn - 1) / bn) < 96 && bn > 128) bn = 128;
      if (tm * ((N + bn - 1) / bn) < 96 && bn > 64) bn = 64;
    }
  }
  // CTA pairs (256-row MMA, half the W traffic per CTA) once there are enough 256-row tiles to go round
  const bool use_pair = option_get(OPT_GEMM_PAIR) != 0 && bn == 256 && M >= 1024;
  const int esz = out_f32 ? 4 : 2;
  const bool out_al = (reinterpret_cast<uintptr_t>(out) % 16 == 0) && ((ldo * esz) % 16 == 0);
  const bool res_al = (resid == nullptr) || ((reinterpret_cast<uintptr_t>(resid) % 16 == 0) && ((ldr * 4) % 16 == 0));
  const bool bias_al = ((bias == nullptr) || (reinterpret_cast<uintptr_t>(bias) % 16 == 0)) &&
                       ((ln_c == nullptr) || (reinterpret_cast<uintptr_t>(ln_c) % 16 == 0));
  bool staged = use_pair && option_get(OPT_GEMM_STAGED) != 0 && out_al && res_al && bias_al;
  // the in-smem residual variant walks whole 128-column halves: ragged N falls back to the direct epilogue
  if (staged && epi == EPI_RESID && out_f32 && (N % 128 != 0)) staged = false;
  CUtensorMap tmA, tmB, tmOut, tmRes;
  if (make_tmap(&tmA, A, M, K, lda, BM, 2)) return 1;
  if (make_tmap(&tmB, W, N, K, ldw, use_pair ? kPairBN / 2 : bn, 2)) return 1;
  if (staged) {
    if (make_tmap(&tmOut, out, M, N, ldo, 32, esz)) return 1;  // one epilogue warp's rows per store
  } else {
    tmOut = tmA;  // unused by the direct epilogue
  }
  if (staged && epi == EPI_RESID && out_f32) {
    if (make_tmap(&tmRes, resid, M, N, ldr, 32, 4)) return 1;
  } else {
    tmRes = tmA;  // unused
  }
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.resid = static_cast<const float*>(resid);
  p.ldr = ldr;
  p.out = out;
  p.ldo = ldo;
  p.ln_stats = ln_stats;
  p.ln_c = ln_c;
  p.ln_eps = ln_eps;
  p.ln_stats_ld = ln_stats_ld > 0 ? ln_stats_ld : 12;
  p.vec_ok = (out_al && res_al && bias_al) ? 1 : 0;
  p.a_f16 = a_f16;
  p.debug = option_get(OPT_GEMM_DEBUG);
  if (use_pair) return dispatch_pair(stream, tmA, tmB, tmOut, tmRes, p, epi, out_f32, staged);
  if (bn == 256) return dispatch_epi<256>(stream, tmA, tmB, p, epi, out_f32);
  if (bn == 128) return dispatch_epi<128>(stream, tmA, tmB, p, epi, out_f32);
  return dispatch_epi<64>(stream, tmA, tmB, p, epi, out_f32);
}

}  // namespace rgqa
