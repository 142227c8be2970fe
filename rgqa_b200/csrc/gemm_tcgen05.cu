// rgqa_b200 — bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), TMA-fed.
//
//   out[M,N] = epilogue( A[M,K] (bf16, row stride lda) · W[N,K]ᵀ (bf16, nn.Linear layout) + bias[N] )
//
// This one kernel carries every nn.Linear of the LXMERT / UNITER forward in the bf16 path
// (reference call sites: src/lxrt/modeling.py:321-323 Q/K/V, :358 and :412 output dense,
// :399 intermediate dense, :578 pooler, src/tasks/gqa_model.py:22-27 answer head).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer   global -> 128B-swizzled smem ring (kStages deep), mbarrier tx-count
//   warp 1      MMA issuer     one lane issues tcgen05.mma (128 x BN x 16), accumulators in TMEM
//   warp 2      TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-11  epilogue       tcgen05.ld 32x32b -> registers, fused bias / erf-GELU / tanh / fp32 residual,
//                              bf16 or fp32 store; overlaps the next tile's main loop
// Tails: TMA zero-fills out-of-bounds rows of A (M tail) and W (N tail); stores are predicated.
#include <cuda.h>

#include <cstdio>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kNumEpiWarps = 8;
constexpr int kNumThreads = (4 + kNumEpiWarps) * 32;

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
  const float* bias;           // [N] fp32 or nullptr
  const float* resid;          // [M, ldr] fp32 residual stream (EPI_RESID) or nullptr
  long long ldr;
  void* out;
  long long ldo;
  int vec_ok;  // rows of out (and resid) are 16-byte aligned -> vector stores allowed
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. far below one bf16 ulp of the output);
// two MUFU ops + a 5-term Horner instead of erff's branchy path: the FFN1 epilogue has ~24 issue
// slots per output element before it, not the MMA, paces the tile.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-z * z * 1.44269504088896340736f);
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_v);
}

template <int EPI>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (EPI == EPI_GELU) return gelu_erf_fast(v);
  if constexpr (EPI == EPI_TANH) return tanhf(v);
  return v;
}

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
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BK;

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

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * BM;
      const int n0 = (t % tiles_n) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (lane == 0) {
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
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
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
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
    const int quarter = e & 3;        // == warp % 4: the TMEM lane quarter this warp may touch
    const int half = e >> 2;          // column half of the tile
    constexpr int kColsPerWarp = BN / 2;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int m0 = (t / tiles_n) * BM;
      const int n0 = (t % tiles_n) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * BN + half * kColsPerWarp);
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += 32) {
        const int col0 = n0 + half * kColsPerWarp + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(t_row + static_cast<uint32_t>(c), v);
        tmem_ld_wait_dep(v);
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        const bool full_chunk = (col0 + 32 <= p.N);
        if (full_chunk && p.vec_ok) {
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = __ldg(b4 + i);
              f[4 * i + 0] += b.x;
              f[4 * i + 1] += b.y;
              f[4 * i + 2] += b.z;
              f[4 * i + 3] += b.w;
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
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = apply_act<EPI>(f[i]);
          if (row_ok) {
            if constexpr (sizeof(OutT) == 2) {
              uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                   static_cast<long long>(row) * p.ldo + col0);
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
              float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                                     static_cast<long long>(row) * p.ldo + col0);
#pragma unroll
              for (int i = 0; i < 8; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
          }
        } else {
          // ragged tail (N % 32 != 0 or unaligned rows): scalar, fully guarded
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = col0 + i;
            if (row_ok && col < p.N) {
              float x = f[i];
              if (p.bias != nullptr) x += __ldg(p.bias + col);
              if constexpr (EPI == EPI_RESID) x += p.resid[static_cast<long long>(row) * p.ldr + col];
              x = apply_act<EPI>(x);
              if constexpr (sizeof(OutT) == 2)
                reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<long long>(row) * p.ldo + col] = __float2bfloat16_rn(x);
              else
                reinterpret_cast<float*>(p.out)[static_cast<long long>(row) * p.ldo + col] = x;
            }
          }
        }
      }
      // all TMEM reads of this accumulator buffer are complete (wait::ld above): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency,
// so the library loads on a CPU-only box for the symbol check).
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row stride ld (elements);
// box = [box_rows, 64 cols], 128-byte swizzle, out-of-bounds elements read as zero.
int make_tmap_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  RGQA_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  RGQA_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA operand must be 16-byte aligned");
  RGQA_CHECK((ld * 2) % 16 == 0, "TMA operand row stride must be a multiple of 16 bytes (ld=%lld)", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RGQA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
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

template <int BN, int EPI, typename OutT>
int launch(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p) {
  using C = Cfg<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI, OutT>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  void* tok = prof_begin(stream);
  kern<<<grid, kNumThreads, C::kSmemBytes, stream>>>(tmA, tmB, p);
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

}  // namespace

int gemm_bf16(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
              const void* resid, long long ldr, void* out, long long ldo, int out_f32, int M, int N, int K, int epi) {
  RGQA_CHECK(M > 0 && N > 0 && K > 0, "gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  RGQA_CHECK(K % BK == 0, "gemm_bf16: K=%d must be a multiple of %d", K, BK);
  RGQA_CHECK(epi != EPI_RESID || resid != nullptr, "gemm_bf16: EPI_RESID needs a residual pointer");
  // Tile width: 256 when it does not waste more than a quarter of the last N tile, else 128.
  const int bn = (N % 256 == 0 || N >= 1024) ? 256 : 128;
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16(&tmA, A, M, K, lda, BM)) return 1;
  if (make_tmap_bf16(&tmB, W, N, K, ldw, bn)) return 1;
  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.resid = static_cast<const float*>(resid);
  p.ldr = ldr;
  p.out = out;
  p.ldo = ldo;
  const int esz = out_f32 ? 4 : 2;
  const bool out_al = (reinterpret_cast<uintptr_t>(out) % 16 == 0) && ((ldo * esz) % 16 == 0);
  const bool res_al = (resid == nullptr) || ((reinterpret_cast<uintptr_t>(resid) % 16 == 0) && ((ldr * 4) % 16 == 0));
  const bool bias_al = (bias == nullptr) || (reinterpret_cast<uintptr_t>(bias) % 16 == 0);
  p.vec_ok = (out_al && res_al && bias_al) ? 1 : 0;
  return bn == 256 ? dispatch_epi<256>(stream, tmA, tmB, p, epi, out_f32)
                   : dispatch_epi<128>(stream, tmA, tmB, p, epi, out_f32);
}

}  // namespace rgqa
