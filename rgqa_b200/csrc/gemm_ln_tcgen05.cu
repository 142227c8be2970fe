// rgqa_b200 — GEMM + bias + fp32 residual + LayerNorm in ONE kernel (bf16 tensor-core path).
//
//   v      = A[M,K] · W[768,K]ᵀ + bias + resid[M,768]          (BertAttOutput / BertOutput dense + residual,
//   out    = (v - mean(v)) * rstd(v) * gamma + beta              src/lxrt/modeling.py:357-361 and :411-415)
//   written twice: bf16 [M,768] (next GEMM's A operand) and fp32 [M,768] (next residual)
//
// Why a separate kernel: LayerNorm needs whole 768-wide rows, and a 128-row tile of fp32 accumulators would need
// 768 TMEM columns (the SM has 512).  A CTA PAIR issuing cta_group::2 MMAs with M = 128 gives each CTA only 64
// rows, and the hardware packs a 64 x N accumulator into N/2 columns x 128 lanes (column half 0 in lanes 0-63,
// half 1 in lanes 64-127).  Three N = 256 MMAs per K step then hold a complete 64 x 768 row block in 384 TMEM
// columns at the full tensor rate — no cross-CTA statistics exchange, no fp32 pre-LayerNorm round trip through
// HBM and no separate LayerNorm kernel.
//
//   warp 0      TMA producer: per k-block its own 64 rows of A (8 KB) + its half (128 rows) of each of the
//               three W tiles (48 KB), 3-stage ring; completion bytes credited to the leader's barrier
//   warp 1      MMA issuer (leader CTA): 3 x 4 tcgen05.mma.cta_group::2 (128 x 256 x 16) per k-block
//   warp 2      TMEM allocator (512 columns)
//   warps 4-7   epilogue, one TMEM lane (= one row, one column half: 384 values) per thread:
//               pass 1  v = acc + bias + resid  -> back into TMEM, row sum
//               pass 1b centred sum of squares (two-pass variance, like torch.nn.LayerNorm)
//               pass 2  normalise, gamma/beta, store bf16 + fp32
//               The fp32 residual arrives by TMA: each warp owns a ring of three 4 KB buffers (32 rows x 32 fp32,
//               128B-swizzled); the first three chunks of a tile are requested while its main loop still runs and
//               every consumed buffer is refilled at once, so ~12 KB per warp stay in flight (a thread owns a ROW,
//               neighbouring lanes own different rows: direct loads would be 32 lines per instruction AND one
//               chunk deep).  Pass 2 reuses the buffers as a swizzled transpose so stores are row-contiguous.
// The accumulator is single-buffered (384 of 512 columns): the next tile's MMAs wait for pass 2, its TMA loads
// do not.  Rows >= M are zero-filled on load and never stored.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace rgqa {
namespace {

constexpr int N_LN = 768;
constexpr int kNT = 3;            // 256-wide W tiles
constexpr int BK = 64;
constexpr int kRowsPerCta = 64;   // cta_group::2, UMMA M = 128
constexpr int kABytes = kRowsPerCta * BK * 2;       // 8 KB
constexpr int kBTileBytes = 128 * BK * 2;           // this CTA's half (128 rows) of one 256-row W tile: 16 KB
constexpr int kStageBytes = kABytes + kNT * kBTileBytes;  // 56 KB
constexpr int kStages = 3;
constexpr int kEpiWarps = 4;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kRing = 3;                            // residual / staging buffers per epilogue warp
constexpr int kBufBytes = 32 * 128;                 // 32 rows x 32 fp32, 128B-swizzled (one TMA box)
constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kRing * kBufBytes + 128 * 4 + 1024 + 512;

struct LnParams {
  int M, K;
  const float* bias;
  const float* resid;
  long long ldr;
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* out_act;
  float* out_res;  // may be null
  int debug;       // measurement only: bit0 no pass-2 global stores, bit1 no pass 1b, bit2 no residual (zeros), bit3 no pass 2 at all
};

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// global column of chunk c (0..11) for a thread in column half `half`: tile nt = c / 4, 32-col piece j = c % 4
__device__ __forceinline__ int chunk_col(int c, int half) { return (c >> 2) * 256 + half * 128 + (c & 3) * 32; }
// TMEM column of the same chunk (each 256-wide tile occupies 128 columns)
__device__ __forceinline__ uint32_t chunk_tcol(int c) { return static_cast<uint32_t>((c >> 2) * 128 + (c & 3) * 32); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_ln_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmRes, const LnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* ring = smem + kStages * kStageBytes;  // [kEpiWarps][kRing][4 KB], 1024-aligned (stage size is 56 KB)
  float* stats = reinterpret_cast<float*>(ring + kEpiWarps * kRing * kBufBytes);  // [128] per-lane partials
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stats + 128);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint64_t* ring_bar = tmem_empty_bar + 1;  // [kEpiWarps][kRing]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ring_bar + kEpiWarps * kRing);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = (p.M + 2 * kRowsPerCta - 1) / (2 * kRowsPerCta);
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 2 * kEpiWarps);
    for (int i = 0; i < kEpiWarps * kRing; ++i) mbar_init(&ring_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int m0 = t * (2 * kRowsPerCta) + static_cast<int>(rank) * kRowsPerCta;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (lane == 0) {
          uint8_t* sa = smem + stage * kStageBytes;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
          tma_load_2d_pair(sa, &tmA, leader_full, kb * BK, m0);
#pragma unroll
          for (int nt = 0; nt < kNT; ++nt)
            tma_load_2d_pair(sa + kABytes + nt * kBTileBytes, &tmB, leader_full, kb * BK,
                             nt * 256 + static_cast<int>(rank) * 128);
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
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kRowsPerCta, 256);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        mbar_wait(tmem_empty_bar, (it & 1u) ^ 1u);  // the previous tile's epilogue has drained the accumulators
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            const uint64_t adesc = umma_smem_desc_sw128(sa);
#pragma unroll
            for (int nt = 0; nt < kNT; ++nt) {
              const uint64_t bdesc = umma_smem_desc_sw128(sa + kABytes + nt * kBTileBytes);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16_pair(tmem_base + static_cast<uint32_t>(nt * 128), adesc + 2u * k, bdesc + 2u * k, idesc,
                               (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar);
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
    // ------------------------------ epilogue (both CTAs): one TMEM lane per thread ------------------------------
    const int q = warp - 4;                 // TMEM lane quarter (== warp % 4)
    const int tl = q * 32 + lane;           // TMEM lane of this thread
    const int half = tl >> 6;               // column half held by this lane
    const int r_local = tl & 63;            // row within the CTA's 64-row block
    uint8_t* my_ring = ring + q * kRing * kBufBytes;
    uint64_t* my_bar = ring_bar + q * kRing;
    const uint32_t leader_tmem_empty = mapa_u32(smem_u32(tmem_empty_bar), 0);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // coalesced side of the staging transpose: lane l handles row (l >> 3) + 4*i, 16-byte piece (l & 7)
    const int crow = lane >> 3, cpiece = lane & 7;
    const int sw = lane & 7;  // swizzle key of this thread's own row (row index within the 32-row buffer = lane)
    uint32_t loads_issued = 0;  // residual chunks requested so far (buffer = n % kRing, parity = (n / kRing) & 1)
    uint32_t loads_used = 0;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      const int row_base = t * (2 * kRowsPerCta) + static_cast<int>(rank) * kRowsPerCta + (q & 1) * 32;  // first row of this warp
      const int row = row_base + lane;
      (void)r_local;
      auto issue_resid = [&](int c) {  // lane 0 only
        const uint32_t b = loads_issued % kRing;
        mbar_expect_tx(&my_bar[b], kBufBytes);
        tma_load_2d(my_ring + b * kBufBytes, &tmRes, &my_bar[b], chunk_col(c, half), row_base);
        ++loads_issued;
      };
      // the first kRing residual chunks fly while this tile's main loop is still running
      if (lane == 0 && !(p.debug & 4)) {
#pragma unroll
        for (int c = 0; c < kRing; ++c) issue_resid(c);
      }
      mbar_wait(tmem_full_bar, it & 1u);
      tc_fence_after();

      // ---------------- pass 1: v = acc + bias + resid -> TMEM, row sum ----------------
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 12; ++c) {
        uint32_t v[32];
        tmem_ld32(t_lane + chunk_tcol(c), v);
        const uint32_t b = loads_used % kRing;
        float4 rr[8];
        if (!(p.debug & 4)) {
        mbar_wait(&my_bar[b], (loads_used / kRing) & 1u);
        ++loads_used;
        const uint8_t* rowp = my_ring + b * kBufBytes + lane * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[i] = *reinterpret_cast<const float4*>(rowp + ((i ^ sw) << 4));
        __syncwarp();  // every lane has its row in registers: the buffer can be refilled
        if (lane == 0 && c + kRing < 12) issue_resid(c + kRing);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait_dep(v);
        const int col0 = chunk_col(c, half);
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(b4 + i);
          const float x0 = (__uint_as_float(v[4 * i + 0]) + bb.x) + rr[i].x;
          const float x1 = (__uint_as_float(v[4 * i + 1]) + bb.y) + rr[i].y;
          const float x2 = (__uint_as_float(v[4 * i + 2]) + bb.z) + rr[i].z;
          const float x3 = (__uint_as_float(v[4 * i + 3]) + bb.w) + rr[i].w;
          sum += (x0 + x1) + (x2 + x3);
          w[4 * i + 0] = __float_as_uint(x0);
          w[4 * i + 1] = __float_as_uint(x1);
          w[4 * i + 2] = __float_as_uint(x2);
          w[4 * i + 3] = __float_as_uint(x3);
        }
        tmem_st32(t_lane + chunk_tcol(c), w);
      }
      tmem_st_wait();
      // row mean: the other column half of this row lives in lane tl ^ 64 (another warp)
      stats[tl] = sum;
      named_bar(1, kEpiWarps * 32);
      const float mean = (sum + stats[tl ^ 64]) * (1.0f / N_LN);
      named_bar(1, kEpiWarps * 32);

      // ---------------- pass 1b: centred sum of squares ----------------
      float sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < ((p.debug & 2) ? 0 : 12); c += 2) {
        uint32_t v0[32], v1[32];
        tmem_ld32(t_lane + chunk_tcol(c), v0);
        tmem_ld32(t_lane + chunk_tcol(c + 1), v1);
        tmem_ld_wait_dep(v0);
        tmem_ld_wait_dep(v1);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d0 = __uint_as_float(v0[i]) - mean, d1 = __uint_as_float(v1[i]) - mean;
          s0 = fmaf(d0, d0, s0);
          s1 = fmaf(d1, d1, s1);
        }
        sq += s0 + s1;
      }
      stats[tl] = sq;
      named_bar(1, kEpiWarps * 32);
      const float rstd = 1.0f / sqrtf((sq + stats[tl ^ 64]) * (1.0f / N_LN) + p.eps);
      named_bar(1, kEpiWarps * 32);

      // ---------------- pass 2: normalise, gamma/beta, row-contiguous bf16 + fp32 stores ----------------
#pragma unroll 1
      for (int c = 0; c < ((p.debug & 8) ? 0 : 12); ++c) {
        uint32_t v[32];
        tmem_ld32(t_lane + chunk_tcol(c), v);
        tmem_ld_wait_dep(v);
        const int col0 = chunk_col(c, half);
        const float4* g4 = reinterpret_cast<const float4*>(p.gamma + col0);
        const float4* e4 = reinterpret_cast<const float4*>(p.beta + col0);
        uint8_t* buf = my_ring + (c % kRing) * kBufBytes;  // all residual chunks of this tile are consumed
        uint8_t* rowp = buf + lane * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 gg = __ldg(g4 + i), ee = __ldg(e4 + i);
          float4 y;
          y.x = (__uint_as_float(v[4 * i + 0]) - mean) * rstd * gg.x + ee.x;
          y.y = (__uint_as_float(v[4 * i + 1]) - mean) * rstd * gg.y + ee.y;
          y.z = (__uint_as_float(v[4 * i + 2]) - mean) * rstd * gg.z + ee.z;
          y.w = (__uint_as_float(v[4 * i + 3]) - mean) * rstd * gg.w + ee.w;
          *reinterpret_cast<float4*>(rowp + ((i ^ sw) << 4)) = y;
        }
        __syncwarp();
        const int col = col0 + cpiece * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = crow + 4 * i;
          const int r = row_base + rl;
          const float4 y = *reinterpret_cast<const float4*>(buf + rl * 128 + ((cpiece ^ (rl & 7)) << 4));
          if (r < p.M && !(p.debug & 1)) {
            if (p.out_res != nullptr) *reinterpret_cast<float4*>(p.out_res + static_cast<long long>(r) * N_LN + col) = y;
            uint2 o;
            o.x = pack_bf16x2(y.x, y.y);
            o.y = pack_bf16x2(y.z, y.w);
            *reinterpret_cast<uint2*>(p.out_act + static_cast<long long>(r) * N_LN + col) = o;
          }
        }
        __syncwarp();  // buffer (c % kRing) is rewritten kRing chunks later (or refilled by the TMA next tile)
      }
      fence_proxy_async_smem();  // staging writes (generic proxy) before the next tile's TMA refills these buffers
      __syncwarp();
      (void)row;
      // accumulators fully consumed: let the leader start the next tile's MMAs
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace

int gemm_bf16_ln(cudaStream_t stream, const void* A, long long lda, const void* W, long long ldw, const float* bias,
                 const float* resid, long long ldr, const float* gamma, const float* beta, float eps, void* out_act,
                 float* out_res, int M, int N, int K) {
  RGQA_CHECK(M > 0 && K > 0, "gemm_bf16_ln: empty problem");
  RGQA_CHECK(N == N_LN, "gemm_bf16_ln: built for N = %d (got %d)", N_LN, N);
  RGQA_CHECK(K % BK == 0, "gemm_bf16_ln: K=%d must be a multiple of %d", K, BK);
  RGQA_CHECK(bias && resid && gamma && beta && out_act, "gemm_bf16_ln: null argument");
  RGQA_CHECK((reinterpret_cast<uintptr_t>(resid) % 16 == 0) && ((ldr * 4) % 16 == 0) &&
                 (reinterpret_cast<uintptr_t>(out_act) % 16 == 0) &&
                 (out_res == nullptr || reinterpret_cast<uintptr_t>(out_res) % 16 == 0) &&
                 (reinterpret_cast<uintptr_t>(bias) % 16 == 0) && (reinterpret_cast<uintptr_t>(gamma) % 16 == 0) &&
                 (reinterpret_cast<uintptr_t>(beta) % 16 == 0),
             "gemm_bf16_ln: operands must be 16-byte aligned");
  CUtensorMap tmA, tmB, tmRes;
  if (make_tmap(&tmA, A, M, K, lda, kRowsPerCta, 2)) return 1;
  if (make_tmap(&tmB, W, N, K, ldw, 128, 2)) return 1;
  if (make_tmap(&tmRes, resid, M, N, ldr, 32, 4)) return 1;  // box: 32 rows x 32 fp32 (128 B), 128B swizzle
  LnParams p;
  p.M = M;
  p.K = K;
  p.bias = bias;
  p.resid = resid;
  p.ldr = ldr;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.out_act = static_cast<__nv_bfloat16*>(out_act);
  p.out_res = out_res;
  p.debug = option_get(OPT_GEMM_DEBUG);
  static bool configured = false;
  if (!configured) {
    RGQA_CUDA_OK(cudaFuncSetAttribute(gemm_ln_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int tiles = (M + 2 * kRowsPerCta - 1) / (2 * kRowsPerCta);
  const int max_pairs = num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  void* tok = prof_begin(stream);
  gemm_ln_pair_kernel<<<2 * pairs, kThreads, kSmemBytes, stream>>>(tmA, tmB, tmRes, p);
  prof_end(tok, stream, 2.0 * M * static_cast<double>(N) * K, PROF_GEMM_TC);
  note_launch();
  RGQA_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace rgqa
